#!/bin/bash
# round 2, GPU call 19: ncu captures of the current binary -- the cluster version of kernel 6 on configs[0] (where do
# 1.9 us per colour step go?) and kernel 5 with its new defaults (the capture profiles/traffic.json refers to);
# launch list of the default step; e2e with the import / analysis scratch kept in the handle
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02c19; mkdir -p $O
timeout 200 python bench.py --other-config-only c1 > $O/c1.json 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gs_small_one_cta -s 3 -c 1 -o $O/small_cluster -f python bench.py --other-config-only c1 > $O/ncu_small.log 2>&1
N="python bench.py --steps 1 --warmup 3 --sweeps 4 --no-cpu-baseline --no-e2e --no-time-to-tol --no-other-configs"
timeout 200 $N > $O/bench_short.json 2> $O/bench_short.err; echo "short rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_default.csv $N > $O/ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gs_sweep_fused -s 14 -c 1 -o $O/fused_rhs3 -f $N > $O/ncu_full3.log 2>&1
timeout 300 python bench.py --no-cpu-baseline --no-time-to-tol --no-other-configs > $O/bench_e2e.json 2> $O/bench_e2e.err
echo "e2e $(grep -o '"warmup_ms": [^]]*]' $O/bench_e2e.json) $(grep -o '"per_step_ms": [^]]*]' $O/bench_e2e.json) $(grep -o '"frac": [0-9.]*' $O/bench_e2e.json | head -1)" | tee $O/summary.txt
cut -c1-300 $O/c1.json
ls -la $O
exit 0
