#!/bin/bash
# round 2, GPU call 13: final library (L2 hints on, kernel 6 with its one-CTA variant, CG SpMV with gather prefetch): whole GPU suite, default line, the ncu
# capture profiles/traffic.json refers to
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02c13; mkdir -p $O
sha256sum coursecomputationalphotography_b200/libgsb200.so > $O/lib.sha
timeout 1800 python -m pytest tests -m gpu -q > $O/pytest_full.log 2>&1; echo "pytest rc=$?" >> $O/pytest_full.log
tail -8 $O/pytest_full.log
timeout 1200 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "default rc=$?"
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-time-to-tol --no-other-configs"
timeout 300 $B --kernel 3 --channels 3 > $O/bench_k3_ch3.json 2>&1
GSB_SMALL_PERSISTENT=0 timeout 120 python bench.py --other-config-only c1 > $O/c1_graph_path.json 2>&1
timeout 120 python bench.py --other-config-only c1 > $O/c1_persistent.json 2>&1
timeout 200 python bench.py --other-config-only cg > $O/cg.json 2>&1
N="python bench.py --steps 1 --warmup 3 --sweeps 4 --no-cpu-baseline --no-e2e --no-time-to-tol --no-other-configs"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_default.csv $N > $O/ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gs_sweep_fused -s 14 -c 1 -o $O/fused_rhs3 -f $N > $O/ncu_full3.log 2>&1
cat $O/c1_graph_path.json $O/c1_persistent.json | cut -c1-400
ls -la $O
