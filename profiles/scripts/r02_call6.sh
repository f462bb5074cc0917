#!/bin/bash
# round 2, GPU call 6: fused sweep kernel v5 (v4 + a waiting CTA retires its current tile first): full suite, leads, default line, ncu
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02c6; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -x -q > $O/pytest_full.log 2>&1; echo "pytest rc=$?" >> $O/pytest_full.log
tail -6 $O/pytest_full.log
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-time-to-tol --no-other-configs"
timeout 300 $B --kernel 3 --channels 3 > $O/bench_k3_ch3.json 2>&1
for lead in 32 64 128 296 600; do GSB_FUSED_LEAD=$lead timeout 300 $B --kernel 5 --channels 3 > $O/bench_k5_ch3_lead$lead.json 2>&1; done
timeout 300 $B --channels 3 --check-every 10 > $O/bench_auto_ch3_ce10.json 2>&1
timeout 300 $B --channels 1 > $O/bench_auto_ch1.json 2>&1
timeout 300 $B --channels 3 --size 1024 --sweeps 400 > $O/bench_auto_1024.json 2>&1
for f in $O/bench_*.json; do echo "$f $(grep -o '"value": [0-9.]*' $f | head -1) $(grep -o '"frac": [0-9.]*' $f | head -1) $(grep -o '"kernel": [0-9]*' $f | head -1)"; done > $O/summary.txt
cat $O/summary.txt
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "default rc=$?"
N="python bench.py --steps 1 --warmup 3 --sweeps 4 --no-cpu-baseline --no-e2e --no-time-to-tol --no-other-configs"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_default.csv $N > $O/ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gs_sweep_fused -s 14 -c 1 -o $O/fused_rhs3 -f $N > $O/ncu_full3.log 2>&1
sha256sum coursecomputationalphotography_b200/libgsb200.so > $O/lib.sha
ls -la $O
