#!/bin/bash
# round 2, GPU call 9: final library -- whole GPU suite, L2-hint variants, the driver's default line, ncu captures
# (fused sweep kernel for traffic.json; the C5 colour-phase kernel for sector amplification; the CG SpMV)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02c9; mkdir -p $O
sha256sum coursecomputationalphotography_b200/libgsb200.so > $O/lib.sha
timeout 1800 python -m pytest tests -m gpu -q > $O/pytest_full.log 2>&1; echo "pytest rc=$?" >> $O/pytest_full.log
tail -8 $O/pytest_full.log
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-time-to-tol --no-other-configs"
for h in 0 1 3 7; do GSB_FUSED_L2HINT=$h timeout 300 $B --kernel 5 --channels 3 > $O/bench_k5_ch3_hint$h.json 2>&1; done
timeout 300 $B --kernel 3 --channels 3 > $O/bench_k3_ch3.json 2>&1
for f in $O/bench_*.json; do echo "$f $(grep -o '"value": [0-9.]*' $f | head -1) $(grep -o '"frac": [0-9.]*' $f | head -1)"; done > $O/summary.txt
cat $O/summary.txt
timeout 1200 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "default rc=$?"
N="python bench.py --steps 1 --warmup 3 --sweeps 4 --no-cpu-baseline --no-e2e --no-time-to-tol --no-other-configs"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_default.csv $N > $O/ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gs_sweep_fused -s 14 -c 1 -o $O/fused_rhs3 -f $N > $O/ncu_full3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gs_phase -s 40 -c 1 -o $O/c5_phase -f python bench.py --other-config-only c5 > $O/ncu_c5.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cg_spmv_dot -s 20 -c 1 -o $O/cg_spmv -f python bench.py --other-config-only cg > $O/ncu_cg.log 2>&1
ls -la $O
