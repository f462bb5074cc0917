#!/bin/bash
# round 2, GPU call 3: fused sweep kernel v3 (item table, control warp polls in parallel): parity, timing, sync floor, ncu
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02c3; mkdir -p $O
timeout 900 python -m pytest tests/test_gs_gpu.py -m gpu -x -q -k "kernels_agree or fused or dependent_launch or masked_blend" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -4 $O/pytest.log
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-time-to-tol"
timeout 300 $B --kernel 3 --channels 3 > $O/bench_k3_ch3.json 2>&1
timeout 300 $B --kernel 5 --channels 3 > $O/bench_k5_ch3.json 2>&1
timeout 300 $B --kernel 5 --channels 1 > $O/bench_k5_ch1.json 2>&1
timeout 300 $B --kernel 4 --channels 1 > $O/bench_k4_ch1.json 2>&1
for dbg in 1 2 3; do GSB_FUSED_DEBUG=$dbg timeout 300 $B --kernel 5 --channels 3 > $O/bench_k5_ch3_debug$dbg.json 2>&1; done
for lead in 100 1200 4000; do GSB_FUSED_LEAD=$lead timeout 300 $B --kernel 5 --channels 3 > $O/bench_k5_ch3_lead$lead.json 2>&1; done
GSB_RING_CTAS=3 timeout 300 $B --kernel 5 --channels 3 > $O/bench_k5_ch3_ctas3.json 2>&1
timeout 300 $B --kernel 5 --channels 3 --check-every 10 > $O/bench_k5_ch3_ce10.json 2>&1
timeout 300 $B --kernel 5 --channels 3 --size 1024 --sweeps 400 > $O/bench_k5_1024.json 2>&1
timeout 300 $B --kernel 3 --channels 3 --size 1024 --sweeps 400 > $O/bench_k3_1024.json 2>&1
for f in $O/bench_*.json; do echo "$f $(grep -o '"value": [0-9.]*' $f | head -1) $(grep -o '"frac": [0-9.]*' $f | head -1)"; done > $O/summary.txt
cat $O/summary.txt
N="python bench.py --steps 1 --warmup 3 --sweeps 4 --no-cpu-baseline --no-e2e --no-time-to-tol --kernel 5"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_k5.csv $N > $O/ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gs_sweep_fused -s 14 -c 1 -o $O/fused_rhs3 -f $N > $O/ncu_full3.log 2>&1
# the whole GPU suite (new this call: CG device loop, masked Gradients, 4096^2 golden test) and the driver's default line
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_full.log 2>&1; echo "pytest rc=$?" >> $O/pytest_full.log
tail -6 $O/pytest_full.log
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "default rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 3 > $O/bench_reference.json 2>&1
ls -la $O
