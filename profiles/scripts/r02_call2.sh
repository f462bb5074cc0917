#!/bin/bash
# round 2, GPU call 2: warp-specialised fused sweep kernel (v2): parity tests, timing vs kernel 3, ncu; default bench line
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02c2; mkdir -p $O
timeout 900 python -m pytest tests/test_gs_gpu.py tests/test_zz_experimental_gpu.py tests/test_configs_gpu.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -5 $O/pytest.log
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-time-to-tol"
for k in 3 5; do for ch in 3 1; do
  timeout 300 $B --kernel $k --channels $ch > $O/bench_k${k}_ch${ch}.json 2>$O/bench_k${k}_ch${ch}.err
done; done
for lead in 64 1200 4000; do
  GSB_FUSED_LEAD=$lead timeout 300 $B --kernel 5 --channels 3 > $O/bench_k5_ch3_lead$lead.json 2>&1
done
GSB_RING_CTAS=3 timeout 300 $B --kernel 5 --channels 3 > $O/bench_k5_ch3_ctas3.json 2>&1
timeout 300 $B --kernel 5 --channels 3 --check-every 10 > $O/bench_k5_ch3_ce10.json 2>&1
timeout 300 $B --kernel 5 --channels 3 --size 1024 --sweeps 400 > $O/bench_k5_1024.json 2>&1
timeout 300 $B --kernel 3 --channels 3 --size 1024 --sweeps 400 > $O/bench_k3_1024.json 2>&1
timeout 300 $B --kernel 5 --channels 1 --size 1024 --sweeps 400 > $O/bench_k5_1024_ch1.json 2>&1
timeout 300 $B --kernel 4 --channels 1 --size 1024 --sweeps 400 > $O/bench_k4_1024_ch1.json 2>&1
for f in $O/bench_*.json; do echo "$f $(grep -o '"value": [0-9.]*' $f | head -1) $(grep -o '"frac": [0-9.]*' $f | head -1)"; done > $O/summary.txt
cat $O/summary.txt
# the default bench line (driver's command), with the time-to-tolerance leg against the reference golden
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "default rc=$?"
N="python bench.py --steps 1 --warmup 3 --sweeps 4 --no-cpu-baseline --no-e2e --no-time-to-tol --kernel 5"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_k5.csv $N > $O/ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gs_sweep_fused -s 14 -c 1 -o $O/fused_rhs3 -f $N > $O/ncu_full3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gs_sweep_fused -s 14 -c 1 -o $O/fused_rhs1 -f $N --channels 1 > $O/ncu_full1.log 2>&1
ls -la $O
