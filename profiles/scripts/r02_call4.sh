#!/bin/bash
# round 2, GPU call 4: fused sweep kernel v4 (polls issued an iteration ahead, batched publish): parity + (pubk, lead) sweep
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02c4; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_full.log 2>&1; echo "pytest rc=$?" >> $O/pytest_full.log
tail -6 $O/pytest_full.log
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-time-to-tol --no-other-configs"
timeout 300 $B --kernel 3 --channels 3 > $O/bench_k3_ch3.json 2>&1
for cfg in "1 64" "1 128" "1 300" "1 700" "1 1300" "2 128" "2 700" "2 1300" "2 2500" "4 2500"; do
  set -- $cfg
  GSB_FUSED_PUBK=$1 GSB_FUSED_LEAD=$2 timeout 300 $B --kernel 5 --channels 3 > $O/bench_k5_ch3_pubk$1_lead$2.json 2>&1
done
GSB_FUSED_DEBUG=3 timeout 300 $B --kernel 5 --channels 3 > $O/bench_k5_ch3_debug3.json 2>&1
timeout 300 $B --kernel 5 --channels 3 --check-every 10 > $O/bench_k5_ch3_ce10.json 2>&1
timeout 300 $B --kernel 5 --channels 1 > $O/bench_k5_ch1.json 2>&1
GSB_FUSED_PUBK=1 GSB_FUSED_LEAD=64 timeout 300 $B --kernel 5 --channels 1 > $O/bench_k5_ch1_pubk1_lead64.json 2>&1
timeout 300 $B --kernel 5 --channels 3 --size 1024 --sweeps 400 > $O/bench_k5_1024.json 2>&1
for f in $O/bench_*.json; do echo "$f $(grep -o '"value": [0-9.]*' $f | head -1) $(grep -o '"frac": [0-9.]*' $f | head -1)"; done > $O/summary.txt
cat $O/summary.txt
N="python bench.py --steps 1 --warmup 3 --sweeps 4 --no-cpu-baseline --no-e2e --no-time-to-tol --no-other-configs --kernel 5"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_k5.csv $N > $O/ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gs_sweep_fused -s 14 -c 1 -o $O/fused_rhs3 -f $N > $O/ncu_full3.log 2>&1
ls -la $O
