#!/bin/bash
# round 2, GPU call 7: where does the fused sweep's DRAM traffic go?  dram bytes per launch (ncu, metrics only) and
# throughput over lead x L2 eviction hints
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02c7; mkdir -p $O
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-time-to-tol --no-other-configs --kernel 5 --channels 3"
N="python bench.py --steps 1 --warmup 3 --sweeps 4 --no-cpu-baseline --no-e2e --no-time-to-tol --no-other-configs --kernel 5 --channels 3"
M="--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:gs_sweep_fused -s 14 -c 2 --csv"
for cfg in "64 0" "64 1" "64 3" "64 7" "600 0" "600 1" "600 3" "600 7" "2400 0" "2400 7" "16 7"; do
  set -- $cfg
  GSB_FUSED_LEAD=$1 GSB_FUSED_L2HINT=$2 timeout 300 $B > $O/bench_lead$1_hint$2.json 2>&1
  GSB_FUSED_LEAD=$1 GSB_FUSED_L2HINT=$2 timeout 300 ncu $M --log-file $O/ncu_lead$1_hint$2.csv $N > /dev/null 2>&1
done
for f in $O/bench_*.json; do echo "$f $(grep -o '"value": [0-9.]*' $f | head -1) $(grep -o '"frac": [0-9.]*' $f | head -1)"; done > $O/summary.txt
cat $O/summary.txt
grep -h "dram__bytes\|gpu__time" $O/ncu_*.csv | head -80
ls $O
