# usage (on the GPU box): bash tools/sweep_ring.sh  -> gpurun_out/sweep_ring.log
mkdir -p gpurun_out
out=gpurun_out/sweep_ring.log; : > $out
for cfg in "4 1 2" "4 1 3" "4 1 4" "4 3 2" "4 3 3" "3 3 2" "3 1 2"; do
  set -- $cfg
  echo "kernel=$1 channels=$2 stages=$3" >> $out
  GSB_RING_STAGES=$3 timeout 120 python bench.py --steps 2 --warmup 3 --sweeps 30 --kernel $1 --channels $2 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); r=d['roofline']; print('   Gnnz/s %.1f frac %.3f launch_ms %.4f'%(d['value'],r['frac'],r['avg_launch_ms']))
" >> $out
done
cat $out
