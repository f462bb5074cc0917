mkdir -p gpurun_out
for c in 3 1; do for st in 2 3 4; do for ct in 2 3 4 6; do
  echo "c=$c stages=$st ctas=$ct" >> gpurun_out/sweep6.log
  GSB_RING_STAGES=$st GSB_RING_CTAS=$ct timeout 120 python bench.py --steps 2 --warmup 3 --sweeps 30 --kernel 3 --channels $c --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); r=d['roofline']; print('   Gnnz/s %.1f frac %.3f launch_ms %.4f'%(d['value'],r['frac'],r['avg_launch_ms']))
" >> gpurun_out/sweep6.log
done; done; done
cat gpurun_out/sweep6.log
