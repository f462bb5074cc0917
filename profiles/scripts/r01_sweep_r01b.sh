# usage (on the GPU box): bash tools/sweep_r01b.sh  -> gpurun_out/sweep_r01b.log
# PDL (programmatic dependent launch) x L2 prefetch of the gather windows, kernel 3/4, k = 3 / 1.
mkdir -p gpurun_out
out=gpurun_out/sweep_r01b.log; : > $out
run() { # label, env..., -- bench args
  label=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  echo "$label :: ${envs[*]} :: $*" >> $out
  env "${envs[@]}" timeout 150 python bench.py --steps 3 --warmup 3 --sweeps 50 --no-e2e --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c "
import json,sys
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); r=d['roofline']; print('   Gnnz/s %.1f frac %.3f launch_ms %.4f ms_step %.3f clocks %s'%(d['value'],r['frac'],r['avg_launch_ms'],d['ms_per_step'],d['clocks'].get('sm_mhz')))
    else:
        print('   ?? '+ln.strip()[:300])
" >> $out
}
run "k3 old"        GSB_PDL=0 GSB_X_PREFETCH=0 -- 
run "k3 pdl"        GSB_PDL=1 GSB_X_PREFETCH=0 -- 
run "k3 xpf1"       GSB_PDL=0 GSB_X_PREFETCH=1 -- 
run "k3 pdl+xpf1"   GSB_PDL=1 GSB_X_PREFETCH=1 -- 
run "k3 pdl+xpf2"   GSB_PDL=1 GSB_X_PREFETCH=2 -- 
run "k3 pdl+xpf1 ce10" GSB_PDL=1 GSB_X_PREFETCH=1 -- --check-every 10
run "k3 win pdl"    GSB_PDL=1 -- --kernel 4
run "k3 pdl+xpf1 3ctas" GSB_PDL=1 GSB_X_PREFETCH=1 GSB_RING_CTAS=3 --
run "k1 old"        GSB_PDL=0 -- --channels 1
run "k1 pdl"        GSB_PDL=1 -- --channels 1
run "k1 ring3 pdl+xpf1" GSB_PDL=1 GSB_X_PREFETCH=1 -- --channels 1 --kernel 3
run "1024 k1 pdl"   GSB_PDL=1 -- --channels 1 --size 1024 --sweeps 500
run "1024 k1 nopdl" GSB_PDL=0 -- --channels 1 --size 1024 --sweeps 500
cat $out
