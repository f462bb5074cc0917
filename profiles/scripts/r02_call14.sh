#!/bin/bash
# round 2, GPU call 14: the strip solver's HALO variant of the ring kernel on one GPU (world 1, GSB_DIST_FORCE_HALO=1:
# no tile is a halo tile, no flag is touched) -- timing against the plain strip path, and one ncu capture of it
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02c14; mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log; tail -2 $O/smoke.log
S="python bench.py --strips --steps 5 --warmup 3 --no-e2e --no-c4 --no-time-to-tol"
timeout 300 $S > $O/bench_strips_n1.json 2>&1
GSB_DIST_FORCE_HALO=1 timeout 300 $S > $O/bench_strips_n1_halo.json 2>&1
for f in $O/bench_*.json; do echo "$f $(grep -o '"value": [0-9.]*' $f | head -1) $(grep -o '"frac": [0-9.]*' $f | head -1)"; done > $O/summary.txt
cat $O/summary.txt
N="python bench.py --strips --steps 1 --warmup 3 --sweeps 4 --no-e2e --no-c4 --no-time-to-tol"
GSB_DIST_FORCE_HALO=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:gs_phase_ring -s 30 -c 1 -o $O/ring_halo_rhs3 -f $N > $O/ncu_halo.log 2>&1
ls -la $O
