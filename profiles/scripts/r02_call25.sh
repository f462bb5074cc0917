#!/bin/bash
# round 2, GPU call 25: the masked blend (time-to-tol system, 5.03 M unknowns) with programmatic dependent launch forced
# on (the automatic rule switches it off above 3 M rows per launch), kernels 5 and 3
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02c25; mkdir -p $O
timeout 300 python bench.py --time-to-tol-only --no-default-eps > $O/ttt_default.json 2>&1
GSB_PDL=1 timeout 300 python bench.py --time-to-tol-only --no-default-eps > $O/ttt_pdl1_k5.json 2>&1
GSB_PDL=1 timeout 300 python bench.py --time-to-tol-only --no-default-eps --kernel 3 > $O/ttt_pdl1_k3.json 2>&1
for f in $O/ttt_*.json; do echo "$f $(grep -o '"sweeps": [0-9]*' $f | head -1) $(grep -o '"ms": [0-9.]*' $f | head -1) $(grep -o '"Gnnz_per_s": [0-9.]*' $f | head -1) $(grep -o '"kernel": [0-9]*' $f | head -1) $(grep -o '"max_abs_vs_reference": [0-9.e-]*' $f | head -1)"; done | tee $O/summary.txt
exit 0
