#!/bin/bash
# round 2, GPU call 15: the masked blend (time-to-tol system, compact unknowns) under kernel 3 and kernel 5
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02c15; mkdir -p $O
for k in 3 5 0; do timeout 300 python bench.py --time-to-tol-only --no-default-eps --kernel $k > $O/ttt_k$k.json 2>&1; done
for f in $O/ttt_*.json; do echo "$f $(grep -o '"sweeps": [0-9]*' $f | head -1) $(grep -o '"ms": [0-9.]*' $f | head -1) $(grep -o '"Gnnz_per_s": [0-9.]*' $f | head -1) $(grep -o '"kernel": [0-9]*' $f | head -1)"; done | tee $O/summary.txt
