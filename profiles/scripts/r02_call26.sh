#!/bin/bash
# round 2, GPU call 26: kernel 2 with L2 residency hints (GSB_STAGED_L2HINT) on BASELINE configs[4] (n = 1e7, 27 entries
# per row, random columns: x = 80 MB fits L2, the CSR stream does not) -- parity, then off / on
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02c26; mkdir -p $O
GSB_STAGED_L2HINT=1 timeout 300 python -m pytest tests/test_gs_gpu.py -m gpu -q -x -k "kernels_agree or multicolor or unsymmetric or zero_diagonal" > $O/pytest_hint.log 2>&1; echo "pytest_hint rc=$?" | tee -a $O/pytest_hint.log
for h in 0 1; do
  GSB_STAGED_L2HINT=$h timeout 300 python bench.py --other-config-only c5 > $O/c5_hint$h.json 2>&1
done
GSB_STAGED_L2HINT=1 timeout 300 python bench.py --other-config-only c5 --c5-n 5000000 > $O/c5_n5e6_hint1.json 2>&1
GSB_STAGED_L2HINT=0 timeout 300 python bench.py --other-config-only c5 --c5-n 5000000 > $O/c5_n5e6_hint0.json 2>&1
for f in $O/c5_*.json; do echo "$f $(grep -o '"ms": [0-9.]*' $f | head -1) $(grep -o '"Gnnz_per_s": [0-9.]*' $f | head -1) $(grep -o '"frac_of_hbm_peak": [0-9.]*' $f) $(grep -o '"kernel": [0-9]*' $f | head -1) $(grep -o '"n_colors": [0-9]*' $f) $(grep -o '"max_abs_vs_xstar_after_20_sweeps": [0-9.e-]*' $f)"; done | tee $O/summary.txt
tail -n 3 $O/pytest_hint.log
exit 0
