#!/bin/bash
# round 2, GPU call 21: CG SpMV as a persistent two-stage ring of bulk copies (cg_spmv_ring) against cg_spmv_dot
# (GSB_CG_RING=0): parity tests, the reference's CG call at 566x752 and 4096^2 with page-locked vectors, kernel times
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02c21; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x -k "cg or spmv or gdf or pano or small" > $O/pytest_cg.log 2>&1; echo "pytest_cg rc=$?" | tee -a $O/pytest_cg.log
for r in 0 1; do
  GSB_CG_RING=$r timeout 300 python bench.py --other-config-only cg > $O/cg_ring$r.json 2>&1
done
timeout 200 python bench.py --other-config-only c1 > $O/c1.json 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:cg_spmv -s 460 -c 24 --csv --log-file $O/launches_cg_ring.csv python bench.py --other-config-only cg > $O/ncu_cg.log 2>&1
{
for f in $O/cg_ring*.json; do echo "$f $(grep -o '"ms_wall_median": [0-9.]*' $f | tr '\n' ' ') $(grep -o '"ms_wall_median_pageable_vectors": [0-9.]*' $f | tr '\n' ' ') $(grep -o '"max_abs_vs_reference": [0-9.e-]*' $f)"; done
echo "c1 $(grep -o '"us_per_solve_wall_median": [0-9.]*' $O/c1.json) $(grep -o '"us_device_sweep_loop": [0-9.]*' $O/c1.json)"
grep -o 'cg_spmv[^"]*"[^"]*","[^"]*","[^"]*","[^"]*","[^"]*","[^"]*","[^"]*","[^"]*","[^"]*","[^"]*"' $O/launches_cg_ring.csv | tail -6
} | tee $O/summary.txt
tail -n 3 $O/pytest_cg.log
tail -n 4 $O/launches_cg_ring.csv | cut -c1-400
exit 0
