#!/bin/bash
# round 2, GPU call 22: after folding cg_check_max into cg_update_xr and deferring the row-length use in kernel 6:
# CG / small-system parity, both legs, an ncu capture of cg_spmv_ring at 4096^2, and the strip arm on one GPU
# (bench.py --strips: the N >= 2 bench code path incl. its e2e leg, world size 1)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02c22; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x -k "cg or spmv or small or kernels_agree" > $O/pytest_cg.log 2>&1; echo "pytest_cg rc=$?" | tee -a $O/pytest_cg.log
timeout 300 python bench.py --other-config-only cg > $O/cg.json 2>&1
timeout 200 python bench.py --other-config-only c1 > $O/c1.json 2>&1
timeout 600 python bench.py --strips --no-time-to-tol --no-c4 --steps 3 --e2e-steps 3 > $O/bench_strips_n1.json 2> $O/bench_strips_n1.err; echo "strips rc=$?" | tee -a $O/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cg_spmv_ring -s 412 -c 1 -o $O/cg_spmv_ring_4096 -f python bench.py --other-config-only cg > $O/ncu_cg.log 2>&1
{
echo "cg $(grep -o '"ms_wall_median": [0-9.]*' $O/cg.json | tr '\n' ' ') $(grep -o '"max_abs_vs_reference": [0-9.e-]*' $O/cg.json)"
echo "c1 $(grep -o '"us_per_solve_wall_median": [0-9.]*' $O/c1.json) $(grep -o '"us_device_sweep_loop": [0-9.]*' $O/c1.json)"
echo "strips $(grep -o '"value": [0-9.]*' $O/bench_strips_n1.json | head -2 | tr '\n' ' ') $(grep -o '"per_step_ms": [^]]*]' $O/bench_strips_n1.json)"
} | tee -a $O/summary.txt
tail -n 3 $O/pytest_cg.log; tail -n 5 $O/bench_strips_n1.err
exit 0
