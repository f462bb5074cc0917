#!/bin/bash
# round 2, GPU call 23: kernel 6 cluster version with the step table and a thread count that gives one pass per colour
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02c23; mkdir -p $O
timeout 300 python -m pytest tests/test_gs_gpu.py -m gpu -q -x -k "small or kernels_agree or known_answer or stop_rule or multi_rhs or zero_diagonal" > $O/pytest_small.log 2>&1; echo "pytest_small rc=$?" | tee -a $O/pytest_small.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/smoke.log
timeout 200 python bench.py --other-config-only c1 > $O/c1.json 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gs_small_cluster -s 3 -c 1 -o $O/small_cluster_v3 -f python bench.py --other-config-only c1 > $O/ncu_small.log 2>&1
echo "c1 $(grep -o '"us_per_solve_wall_median": [0-9.]*' $O/c1.json) $(grep -o '"us_device_sweep_loop": [0-9.]*' $O/c1.json) $(grep -o '"sweeps": [0-9]*' $O/c1.json | head -1) $(grep -o '"max_abs_vs_reference": [0-9.e-]*' $O/c1.json)" | tee $O/summary.txt
tail -n 3 $O/pytest_small.log; tail -n 2 $O/smoke.log
exit 0
