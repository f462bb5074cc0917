#!/bin/bash
# round 2, GPU call 20: kernel 6 cluster version with st.async + mbarriers (no fence / cluster barrier in the loop):
# parity, configs[0] against the cluster.sync() version (GSB_SMALL_PIPE=4); per-kernel times of the CG leg (launch list)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02c20; mkdir -p $O
timeout 300 python -m pytest tests/test_gs_gpu.py -m gpu -q -x -k "small or kernels_agree or known_answer or stop_rule or multi_rhs or zero_diagonal" > $O/pytest_small.log 2>&1; echo "pytest_small rc=$?" | tee -a $O/pytest_small.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/smoke.log
for p in 4 3; do
  GSB_SMALL_PIPE=$p timeout 200 python bench.py --other-config-only c1 > $O/c1_pipe$p.json 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gs_small_cluster -s 3 -c 1 -o $O/small_cluster_v2 -f python bench.py --other-config-only c1 > $O/ncu_small.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:cg_ -c 1700 --csv --log-file $O/launches_cg.csv python bench.py --other-config-only cg > $O/ncu_cg.log 2>&1
{
for f in $O/c1_pipe*.json; do echo "$f $(grep -o '"us_per_solve_wall_median": [0-9.]*' $f) $(grep -o '"us_device_sweep_loop": [0-9.]*' $f) $(grep -o '"sweeps": [0-9]*' $f | head -1) $(grep -o '"max_abs_vs_reference": [0-9.e-]*' $f)"; done
} | tee $O/summary.txt
tail -n 3 $O/pytest_small.log; tail -n 2 $O/smoke.log; tail -n 3 $O/ncu_cg.log | cut -c1-300
exit 0
