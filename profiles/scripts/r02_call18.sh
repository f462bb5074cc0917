#!/bin/bash
# round 2, GPU call 18: kernel 6 on a thread-block cluster (GSB_SMALL_PIPE=3) -- parity and configs[0]; kernel 5 with the
# gather-window L2 prefetch (GSB_FUSED_L2HINT bit 64) against the call-17 knobs, every setting twice (box noise ~1 %)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02c18; mkdir -p $O
Q="--no-e2e --no-cpu-baseline --no-time-to-tol --no-other-configs --steps 5 --warmup 3"
timeout 600 python -m pytest tests/test_gs_gpu.py -m gpu -q -x -k "small or kernels_agree or known_answer or stop_rule or multi_rhs" > $O/pytest_small.log 2>&1; echo "pytest_small rc=$?" | tee -a $O/pytest_small.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/smoke.log
GSB_FUSED_L2HINT=123 timeout 600 python -m pytest tests/test_gs_gpu.py -m gpu -q -x -k "fused or kernels_agree" > $O/pytest_fused_h123.log 2>&1; echo "pytest_fused_h123 rc=$?" | tee -a $O/pytest_fused_h123.log
for p in 2 3; do
  GSB_SMALL_PIPE=$p timeout 200 python bench.py --other-config-only c1 > $O/c1_pipe$p.json 2>&1
done
for rep in a b; do
  for h in 3 59 123 67 91; do
    GSB_FUSED_L2HINT=$h timeout 200 python bench.py $Q > $O/bench_k5_h${h}_$rep.json 2> $O/bench_k5_h${h}_$rep.err
  done
done
timeout 300 python bench.py --no-cpu-baseline --no-time-to-tol --no-other-configs > $O/bench_e2e.json 2> $O/bench_e2e.err
{
for f in $O/bench_k5_h*.json; do echo "$f $(grep -o '"value": [0-9.]*' $f | head -1) $(grep -o '"frac": [0-9.]*' $f | head -1)"; done
for f in $O/c1_pipe*.json; do echo "$f $(grep -o '"us_per_solve_wall_median": [0-9.]*' $f) $(grep -o '"us_device_sweep_loop": [0-9.]*' $f) $(grep -o '"sweeps": [0-9]*' $f | head -1) $(grep -o '"max_abs_vs_reference": [0-9.e-]*' $f)"; done
echo "e2e $(grep -o '"warmup_ms": [^]]*]' $O/bench_e2e.json) $(grep -o '"per_step_ms": [^]]*]' $O/bench_e2e.json)"
} | tee $O/summary.txt
tail -n 3 $O/pytest_small.log; tail -n 3 $O/pytest_fused_h123.log; tail -n 2 $O/smoke.log
exit 0
