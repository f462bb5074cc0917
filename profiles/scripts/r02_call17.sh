#!/bin/bash
# round 2, GPU call 17: kernel 5 fence / L2-prefetch knobs (GSB_FUSED_L2HINT bits 8, 16, 32), the pipelined one-CTA
# small-system kernel (GSB_SMALL_PIPE 0/1/2) on BASELINE configs[0], upload-of-b overlap on / off in the e2e leg
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02c17; mkdir -p $O
Q="--no-e2e --no-cpu-baseline --no-time-to-tol --no-other-configs --steps 5 --warmup 3"
# parity first: the solver tests with the defaults, then the fused-sweep tests with every new knob on
timeout 600 python -m pytest tests/test_gs_gpu.py -m gpu -q -x > $O/pytest_gs.log 2>&1; echo "pytest_gs rc=$?" | tee -a $O/pytest_gs.log
GSB_FUSED_L2HINT=59 timeout 600 python -m pytest tests/test_gs_gpu.py -m gpu -q -x -k "fused or kernels_agree" > $O/pytest_fused_h59.log 2>&1; echo "pytest_fused_h59 rc=$?" | tee -a $O/pytest_fused_h59.log
for h in 3 11 19 27 59; do
  GSB_FUSED_L2HINT=$h timeout 200 python bench.py $Q > $O/bench_k5_h$h.json 2> $O/bench_k5_h$h.err
done
for p in 0 1 2; do
  GSB_SMALL_PIPE=$p timeout 200 python bench.py --other-config-only c1 > $O/c1_pipe$p.json 2>&1
done
for ov in 0 1; do
  GSB_B_OVERLAP=$ov timeout 300 python bench.py --no-cpu-baseline --no-time-to-tol --no-other-configs --e2e-steps 8 > $O/bench_e2e_ov$ov.json 2> $O/bench_e2e_ov$ov.err
done
{
for f in $O/bench_k5_h*.json; do echo "$f $(grep -o '"value": [0-9.]*' $f | head -1) $(grep -o '"frac": [0-9.]*' $f | head -1)"; done
for f in $O/c1_pipe*.json; do echo "$f $(grep -o '"us_per_solve_wall_median": [0-9.]*' $f) $(grep -o '"us_device_sweep_loop": [0-9.]*' $f) $(grep -o '"sweeps": [0-9]*' $f | head -1)"; done
for f in $O/bench_e2e_ov*.json; do echo "$f $(grep -o '"per_step_ms": [^]]*]' $f)"; done
} | tee $O/summary.txt
tail -3 $O/pytest_gs.log $O/pytest_fused_h59.log
