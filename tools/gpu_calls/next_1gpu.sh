# First 1-GPU call of the next round: everything that changed after the last GPU call of round 1 is unmeasured.
#   gpurun --timeout 900 -- 'bash tools/gpu_calls/next_1gpu.sh'
mkdir -p gpurun_out
(timeout 700 python -m pytest tests -m gpu -q 2>&1 | tail -25) > gpurun_out/n1_pytest.log
cat gpurun_out/n1_pytest.log
(timeout 600 python bench.py 2>gpurun_out/n1_bench.err | tail -1) > gpurun_out/n1_bench.json
python - <<'PY'
import json
d = json.loads(open("gpurun_out/n1_bench.json").read().strip().splitlines()[-1])
e = d["e2e"]
print("value %.1f frac %.3f | e2e %.1f Gnnz/s: step %.0f ms, import %.0f ms, analysis %.0f ms | cpu %.2f" %
      (d["value"], d["roofline"]["frac"], e["value"], e["ms_per_step"], e["import_ms"], e["analysis_ms"],
       d["cpu_baseline"]["value"]))
PY
# strip path at world 1 with / without the halo variant: must be equal (211 vs 211.5 us per phase in round 1)
for f in 0 1; do
  (GSB_PDL=0 GSB_TRACE_PHASES=1 GSB_DIST_FORCE_HALO=$f timeout 300 python bench.py --steps 2 --warmup 3 --sweeps 100 --no-e2e --no-cpu-baseline --strips 2>&1 | grep "gsb trace" | tail -1)
done
