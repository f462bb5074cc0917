# 1-GPU call: strip-solver worker at world 1 (with traceback), the canonical bench line, and the strip path at
# world 1 with / without the halo kernel variant (phase trace) to separate kernel cost from exchange cost
mkdir -p gpurun_out
export GSB_WORKER_LOG=gpurun_out/c3_worker
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29811 tests/dist_worker.py gpu 2>&1 | grep -v "^W10\|^$" | tail -60) > gpurun_out/c3_worker1.log
cat gpurun_out/c3_worker1.log | tail -30
(timeout 900 python bench.py 2>gpurun_out/c3_bench.err | tail -1) > gpurun_out/c3_bench.json
cut -c1-1500 gpurun_out/c3_bench.json
for v in "0 0" "1 0" "1 1"; do
  set -- $v
  echo "== strips=$1 force_halo=$2" >> gpurun_out/c3_strips.log
  if [ "$1" = "1" ]; then extra="--strips"; else extra=""; fi
  (GSB_PDL=0 GSB_TRACE_PHASES=1 GSB_DIST_FORCE_HALO=$2 timeout 300 python bench.py --steps 3 --warmup 3 --sweeps 100 --no-e2e --no-cpu-baseline $extra 2>&1 | grep "gsb trace\|^{" | cut -c1-400 | tail -4) >> gpurun_out/c3_strips.log
done
cat gpurun_out/c3_strips.log
