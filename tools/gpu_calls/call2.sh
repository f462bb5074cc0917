# 2-GPU call: strip-solver parity (both halo transports, both stop-rule reductions) + N = 2 bench variants
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/c2_gpus.log
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 tests/dist_worker.py gpu 2>&1 | tail -40) > gpurun_out/c2_worker2.log
cat gpurun_out/c2_worker2.log
(timeout 600 python -m pytest tests/test_dist.py -m gpu -x -q 2>&1 | tail -30) > gpurun_out/c2_pytest.log
cat gpurun_out/c2_pytest.log
port=29720
run() { # label envs -- args
  label=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  port=$((port+1))
  (env "${envs[@]}" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 2 --steps 3 --warmup 3 --sweeps 100 --no-e2e "$@" 2>&1 | tail -1) > gpurun_out/c2_$label.log
  python - <<PY
import json
ln=open("gpurun_out/c2_$label.log").read().strip().splitlines()[-1]
try:
    d=json.loads(ln); print("$label: Gnnz/s %.1f  ms/step %.3f  launch_ms %.4f  %s | %s"%(d["value"],d["ms_per_step"],d["roofline"]["avg_launch_ms"],d["config"].get("halo","")[:40],d["config"].get("stop_rule_allreduce","")))
except Exception as e:
    print("$label: ??", ln[:400])
PY
}
run new_ce1        GSB_X=0 --
run new_nopdl_ce1  GSB_PDL=0 --
run new_pdl1_ce1   GSB_PDL=1 --
run old_ce1        GSB_PDL=0 GSB_DIST_EPS=nccl --
run new_ce10       GSB_X=0 -- --check-every 10
run nccl_ce1       GSB_PDL=0 GSB_DIST_TRANSPORT=nccl --
