# 2-GPU call: halo-variant cost at world 1, strip-solver parity at world 2 (with tracebacks), N = 2 bench + phase trace
mkdir -p gpurun_out
: > gpurun_out/c5_halo.log
run1() { # label force
  echo "== $1 force_halo=$2" >> gpurun_out/c5_halo.log
  (GSB_PDL=0 GSB_TRACE_PHASES=1 GSB_DIST_FORCE_HALO=$2 timeout 300 python bench.py --steps 2 --warmup 3 --sweeps 100 --no-e2e --no-cpu-baseline --strips 2>&1 | grep "gsb trace" | tail -1) >> gpurun_out/c5_halo.log
}
run1 "world 1 plain" 0
run1 "world 1 halo variant (role split)" 1
cat gpurun_out/c5_halo.log
export GSB_WORKER_LOG=$PWD/gpurun_out/c5_worker
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29911 tests/dist_worker.py gpu 2>&1 | grep -v "^W10\|^$" | grep -B2 -A25 "FAILED:\|strips ok\|stop rule ok" | head -80) > gpurun_out/c5_worker2.log
cat gpurun_out/c5_worker2.log
unset GSB_WORKER_LOG
port=29920
run() { # label envs -- args
  label=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  port=$((port+1))
  (env "${envs[@]}" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 2 --steps 3 --warmup 3 --sweeps 100 --no-e2e "$@" 2>&1 | grep "gsb trace\|^{" | tail -3) > gpurun_out/c5_$label.log
  grep "gsb trace" gpurun_out/c5_$label.log | tail -2
  python - <<PY
import json
ln=[l for l in open("gpurun_out/c5_$label.log").read().strip().splitlines() if l.startswith("{")]
try:
    d=json.loads(ln[-1]); print("$label: Gnnz/s %.1f  ms/step %.3f  launch_ms %.4f | %s"%(d["value"],d["ms_per_step"],d["roofline"]["avg_launch_ms"],d["config"].get("stop_rule_allreduce","")))
except Exception as e:
    print("$label: ??", ln[-1][:300] if ln else "no output")
PY
}
run auto_ce1   GSB_X=0 --
run pdl1_ce1   GSB_PDL=1 --
run trace_ce1  GSB_PDL=0 GSB_TRACE_PHASES=1 --
run auto_ce10  GSB_X=0 -- --check-every 10
run k1_ce1     GSB_X=0 -- --channels 1
