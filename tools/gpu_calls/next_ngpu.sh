# N-GPU call of the next round (N = 2 or 8):  gpurun --gpus N --timeout 300 -- 'bash tools/gpu_calls/next_ngpu.sh N'
# Strip-solver parity first (short timeout: a hang must not eat the budget), then the bench with and without the
# phase trace.  Round-1 reference points: N = 2 974 Gnnz/s, N = 8 2756 (ce 1) / 3564 (ce 10).
N=${1:-2}
mkdir -p gpurun_out
export GSB_WORKER_LOG=$PWD/gpurun_out/nn_worker
if [ "$N" = "2" ]; then
  (timeout 150 python -u -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29961 tests/dist_worker.py gpu 2>&1 | grep "FAILED\|strips ok\|stop rule ok\|Error\|assert" | head -30)
  cat gpurun_out/nn_worker.rank* 2>/dev/null | tail -30
fi
unset GSB_WORKER_LOG
port=29970
for ce in 1 10; do
  port=$((port+1))
  (timeout 60 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 3 --warmup 3 --sweeps 100 --check-every $ce --no-e2e 2>&1 | grep "^{" | tail -1) > gpurun_out/nn_n${N}_ce${ce}.json
  cut -c1-150 gpurun_out/nn_n${N}_ce${ce}.json
done
port=$((port+1))
(GSB_PDL=0 GSB_TRACE_PHASES=1 timeout 60 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 2 --warmup 3 --sweeps 100 --no-e2e 2>&1 | grep "gsb trace" | sort | tail -$N) > gpurun_out/nn_n${N}_trace.log
cat gpurun_out/nn_n${N}_trace.log
