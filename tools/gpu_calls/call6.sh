# 2-GPU call (short): strip-solver parity incl. the real-epsilon stop at world 2; one N = 2 bench line
mkdir -p gpurun_out
export GSB_WORKER_LOG=$PWD/gpurun_out/c6_worker
(timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29931 tests/dist_worker.py gpu 2>&1 | grep "FAILED\|strips ok\|stop rule ok\|Error\|error\|assert" | head -30) > gpurun_out/c6_worker2.log
echo "worker exit: ${PIPESTATUS[0]}" >> gpurun_out/c6_worker2.log
cat gpurun_out/c6_worker2.log; cat gpurun_out/c6_worker.rank* 2>/dev/null | tail -30
unset GSB_WORKER_LOG
(timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29932 bench.py --gpus 2 --steps 3 --warmup 3 --sweeps 100 --no-e2e 2>&1 | grep "^{" | tail -1) > gpurun_out/c6_n2_ce1.json
cut -c1-200 gpurun_out/c6_n2_ce1.json
