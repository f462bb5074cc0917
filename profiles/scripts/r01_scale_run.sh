# usage (on an 8-GPU box): bash tools/scale_run.sh -> gpurun_out/scale_*.log
mkdir -p gpurun_out
nvidia-smi -L | head -8 > gpurun_out/scale_gpus.log
port=29600
for n in 8 4 2; do
  for tr in peer nccl; do
    for ce in 1 10; do
      port=$((port+1))
      (GSB_DIST_TRANSPORT=$tr timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 3 --warmup 3 --sweeps 100 --check-every $ce --no-e2e 2>&1 | tail -1) > gpurun_out/scale_n${n}_${tr}_ce${ce}.log
    done
  done
done
(timeout 300 python bench.py --steps 3 --warmup 3 --sweeps 100 --no-e2e --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/scale_n1_ce1.log
(timeout 300 python bench.py --steps 3 --warmup 3 --sweeps 100 --check-every 10 --no-e2e --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/scale_n1_ce10.log
(timeout 300 python -m pytest tests/test_dist.py -m gpu -q 2>&1 | tail -3) > gpurun_out/scale_pytest.log
# 16384^2 single channel on 8 GPUs (BASELINE configs[3])
port=$((port+1))
(timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 8 --size 16384 --channels 1 --steps 2 --warmup 3 --sweeps 50 --no-e2e 2>&1 | tail -1) > gpurun_out/scale_c4_n8.log
for f in gpurun_out/scale_*.log; do echo "== $f"; cut -c1-260 $f | tail -2; done
