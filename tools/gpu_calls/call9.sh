# 8-GPU call (short): N = 8 strong scaling of the bench workload, stop rule every sweep and every 10th
mkdir -p gpurun_out
(timeout 45 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29951 bench.py --gpus 8 --steps 3 --warmup 3 --sweeps 100 --no-e2e 2>&1 | grep "^{" | tail -1) > gpurun_out/c9_n8_ce1.json
cut -c1-160 gpurun_out/c9_n8_ce1.json
(timeout 40 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29952 bench.py --gpus 8 --steps 3 --warmup 3 --sweeps 100 --check-every 10 --no-e2e 2>&1 | grep "^{" | tail -1) > gpurun_out/c9_n8_ce10.json
cut -c1-160 gpurun_out/c9_n8_ce10.json
