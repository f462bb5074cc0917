#!/bin/bash
# round 2, GPU call 5 (2 GPUs): single-process multi-device mode, C++ drop-in on 2 devices, strips parity, bench --gpus 2
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02c10; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 900 python -m pytest tests/test_mgpu_gpu.py tests/test_cpp_dropin_gpu.py tests/test_dist.py -m gpu -q > $O/pytest_mgpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_mgpu.log
tail -15 $O/pytest_mgpu.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $T bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 rc=$?"
tail -c 600 $O/bench_n2.err
GSB_FUSED_END=1 timeout 300 $T bench.py --gpus 2 --steps 5 --warmup 3 --no-c4 --no-e2e --no-time-to-tol > $O/bench_n2_fend.json 2>&1
for f in $O/bench_*.json; do echo "$f $(grep -o '"value": [0-9.]*' $f | head -1) $(grep -o '"parity_bitwise_vs_1gpu": [a-z]*' $f | head -2 | tr '\n' ' ')"; done > $O/summary.txt
cat $O/summary.txt
ls -la $O
