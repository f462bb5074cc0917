#!/bin/bash
# round 2, GPU call 11 (8 GPUs): final library at N = 8 -- the two host-API tests that were below the threshold,
# the default scaling line (separate end-of-sweep kernel) with the masked time-to-tol, config4 and e2e legs, N = 4
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02c11; mkdir -p $O
timeout 600 python -m pytest tests/test_mgpu_gpu.py tests/test_cpp_dropin_gpu.py -m gpu -q -k "8" > $O/pytest_n8.log 2>&1; echo "pytest rc=$?" >> $O/pytest_n8.log
tail -5 $O/pytest_n8.log
T8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
T4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522"
timeout 600 $T8 bench.py --gpus 8 --steps 10 --warmup 3 > $O/bench_n8.json 2> $O/bench_n8.err; echo "bench n8 rc=$?"
tail -c 300 $O/bench_n8.err
GSB_RING_CTAS=3 timeout 300 $T8 bench.py --gpus 8 --steps 10 --warmup 3 --no-c4 --no-e2e --no-time-to-tol > $O/bench_n8_ctas3.json 2>&1
timeout 300 $T4 bench.py --gpus 4 --steps 10 --warmup 3 --no-c4 --no-e2e --no-time-to-tol > $O/bench_n4.json 2> $O/bench_n4.err
for f in $O/bench_*.json; do echo "$f $(grep -o '"value": [0-9.]*' $f | head -1) $(grep -o '"parity_bitwise_vs_1gpu": [a-z]*' $f | tr '\n' ' ')"; done > $O/summary.txt
cat $O/summary.txt
ls -la $O
