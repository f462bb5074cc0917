#!/bin/bash
# round 2, GPU call 28 (2 GPUs): the multi-device tests and bench --gpus 2 with the round's final library
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02c28; mkdir -p $O
timeout 500 python -m pytest tests/test_mgpu_gpu.py tests/test_cpp_dropin_gpu.py tests/test_dist.py -m gpu -q -x > $O/pytest_mgpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_mgpu.log
tail -5 $O/pytest_mgpu.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $T bench.py --gpus 2 --steps 5 --warmup 3 --no-time-to-tol --c4-size 8192 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 rc=$?"
tail -c 400 $O/bench_n2.err
for f in $O/bench_*.json; do echo "$f $(grep -o '"value": [0-9.]*' $f | head -2 | tr '\n' ' ') $(grep -o '"parity_bitwise_vs_1gpu": [a-z]*' $f | head -2 | tr '\n' ' ') $(grep -o '"per_step_ms": [^]]*]' $f | head -1)"; done | tee $O/summary.txt
exit 0
