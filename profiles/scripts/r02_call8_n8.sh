#!/bin/bash
# round 2, GPU call 8 (8 GPUs): parity at world 8 (both modes), the driver's scaling line at N = 8 and 4, end-of-sweep variants, phase trace
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02c8; mkdir -p $O
timeout 900 python -m pytest tests/test_mgpu_gpu.py tests/test_cpp_dropin_gpu.py tests/test_dist.py -m gpu -q -k "8" > $O/pytest_n8.log 2>&1; echo "pytest rc=$?" >> $O/pytest_n8.log
tail -8 $O/pytest_n8.log
T8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
T4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522"
timeout 600 $T8 bench.py --gpus 8 --steps 10 --warmup 3 > $O/bench_n8.json 2> $O/bench_n8.err; echo "bench n8 rc=$?"
tail -c 400 $O/bench_n8.err
GSB_FUSED_END=0 timeout 300 $T8 bench.py --gpus 8 --steps 10 --warmup 3 --no-c4 --no-e2e > $O/bench_n8_nofend.json 2>&1
GSB_PDL=0 timeout 300 $T8 bench.py --gpus 8 --steps 10 --warmup 3 --no-c4 --no-e2e > $O/bench_n8_pdl0.json 2>&1
timeout 300 $T8 bench.py --gpus 8 --steps 10 --warmup 3 --no-c4 --no-e2e --check-every 10 > $O/bench_n8_ce10.json 2>&1
GSB_TRACE_PHASES=1 GSB_PDL=0 timeout 300 $T8 bench.py --gpus 8 --steps 2 --warmup 3 --no-c4 --no-e2e > $O/bench_n8_trace.json 2> $O/trace_n8.txt
timeout 600 $T4 bench.py --gpus 4 --steps 10 --warmup 3 --no-e2e > $O/bench_n4.json 2> $O/bench_n4.err
for f in $O/bench_*.json; do echo "$f $(grep -o '"value": [0-9.]*' $f | head -1) $(grep -o '"parity_bitwise_vs_1gpu": [a-z]*' $f | tr '\n' ' ')"; done > $O/summary.txt
cat $O/summary.txt; grep "gsb trace" $O/trace_n8.txt | head -8
ls -la $O
