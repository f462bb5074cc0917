#!/bin/bash
# round 2, GPU call 27: the round's library as committed (kernel 6 cluster version, CG ring SpMV, kernel 2 L2 hints, kernel 5 defaults and policy): whole GPU suite, smoke, default line, reference arm
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02c27; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/pytest_full.log 2>&1; echo "pytest rc=$?" >> $O/pytest_full.log
tail -6 $O/pytest_full.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log; tail -2 $O/smoke.log
timeout 1200 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "default rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 3 > $O/bench_reference.json 2>&1
ls -la $O
