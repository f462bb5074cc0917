#!/bin/bash
# round 2, GPU call 29 (last): allocation-free steady state (fixed scratch slots, scan scratch) -- whole GPU suite, smoke, default line
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02c29; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q > $O/pytest_full.log 2>&1; echo "pytest rc=$?" >> $O/pytest_full.log
tail -4 $O/pytest_full.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log; tail -2 $O/smoke.log
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "default rc=$?"
grep -o '"device_allocs_and_frees_per_step": [0-9.]*' $O/bench_default.json; grep -o '"per_step_ms": [^]]*]' $O/bench_default.json | head -1
exit 0
