#!/bin/bash
# round 2, GPU call 30: ncu capture of kernel 5 from the final sources (the capture profiles/traffic.json refers to)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02c30; mkdir -p $O
N="python bench.py --steps 1 --warmup 3 --sweeps 4 --no-cpu-baseline --no-e2e --no-time-to-tol --no-other-configs"
timeout 100 ncu --set full --clock-control none --import-source on -k regex:gs_sweep_fused -s 14 -c 1 -o $O/fused_rhs3 -f $N > $O/ncu_full3.log 2>&1
ls -la $O
exit 0
