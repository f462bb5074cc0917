# 1-GPU call: full GPU test suite, the canonical bench line, ncu launch list + full capture of the ring kernel
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -12) > gpurun_out/c8_pytest.log
cat gpurun_out/c8_pytest.log
(timeout 600 python bench.py 2>gpurun_out/c8_bench.err | tail -1) > gpurun_out/c8_bench.json
cut -c1-300 gpurun_out/c8_bench.json
CMD="python bench.py --steps 2 --warmup 3 --sweeps 10 --no-e2e --no-cpu-baseline"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01b.csv $CMD > gpurun_out/c8_ncu1.log 2>&1
tail -2 gpurun_out/c8_ncu1.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gs_phase_ring -s 30 -c 2 -o gpurun_out/prof_ring_rhs3_b $CMD > gpurun_out/c8_ncu2.log 2>&1
tail -2 gpurun_out/c8_ncu2.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gs_phase_ring -s 30 -c 2 -o gpurun_out/prof_ringwin_rhs1_b $CMD --channels 1 > gpurun_out/c8_ncu3.log 2>&1
tail -2 gpurun_out/c8_ncu3.log
ls -la gpurun_out/*.ncu-rep
