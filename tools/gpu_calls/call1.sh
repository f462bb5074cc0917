mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/c1_gpus.log
# bitwise: programmatic dependent launch and the L2 prefetch must not change a single bit
for sz in "4096 3 12" "2048 1 12" "1500 2 9"; do
  set -- $sz
  a=$(GSB_PDL=0 GSB_X_PREFETCH=0 timeout 200 python tools/x_hash.py $1 $2 $3 2>&1 | tail -1)
  b=$(GSB_PDL=1 GSB_X_PREFETCH=1 timeout 200 python tools/x_hash.py $1 $2 $3 2>&1 | tail -1)
  c=$(GSB_PDL=1 GSB_X_PREFETCH=1 timeout 200 python tools/x_hash.py $1 $2 $3 3 2>&1 | tail -1)
  echo "$sz :: $a" >> gpurun_out/c1_hash.log; echo "$sz :: $b" >> gpurun_out/c1_hash.log; echo "$sz :: $c" >> gpurun_out/c1_hash.log
done
cat gpurun_out/c1_hash.log
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/c1_pytest.log
cat gpurun_out/c1_pytest.log
bash tools/sweep_r01b.sh
