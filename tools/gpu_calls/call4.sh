# 1-GPU call: what does the halo variant of the ring kernel cost per tile?  (world 1, no tile is a halo tile)
mkdir -p gpurun_out
: > gpurun_out/c4_halo.log
run() { # label lib force
  echo "== $1 force_halo=$3" >> gpurun_out/c4_halo.log
  (GSB_LIB_PATH=$2 GSB_PDL=0 GSB_TRACE_PHASES=1 GSB_DIST_FORCE_HALO=$3 timeout 300 python bench.py --steps 2 --warmup 3 --sweeps 100 --no-e2e --no-cpu-baseline --strips 2>&1 | grep "gsb trace" | tail -1) >> gpurun_out/c4_halo.log
}
run "plain (no halo variant)" "" 0
run "B arithmetic interior numbering" "" 1
run "A table look-up per tile" $PWD/tools/exp/libgsb200_A.so 1
run "C arithmetic, push/fence code compiled out" $PWD/tools/exp/libgsb200_C.so 1
cat gpurun_out/c4_halo.log
