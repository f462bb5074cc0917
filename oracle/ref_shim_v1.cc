// TEST INFRASTRUCTURE ONLY.  Builds the lab3 (v1) reference container in place:
//   /root/reference/labs/lab3/src/OpenCVHW1/sparse-matrix.h
// Flags (oracle/Makefile): -std=c++17 -O2 -include cstring -include cmath -fno-access-control
#include REF_V1_HEADER
#define REFNS ::
#define SHIM(name) ref1_##name
#include "ref_shim.inc"
