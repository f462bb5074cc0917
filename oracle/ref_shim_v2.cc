// TEST INFRASTRUCTURE ONLY.  Builds the lab8/project (v2) reference container in place:
//   /root/reference/labs/lab8/src/OpenCVHW1/sparse-matrix.h
// (project/src/PhotoMontage/sparse-matrix.h is identical modulo whitespace.)
// Flags (oracle/Makefile): -std=c++17 -O2 -include cstring -include cmath -fno-access-control
#define USE_NAME_SPACE refv2
#include REF_V2_HEADER
#define REFNS refv2::
#define SHIM(name) ref2_##name
#define REF_HAS_V2 1
#include "ref_shim.inc"
