"""gs-b200: the reference's SparseMatrix + Gauss-Seidel hot path on B200 (sm_100a).

`SparseMatrix` mirrors the reference class (labs/lab8/src/OpenCVHW1/sparse-matrix.h) over the C ABI of
libgsb200.so (include/gsb200.h).  There is no CPU implementation in this package: without the built
library and a CUDA device every compute call raises.
"""
from . import _lib, gdf, pano
from ._lib import GsbError, GsOptions, GsStats, load
from .sparse_matrix import (SparseMatrix, dotProd, manhattonDist, poisson_rhs, vecadd, vecmul, vecsub, veclen2,
                            writeback_u8)

__all__ = ["SparseMatrix", "manhattonDist", "dotProd", "veclen2", "vecadd", "vecsub", "vecmul", "poisson_rhs",
           "writeback_u8", "GsbError", "GsOptions", "GsStats", "load", "_lib", "gdf", "pano"]
