/* gsb200.h -- C ABI of libgsb200.so: the B200 (sm_100a) implementation of the reference's
 * SparseMatrix container + Gauss-Seidel hot path.
 *
 * "v1" = labs/lab3/src/OpenCVHW1/sparse-matrix.h, "v2" = labs/lab8/src/OpenCVHW1/sparse-matrix.h
 * (== project/src/PhotoMontage/sparse-matrix.h) of linwe2012/CourseComputationalPhotography.
 * Every entry point names the reference member it replaces.  The reference has no FFI layer
 * (it is a header-only C++17 template), so this is the boundary a maintainer would bind the
 * class's method bodies to; include/sparse-matrix.h is that binding.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types.
 *   - every function returns a gsb_status (0 = ok); gsb_last_error() gives the text of the
 *     last failure on the calling thread.
 *   - pointers are HOST pointers unless the function name ends in _dev, in which case vector
 *     arguments are DEVICE pointers on the matrix's device and the call is asynchronous on
 *     the library stream only until it needs the stop-rule scalar (it then synchronises).
 *   - element type: the reference instantiates SparseMatrix<int> (lab3) and
 *     SparseMatrix<double> (lab8/project).  `vtype` selects which; solver arithmetic is
 *     always FP64 exactly as the reference promotes T to double (v2 :364-373).
 *   - index type: int32 (reference default IndexType=int, v2 :107,119).
 *   - there is NO CPU fallback: without a CUDA device every compute entry point returns
 *     GSB_ERR_NO_DEVICE.
 */
#ifndef GSB200_H
#define GSB200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum gsb_status {
    GSB_OK = 0,
    GSB_ERR_ARG = 1,        /* null pointer, negative size, nrhs out of range ...           */
    GSB_ERR_SHAPE = 2,      /* vector length / matrix shape mismatch, non-square solve       */
    GSB_ERR_UNSORTED = 3,   /* sorted-COO contract violated (rows not non-decreasing)        */
    GSB_ERR_CUDA = 4,       /* a CUDA runtime call or kernel failed                          */
    GSB_ERR_NCCL = 5,       /* NCCL missing or a NCCL call failed                            */
    GSB_ERR_NO_DEVICE = 6,  /* no usable CUDA device                                         */
    GSB_ERR_ALLOC = 7,      /* host or device allocation failed                              */
    GSB_ERR_COLORING = 8,   /* supplied/requested ordering is not a proper colouring         */
    GSB_ERR_OVERFLOW = 9,   /* sizes exceed the int32 index range                            */
    GSB_ERR_STATE = 10      /* handle not in the state the call needs (e.g. empty matrix)    */
} gsb_status;

enum { GSB_F64 = 0, GSB_I32 = 1 };                       /* vtype */
enum { GSB_ORDER_AUTO = 0, GSB_ORDER_REDBLACK = 1, GSB_ORDER_MULTICOLOR = 2, GSB_ORDER_USER = 3 };

typedef struct gsb_matrix gsb_matrix; /* opaque: device-resident slack CSR + solver format */
typedef struct gsb_dist gsb_dist;     /* opaque: one rank of a row-strip multi-GPU solve    */

const char *gsb_last_error(void);
int gsb_version(void);
int gsb_device_count(int *count);
int gsb_set_device(int device);       /* device used by handles created afterwards (default 0) */
void *gsb_stream(void);               /* the library's cudaStream_t on the current device      */

/* ---------------------------------------------------------------------------------------
 * Container: assembly
 * ------------------------------------------------------------------------------------- */
int gsb_matrix_create(gsb_matrix **out, int vtype);
int gsb_matrix_destroy(gsb_matrix *m);

/* A1  initializeFromVector(rows, cols&&, vals&&)            v1 :209-255, v2 :265-319
 * Sorted COO (by row, then col; explicit zeros allowed) -> slack CSR, on the device.
 * n_rows = rows[n-1]+1, n_cols = max(cols)+1.  Layout is bit-exact with the reference:
 * row_begin[r] = #entries with row<r, row_nnz[r] = nonzeros of r, row_left[r] = explicit
 * zeros of r, live slots = the nonzeros in input order, slack slots keep the input entry
 * that was at that position.  vals: double* (GSB_F64) or int32_t* (GSB_I32). */
int gsb_matrix_assemble_sorted_coo(gsb_matrix *m, const int *rows, const int *cols, const void *vals,
                                   int64_t n);

/* initializeFromTriplets(Triplet*, cnt)                      v2 :249-263 (crashes upstream)
 * Unsorted COO -> compact CSR by a device radix sort on (row, col): the last duplicate of a
 * coordinate wins (the reference's repeated insert() semantics), zeros are dropped, columns
 * ascend, no slack. */
int gsb_matrix_assemble_coo(gsb_matrix *m, const int *rows, const int *cols, const void *vals, int64_t n,
                            int n_rows, int n_cols);

/* A3  initializeFromEigenRowMajor(values,n_values,row_offset,n_row_offset,col_offset,
 *                                 n_col_offset,non_zeros,n_non_zeros)      v2 :537-620
 * CSR import; nnz_per_row == NULL is the compressed (Poisson) case.  Reproduces the
 * reference's treatment of trailing empty rows (:608-614) and of given per-row counts. */
int gsb_matrix_import_csr(gsb_matrix *m, const void *values, int n_values, const int *row_off,
                          int n_row_off, const int *col_idx, int n_col_off, const int *nnz_per_row,
                          int n_nnz_per_row);

/* Raw layout upload: the five reference arrays as they stand on the host (used by the C++
 * mirror after insert() edits the host copy; A5 v2 :183-247 stays on the host). */
int gsb_matrix_upload(gsb_matrix *m, const void *values, const int *cols, int64_t store,
                      const int *row_begin, const int *row_nnz, const int *row_left, int n_rows, int n_cols);

int gsb_matrix_shape(const gsb_matrix *m, int64_t *store, int *n_rows, int *n_cols, int64_t *nnz);
/* Copy the five reference arrays back (any pointer may be NULL). values: vtype elements. */
int gsb_matrix_download(const gsb_matrix *m, void *values, int *cols, int *row_begin, int *row_nnz,
                        int *row_left);

/* Stateless forms of A1/A3 (create + assemble + download + destroy). Output arrays: values/cols
 * sized n (A1) or n_values (A3); row arrays sized n_rows. */
int gsb_csr_from_sorted_coo(int vtype, const int *rows, const int *cols, const void *vals, int64_t n,
                            void *values_out, int *cols_out, int *row_begin, int *row_nnz, int *row_left,
                            int *n_rows, int *n_cols);
int gsb_csr_import(int vtype, const void *values, int n_values, const int *row_off, int n_row_off,
                   const int *col_idx, int n_col_off, const int *nnz_per_row, int n_nnz_per_row,
                   void *values_out, int *cols_out, int *row_begin, int *row_nnz, int *row_left);

/* A4  at(row,col) / coeff(row,col)                          v2 :162-178, :627-645
 * Batched lookup on the device copy: out[k] = at(rows[k], cols[k]) as double. */
int gsb_matrix_at(const gsb_matrix *m, const int *rows, const int *cols, int64_t count, double *out);

/* ---------------------------------------------------------------------------------------
 * Ordering: colouring + colour-major reorder of the solver format
 * ------------------------------------------------------------------------------------- */
/* ordering: GSB_ORDER_*.  user_colors (host, n_rows ints, colours 0..k-1) only for _USER.
 * AUTO = 5-point grid probe -> red-black, else greedy multicolour.  Called implicitly (AUTO)
 * by the first solve if the caller did not. */
int gsb_matrix_analyze(gsb_matrix *m, int ordering, const int *user_colors);
int gsb_matrix_coloring(const gsb_matrix *m, int *n_colors, int *ordering_used, int *grid_width);
/* perm[new] = old row; colors[old row] = colour.  Either may be NULL. */
int gsb_matrix_ordering(const gsb_matrix *m, int *perm, int *colors);

/* ---------------------------------------------------------------------------------------
 * Solver (A6-A8)
 * ------------------------------------------------------------------------------------- */
typedef struct gsb_gs_options {
    int ordering;        /* GSB_ORDER_* used if the matrix has not been analysed yet (default AUTO)  */
    int check_every;     /* evaluate the stop rule every k-th sweep; 1 = every sweep (reference)     */
    int batch_sweeps;    /* sweeps enqueued between host reads of the stop flag; 0 = library default */
    int use_graph;       /* 1 = replay a captured CUDA graph per batch; 0 = plain launches; -1 auto  */
    int kernel;          /* 0 = auto; 1 = row-per-thread direct; 2 = staged tiles (bulk copies);
                            3 = persistent ring; 4 = ring + shared-memory gather windows;
                            5 = two-colour systems: both colours in one launch per sweep (L2 wavefront);
                            6 = small systems: the whole solve in one persistent launch (grid barriers)  */
    int compute_residual;/* 1 = also return ||b - A x||_2 per right-hand side in stats                */
    int fused_lead;      /* kernel 5: tiles colour 0 runs ahead of colour 1 beyond the dependency
                            distance; 0 = library default (one tile per resident CTA)                 */
    int reserved[1];
} gsb_gs_options;

typedef struct gsb_gs_stats {
    int sweeps;          /* cnt at exit (v2 :377)                                                     */
    int n_colors;
    int ordering_used;
    int kernel_used;     /* 1..6 as above; strip solver: +10 halo exchange fused into the phase kernels,
                            +20 more when the stop-rule all-reduce is fused into the end-of-sweep kernel */
    int64_t kernel_launches; /* launches of this library's kernels enqueued by the call              */
    double last_eps[4];  /* L1 norm of the last evaluated sweep update, per right-hand side (v2 :376) */
    double residual_l2[4];
    double solve_ms;     /* device time of the sweep loop (CUDA events on the library stream)         */
    double setup_ms;     /* analysis/reorder time if it ran inside this call, else 0                  */
} gsb_gs_stats;

void gsb_gs_default_options(gsb_gs_options *o);

/* A6  gaussSeidel(b, epsilon = 1e-6, max_iteration = 1000)   v1 :275-305, v2 :350-380
 * x0 = 1.0; zero/absent diagonal rows are skipped; stop when the L1 norm of a sweep's update
 * is <= epsilon or after max_iteration sweeps.  Ordering differs from the reference's
 * lexicographic sweep (red-black / multicolour), so iterates differ; the fixed point does
 * not.  nrhs in 1..4 right-hand sides share one pass over the matrix (b and x are
 * nrhs * n_rows doubles, one vector after another); the loop stops when every right-hand
 * side meets epsilon.  opts and stats may be NULL. */
int gsb_gauss_seidel(gsb_matrix *m, const double *b, int nrhs, double epsilon, int max_iteration,
                     const gsb_gs_options *opts, double *x_out, gsb_gs_stats *stats);
int gsb_gauss_seidel_dev(gsb_matrix *m, const double *b_dev, int nrhs, double epsilon, int max_iteration,
                         const gsb_gs_options *opts, double *x_dev, gsb_gs_stats *stats);
/* EXTENSION (not in the reference API): same with an initial guess x0. */
int gsb_gauss_seidel_x0(gsb_matrix *m, const double *b, const double *x0, int nrhs, double epsilon,
                        int max_iteration, const gsb_gs_options *opts, double *x_out, gsb_gs_stats *stats);

/* A7  applyToVector(in, out)                                 v1 :307-318, v2 :382-393
 * out = A * in over the live entries in storage order (bit-exact with the reference's
 * unfused multiply-add). in: n_cols doubles, out: n_rows doubles. */
int gsb_spmv(gsb_matrix *m, const double *in, double *out);
int gsb_spmv_dev(gsb_matrix *m, const double *in_dev, double *out_dev);
/* ||b - A x||_2 (the residual norm reported beside every parity number) */
int gsb_residual_l2(gsb_matrix *m, const double *b, const double *x, double *out);
int gsb_residual_l2_dev(gsb_matrix *m, const double *b_dev, const double *x_dev, double *out);

/* A8  manhattonDist / dotProd / veclen2 / vecadd / vecsub / vecmul   v2 :45-105 */
int gsb_l1_dist(const double *a, const double *b, int64_t n, double *out);
int gsb_dot(const double *a, const double *b, int64_t n, double *out);
int gsb_axpy(const double *a, const double *b, double scale_b, int64_t n, double *out); /* a+scale*b */
int gsb_vecmul(const double *a, const double *b, int64_t n, double *out);               /* a .* b    */

/* N1  conjugateGradient(b, eps, max_iter, initialize)        v2 :396-434
 *     conjugateGradientEigen(b, eps, max_iter) (Jacobi-PCG)  v2 :472-535
 * x0 may be NULL (zero start).  iters = loop count at exit. */
int gsb_conjugate_gradient(gsb_matrix *m, const double *b, double epsilon, int max_iteration,
                           const double *x0, double *x_out, int *iters);
/* EXTENSION: nrhs in 1..4 right-hand sides (colour channels) in one call -- every SpMV reads the matrix once for
 * all of them; each right-hand side stops on its own (iters: nrhs counters).  b, x0, x_out: nrhs * n doubles. */
int gsb_conjugate_gradient_multi(gsb_matrix *m, const double *b, int nrhs, double epsilon, int max_iteration,
                                 const double *x0, double *x_out, int *iters);
int gsb_conjugate_gradient_jacobi(gsb_matrix *m, const double *b, double epsilon, int max_iteration,
                                  double *x_out, int *iters);

/* ---------------------------------------------------------------------------------------
 * Poisson system (A9, A10; "next" rows N2/N4): built on the device, no Eigen
 * ------------------------------------------------------------------------------------- */
/* A9  SolveChannel front half: PhotoMontage.cpp:541-597 == hw8_pa.cc:911-967.
 * Builds A^T*A of the forward-difference system for a W x H image straight into `m`
 * (compressed CSR, ascending columns, == what initializeFromEigenRowMajor receives). */
int gsb_poisson_matrix(gsb_matrix *m, int W, int H);
/* A^T*b for nch channels: gx, gy = nch*H*W float32 gradients (host), constraint[nch];
 * b_out = nch*W*H doubles (host). */
int gsb_poisson_rhs(int W, int H, int nch, const float *gx, const float *gy, const double *constraint,
                    double *b_out);
int gsb_poisson_rhs_dev(int W, int H, int nch, const float *gx_dev, const float *gy_dev,
                        const double *constraint, double *b_dev);
/* Same for the image rows [y0, y1) of a strip (multi-GPU): gx_rows / gy_rows hold image rows
 * [max(y0-1,0), y1) of the gradients, nch planes; b_out = nch * W*(y1-y0) doubles. */
int gsb_poisson_rhs_rows(int W, int H, int y0, int y1, int nch, const float *gx_rows, const float *gy_rows,
                         const double *constraint, double *b_out);
int gsb_poisson_rhs_rows_dev(int W, int H, int y0, int y1, int nch, const float *gx_rows_dev,
                             const float *gy_rows_dev, const double *constraint, double *b_dev);
/* A10 write-back: uchar(clamp(x, 0, 255)), truncating. PhotoMontage.cpp:617-626 */
int gsb_writeback_u8(const double *x, int64_t n, unsigned char *out);
int gsb_writeback_u8_dev(const double *x_dev, int64_t n, unsigned char *out_dev);

/* ---------------------------------------------------------------------------------------
 * Gradient-domain fusion driver ("next" rows N2 + N4): the callers either side of the solve,
 * so that the whole stage -- images / gradients in, fused image out -- stays on the device.
 *
 * Layouts are the reference's cv::Mat layouts:
 *   images  n_images x (H x W x 3) bytes, channel-interleaved (CV_8UC3, continuous), back to back
 *   labels  H x W bytes: ResultLabel.at<uchar>(y, x) < n_images
 *   gx, gy  3 planes of H x W float32 (plane c = component c of the reference's Vec3f gradient)
 *   x / init 3 planes of H x W doubles (the reference solves one channel at a time)
 *   out     H x W x 3 bytes, channel-interleaved
 * ------------------------------------------------------------------------------------- */
#define GSB_GDF_GS 0 /* gaussSeidel, the three channels fused into one pass over the matrix   */
#define GSB_GDF_CG 1 /* conjugateGradient per channel -- what the reference's drivers call     */
typedef struct gsb_gdf_options {
    int solver;          /* GSB_GDF_GS (default) or GSB_GDF_CG                                        */
    int max_iteration;   /* GS: 1000 (v2 :350); CG: the reference's `iterations` (50 with an init)    */
    double epsilon;      /* GS: 1e-6 (v2 :350); CG: 1e-10 (PhotoMontage.cpp:613, hw8_pa.cc:972)       */
    gsb_gs_options gs;   /* ordering / kernel / check_every of the GS path                            */
    int reserved[4];
} gsb_gdf_options;
typedef struct gsb_gdf_stats {
    int iterations[4];     /* GS: sweeps (same for all channels); CG: loop count per channel          */
    double last_eps[4];    /* GS: L1 norm of the last evaluated sweep update per channel              */
    double residual_l2[4]; /* ||A^T b - A^T A x||_2 per channel at exit                               */
    double solve_ms;       /* device time of the solver loop(s)                                       */
    double total_ms;       /* wall time of the call                                                   */
} gsb_gdf_stats;
void gsb_gdf_default_options(gsb_gdf_options *o);
/* N2  GradientAt of the labelled source image   PhotoMontage.cpp:399-408, :419-425
 * (last row / column, which the reference never reads, are written as 0) */
int gsb_gdf_gradients(const unsigned char *images, int n_images, const unsigned char *labels, int W, int H,
                      float *gx, float *gy);
/* N4  fast_init_value: init = the composite   PhotoMontage.cpp:599-610 */
int gsb_gdf_composite(const unsigned char *images, int n_images, const unsigned char *labels, int W, int H,
                      double *x0);
/* SolveChannel for the three channels with the gradients given (hw8_pa.cc:902-986, called at :808-810;
 * PhotoMontage.cpp:535-628): A^T b from gx / gy / constraint[3], the Poisson matrix built on the device (kept
 * and reused by later calls of the same size), solve, clamp and interleave.  init = 3*W*H doubles or NULL
 * (GS: the reference's 1.0; CG: zero).  An initial guess for GS is an extension of the reference API. */
int gsb_gdf_solve(int W, int H, const float *gx, const float *gy, const double *constraint, const double *init,
                  const gsb_gdf_options *opts, unsigned char *out, gsb_gdf_stats *stats);
/* BuildSolveGradientFusion (PhotoMontage.cpp:410-433): gradients picked by the label map, constraint =
 * Images[0](0,0), fast_init != 0 starts from the composite (:599-610), three SolveChannel calls, result image. */
int gsb_gdf_fuse(const unsigned char *images, int n_images, const unsigned char *labels, int W, int H,
                 int fast_init, const gsb_gdf_options *opts, unsigned char *out, gsb_gdf_stats *stats);
/* drops the Poisson matrix the driver keeps between calls */
int gsb_gdf_release(void);

/* ---------------------------------------------------------------------------------------
 * lab8 panorama: producers of the right-hand side ("next" row N3), labs/lab8/src/OpenCVHW1/hw8_pa.cc.
 * All buffers are continuous cv::Mat layouts: images H x W x 3 bytes, gradients H x W x 3 float32
 * (CV_32FC3), masks H x W bytes.  The warps / erosions between these steps are OpenCV calls and stay
 * with the caller.  The merge functions are the reference's per-row scans (skip to the source mask,
 * skip what the target already covers, copy the rest of the run).
 * ------------------------------------------------------------------------------------- */
/* MaskImage :443-466 */
int gsb_pano_mask_image(const unsigned char *src, const unsigned char *mask, int W, int H, unsigned char *out);
/* struct Gradients(m) :604-636 -- GradientAt :314-323 for y < H-1, x < W-1 (0 elsewhere) */
int gsb_pano_gradients(const unsigned char *img, int W, int H, float *gx, float *gy);
/* struct Gradients, second constructor (mask-driven)            hw8_pa.cc:638-676, ZeroGradientAt :325-334
 * per row: ZeroGradientAt left of the first mask pixel, GradientAt from there to the row end, 0 elsewhere */
int gsb_pano_gradients_masked(const unsigned char *img, const unsigned char *mask, int W, int H, float *gx,
                              float *gy);
/* MergeImage2<float>(target, src, target_mask, src_outer_mask, src_inner_mask) :338-385; target in place */
int gsb_pano_merge2_f32(float *target, const float *src, const unsigned char *target_mask,
                        const unsigned char *src_outer_mask, const unsigned char *src_inner_mask, int W, int H);
/* MergeImage<uchar, channel>(target, src, target_mask, src_mask, SkipHowMany) :387-441; channel 3 or 1;
 * target == target_mask is allowed (the reference's mask merge, :768) */
int gsb_pano_merge_u8(unsigned char *target, const unsigned char *src, const unsigned char *target_mask,
                      const unsigned char *src_mask, int channel, double skip_how_many, int W, int H);
/* EnforceGradientBound(dx, dy, src, mask) :468-498; dx, dy in place */
int gsb_pano_enforce_gradient_bound(float *dx, float *dy, const unsigned char *src, const unsigned char *mask,
                                    int W, int H);
/* one iteration of the stitch loop :740-768 after its warps: raw, dx, dy, mask updated in place from the
 * warped source image and its two eroded masks (erode_mask = inner, erode_mask2 = outer) */
int gsb_pano_merge_step(unsigned char *raw, float *dx, float *dy, unsigned char *mask, const unsigned char *warped,
                        const unsigned char *erode_mask, const unsigned char *erode_mask2, int W, int H);
/* CV_32FC3 -> three planes: the gradient layout gsb_gdf_solve / gsb_poisson_rhs take */
int gsb_pano_split_planes_f32(const float *interleaved, int W, int H, float *planes);

/* ---------------------------------------------------------------------------------------
 * Multi-GPU row strips: one rank per GPU, halo exchange per colour phase
 * ------------------------------------------------------------------------------------- */
#define GSB_UNIQUE_ID_BYTES 128
/* rank 0 fills id (ncclGetUniqueId); the caller broadcasts it to the other ranks by any
 * means (torch.distributed in bench.py) before gsb_dist_init. */
int gsb_dist_unique_id(unsigned char id[GSB_UNIQUE_ID_BYTES]);
int gsb_dist_init(gsb_dist **out, const unsigned char id[GSB_UNIQUE_ID_BYTES], int rank, int world,
                  int device);
int gsb_dist_finalize(gsb_dist *d);
/* Reference-faithful full-grid Poisson matrix (A9) for the image rows [y0, y1) owned by this
 * rank, generated on the device (config C4: 16384^2 never exists on one host). */
int gsb_dist_poisson_strip(gsb_dist *d, int W, int H, int y0, int y1);
/* General form: this rank's rows [row0, row0+n_local) of a square n_global system, compressed
 * CSR with GLOBAL column indices; colours = parity of (col % grid_width + col / grid_width). */
int gsb_dist_matrix_rows(gsb_dist *d, const double *values, const int *row_off, const int *col_idx,
                         int64_t row0, int n_local, int64_t n_global, int grid_width);
/* b_dev / x_dev: nrhs vectors of n_local doubles on this rank's device. Collective. */
int gsb_dist_gauss_seidel_dev(gsb_dist *d, const double *b_dev, int nrhs, double epsilon,
                              int max_iteration, const gsb_gs_options *opts, double *x_dev,
                              gsb_gs_stats *stats);
int gsb_dist_residual_l2_dev(gsb_dist *d, const double *b_dev, const double *x_dev, double *out);

/* ---------------------------------------------------------------------------------------
 * Single-process multi-device solve: N devices of one box behind ONE blocking call, the call
 * model of the reference (one host thread, synchronous solve: project/src/PhotoMontage/
 * main.cpp:581).  No NCCL, no torch, no IPC -- the devices reach each other through peer access.
 * ------------------------------------------------------------------------------------- */
/* Process-wide device list of the HOST entry points (gsb_gauss_seidel, i.e. SparseMatrix::
 * gaussSeidel of the drop-in header): with two or more devices a solve of a matrix that takes
 * a two-colouring (red-black grids, the caller's two colours) is split into row strips, one per
 * device; x is bit-identical to the single-device solve.  Matrices with more colours, tiny
 * systems and the x0 extension stay on one device (stats.kernel_used < 10 says so).
 * n = 0 or 1: single device.  Default: the environment variable GSB_DEVICES ("0,1,2,3"). */
int gsb_set_devices(const int *devices, int n);
int gsb_get_devices(int *devices, int cap); /* returns the count */
/* The same machinery as an explicit object (bench.py, tests): */
typedef struct gsb_dist_group gsb_dist_group;
int gsb_dist_init_local(gsb_dist_group **out, const int *devices, int n);
int gsb_dist_group_finalize(gsb_dist_group *g);
int gsb_dist_group_size(const gsb_dist_group *g);
/* shard an assembled matrix (any assembly entry point) by rows; analyses it if necessary */
int gsb_dist_group_matrix(gsb_dist_group *g, gsb_matrix *m);
/* or generate the reference-faithful W x H Poisson system strip by strip on the devices (C4) */
int gsb_dist_group_poisson(gsb_dist_group *g, int W, int H);
/* b / x_out: HOST vectors, nrhs * n doubles, as gsb_gauss_seidel takes them */
int gsb_dist_group_gauss_seidel(gsb_dist_group *g, const double *b, int nrhs, double epsilon,
                                int max_iteration, const gsb_gs_options *opts, double *x_out,
                                gsb_gs_stats *stats);
int gsb_dist_group_residual_l2(gsb_dist_group *g, const double *b, const double *x, double *out);
/* one-process-per-GPU mode: the caller's two-colouring of ALL n_global rows (bytes 0/1) for
 * gsb_dist_matrix_rows, instead of pixel parity (compact / masked systems); n_global = 0 clears */
int gsb_dist_set_colors(gsb_dist *d, const unsigned char *colors, int64_t n_global);

/* Pinned host memory for callers that want full-rate PCIe copies (bench.py e2e leg). */
int gsb_host_alloc(void **ptr, int64_t bytes);
int gsb_host_free(void *ptr);

/* Diagnostics: how many device allocations / frees (cudaMalloc / cudaFree) the library has made in this process so
   far.  A steady-state import + analyse + solve of an unchanged shape makes none (bench.py reports the count per e2e
   step). */
int gsb_alloc_counters(int64_t *device_allocs, int64_t *device_frees);

#ifdef __cplusplus
}
#endif
#endif /* GSB200_H */
