/*
 * rtat_b200.h — C ABI of the B200-native BLAS back end for rtatblas' GEMM execution path.
 *
 * This is the drop-in boundary.  The reference reaches the GPU only through the `rtat::gpu::`
 * alias table (reference src/gpu-api/gpu-api.h:87-113), which forwards to cuBLAS-v2 entry points.
 * Every function below replaces exactly one of those aliases and keeps the cuBLAS-v2 calling
 * convention: column-major storage, `int` dimensions, alpha/beta passed as HOST pointers, all work
 * enqueued asynchronously on the stream bound to the handle, an integer status returned
 * (0 == success; numbering follows cublasStatus_t so existing `== BLAS_STATUS_SUCCESS` checks hold).
 *
 *   reference alias (gpu-api.h line)          replacement
 *   ---------------------------------------   -------------------------------
 *   gpu::blasCreate        (:91)              rtat_b200_create
 *   gpu::blasDestroy       (:92)              rtat_b200_destroy
 *   gpu::blasSetStream     (:102)             rtat_b200_set_stream
 *   gpu::blasGetStream     (:101)             rtat_b200_get_stream
 *   gpu::blasDgemm         (:94)              rtat_b200_dgemm
 *   gpu::blasSgemm         (:98)              rtat_b200_sgemm
 *   gpu::blasDgeam         (:93)              rtat_b200_dgeam
 *   gpu::blasSgeam         (:97)              rtat_b200_sgeam
 *   gpu::blasDsyrk         (:96)              rtat_b200_dsyrk
 *   gpu::blasSsyrk         (:100)             rtat_b200_ssyrk
 *   gpu::blasDtrsm         (:95)              rtat_b200_dtrsm
 *   gpu::blasStrsm         (:99)              rtat_b200_strsm
 *   gpu::randGenerateUniform[Double] (:119-120)  rtat_b200_fill_uniform_{s,d}   (seeded; drivers only)
 *
 * There is no CPU fallback and no vendor-BLAS call behind these functions: they launch hand-written
 * sm_100a kernels or fail with a non-zero status.
 *
 * No torch / C++ types cross this boundary: plain pointers and sizes only.  `stream` arguments are
 * `cudaStream_t` values passed as `void*` so that the header has no CUDA include dependency.
 */
#ifndef RTAT_B200_H
#define RTAT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RTAT_B200_VERSION 100 /* 0.1.0 */

/* ---- status codes (cublasStatus_t numbering) -------------------------------------------------- */
enum {
  RTAT_B200_STATUS_SUCCESS = 0,
  RTAT_B200_STATUS_NOT_INITIALIZED = 1,
  RTAT_B200_STATUS_ALLOC_FAILED = 3,
  RTAT_B200_STATUS_INVALID_VALUE = 7,
  RTAT_B200_STATUS_ARCH_MISMATCH = 8,
  RTAT_B200_STATUS_EXECUTION_FAILED = 13,
  RTAT_B200_STATUS_INTERNAL_ERROR = 14,
  RTAT_B200_STATUS_NOT_SUPPORTED = 15
};

/* ---- operation flags (cublasOperation_t numbering; reference gpu-api.h:112-113) --------------- */
enum { RTAT_B200_OP_N = 0, RTAT_B200_OP_T = 1 };
/* ---- triangle / side / diagonal flags (cublasFillMode_t, cublasSideMode_t, cublasDiagType_t numbering;
 *      reference gpu-api.h:103-111) ------------------------------------------------------------- */
enum { RTAT_B200_FILL_LOWER = 0, RTAT_B200_FILL_UPPER = 1 };
enum { RTAT_B200_SIDE_LEFT = 0, RTAT_B200_SIDE_RIGHT = 1 };
enum { RTAT_B200_DIAG_NON_UNIT = 0, RTAT_B200_DIAG_UNIT = 1 };

typedef struct rtat_b200_context *rtat_b200_handle_t;

int rtat_b200_version(void);
const char *rtat_b200_status_string(int status);

/* ---- handle management (replaces cublasCreate/Destroy/SetStream/GetStream) --------------------- */
int rtat_b200_create(rtat_b200_handle_t *handle);
int rtat_b200_destroy(rtat_b200_handle_t handle);
int rtat_b200_set_stream(rtat_b200_handle_t handle, void *stream /* cudaStream_t */);
int rtat_b200_get_stream(rtat_b200_handle_t handle, void **stream /* cudaStream_t* */);

/* ---- GEMM:  C = alpha * op(A) * op(B) + beta * C   (replaces cublasDgemm / cublasSgemm) --------
 * op(A) is m x k, op(B) is k x n, C is m x n, all column-major with leading dimensions lda/ldb/ldc.
 * beta == 0 means C is write-only (never read, so NaN/uninitialised scratch is fine — the
 * reference hands uninitialised scratch as C: matrixop.h:433-436).
 * dgemm runs FP64 DMMA tiles.  sgemm is FP32-faithful: 3xTF32 split products on tcgen05 tensor
 * cores with FP32 accumulation, or an FP32 FMA kernel for operands TMA cannot address. */
int rtat_b200_dgemm(rtat_b200_handle_t handle, int transa, int transb, int m, int n, int k,
                    const double *alpha, const double *A, int lda, const double *B, int ldb,
                    const double *beta, double *C, int ldc);
int rtat_b200_sgemm(rtat_b200_handle_t handle, int transa, int transb, int m, int n, int k,
                    const float *alpha, const float *A, int lda, const float *B, int ldb,
                    const float *beta, float *C, int ldc);

/* ---- GEAM:  C = alpha * op(A) + beta * op(B)   (replaces cublasDgeam / cublasSgeam) ------------
 * C is m x n.  The reference always calls this with transb == N and C aliasing B (in place,
 * ldb == ldc; matrixop.h:302,340-341); that aliasing is supported.  C aliasing A is supported only
 * for transa == N with lda == ldc.  beta == 0 means B is never read; alpha == 0 means A is never
 * read.  Results are bit-identical to cuBLAS 12.9 geam: c = fma(beta, b, RN(alpha*a)) in the
 * general case — pinned against the reference executor's outputs, see tests/golden. */
int rtat_b200_dgeam(rtat_b200_handle_t handle, int transa, int transb, int m, int n,
                    const double *alpha, const double *A, int lda, const double *beta,
                    const double *B, int ldb, double *C, int ldc);
int rtat_b200_sgeam(rtat_b200_handle_t handle, int transa, int transb, int m, int n,
                    const float *alpha, const float *A, int lda, const float *beta, const float *B,
                    int ldb, float *C, int ldc);

/* ---- SYRK:  C = alpha * op(A) * op(A)^T + beta * C  on the `uplo` triangle   (replaces cublasDsyrk / cublasSsyrk)
 * C is n x n; op(A) is n x k (trans == N: A is n x k, lda >= n; trans == T: A is k x n, lda >= k).  Only the
 * requested triangle of C (diagonal included) is read or written; beta == 0 means it is write-only;
 * alpha == 0 or k == 0 scales the triangle by beta without reading A.  Runs the GEMM kernels' triangular
 * tile schedule (syrk.cu), so precision and alignment rules are those of rtat_b200_dgemm / _sgemm. */
int rtat_b200_dsyrk(rtat_b200_handle_t handle, int uplo, int trans, int n, int k, const double *alpha,
                    const double *A, int lda, const double *beta, double *C, int ldc);
int rtat_b200_ssyrk(rtat_b200_handle_t handle, int uplo, int trans, int n, int k, const float *alpha,
                    const float *A, int lda, const float *beta, float *C, int ldc);

/* ---- TRSM:  op(A) X = alpha B (side LEFT) or X op(A) = alpha B (side RIGHT), X overwrites B   (replaces
 *      cublasDtrsm / cublasStrsm) ------------------------------------------------------------------
 * B is m x n; A is triangular (`uplo`), m x m for LEFT, n x n for RIGHT; only that triangle is read, and not its
 * diagonal when diag == UNIT.  alpha == 0 zeroes B without reading A.  Recursive blocked solve whose updates run in
 * the GEMM kernels above (trsm.cu).  Only the 32 x 32 diagonal sub-blocks are inverted (one pre-pass per call, by
 * substitution on the identity) and applied as small products; everything larger is substitution. */
int rtat_b200_dtrsm(rtat_b200_handle_t handle, int side, int uplo, int trans, int diag, int m, int n,
                    const double *alpha, const double *A, int lda, double *B, int ldb);
int rtat_b200_strsm(rtat_b200_handle_t handle, int side, int uplo, int trans, int diag, int m, int n,
                    const float *alpha, const float *A, int lda, float *B, int ldb);

/* ---- seeded uniform fill in (lo, hi]  (replaces the unseeded cuRAND fill of the drivers) -------
 * Counter-based (SplitMix64 of seed + element index): the same (seed, index) gives the same value
 * on any grid, so hosts can regenerate inputs without copying them. */
int rtat_b200_fill_uniform_d(rtat_b200_handle_t handle, double *x, size_t n, uint64_t seed,
                             double lo, double hi);
int rtat_b200_fill_uniform_s(rtat_b200_handle_t handle, float *x, size_t n, uint64_t seed, float lo,
                             float hi);

/* ---- tuning / introspection (no reference counterpart) ------------------------------------------
 * rtat_b200_set_option:  "dgemm.variant", "sgemm.variant", "geam.variant" select a kernel family
 *   explicitly (0 = automatic).  Used by the tests to force every code path and by the planner's
 *   fused/explicit plan dimension.  The other keys are switches for A/B measurements and tests; unknown
 *   keys are stored and ignored.  (Defaults in brackets.)
 *     dgemm.tile [0]            128 / 64 / 192 (= 128x64 tiles): force the tile configuration of the persistent DGEMM
 *     dgemm.mid [1]             0: the automatic choice never takes 128x64 tiles
 *     dgemm.skinny [1]          0: thin strips and GEMV-shaped calls (at most 4 rows / columns of C) stay on the tile kernels
 *     dgemm.skinny_fused [1]    0: the skinny kernel's partial sums are added by a second launch instead of the last CTA
 *     dgemm.skinny_pipe [0]     1 / 2: S tiles of the skinny kernel fetched one tile ahead always / never (0 = by shape)
 *     dgemm.group [8]           tile rows the grouped raster of the persistent DGEMM sweeps across n together
 *     dgemm.ktail [1]           0: the last k-tile of a k that is not a multiple of 16 runs all four DMMA steps
 *     dgemm.streamk [1]         0: data-parallel schedule only.  Stream-K note: the CTA that owns a tile's first k-tile
 *                               waits for partial tiles from higher-numbered CTAs of the same launch.  CTAs are issued in
 *                               order and the lower ones can always finish, so a launch that is only partly resident
 *                               (other work holding SMs) still completes; the wait is bounded all the same (20 s, then
 *                               the kernel traps and the call surfaces a CUDA error instead of hanging the GPU).  Callers
 *                               that run several handles on concurrent streams and want no such dependency at all set
 *                               this option to 0.
 *     dgemm.tma3d [1]           0: eight 2-D TMA boxes per M/N-contiguous operand stage instead of one 3-D box
 *     dgemm.peel [1]            0: keep remainders of 1..16 rows / columns in the 128x128 launch instead of a second
 *                               launch on 64x64 tiles
 *     dgemm.prefetch_c [1]      0: beta != 0 without the L2 prefetch of the C tile at the start of its main loop
 *     dgemm.l2_hints [0]        1: evict_last / evict_first on the A / B loads (measured: more DRAM traffic)
 *     dgemm.force_unaligned [0] 1: cp.async producers even when TMA could address the operands
 *     sgemm.presplit [1]        0: split op(B) into big / small TF32 parts per tile instead of once per call; 1: once per
 *                               call from 768^3 multiply-adds on (or when TMA cannot address B as stored); 2: always
 *     sgemm.pair [1]            0: single-CTA tcgen05 kernel instead of CTA pairs (cta_group::2)
 *     sgemm.pad_a [1]           0: an op(A) TMA cannot address goes to the FP32 SIMT kernel instead of an aligned copy
 *     geam.transpose [0]        transposing geam: 1 = TMA pipeline, 2 = register-blocked kernel (the default choice),
 *                               3 = scalar tile kernel; geam.order = 1 / geam.upc = 2: CTA order / tile height of kernel 2
 *     geam.persistent [0]       1: grid-stride streaming copies on 8 CTAs per SM instead of one chunk per CTA
 *     geam.even_chunks [1]      0: fixed 2048-element (16 B vectors: 2048-vector) chunks instead of equal parts of a column
 *     geam.force_cpasync [0]    1: cp.async producers in the TMA transposer even for a TMA-addressable source
 *     geam.tune, geam.rounding  launch-shape and rounding experiments of geam (see csrc/geam.cu)
 *     trsm.base [0]             diagonal-block size of the recursive TRSM (0 = 128)
 *     host.panels, host.lead    override the schedule of the host-buffer entry points; host.trace = 1 prints
 *                               the per-stage timeline of every call to stderr
 *     mg.signal [1], mg.spin_wait [0]   flag transport of the multi-GPU host entry (rtat_b200_mg.h)
 * rtat_b200_last_kernel: name of the kernel variant the most recent call on this handle launched.
 * rtat_b200_launch_count: number of kernel launches issued through this handle so far. */
int rtat_b200_set_option(rtat_b200_handle_t handle, const char *key, int value);
int rtat_b200_get_option(rtat_b200_handle_t handle, const char *key, int *value);
const char *rtat_b200_last_kernel(rtat_b200_handle_t handle);
uint64_t rtat_b200_launch_count(rtat_b200_handle_t handle);

/* Schedule the host-buffer entry points below would use for a problem (pure host arithmetic: callable without a GPU;
 * csrc/host_schedule.h).  world = 0: rtat_b200_{d,s}gemm_host; world >= 1: rtat_b200_mg_dgemm_host on `world` GPUs with
 * n = this rank's column count.  opt_panels / opt_lead as the options host.panels / host.lead (0 = automatic).
 * out = {blocks, lead, npanels, nb, col0[blocks], cols[blocks], k0[npanels], kc[npanels]}: column blocks of op(B) / C,
 * how many leading ones take the K-panel phase, and the K-panels of A.  Returns the number of ints written, -1 on a bad
 * argument or when cap is too small (80 always suffices). */
int rtat_b200_host_schedule(int elem_size, int transa, int m, int n, int k, int lda, int world, int opt_panels, int opt_lead, int *out,
                            int cap);

/* ---- host-buffer entry points ---------------------------------------------------------------------
 * Same contract as dgemm/sgemm but A, B, C are HOST pointers (pinned or pageable).  Operands are
 * staged through the handle's device arena (grown on demand) as a pipeline on three streams: large
 * problems first stream A in K-panels against the leading column blocks of op(B) / C (so the math starts
 * after the first panel instead of after all of A), then run whole-K column blocks — H2D of block j+1,
 * the GEMM of block j (on the handle's stream) and D2H of block j-1 overlap; the call returns after C is
 * valid on the host.  This is what a caller that
 * keeps its matrices in host memory sees, and what bench.py times as the end-to-end number. */
int rtat_b200_dgemm_host(rtat_b200_handle_t handle, int transa, int transb, int m, int n, int k,
                         const double *alpha, const double *A, int lda, const double *B, int ldb,
                         const double *beta, double *C, int ldc);
int rtat_b200_sgemm_host(rtat_b200_handle_t handle, int transa, int transb, int m, int n, int k,
                         const float *alpha, const float *A, int lda, const float *B, int ldb,
                         const float *beta, float *C, int ldc);

#ifdef __cplusplus
}
#endif
#endif /* RTAT_B200_H */
