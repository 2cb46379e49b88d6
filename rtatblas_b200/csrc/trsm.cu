// TRSM:  solve  op(A) X = alpha B  (left)  or  X op(A) = alpha B  (right)  in place of B   (replaces cublasDtrsm /
// cublasStrsm behind MatrixTrs / MatrixTrsAlloc, reference src/matrix_ops/matrixop.h:78-109, :442-534).
//
// Recursive blocked solve: the triangular dimension is halved (at multiples of 128) until a block fits the panel
// kernel; between the two halves one GEMM folds the solved half into the other half's right-hand side,
//     left,  op(A) lower:  X1 = A11 \ (alpha B1);  B2 = alpha B2 - A21 X1;  X2 = A22 \ B2
// so for m >> 128 practically all flops run in the DMMA / tcgen05 GEMM kernels on large k (the recursion hands the
// GEMMs k = m/2, m/4, ... instead of the k = 128 panels of a left-looking sweep).
//
// Diagonal blocks (<= 128 rows) are solved by trsm_panel_kernel: one CTA holds a 128 x 64 panel of right-hand sides
// in shared memory and walks its four 32-row sub-blocks; per sub-block every thread owns a 2 x 4 patch of the
// 32 x 64 result, so a step is two small dense products (fold the solved sub-blocks in, then apply the sub-block's
// inverse) instead of a 32-step dependent chain per vector.  Only the 32 x 32 DIAGONAL sub-blocks are inverted (all
// of them in one pre-pass launch per call, by substitution on the identity); everything off those sub-blocks is
// plain substitution, which keeps the error growth that of a block substitution with well-conditioned 32 x 32 pivots.
// Only the `uplo` triangle of A is read (not even its diagonal when diag == UNIT).
#include "common.cuh"

namespace rtat_b200 {

namespace trsm {

constexpr int NB = 32;        // inverted diagonal sub-block
constexpr int BASE = 128;     // rows of one panel-kernel launch (default recursion cut-off, and its maximum)
constexpr int V = 64;         // right-hand-side vectors per CTA
constexpr int XS = V + 4;     // row stride of the panel in shared memory (rows stay 16 B aligned for float and double)
constexpr int FS = NB + 4;    // row stride of the staged F rows (4-element vector loads stay 16 B aligned for float and double)
constexpr int PANEL_THREADS = 256;   // all of them stage and store; the first COMPUTE_THREADS own the 32 x 64 sub-block results
constexpr int COMPUTE_THREADS = 128;

// F is the triangular matrix normalised so that the right-hand-side vectors are columns of B (LEFT) or rows of B
// (!LEFT, where F = op(A)^T):  F(i, j) = tr ? A(j, i) : A(i, j);  lower_f: F is lower triangular.

// Pre-pass: inverse of every NB x NB diagonal block of F (identity-padded past na), by substitution on the columns
// of the identity; block b is stored k-major: inv[(b * NB + k) * NB + i] = Finv_b(i, k).
template <typename T>
__global__ void __launch_bounds__(NB)
trsm_diag_inverse_kernel(const T *__restrict__ A, int lda, int na, int tr, int lower_f, int unit, T *__restrict__ inv) {
  __shared__ T sF[NB][NB + 1];
  const int b = blockIdx.x, t = threadIdx.x, r0 = b * NB;
  // element (i, j) of the block; lanes run along A's contiguous dimension.  All 32 loads first, then the stores.
  T col[NB];
#pragma unroll
  for (int e = 0; e < NB; e++) {
    const int i = tr ? e : t, j = tr ? t : e;
    const int gi = r0 + i, gj = r0 + j;
    col[e] = (i == j) ? T(1) : T(0);
    if (gi < na && gj < na && ((lower_f ? i > j : i < j) || (i == j && !unit))) col[e] = tr ? A[(size_t)gi * lda + gj] : A[(size_t)gj * lda + gi];
  }
#pragma unroll
  for (int e = 0; e < NB; e++) sF[tr ? e : t][tr ? t : e] = col[e];
  __syncthreads();
  __shared__ T sRec[NB];
  sRec[t] = T(1) / sF[t][t];  // one division per diagonal entry, in parallel, instead of 32 in every thread's chain
  __syncthreads();
  T x[NB];
#pragma unroll
  for (int i = 0; i < NB; i++) x[i] = (i == t) ? T(1) : T(0);
  if (lower_f) {
#pragma unroll
    for (int i = 0; i < NB; i++) {
      x[i] *= sRec[i];
#pragma unroll
      for (int j = i + 1; j < NB; j++) x[j] -= sF[j][i] * x[i];
    }
  } else {
#pragma unroll
    for (int i = NB - 1; i >= 0; i--) {
      x[i] *= sRec[i];
#pragma unroll
      for (int j = 0; j < i; j++) x[j] -= sF[j][i] * x[i];
    }
  }
  T *out = inv + ((size_t)b * NB + t) * NB;
#pragma unroll
  for (int i = 0; i < NB; i++) out[i] = x[i];
}

// 4- or 8-byte asynchronous copy global -> shared; valid == false writes zeros without touching the source
template <typename T>
__device__ __forceinline__ void cp_async_elem(T *dst_smem, const T *src, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst_smem);
  const int bytes = valid ? (int)sizeof(T) : 0;
  if (sizeof(T) == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(src), "r"(bytes) : "memory");
  else asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}

// aligned vector load from shared memory: 4 consecutive elements
__device__ __forceinline__ void lds4(const double *p, double (&v)[4]) {
  const double2 a = *reinterpret_cast<const double2 *>(p), b = *reinterpret_cast<const double2 *>(p + 2);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
__device__ __forceinline__ void lds4(const float *p, float (&v)[4]) { const float4 t = *reinterpret_cast<const float4 *>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }

// acc[4][4] += sign * F[kk][row..row+3] (x) X[kk][vec..vec+3] over kk in [0, nk), nk a multiple of 8.  The operands of
// eight k-steps are fetched into registers before the first FMA of the group, and each thread carries 16 independent
// accumulators: one warp per scheduler then has enough FMAs in flight to cover both the shared-memory and the FP64
// latency (a 2 x 4 patch on twice the threads was latency-bound at a third of the pipe's rate).
template <typename T, bool SUBTRACT>
__device__ __forceinline__ void panel_product(T (&acc)[4][4], const T *F, const T *X, int nk) {
  constexpr int U = 8;
  for (int kk0 = 0; kk0 < nk; kk0 += U) {
    T f[U][4], x[U][4];
#pragma unroll
    for (int u = 0; u < U; u++) {
      lds4(F + (kk0 + u) * FS, f[u]);
      lds4(X + (kk0 + u) * XS, x[u]);
    }
#pragma unroll
    for (int u = 0; u < U; u++)
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int c = 0; c < 4; c++) {
          if (SUBTRACT) acc[a][c] -= f[u][a] * x[u][c];
          else acc[a][c] += f[u][a] * x[u][c];
        }
  }
}

constexpr int F_ROWS = NB * (0 + 1 + 2 + 3);  // staged rows of F over the four steps: 0 + 32 + 64 + 96

// Solves F x = alpha b for V vectors per CTA, F = the s x s (s <= BASE) diagonal block at A whose NB x NB diagonal
// sub-block inverses start at `inv`.  Everything the CTA needs — the panel, the strictly triangular part of F in
// step order, the four inverses — is requested with cp.async at entry, one commit group per step, so the global
// latency is paid once and later steps' data arrives behind the earlier steps' arithmetic.
template <typename T, bool LEFT>
__global__ void __launch_bounds__(PANEL_THREADS)
trsm_panel_kernel(const T *__restrict__ A, int lda, T *B, int ldb, int s, int nvec, T alpha, int tr, int lower_f,
                  const T *__restrict__ inv) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T *sX = reinterpret_cast<T *>(smem_raw);   // [BASE][XS]: the panel, row = index along the triangular dimension
  T *sF = sX + BASE * XS;                    // [F_ROWS][FS]: step t's rows at NB * t (t - 1) / 2: sF[kk][i] = F(r0 + i, k0 + kk)
  T *sD = sF + F_ROWS * FS;                  // [nsub][NB][FS]: sD[sb][k][i] = Finv_sb(i, k)
  const int tid = threadIdx.x;
  const int v0 = blockIdx.x * V;
  const int nsub = (s + NB - 1) / NB, rows = nsub * NB;

  // ---- asynchronous copies, one commit group per step: what step t needs (its 32 panel rows, its inverse, its rows of F)
  // lands while the steps before it compute.  Index arithmetic is shifts and masks only.
  for (int step = 0; step < nsub; step++) {
    const int sb = lower_f ? step : nsub - 1 - step, r0 = sb * NB;
    if (LEFT) {  // r fastest: columns of B are contiguous
      for (int idx = tid; idx < NB * V; idx += PANEL_THREADS) {
        const int r = r0 + idx % NB, v = idx / NB;
        const bool ok = r < s && v0 + v < nvec;
        cp_async_elem(sX + r * XS + v, ok ? B + (size_t)(v0 + v) * ldb + r : B, ok);
      }
    } else {     // v fastest: rows of B run along its contiguous dimension
      for (int idx = tid; idx < NB * V; idx += PANEL_THREADS) {
        const int v = idx % V, r = r0 + idx / V;
        const bool ok = r < s && v0 + v < nvec;
        cp_async_elem(sX + r * XS + v, ok ? B + (size_t)r * ldb + v0 + v : B, ok);
      }
    }
    for (int idx = tid; idx < NB * NB; idx += PANEL_THREADS)
      cp_async_elem(sD + (sb * NB + idx / NB) * FS + idx % NB, inv + (size_t)sb * NB * NB + idx, true);
    const int k0 = lower_f ? 0 : r0 + NB, nk = step * NB;  // the rows solved before this step
    T *dst = sF + (NB * step * (step - 1) / 2) * FS;
    for (int idx = tid; idx < nk * NB; idx += PANEL_THREADS) {
      // fastest index = A's contiguous one: i for F(i, k) = A(i, k), k (in runs of NB) for F(i, k) = A(k, i)
      const int lo = idx % NB, mid = (idx / NB) % NB, hi = idx / (NB * NB);
      const int i = tr ? mid : lo, kk = tr ? hi * NB + lo : idx / NB;
      const int gi = r0 + i, gk = k0 + kk;
      const bool ok = gi < s && gk < s;
      cp_async_elem(dst + kk * FS + i, ok ? (tr ? A + (size_t)gi * lda + gk : A + (size_t)gk * lda + gi) : A, ok);
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  }
  // blocks until at most `pending` of this thread's commit groups are still in flight
  auto wait_groups = [](int pending) {
    if (pending >= 3) asm volatile("cp.async.wait_group 3;\n" ::: "memory");
    else if (pending == 2) asm volatile("cp.async.wait_group 2;\n" ::: "memory");
    else if (pending == 1) asm volatile("cp.async.wait_group 1;\n" ::: "memory");
    else asm volatile("cp.async.wait_group 0;\n" ::: "memory");
  };
  wait_groups(nsub - 1);
  __syncthreads();

  // compute thread -> 4 rows x 4 vectors of the 32 x 64 sub-block result; a warp spans all 32 rows and 16 vectors
  const bool computes = tid < COMPUTE_THREADS;
  const int lane = tid % 32;
  const int row = 4 * (lane / 4);                            // 0..28
  const int vec = 16 * ((tid / 32) % 4) + 4 * (lane % 4);    // 0..60

  for (int step = 0; step < nsub; step++) {
    const int sb = lower_f ? step : nsub - 1 - step, r0 = sb * NB;
    const int k0 = lower_f ? 0 : r0 + NB, nk = step * NB;
    const T *F = sF + (NB * step * (step - 1) / 2) * FS;
    const T *D = sD + sb * NB * FS;
    T acc[4][4];
    if (computes) {
      // R = alpha b_sb - F[sb, solved] X[solved]   (alpha was folded into the solved X already)
#pragma unroll
      for (int a = 0; a < 4; a++) {
        lds4(sX + (r0 + row + a) * XS + vec, acc[a]);
#pragma unroll
        for (int c = 0; c < 4; c++) acc[a][c] *= alpha;
      }
      panel_product<T, true>(acc, F + row, sX + k0 * XS + vec, nk);
      // only their owner reads the b_sb entries -> publish R in their place
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int c = 0; c < 4; c++) sX[(r0 + row + a) * XS + vec + c] = acc[a][c];
    }
    __syncthreads();
    if (computes) {
      // X_sb = Finv_sb R
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int c = 0; c < 4; c++) acc[a][c] = T(0);
      panel_product<T, false>(acc, D + row, sX + r0 * XS + vec, NB);
    }
    __syncthreads();  // all reads of R done
    if (computes) {
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int c = 0; c < 4; c++) sX[(r0 + row + a) * XS + vec + c] = acc[a][c];
    }
    wait_groups(nsub - 2 - step);  // the next step's group (no-op after the last one)
    __syncthreads();               // X_sb and the next step's staged data visible to everyone
  }

  // ---- panel out ----
  if (LEFT) {
    for (int idx = tid; idx < BASE * V; idx += PANEL_THREADS) {
      const int r = idx % BASE, v = idx / BASE;
      if (r < s && v0 + v < nvec) B[(size_t)(v0 + v) * ldb + r] = sX[r * XS + v];
    }
  } else {
    for (int idx = tid; idx < rows * V; idx += PANEL_THREADS) {
      const int v = idx % V, r = idx / V;
      if (r < s && v0 + v < nvec) B[(size_t)r * ldb + v0 + v] = sX[r * XS + v];
    }
  }
}

template <typename T>
constexpr size_t panel_smem_bytes() { return sizeof(T) * (size_t(BASE) * XS + size_t(F_ROWS) * FS + size_t(BASE) * FS); }

template <typename T>
struct Problem {
  rtat_b200_context *h;
  bool left, lower, trans, unit;
  int m, n;
  const T *A;
  int lda;
  T *B;
  int ldb;
  int (*gemm)(rtat_b200_context *, int, int, int, int, int, T, const T *, int, const T *, int, T, T *, int);
  int base;       // diagonal blocks up to this size are solved by one trsm_panel_kernel launch
  const T *inv;   // inverted NB x NB diagonal blocks of the whole triangle
  bool tr, lower_f;
};

template <typename T, bool LEFT>
static int launch_panel(const Problem<T> &p, int o, int s, T alpha) {
  auto kern = trsm_panel_kernel<T, LEFT>;
  static std::atomic<uint64_t> configured{0};  // per template instantiation
  if (int st = configure_dynamic_smem(kern, panel_smem_bytes<T>(), p.h->device, configured)) return st;
  const int nvec = LEFT ? p.n : p.m;
  const T *Ablk = p.A + (size_t)o * p.lda + o;
  T *Bblk = LEFT ? p.B + o : p.B + (size_t)o * p.ldb;
  kern<<<(nvec + V - 1) / V, PANEL_THREADS, panel_smem_bytes<T>(), p.h->stream>>>(Ablk, p.lda, Bblk, p.ldb, s, nvec, alpha, p.tr, p.lower_f,
                                                                                   p.inv + (size_t)(o / NB) * NB * NB);
  return check_launch(p.h, "trsm_panel_128x64");
}

// solves the part of the system that lives on [o, o + s) of the triangular dimension
template <typename T>
static int solve(const Problem<T> &p, int o, int s, T alpha) {
  if (s <= p.base) return p.left ? launch_panel<T, true>(p, o, s, alpha) : launch_panel<T, false>(p, o, s, alpha);
  const int blocks = (s + p.base - 1) / p.base;
  const int s1 = (blocks + 1) / 2 * p.base, s2 = s - s1, o1 = o, o2 = o + s1;
  const bool lower_eff = p.lower != p.trans;           // shape of op(A)
  const bool first_first = p.left ? lower_eff : !lower_eff;
  // the off-diagonal block of A that couples the halves lies below the diagonal for a lower A, above for an upper A
  const T *Aoff = p.lower ? p.A + (size_t)o1 * p.lda + o2   // A(o2.., o1..): s2 x s1
                          : p.A + (size_t)o2 * p.lda + o1;  // A(o1.., o2..): s1 x s2
  const int op = p.trans ? RTAT_B200_OP_T : RTAT_B200_OP_N;
  int st;
  if (first_first) {
    if ((st = solve(p, o1, s1, alpha))) return st;
    if (p.left)  // B2 = alpha B2 - op(A)[2,1] X1
      st = p.gemm(p.h, op, RTAT_B200_OP_N, s2, p.n, s1, T(-1), Aoff, p.lda, p.B + o1, p.ldb, alpha, p.B + o2, p.ldb);
    else         // B2 = alpha B2 - X1 op(A)[1,2]
      st = p.gemm(p.h, RTAT_B200_OP_N, op, p.m, s2, s1, T(-1), p.B + (size_t)o1 * p.ldb, p.ldb, Aoff, p.lda, alpha, p.B + (size_t)o2 * p.ldb, p.ldb);
    if (st) return st;
    return solve(p, o2, s2, T(1));
  }
  if ((st = solve(p, o2, s2, alpha))) return st;
  if (p.left)  // B1 = alpha B1 - op(A)[1,2] X2
    st = p.gemm(p.h, op, RTAT_B200_OP_N, s1, p.n, s2, T(-1), Aoff, p.lda, p.B + o2, p.ldb, alpha, p.B + o1, p.ldb);
  else         // B1 = alpha B1 - X2 op(A)[2,1]
    st = p.gemm(p.h, RTAT_B200_OP_N, op, p.m, s1, s2, T(-1), p.B + (size_t)o2 * p.ldb, p.ldb, Aoff, p.lda, alpha, p.B + (size_t)o1 * p.ldb, p.ldb);
  if (st) return st;
  return solve(p, o1, s1, T(1));
}

template <typename T>
static int run(rtat_b200_context *h, int side, int uplo, int trans, int diag, int m, int n, T alpha, const T *A, int lda, T *B, int ldb,
               int (*gemm)(rtat_b200_context *, int, int, int, int, int, T, const T *, int, const T *, int, T, T *, int), const char *name) {
  if (alpha == T(0)) {  // BLAS: A is not referenced, B = 0
    if (cudaMemset2DAsync(B, sizeof(T) * ldb, 0, sizeof(T) * m, n, h->stream) != cudaSuccess) return RTAT_B200_STATUS_EXECUTION_FAILED;
    h->last_kernel = "memset2d";
    return RTAT_B200_STATUS_SUCCESS;
  }
  Problem<T> p{h, side == RTAT_B200_SIDE_LEFT, uplo == RTAT_B200_FILL_LOWER, trans != RTAT_B200_OP_N, diag == RTAT_B200_DIAG_UNIT,
               m, n, A, lda, B, ldb, gemm, 0, nullptr, false, false};
  // trsm.base: recursion cut-off (a multiple of NB, at most BASE); smaller = more of the work in the GEMM kernels
  p.base = opt(h, "trsm.base", 0);
  if (p.base <= 0 || p.base > BASE) p.base = BASE;
  p.base = (p.base + NB - 1) / NB * NB;
  p.tr = p.left ? p.trans : !p.trans;
  p.lower_f = p.tr ? !p.lower : p.lower;
  const int na = p.left ? m : n, nblk = (na + NB - 1) / NB;
  const size_t need = sizeof(T) * (size_t)nblk * NB * NB;
  if (need > h->trsm_inv_bytes) {
    if (h->trsm_inv) {
      cudaStreamSynchronize(h->stream);  // an earlier call may still be reading it
      cudaFree(h->trsm_inv);
      h->trsm_inv = nullptr;
      h->trsm_inv_bytes = 0;
    }
    const size_t want = (need + (size_t(1) << 20) - 1) & ~((size_t(1) << 20) - 1);
    if (cudaMalloc(&h->trsm_inv, want) != cudaSuccess) {
      (void)cudaGetLastError();
      return RTAT_B200_STATUS_ALLOC_FAILED;
    }
    h->trsm_inv_bytes = want;
  }
  p.inv = static_cast<const T *>(h->trsm_inv);
  h->scratch_in_flight = true;
  trsm_diag_inverse_kernel<T><<<nblk, NB, 0, h->stream>>>(A, lda, na, p.tr, p.lower_f, p.unit, static_cast<T *>(h->trsm_inv));
  int st = check_launch(h, "trsm_diag_inverse_32");
  if (st) return st;
  st = solve(p, 0, na, alpha);
  if (!st) h->last_kernel = name;
  return st;
}

}  // namespace trsm

int dtrsm(rtat_b200_context *h, int side, int uplo, int trans, int diag, int m, int n, double alpha, const double *A, int lda, double *B,
          int ldb) {
  return trsm::run<double>(h, side, uplo, trans, diag, m, n, alpha, A, lda, B, ldb, dgemm, "dtrsm_recursive_dmma");
}

int strsm(rtat_b200_context *h, int side, int uplo, int trans, int diag, int m, int n, float alpha, const float *A, int lda, float *B,
          int ldb) {
  return trsm::run<float>(h, side, uplo, trans, diag, m, n, alpha, A, lda, B, ldb, sgemm, "strsm_recursive_3xtf32");
}

}  // namespace rtat_b200
