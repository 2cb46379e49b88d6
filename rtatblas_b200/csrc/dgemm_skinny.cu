// Skinny DGEMM: the thin remainder strips the persistent kernel peels off (dgemm_ws.cu: m % 128 or n % tile in 1..16).
//
//   Y (r x N) = alpha * S (r x K) * X (K x N) + beta * Y,   r <= 4
//
// with generic element strides, so that one kernel serves the row strip of C (S = the last rows of op(A), X = op(B),
// Y = those rows of C) and, transposed, the column strip (S = the last columns of op(B), X = op(A)^T, Y = those columns
// of C read as rows).  Part of the replacement of cublasDgemm behind MatrixMult (reference src/matrix_ops/matrixop.h:15-46).
//
// Such a strip is a handful of GEMVs: the work is reading X once (config 3: 1 x 3001 x 2049, 49 MB), which the tile
// kernel cannot do cheaply — a 64x64 tile per 64 columns of X, each cut across six CTAs by stream-K with a 32 KB partial
// tile parked and fetched per cut, 30-35 us for one row of C.  Here every thread owns output columns, S lives in shared
// memory, X streams through registers, FP64 FMA pipe (r FMAs per element of X).  Measured against the tile kernel on
// 1..16 x 3001 x 2049 and 8192-wide strips (probe/perf_skinny2.py): r = 1 1.3-2.4x faster, r = 4 1.0-1.4x, r = 8 and 16
// slower (the FMA pipe delivers ~6 TFLOP/s here), so strips wider than 4 stay on the tile kernel.
//
//   X contiguous along N (xs_j == 1): thread j of a CTA owns column j; a warp reads 256 contiguous bytes of one k-row.
//   X contiguous along K (xs_k == 1): a warp owns JW columns, lanes stride over k (256 contiguous bytes of one column
//                                     per load) and the r x JW sums are folded with shuffles.
//
// K is cut into chunks, one CTA per (column block, chunk), so that a strip a few thousand columns wide still fills the
// machine; partial sums go to the handle's arena and the LAST CTA of a column block to arrive (a counter per block) adds
// them IN CHUNK ORDER — deterministic whoever is last, like the stream-K fix-up — and applies alpha / beta.  One chunk:
// the CTA finishes Y itself.  (Round 2's first version reduced in a second launch: 3 us of a 15 us strip.)
#include "common.cuh"

namespace rtat_b200 {

int ensure_arena(rtat_b200_context *h, size_t bytes);

namespace skinny {

template <typename T>
struct Params {
  int r, N, K;
  const T *S;
  long long ss_i, ss_k;
  const T *X;
  long long xs_k, xs_j;
  T *Y;
  long long ys_i, ys_j;
  T alpha, beta;
  int kc, nchunks, jblocks;
  T *partial;     // [nchunks][R][N] when nchunks > 1
  unsigned *counters;  // one per column block, zero between launches: the CTA that finds nchunks - 1 adds the partial sums
};

constexpr int THREADS = 128;
constexpr int SKINNY_MAX_ROWS = 4;
constexpr int MAX_COUNTERS = 4096;

// S tile: R x KC elements, zero beyond the strip's r rows and the unit's valid k.  A thread's share is fetched into
// registers one unit ahead (the global latency of the tile would otherwise sit between every two units: a CTA computes a
// unit in a few hundred cycles) and stored into the other of two shared-memory buffers.  The faster index of the fetch
// follows S's contiguous direction.
template <typename T, int R, int KC>
struct STile {
  static constexpr int PER = (R * KC + THREADS - 1) / THREADS;
  T v[PER];
  __device__ __forceinline__ void fetch(const Params<T> &p, int k0, int k_end) {
#pragma unroll
    for (int t = 0; t < PER; t++) {
      const int idx = threadIdx.x + t * THREADS;
      int i, kk;
      if (p.ss_k == 1) { kk = idx % KC; i = idx / KC; } else { i = idx % R; kk = idx / R; }
      v[t] = T(0);
      if (idx < R * KC && i < p.r && k0 + kk < k_end) v[t] = __ldg(p.S + i * p.ss_i + (long long)(k0 + kk) * p.ss_k);
    }
  }
  template <bool K_MAJOR /* sm[kk][i] (true) or sm[i][kk] */>
  __device__ __forceinline__ void store(const Params<T> &p, T *sm) const {
#pragma unroll
    for (int t = 0; t < PER; t++) {
      const int idx = threadIdx.x + t * THREADS;
      int i, kk;
      if (p.ss_k == 1) { kk = idx % KC; i = idx / KC; } else { i = idx % R; kk = idx / R; }
      if (idx < R * KC) sm[K_MAJOR ? kk * R + i : i * KC + kk] = v[t];
    }
  }
};

template <typename T>
__device__ __forceinline__ void finish(const Params<T> &p, int i, int j, T sum) {
  T *y = p.Y + i * p.ys_i + j * p.ys_j;
  T v = p.alpha * sum;
  if (p.beta != T(0)) v += p.beta * *y;
  *y = v;
}

// sum of output (i, j) over the chunks, in chunk order; loads in batches of 8 (one memory latency per batch, not per chunk)
template <typename T, int R>
__device__ __forceinline__ T ordered_sum(const Params<T> &p, int i, int j) {
  const T *src = p.partial + (size_t)i * p.N + j;
  const size_t step = (size_t)R * p.N;
  T sum = T(0);
  int c = 0;
  for (; c + 8 <= p.nchunks; c += 8) {
    T v[8];
#pragma unroll
    for (int u = 0; u < 8; u++) v[u] = __ldcg(src + (size_t)(c + u) * step);
#pragma unroll
    for (int u = 0; u < 8; u++) sum += v[u];
  }
  for (; c < p.nchunks; c++) sum += __ldcg(src + (size_t)c * step);
  return sum;
}

// After a CTA has stored its partial sums: the last of a column block's nchunks CTAs to arrive adds them up — in chunk
// order whoever it is, so the result does not depend on the arrival order — and resets the counter for the next launch.
template <typename T>
__device__ __forceinline__ bool last_of_column_block(const Params<T> &p, int jb) {
  __shared__ unsigned arrived;
  if (!p.counters) return false;  // separate reduction kernel (more column blocks than counters)
  __threadfence();   // this thread's partial sums before the counter
  __syncthreads();
  if (threadIdx.x == 0) {
    arrived = atomicAdd(p.counters + jb, 1u);
    if (arrived == (unsigned)p.nchunks - 1) p.counters[jb] = 0;
  }
  __syncthreads();
  const bool last = arrived == (unsigned)p.nchunks - 1;
  if (last) __threadfence();  // the other CTAs' partial sums after the counter
  return last;
}

// ---- X contiguous along N ---------------------------------------------------------------------------------------
template <typename T, int R, int KC, bool PIPE>
__global__ void __launch_bounds__(THREADS) skinny_xj_kernel(const Params<T> p) {
  __shared__ __align__(16) T sm[PIPE ? 2 : 1][KC * R];
  const int jb = blockIdx.x % p.jblocks, chunk = blockIdx.x / p.jblocks;
  const int k_begin = chunk * p.kc, k_end = min(k_begin + p.kc, p.K);
  const int j = jb * THREADS + threadIdx.x;
  const bool active = j < p.N;
  T acc[R];
#pragma unroll
  for (int i = 0; i < R; i++) acc[i] = T(0);
  STile<T, R, KC> next;
  if (PIPE) next.fetch(p, k_begin, k_end);
  int buf = 0;
  for (int k0 = k_begin; k0 < k_end; k0 += KC, buf ^= PIPE ? 1 : 0) {
    const int kc = min(KC, k_end - k0);
    if (PIPE) {
      next.template store<true>(p, sm[buf]);
      __syncthreads();  // one barrier per unit: buffer `buf` was last read two units ago
      if (k0 + KC < k_end) next.fetch(p, k0 + KC, k_end);
    } else {
      __syncthreads();  // the previous unit's tile has been read
      next.fetch(p, k0, k_end);
      next.template store<true>(p, sm[0]);
      __syncthreads();
    }
    if (!active) continue;
    // X in batches of U, the loads of batch b + 1 issued before the FMAs of batch b: written as a straight batch (loads, then
    // FMAs) ptxas kept 3-5 loads in flight and stalled the first FMA on them (32 registers; SASS, r = 1); a GEMV lives on
    // memory-level parallelism
    const T *xp = p.X + (long long)k0 * p.xs_k + j;
    constexpr int U = 8, NB = KC / U;
    static_assert(NB % 2 == 0, "batches are processed in pairs");
    auto load = [&](T (&x)[U], int kk) {
#pragma unroll
      for (int u = 0; u < U; u++) x[u] = (kk + u < kc) ? __ldg(xp + (long long)(kk + u) * p.xs_k) : T(0);
    };
    auto fmas = [&](const T (&x)[U], int kk) {
#pragma unroll
      for (int u = 0; u < U; u++)
#pragma unroll
        for (int i = 0; i < R; i++) acc[i] = fma(sm[buf][(kk + u) * R + i], x[u], acc[i]);  // rows of sm beyond kc are zero
    };
    T xa[U], xb[U];
    load(xa, 0);
#pragma unroll
    for (int b = 0; b < NB; b += 2) {
      if (b * U >= kc) break;
      load(xb, (b + 1) * U);
      fmas(xa, b * U);
      if (b + 2 < NB) load(xa, (b + 2) * U);
      fmas(xb, (b + 1) * U);
    }
  }
  if (p.nchunks == 1) {
    if (!active) return;
#pragma unroll
    for (int i = 0; i < R; i++)
      if (i < p.r) finish(p, i, j, acc[i]);
    return;
  }
  if (active) {
#pragma unroll
    for (int i = 0; i < R; i++)
      if (i < p.r) __stcg(p.partial + ((size_t)chunk * R + i) * p.N + j, acc[i]);
  }
  if (!last_of_column_block(p, jb) || !active) return;
#pragma unroll
  for (int i = 0; i < R; i++)
    if (i < p.r) finish(p, i, j, ordered_sum<T, R>(p, i, j));
}

// ---- X contiguous along K ---------------------------------------------------------------------------------------
template <typename T, int R, int KC, int JW, bool PIPE>
__global__ void __launch_bounds__(THREADS) skinny_xk_kernel(const Params<T> p) {
  __shared__ __align__(16) T sm[PIPE ? 2 : 1][R * KC];
  const int jb = blockIdx.x % p.jblocks, chunk = blockIdx.x / p.jblocks;
  const int k_begin = chunk * p.kc, k_end = min(k_begin + p.kc, p.K);
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int j0 = (jb * (THREADS / 32) + warp) * JW;
  const bool active = j0 < p.N;
  T acc[JW][R];
#pragma unroll
  for (int jj = 0; jj < JW; jj++)
#pragma unroll
    for (int i = 0; i < R; i++) acc[jj][i] = T(0);
  const T *xp[JW];
#pragma unroll
  for (int jj = 0; jj < JW; jj++) xp[jj] = p.X + (long long)min(j0 + jj, p.N - 1) * p.xs_j;  // clamped: never stored
  T xn[KC / 32][JW];
  auto load_x = [&](int k0) {
#pragma unroll
    for (int t = 0; t < KC / 32; t++)
#pragma unroll
      for (int jj = 0; jj < JW; jj++) xn[t][jj] = (k0 + lane + 32 * t < k_end) ? __ldg(xp[jj] + k0 + lane + 32 * t) : T(0);
  };
  if (active) load_x(k_begin);
  STile<T, R, KC> next;
  if (PIPE) next.fetch(p, k_begin, k_end);
  int buf = 0;
  for (int k0 = k_begin; k0 < k_end; k0 += KC, buf ^= PIPE ? 1 : 0) {
    const int kc = min(KC, k_end - k0);
    if (PIPE) {
      next.template store<false>(p, sm[buf]);
      __syncthreads();
      if (k0 + KC < k_end) next.fetch(p, k0 + KC, k_end);
    } else {
      __syncthreads();
      next.fetch(p, k0, k_end);
      next.template store<false>(p, sm[0]);
      __syncthreads();
    }
    if (!active) continue;
    // this unit's KC / 32 x JW elements of X were asked for one unit ago (xn); the next unit's go out before this one's FMAs
    T xc[KC / 32][JW];
#pragma unroll
    for (int t = 0; t < KC / 32; t++)
#pragma unroll
      for (int jj = 0; jj < JW; jj++) xc[t][jj] = xn[t][jj];
    if (k0 + KC < k_end) load_x(k0 + KC);
#pragma unroll
    for (int t = 0; t < KC / 32; t++) {
      const int kk = lane + 32 * t;
#pragma unroll
      for (int i = 0; i < R; i++) {
        const T sv = sm[buf][i * KC + kk];  // zero beyond kc, and so is xc
#pragma unroll
        for (int jj = 0; jj < JW; jj++) acc[jj][i] = fma(sv, xc[t][jj], acc[jj][i]);
      }
    }
  }
  if (active) {
#pragma unroll
    for (int jj = 0; jj < JW; jj++)
#pragma unroll
      for (int i = 0; i < R; i++) {
        T v = acc[jj][i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        acc[jj][i] = v;
      }
    if (lane == 0) {
#pragma unroll
      for (int jj = 0; jj < JW; jj++) {
        const int j = j0 + jj;
        if (j < p.N) {
#pragma unroll
          for (int i = 0; i < R; i++) {
            if (i < p.r) {
              if (p.nchunks == 1) finish(p, i, j, acc[jj][i]);
              else __stcg(p.partial + ((size_t)chunk * R + i) * p.N + j, acc[jj][i]);
            }
          }
        }
      }
    }
  }
  if (p.nchunks == 1) return;
  if (!last_of_column_block(p, jb)) return;
  // this CTA's (THREADS / 32) * JW columns x r rows: one output per thread
  constexpr int COLS = (THREADS / 32) * JW;
  for (int o = threadIdx.x; o < COLS * R; o += THREADS) {
    const int i = o / COLS, j = jb * COLS + o % COLS;
    if (i < p.r && j < p.N) finish(p, i, j, ordered_sum<T, R>(p, i, j));
  }
}

// partial sums -> Y, chunks added in order: the fallback when a launch has more column blocks than counters
template <typename T>
__global__ void __launch_bounds__(THREADS) skinny_reduce_kernel(const Params<T> p, int R) {
  const long long idx = (long long)blockIdx.x * THREADS + threadIdx.x;
  if (idx >= (long long)p.r * p.N) return;
  const int i = int(idx / p.N), j = int(idx % p.N);
  const T *src = p.partial + (size_t)i * p.N + j;
  const size_t step = (size_t)R * p.N;
  T sum = T(0);
  for (int c = 0; c < p.nchunks; c++) sum += __ldcg(src + (size_t)c * step);
  finish(p, i, j, sum);
}

template <typename T, int R>
static int run(rtat_b200_context *h, Params<T> p) {
  const bool xj = p.xs_j == 1;
  // k per shared-memory S tile (R x KC doubles): short tiles give the N-contiguous GEMVs more CTAs (1 x 8192 x 2048:
  // 42 -> 31 us); the K-contiguous kernel folds its sums with shuffles once per chunk and wants chunks of 512 k or more
  // (256: 37 -> 49 us)
  constexpr int KCJ = R <= 2 ? 64 : 128, KCK = 128;
  constexpr int JW = 4;                         // columns per warp of the K-contiguous kernel
  const int KC = xj ? KCJ : KCK;
  const int cols_per_cta = xj ? THREADS : (THREADS / 32) * JW;
  p.jblocks = (p.N + cols_per_cta - 1) / cols_per_cta;
  // chunks of whole S tiles: enough CTAs for 6-8 per SM
  const int units = (p.K + KC - 1) / KC;
  int nchunks = ((xj ? 8 : 6) * h->sm_count + p.jblocks - 1) / p.jblocks;
  if (nchunks > units) nchunks = units;
  if (!xj && nchunks > (units + 3) / 4) nchunks = (units + 3) / 4;
  if (nchunks < 1) nchunks = 1;
  p.kc = ((units + nchunks - 1) / nchunks) * KC;
  p.nchunks = (p.K + p.kc - 1) / p.kc;
  bool fused = true;
  if (p.nchunks > 1) {
    const size_t bytes = sizeof(T) * (size_t)p.nchunks * R * p.N;
    if (bytes > (size_t(256) << 20)) return RTAT_B200_STATUS_NOT_SUPPORTED;
    if (int st = ensure_arena(h, bytes)) return st;
    p.partial = static_cast<T *>(h->arena);
    fused = p.jblocks <= MAX_COUNTERS && opt(h, "dgemm.skinny_fused", 1) != 0;
    if (fused && !h->skinny_counters) {
      if (cudaMalloc(&h->skinny_counters, MAX_COUNTERS * sizeof(unsigned)) != cudaSuccess ||
          cudaMemsetAsync(h->skinny_counters, 0, MAX_COUNTERS * sizeof(unsigned), h->stream) != cudaSuccess) {
        (void)cudaGetLastError();
        return RTAT_B200_STATUS_ALLOC_FAILED;
      }
    }
    p.counters = fused ? static_cast<unsigned *>(h->skinny_counters) : nullptr;
  }
  const long long grid = (long long)p.jblocks * p.nchunks;
  if (grid > 0x7fffffffLL) return RTAT_B200_STATUS_NOT_SUPPORTED;
  const Params<T> &pk = p;
  // S tile fetched one unit ahead into a second buffer (PIPE) or in place between two barriers: option dgemm.skinny_pipe
  // (0 = choose, 1 / 2 = on / off)
  int pipe = opt(h, "dgemm.skinny_pipe", 0);
  // measured with the X loads pipelined (probe/perf_gemv_vs_cublas.py, us, on / off): N-contiguous X 1 x 8192 x 4096 45.2 / 47.2,
  // 32768 x 1 x 32768 1170 / 1196; K-contiguous X 4 x 8192 x 2048 37.0 / 41.4, 1 x 8192 x 2048 26.8 / 26.8, 32768 x 2 x 32768
  // 1374 / 1281
  if (pipe == 0) pipe = (xj || R >= 4) ? 1 : 2;
  if (xj) {
    if (pipe == 1) skinny_xj_kernel<T, R, KCJ, true><<<(unsigned)grid, THREADS, 0, h->stream>>>(pk);
    else skinny_xj_kernel<T, R, KCJ, false><<<(unsigned)grid, THREADS, 0, h->stream>>>(pk);
  } else {
    if (pipe == 1) skinny_xk_kernel<T, R, KCK, JW, true><<<(unsigned)grid, THREADS, 0, h->stream>>>(pk);
    else skinny_xk_kernel<T, R, KCK, JW, false><<<(unsigned)grid, THREADS, 0, h->stream>>>(pk);
  }
  constexpr bool DBL = sizeof(T) == 8;
  int st = check_launch(h, DBL ? (xj ? "dgemm_skinny_fma_xn" : "dgemm_skinny_fma_xk") : (xj ? "sgemm_skinny_fma_xn" : "sgemm_skinny_fma_xk"));
  if (st || p.nchunks == 1 || fused) return st;
  const long long outs = (long long)p.r * p.N;
  skinny_reduce_kernel<T><<<(unsigned)((outs + THREADS - 1) / THREADS), THREADS, 0, h->stream>>>(p, R);
  return check_launch(h, DBL ? (xj ? "dgemm_skinny_fma_xn+reduce" : "dgemm_skinny_fma_xk+reduce")
                             : (xj ? "sgemm_skinny_fma_xn+reduce" : "sgemm_skinny_fma_xk+reduce"));
}

}  // namespace skinny

// Y (r x N) = alpha S X + beta Y with element strides; NOT_SUPPORTED when the strip is not skinny or X has no unit stride.
template <typename T>
static int skinny_entry(rtat_b200_context *h, int r, int N, int K, T alpha, const T *S, long long ss_i, long long ss_k, const T *X, long long xs_k,
                        long long xs_j, T beta, T *Y, long long ys_i, long long ys_j) {
  // Wider strips are FMA-bound here (FP64, r = 8: 6 TFLOP/s, slower than the tile kernel's padded DMMAs): measured break-even r = 4
  if (r < 1 || r > skinny::SKINNY_MAX_ROWS || N < 1 || K < 1 || (xs_k != 1 && xs_j != 1)) return RTAT_B200_STATUS_NOT_SUPPORTED;
  skinny::Params<T> p{};
  p.r = r; p.N = N; p.K = K;
  p.S = S; p.ss_i = ss_i; p.ss_k = ss_k;
  p.X = X; p.xs_k = xs_k; p.xs_j = xs_j;
  p.Y = Y; p.ys_i = ys_i; p.ys_j = ys_j;
  p.alpha = alpha; p.beta = beta;
  if (r == 1) return skinny::run<T, 1>(h, p);
  if (r == 2) return skinny::run<T, 2>(h, p);
  return skinny::run<T, 4>(h, p);
}

int dgemm_skinny(rtat_b200_context *h, int r, int N, int K, double alpha, const double *S, long long ss_i, long long ss_k, const double *X,
                 long long xs_k, long long xs_j, double beta, double *Y, long long ys_i, long long ys_j) {
  return skinny_entry<double>(h, r, N, K, alpha, S, ss_i, ss_k, X, xs_k, xs_j, beta, Y, ys_i, ys_j);
}

// FP32: plain FP32 FMAs (the arithmetic of the reference's cublasSgemm in its default math mode)
int sgemm_skinny(rtat_b200_context *h, int r, int N, int K, float alpha, const float *S, long long ss_i, long long ss_k, const float *X,
                 long long xs_k, long long xs_j, float beta, float *Y, long long ys_i, long long ys_j) {
  return skinny_entry<float>(h, r, N, K, alpha, S, ss_i, ss_k, X, xs_k, xs_j, beta, Y, ys_i, ys_j);
}

}  // namespace rtat_b200
