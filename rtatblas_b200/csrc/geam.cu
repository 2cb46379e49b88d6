// GEAM kernels:  C = alpha * op(A) + beta * op(B)   (column-major, m x n)
//
// These replace cublas{D,S}geam behind the reference's MatrixMove (transpose / pad / copy,
// reference src/matrix_ops/matrixop.h:307-344) and MatrixAccumulate (in-place fold of a scratch
// product into C, matrixop.h:256-305).  Both are pure HBM streaming: algorithmic traffic is one read
// of every operand that is used and one write of C.
//
//   geam_stream_kernel : no operand transposed.  Each thread moves 16 B vectors (or scalars when an
//                        operand's columns are not 16 B aligned, e.g. ld = 4097 doubles), several
//                        per thread in flight, the grid sweeping each column's chunks in memory order.
//   geam_transpose_tma : (geam_tma.cu) A transposed, operands TMA-addressable: TMA load -> swizzled smem ->
//                        conflict-free transposed smem-to-smem pass -> TMA store.
//   geam_tile_kernel   : at least one operand transposed (any alignment).  64x64 tiles are staged through padded
//                        shared memory (odd pitch => conflict-free transposed reads); global loads
//                        run along the operand's contiguous dimension and global stores along C's.
//
// Arithmetic (pinned bit-for-bit against cuBLAS 12.9 geam through tests/golden): beta == 0 =>
// c = alpha*a with B never read; alpha == 0 => c = beta*b with A never read; otherwise
// c = fma(beta, b, RN(alpha*a)).
#include <type_traits>
#include "common.cuh"

namespace rtat_b200 {

template <typename T>
bool geam_transpose_tma_supported(int m, int n, const T *A, int lda, const T *B, int ldb, const T *C, int ldc, bool use_b);
template <typename T>
int geam_transpose_tma(rtat_b200_context *h, int m, int n, T alpha, const T *A, int lda, T beta, const T *B, int ldb, T *C, int ldc);

template <typename T>
__device__ __forceinline__ T mul_rn(T a, T b);
template <>
__device__ __forceinline__ double mul_rn<double>(double a, double b) { return __dmul_rn(a, b); }
template <>
__device__ __forceinline__ float mul_rn<float>(float a, float b) { return __fmul_rn(a, b); }
template <typename T>
__device__ __forceinline__ T add_rn(T a, T b);
template <>
__device__ __forceinline__ double add_rn<double>(double a, double b) { return __dadd_rn(a, b); }
template <>
__device__ __forceinline__ float add_rn<float>(float a, float b) { return __fadd_rn(a, b); }
template <typename T>
__device__ __forceinline__ T fma_rn(T a, T b, T c);
template <>
__device__ __forceinline__ double fma_rn<double>(double a, double b, double c) { return __fma_rn(a, b, c); }
template <>
__device__ __forceinline__ float fma_rn<float>(float a, float b, float c) { return __fmaf_rn(a, b, c); }

// MODE: 0 = alpha*a only, 1 = beta*b only, 2 = mul,mul,add  3 = fma(alpha,a,beta*b)  4 = fma(beta,b,alpha*a)
template <typename T, int MODE>
__device__ __forceinline__ T axpby(T alpha, T a, T beta, T b) {
  if (MODE == 0) return mul_rn(alpha, a);
  if (MODE == 1) return mul_rn(beta, b);
  if (MODE == 2) return add_rn(mul_rn(alpha, a), mul_rn(beta, b));
  if (MODE == 3) return fma_rn(alpha, a, mul_rn(beta, b));
  return fma_rn(beta, b, mul_rn(alpha, a));
}

template <typename T>
struct Vec16;
template <>
struct Vec16<double> { using type = double2; static constexpr int N = 2; };
template <>
struct Vec16<float> { using type = float4; static constexpr int N = 4; };

// 256-bit global accesses (LDG.E.256 / STG.E.256, new on sm_100): 4 doubles or 8 floats per instruction
template <typename T>
struct Vec32;
template <>
struct Vec32<double> {
  static constexpr int N = 4;
  // HINT: 0 = default caching, 1 = no L1 allocation + L2 evict-first (data used once), 2 = .cs (streaming) on both sides
  template <int HINT>
  static __device__ __forceinline__ void load(double (&v)[4], const double *p) {
    if (HINT == 1) asm volatile("ld.global.L1::no_allocate.L2::evict_first.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
    else if (HINT == 2) asm volatile("ld.global.cs.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
    else asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
  }
  template <int HINT>
  static __device__ __forceinline__ void store(double *p, const double (&v)[4]) {
    if (HINT == 1) asm volatile("st.global.L1::no_allocate.L2::evict_first.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
    else if (HINT == 2) asm volatile("st.global.cs.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
    else asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
  }
};
template <>
struct Vec32<float> {
  static constexpr int N = 8;
#define RTAT_LD8(Q) asm volatile("ld.global" Q ".v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p))
#define RTAT_ST8(Q) asm volatile("st.global" Q ".v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory")
  template <int HINT>
  static __device__ __forceinline__ void load(float (&v)[8], const float *p) {
    if (HINT == 1) RTAT_LD8(".L1::no_allocate.L2::evict_first");
    else if (HINT == 2) RTAT_LD8(".cs");
    else RTAT_LD8("");
  }
  template <int HINT>
  static __device__ __forceinline__ void store(float *p, const float (&v)[8]) {
    if (HINT == 1) RTAT_ST8(".L1::no_allocate.L2::evict_first");
    else if (HINT == 2) RTAT_ST8(".cs");
    else RTAT_ST8("");
  }
#undef RTAT_LD8
#undef RTAT_ST8
};

// streaming kernel on 32 B vectors (every operand 32 B aligned, ld and m multiples of 32 B): same chunk
// walk as geam_stream_kernel below
template <typename T, int MODE, int UNR, int HINT>
__global__ void __launch_bounds__(256)
geam_stream256_kernel(int m, int n, T alpha, const T *__restrict__ A, int lda, T beta, const T *B, int ldb, T *C, int ldc,
                      int chunks_per_col) {
  constexpr int VN = Vec32<T>::N;
  constexpr int CHUNK = 256 * VN * UNR;
  const long long total = (long long)chunks_per_col * n;
  for (long long w = blockIdx.x; w < total; w += gridDim.x) {
    const int j = int(w / chunks_per_col);
    const int i0 = int(w % chunks_per_col) * CHUNK + threadIdx.x * VN;
    const T *a = A + (size_t)j * lda, *b = B + (size_t)j * ldb;
    T *c = C + (size_t)j * ldc;
    T va[UNR][VN], vb[UNR][VN];
#pragma unroll
    for (int u = 0; u < UNR; u++) {
      const int i = i0 + u * 256 * VN;
      if (i < m) {
        if (MODE != 1) Vec32<T>::template load<HINT>(va[u], a + i);
        if (MODE != 0) Vec32<T>::template load<HINT>(vb[u], b + i);
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; u++) {
      const int i = i0 + u * 256 * VN;
      if (i < m) {
        T out[VN];
#pragma unroll
        for (int e = 0; e < VN; e++) out[e] = axpby<T, MODE>(alpha, MODE != 1 ? va[u][e] : T(0), beta, MODE != 0 ? vb[u][e] : T(0));
        Vec32<T>::template store<HINT>(c + i, out);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// streaming kernel (no transposes).  Work item = one chunk of 256*VN*UNR consecutive elements of one
// column; consecutive CTAs take consecutive chunks, so for tight leading dimensions the whole grid
// sweeps memory linearly (DRAM-page friendly), and every thread keeps UNR 16 B requests in flight.
// ---------------------------------------------------------------------------------------------
template <typename T, int MODE, bool VEC, int UNR>
__global__ void __launch_bounds__(256)
geam_stream_kernel(int m, int n, T alpha, const T *__restrict__ A, int lda, T beta, const T *B, int ldb, T *C, int ldc,
                   int chunks_per_col, int chunk_len) {
  constexpr int VN = VEC ? Vec16<T>::N : 1;
  using V = typename Vec16<T>::type;
  const long long total = (long long)chunks_per_col * n;
  for (long long w = blockIdx.x; w < total; w += gridDim.x) {
    const int j = int(w / chunks_per_col);
    // chunk_len <= 256 * VN * UNR elements of the column, the same for every chunk (the dispatcher cuts a column into equal
    // parts: m = 4097 on 2048-element chunks left every third CTA a single element)
    const int i0 = int(w % chunks_per_col) * chunk_len + threadIdx.x * VN;
    const int m_end = min(m, int(w % chunks_per_col) * chunk_len + chunk_len);
    const T *a = A + (size_t)j * lda, *b = B + (size_t)j * ldb;
    T *c = C + (size_t)j * ldc;
    if (VEC) {
      V va[UNR], vb[UNR];
#pragma unroll
      for (int u = 0; u < UNR; u++) {
        const int i = i0 + u * 256 * VN;
        if (i < m_end) {  // m % VN == 0 and chunk_len % VN == 0 are guaranteed by the dispatcher for the vector path
          if (MODE != 1) va[u] = __ldcs(reinterpret_cast<const V *>(a + i));
          if (MODE != 0) vb[u] = *reinterpret_cast<const V *>(b + i);
        }
      }
#pragma unroll
      for (int u = 0; u < UNR; u++) {
        const int i = i0 + u * 256 * VN;
        if (i < m_end) {
          V out;
          T *o = reinterpret_cast<T *>(&out);
          const T *pa = reinterpret_cast<const T *>(&va[u]);
          const T *pb = reinterpret_cast<const T *>(&vb[u]);
#pragma unroll
          for (int e = 0; e < VN; e++) o[e] = axpby<T, MODE>(alpha, MODE != 1 ? pa[e] : T(0), beta, MODE != 0 ? pb[e] : T(0));
          __stcs(reinterpret_cast<V *>(c + i), out);
        }
      }
    } else {
      T va[UNR], vb[UNR];
#pragma unroll
      for (int u = 0; u < UNR; u++) {
        const int i = i0 + u * 256;
        if (i < m_end) {
          if (MODE != 1) va[u] = __ldcs(a + i);
          if (MODE != 0) vb[u] = b[i];
        }
      }
#pragma unroll
      for (int u = 0; u < UNR; u++) {
        const int i = i0 + u * 256;
        if (i < m_end) __stcs(c + i, axpby<T, MODE>(alpha, MODE != 1 ? va[u] : T(0), beta, MODE != 0 ? vb[u] : T(0)));
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// tile kernel (some operand transposed).  TILE x TILE elements, 256 threads.
// Operand X contributes x(i,j) = X[j + i*ldx] when transposed (X is n x m), X[i + j*ldx] otherwise.
// ---------------------------------------------------------------------------------------------
template <typename T, int MODE, bool TA, bool TB, int TILE>
__global__ void __launch_bounds__(256)
geam_tile_kernel(int m, int n, T alpha, const T *__restrict__ A, int lda, T beta, const T *B, int ldb,
                 T *C, int ldc, int tiles_m, int tiles_n) {
  constexpr int PITCH = TILE + 1;
  constexpr int YS = 256 / TILE;        // thread rows per pass
  constexpr int PER = TILE / YS;        // elements per thread
  constexpr bool USE_A = MODE != 1, USE_B = MODE != 0;
  __shared__ T sa[(TA && USE_A) ? TILE * PITCH : 1];
  __shared__ T sb[(TB && USE_B) ? TILE * PITCH : 1];
  const int tx = threadIdx.x % TILE, ty = threadIdx.x / TILE;  // ty in 0..YS-1
  for (long long tile = blockIdx.x; tile < (long long)tiles_m * tiles_n; tile += gridDim.x) {
    int i0 = int(tile % tiles_m) * TILE, j0 = int(tile / tiles_m) * TILE;
    // stage transposed operands: global reads run along the operand's own contiguous dimension (j)
    if (TA && USE_A) {
      int j = j0 + tx;
#pragma unroll 8
      for (int r = ty; r < TILE; r += YS) {
        int i = i0 + r;
        if (i < m && j < n) sa[r * PITCH + tx] = A[(size_t)i * lda + j];
      }
    }
    if (TB && USE_B) {
      int j = j0 + tx;
#pragma unroll 8
      for (int r = ty; r < TILE; r += YS) {
        int i = i0 + r;
        if (i < m && j < n) sb[r * PITCH + tx] = B[(size_t)i * ldb + j];
      }
    }
    if ((TA && USE_A) || (TB && USE_B)) __syncthreads();
    {
      int i = i0 + tx;
      if (i < m) {
        T va[PER], vb[PER];
#pragma unroll
        for (int r = 0; r < PER; r++) {
          int jj = ty + YS * r, j = j0 + jj;
          if (j < n) {
            if (USE_A) va[r] = TA ? sa[tx * PITCH + jj] : A[(size_t)j * lda + i];
            if (USE_B) vb[r] = TB ? sb[tx * PITCH + jj] : B[(size_t)j * ldb + i];
          }
        }
#pragma unroll
        for (int r = 0; r < PER; r++) {
          int j = j0 + ty + YS * r;
          if (j < n) C[(size_t)j * ldc + i] = axpby<T, MODE>(alpha, USE_A ? va[r] : T(0), beta, USE_B ? vb[r] : T(0));
        }
      }
    }
    if ((TA && USE_A) || (TB && USE_B)) __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// register-blocked transpose (A transposed, B not):  C (m x n) = alpha * A^T + beta * B,  A is n x m.
// One CTA = one tile of TJ (A's contiguous dimension, j) x TI (C's contiguous dimension, i) elements, launched as a plain
// grid of small CTAs (several resident per SM, the hardware scheduler balancing them), every global access a full
// 16-byte vector where the operand allows it:
//   phase 1  a thread loads E = 16 / sizeof(T) vectors from E consecutive rows i of A (a warp covers 512 contiguous bytes
//            of each row), transposes the E x E block in registers and stores E 16-byte UNITS — E consecutive i of one
//            column j — to shared memory.  Units of column j sit at j * UPC + (u ^ swz(j)): the XOR makes the eight lanes of
//            a quarter warp, which write eight different columns, hit eight different 16-byte bank groups.
//   phase 2  a warp reads whole columns of units back (a permutation of consecutive addresses: conflict-free), applies
//            alpha / beta and stores 512 contiguous bytes of a column of C per instruction.
// VEC_IN / VEC_OUT = false: that side is accessed element by element, still one contiguous run per warp instruction
// (odd leading dimensions, bases off 16 bytes, extents that are not a multiple of E): then a lane owns columns
// lane + 32 e instead of E lane + e (phase 1) or reads single elements of the units (phase 2).
// ---------------------------------------------------------------------------------------------
template <typename T>
struct Unit16 { T v[16 / sizeof(T)]; };

#ifndef RTAT_GEAM_TR_MINB
#define RTAT_GEAM_TR_MINB 1  // probe builds raise this (resident CTAs per SM the register allocation must allow)
#endif
template <typename T, bool USE_B, bool VEC_IN, bool VEC_OUT, int UPC>
__global__ void __launch_bounds__(256, RTAT_GEAM_TR_MINB)
geam_transpose_reg_kernel(int m, int n, T alpha, const T *__restrict__ A, int lda, T beta, const T *B, int ldb, T *C, int ldc,
                          int tiles_m, int tiles_n, int order) {
  constexpr int E = 16 / sizeof(T);
  constexpr int TJ = 32 * E;                       // 64 doubles / 128 floats: 512 B of every row of A
  constexpr int TI = UPC * E;                      // UPC units per column: 32 or 64 elements
  static_assert(UPC % 8 == 0 && TI % 32 == 0 && UPC <= 32, "swizzle period; scalar phase 2 moves 32 elements of a column per instruction");
  using V = typename Vec16<T>::type;
  __shared__ __align__(16) T s[TJ * TI];           // 16 or 32 KB
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  // order 0: consecutive CTAs walk down C's contiguous dimension (adjacent 512 B pieces of the same columns of C);
  // order 1: along A's contiguous dimension (adjacent pieces of the same rows of A)
  const int ti = order ? blockIdx.x / tiles_n : blockIdx.x % tiles_m, tj = order ? blockIdx.x % tiles_n : blockIdx.x / tiles_m;
  const int i0 = ti * TI, j0 = tj * TJ;
  auto swz = [](int jl) { return (VEC_IN ? jl / E : jl) & 7; };
  auto col_of = [&](int e) { return VEC_IN ? E * lane + e : lane + 32 * e; };

  // ---- phase 1
#pragma unroll
  for (int g = warp; g < UPC; g += 8) {
    T a[E][E];
#pragma unroll
    for (int r = 0; r < E; r++) {
      const int i = i0 + g * E + r;
      if (VEC_IN) {
        const int j = j0 + E * lane;
        V v = V();
        if (i < m && j < n) v = __ldcs(reinterpret_cast<const V *>(A + (size_t)i * lda + j));
        const T *pv = reinterpret_cast<const T *>(&v);
#pragma unroll
        for (int e = 0; e < E; e++) a[r][e] = pv[e];
      } else {
#pragma unroll
        for (int e = 0; e < E; e++) {
          const int j = j0 + lane + 32 * e;
          a[r][e] = (i < m && j < n) ? __ldcs(A + (size_t)i * lda + j) : T(0);
        }
      }
    }
#pragma unroll
    for (int e = 0; e < E; e++) {
      const int jl = col_of(e);
      Unit16<T> u;
#pragma unroll
      for (int r = 0; r < E; r++) u.v[r] = a[r][e];
      *reinterpret_cast<V *>(&s[(jl * UPC + (g ^ swz(jl))) * E]) = *reinterpret_cast<V *>(&u);
    }
  }
  __syncthreads();
  // ---- phase 2
  if (VEC_OUT) {
    constexpr int CPI = 32 / UPC;            // columns per warp instruction
    constexpr int ITER = TJ / CPI / 8;       // 8
    const int u = lane % UPC;
    const int i = i0 + u * E;
    V bv[ITER];
    if (USE_B) {
#pragma unroll
      for (int t = 0; t < ITER; t++) {
        const int j = j0 + (t * 8 + warp) * CPI + lane / UPC;
        if (j < n && i < m) bv[t] = *reinterpret_cast<const V *>(B + (size_t)j * ldb + i);
      }
    }
#pragma unroll
    for (int t = 0; t < ITER; t++) {
      const int jl = (t * 8 + warp) * CPI + lane / UPC, j = j0 + jl;
      if (j < n && i < m) {
        V av = *reinterpret_cast<const V *>(&s[(jl * UPC + (u ^ swz(jl))) * E]);
        const T *pa = reinterpret_cast<const T *>(&av);
        const T *pb = reinterpret_cast<const T *>(&bv[t]);
        V out;
        T *po = reinterpret_cast<T *>(&out);
#pragma unroll
        for (int e = 0; e < E; e++) po[e] = axpby<T, USE_B ? 4 : 0>(alpha, pa[e], beta, USE_B ? pb[e] : T(0));
        __stcs(reinterpret_cast<V *>(C + (size_t)j * ldc + i), out);
      }
    }
  } else {
    constexpr int RPC = TI / 32;             // warp instructions per column (2)
    constexpr int ITER = TJ * RPC / 8;       // 16 / 32
    constexpr int BATCH = 8;
#pragma unroll
    for (int t0 = 0; t0 < ITER; t0 += BATCH) {
      T bv[BATCH];
      if (USE_B) {
#pragma unroll
        for (int t = 0; t < BATCH; t++) {
          const int x = (t0 + t) * 8 + warp, jl = x / RPC, il = (x % RPC) * 32 + lane;
          if (j0 + jl < n && i0 + il < m) bv[t] = B[(size_t)(j0 + jl) * ldb + i0 + il];
        }
      }
#pragma unroll
      for (int t = 0; t < BATCH; t++) {
        const int x = (t0 + t) * 8 + warp, jl = x / RPC, il = (x % RPC) * 32 + lane;
        if (j0 + jl < n && i0 + il < m) {
          const T av = s[(jl * UPC + ((il / E) ^ swz(jl))) * E + il % E];
          __stcs(C + (size_t)(j0 + jl) * ldc + i0 + il, axpby<T, USE_B ? 4 : 0>(alpha, av, beta, USE_B ? bv[t] : T(0)));
        }
      }
    }
  }
}

template <typename T, int MODE>
static int launch_geam(rtat_b200_context *h, int transa, int transb, int m, int n, T alpha, const T *A,
                       int lda, T beta, const T *B, int ldb, T *C, int ldc) {
  const bool ta = transa == RTAT_B200_OP_T && MODE != 1, tb = transb == RTAT_B200_OP_T && MODE != 0;
  cudaStream_t s = h->stream;
  const int max_ctas = h->sm_count * 8;
  if (!ta && !tb) {
    constexpr int VN = Vec16<T>::N;
    auto aligned = [&](const T *p, int ld) { return (reinterpret_cast<uintptr_t>(p) % 16 == 0) && (ld % VN == 0); };
    bool vec = (m % VN == 0) && aligned(C, ldc) && (MODE == 1 || aligned(A, lda)) && (MODE == 0 || aligned(B, ldb));
    if (opt(h, "geam.variant") == 1) vec = false;
    constexpr int UNR_V = 4, UNR_S = 8;
    {
      constexpr int VN32 = Vec32<T>::N;
      auto aligned32 = [&](const T *p, int ld) { return (reinterpret_cast<uintptr_t>(p) % 32 == 0) && (ld % VN32 == 0); };
      bool vec32 = (m % VN32 == 0) && aligned32(C, ldc) && (MODE == 1 || aligned32(A, lda)) && (MODE == 0 || aligned32(B, ldb)) &&
                   opt(h, "geam.variant") == 0;
      if (vec32) {
        const int tune = opt(h, "geam.tune");
        auto go = [&](auto unr_tag, int ctas_per_sm, auto hint_tag) {
          constexpr int UNR = decltype(unr_tag)::value;
          constexpr int HINT = decltype(hint_tag)::value;
          const int chunk32 = 256 * VN32 * UNR;
          const int cpc = (m + chunk32 - 1) / chunk32;
          const long long tot = (long long)cpc * n;
          const long long cap = (long long)h->sm_count * ctas_per_sm;
          const int g = (int)(tot < cap ? tot : cap);
          geam_stream256_kernel<T, MODE, UNR, HINT><<<g, 256, 0, s>>>(m, n, alpha, A, lda, beta, B, ldb, C, ldc, cpc);
          return check_launch(h, "geam_stream_vec32");
        };
        // measured on 8192^2 f64 (GB/s): UNR 4 x 8 CTAs/SM 6333 | 1 x 16: 6252 | 2 x 8: 6082 | 2 x 4: 6051 | 1 x 8: 5975
        if (tune == 1) return go(std::integral_constant<int, 2>{}, 8, std::integral_constant<int, 0>{});
        if (tune == 2) return go(std::integral_constant<int, 1>{}, 16, std::integral_constant<int, 0>{});
        if (tune == 3) return go(std::integral_constant<int, 4>{}, 8, std::integral_constant<int, 1>{});
        if (tune == 4) return go(std::integral_constant<int, 4>{}, 8, std::integral_constant<int, 2>{});
        if (tune == 5) return go(std::integral_constant<int, 2>{}, 16, std::integral_constant<int, 1>{});
        if (tune == 6) return go(std::integral_constant<int, 8>{}, 4, std::integral_constant<int, 1>{});
        // one chunk per CTA, as many CTAs as chunks (the hardware scheduler hands them out): 8 / 16 / 32 KB per CTA
        if (tune == 7) return go(std::integral_constant<int, 1>{}, 1 << 20, std::integral_constant<int, 0>{});
        if (tune == 8) return go(std::integral_constant<int, 2>{}, 1 << 20, std::integral_constant<int, 0>{});
        if (tune == 9) return go(std::integral_constant<int, 4>{}, 1 << 20, std::integral_constant<int, 0>{});
        if (tune == 10) return go(std::integral_constant<int, 1>{}, 1 << 20, std::integral_constant<int, 2>{});
        if (tune == 11) return go(std::integral_constant<int, 2>{}, 1 << 20, std::integral_constant<int, 2>{});
        if (tune == 12) return go(std::integral_constant<int, 4>{}, 8, std::integral_constant<int, 0>{});  // the persistent grid of round 1
        // default: one 8 KB chunk per CTA.  8192^2 f64: 6784 GB/s against 6429 for the persistent grid and 6655-6810 for cublasDgeam
        return go(std::integral_constant<int, 1>{}, 1 << 20, std::integral_constant<int, 0>{});
      }
    }
    const int chunk_max = 256 * (vec ? VN * UNR_V : UNR_S);
    const int chunks_per_col = (m + chunk_max - 1) / chunk_max;
    // equal parts of a column, rounded up to whole warps of vectors
    const int gran = 32 * (vec ? VN : 1);
    const int chunk_len = opt(h, "geam.even_chunks", 1) ? ((m + chunks_per_col - 1) / chunks_per_col + gran - 1) / gran * gran : chunk_max;
    const long long total = (long long)chunks_per_col * n;
    // one chunk per CTA (see the 32-byte path above); geam.persistent = 1 restores the grid-stride walk over 8 CTAs per SM
    const long long cap = opt(h, "geam.persistent") ? (long long)max_ctas : 0x7fffffffLL;
    const int grid = (int)(total < cap ? total : cap);
    if (vec) {
      geam_stream_kernel<T, MODE, true, UNR_V><<<grid, 256, 0, s>>>(m, n, alpha, A, lda, beta, B, ldb, C, ldc, chunks_per_col, chunk_len);
      return check_launch(h, "geam_stream_vec16");
    }
    geam_stream_kernel<T, MODE, false, UNR_S><<<grid, 256, 0, s>>>(m, n, alpha, A, lda, beta, B, ldb, C, ldc, chunks_per_col, chunk_len);
    return check_launch(h, "geam_stream_scalar");
  }
  const int TL = (ta && tb) ? 32 : 64;
  int tiles_m = (m + TL - 1) / TL, tiles_n = (n + TL - 1) / TL;
  long long tiles = (long long)tiles_m * tiles_n;
  int grid = (int)(tiles < max_ctas ? tiles : max_ctas);
  // geam.transpose: 0 = choose, 1 = TMA pipeline (geam_tma.cu), 2 = register-blocked kernel above, 3 = scalar tile kernel
  const int tsel = opt(h, "geam.transpose");
  if (ta && !tb && (MODE == 0 || MODE == 4) && opt(h, "geam.variant") != 1 && (tsel == 0 || tsel == 2)) {
    constexpr int E = 16 / (int)sizeof(T);
    constexpr int TJ = 32 * E;
    // units (16 B) per column of the tile: tiles of 64 x 32 doubles / 128 x 32 floats (16 KB) by default, twice that with
    // geam.upc = 2 (measured on 8192^2: see DESIGN.md §3)
    constexpr int UPC_S = 32 / E, UPC_L = 64 / E;
    const bool large = opt(h, "geam.upc") == 2;
    const int TI = (large ? UPC_L : UPC_S) * E;
    auto al16 = [&](const T *p, int ld) { return reinterpret_cast<uintptr_t>(p) % 16 == 0 && ld % E == 0; };
    const bool vin = al16(A, lda) && n % E == 0;
    const bool vout = al16(C, ldc) && m % E == 0 && (MODE == 0 || al16(B, ldb));
    const int tm = (m + TI - 1) / TI, tn = (n + TJ - 1) / TJ;
    if ((long long)tm * tn <= 0x7fffffffLL) {
      const int order = opt(h, "geam.order");
      constexpr bool UB = MODE != 0;
#define RTAT_TR(VI, VO)                                                                                                                 \
  if (large) geam_transpose_reg_kernel<T, UB, VI, VO, UPC_L><<<tm * tn, 256, 0, s>>>(m, n, alpha, A, lda, beta, B, ldb, C, ldc, tm, tn, order); \
  else geam_transpose_reg_kernel<T, UB, VI, VO, UPC_S><<<tm * tn, 256, 0, s>>>(m, n, alpha, A, lda, beta, B, ldb, C, ldc, tm, tn, order)
      if (vin && vout) { RTAT_TR(true, true); return check_launch(h, "geam_transpose_reg_v16_v16"); }
      if (vin) { RTAT_TR(true, false); return check_launch(h, "geam_transpose_reg_v16_s"); }
      if (vout) { RTAT_TR(false, true); return check_launch(h, "geam_transpose_reg_s_v16"); }
      RTAT_TR(false, false);
      return check_launch(h, "geam_transpose_reg_s_s");
#undef RTAT_TR
    }
  }
  if (ta && !tb && (MODE == 0 || MODE == 4) && opt(h, "geam.variant") != 1 && tsel != 3 &&
      geam_transpose_tma_supported<T>(m, n, A, lda, B, ldb, C, ldc, MODE != 0)) {
    int st = geam_transpose_tma<T>(h, m, n, alpha, A, lda, MODE == 0 ? T(0) : beta, B, ldb, C, ldc);
    if (st != RTAT_B200_STATUS_NOT_SUPPORTED) return st;
  }
  if (ta && !tb) {
    geam_tile_kernel<T, MODE, true, false, 64><<<grid, 256, 0, s>>>(m, n, alpha, A, lda, beta, B, ldb, C, ldc, tiles_m, tiles_n);
    return check_launch(h, "geam_tile_tn");
  } else if (!ta && tb) {
    geam_tile_kernel<T, MODE, false, true, 64><<<grid, 256, 0, s>>>(m, n, alpha, A, lda, beta, B, ldb, C, ldc, tiles_m, tiles_n);
    return check_launch(h, "geam_tile_nt");
  }
  geam_tile_kernel<T, MODE, true, true, 32><<<grid, 256, 0, s>>>(m, n, alpha, A, lda, beta, B, ldb, C, ldc, tiles_m, tiles_n);
  return check_launch(h, "geam_tile_tt");
}

template <typename T>
int geam(rtat_b200_context *h, int transa, int transb, int m, int n, T alpha, const T *A, int lda,
         T beta, const T *B, int ldb, T *C, int ldc) {
  if (alpha == T(0) && beta == T(0)) {
    // neither operand is referenced: C = 0 (MatrixSyrkAlloc clears its scratch this way, reference matrixop.h:612)
    if (cudaMemset2DAsync(C, sizeof(T) * ldc, 0, sizeof(T) * m, n, h->stream) != cudaSuccess) return RTAT_B200_STATUS_EXECUTION_FAILED;
    h->last_kernel = "memset2d";
    return RTAT_B200_STATUS_SUCCESS;
  }
  if (beta == T(0)) return launch_geam<T, 0>(h, transa, transb, m, n, alpha, A, lda, beta, B, ldb, C, ldc);
  if (alpha == T(0)) return launch_geam<T, 1>(h, transa, transb, m, n, alpha, A, lda, beta, B, ldb, C, ldc);
  switch (opt(h, "geam.rounding", 4)) {  // 4 = cuBLAS behaviour; 2/3 kept for experiments
    case 2: return launch_geam<T, 2>(h, transa, transb, m, n, alpha, A, lda, beta, B, ldb, C, ldc);
    case 3: return launch_geam<T, 3>(h, transa, transb, m, n, alpha, A, lda, beta, B, ldb, C, ldc);
    default: return launch_geam<T, 4>(h, transa, transb, m, n, alpha, A, lda, beta, B, ldb, C, ldc);
  }
}

template int geam<double>(rtat_b200_context *, int, int, int, int, double, const double *, int, double, const double *, int, double *, int);
template int geam<float>(rtat_b200_context *, int, int, int, int, float, const float *, int, float, const float *, int, float *, int);

// ---------------------------------------------------------------------------------------------
// seeded counter-based uniform fill (drivers / tests only)
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

template <typename T>
__global__ void fill_uniform_kernel(T *x, size_t n, uint64_t seed, T lo, T hi) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += stride) {
    uint64_t r = splitmix64(seed * 0xD1342543DE82EF95ull + i);
    // 53 random bits -> u in (0, 1]
    double u = (double)((r >> 11) + 1) * (1.0 / 9007199254740992.0);
    x[i] = (T)((double)lo + ((double)hi - (double)lo) * u);
  }
}

template <typename T>
int fill_uniform(rtat_b200_context *h, T *x, size_t n, uint64_t seed, T lo, T hi) {
  size_t blocks = (n + 255) / 256;
  size_t cap = (size_t)h->sm_count * 16;
  fill_uniform_kernel<T><<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, h->stream>>>(x, n, seed, lo, hi);
  return check_launch(h, "fill_uniform");
}
template int fill_uniform<double>(rtat_b200_context *, double *, size_t, uint64_t, double, double);
template int fill_uniform<float>(rtat_b200_context *, float *, size_t, uint64_t, float, float);

}  // namespace rtat_b200
