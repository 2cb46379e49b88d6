// Probe: cuBLAS 12.9 throughput on the BASELINE configs (the reference path's vendor kernels).
// Not part of the product; gives the "90 % of cuBLAS" bar and kernel names (run under ncu).
// Build: nvcc -O3 -o cublas_ref cublas_ref.cu -lcublas
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cublas_v2.h>

template <typename F>
float best_ms(F f, int reps) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); f(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int i = 0; i < reps; i++) {
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
  }
  return best;
}

__global__ void fill(double *p, size_t n, unsigned seed) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  for (; i < n; i += (size_t)gridDim.x * blockDim.x) {
    unsigned long long x = (i + 1) * 0x9E3779B97F4A7C15ull + seed; x ^= x >> 31; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 29;
    p[i] = (double)(x >> 11) * (2.0 / 9007199254740992.0) - 1.0;
  }
}
__global__ void fillf(float *p, size_t n, unsigned seed) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  for (; i < n; i += (size_t)gridDim.x * blockDim.x) {
    unsigned long long x = (i + 1) * 0x9E3779B97F4A7C15ull + seed; x ^= x >> 31; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 29;
    p[i] = (float)((double)(x >> 11) * (2.0 / 9007199254740992.0) - 1.0);
  }
}

int main(int argc, char **argv) {
  int reps = argc > 1 ? atoi(argv[1]) : 5;
  cublasHandle_t h; cublasCreate(&h);
  size_t big = (size_t)8192 * 8192;
  double *A, *B, *C; cudaMalloc(&A, big * 8); cudaMalloc(&B, big * 8); cudaMalloc(&C, big * 8);
  fill<<<1024, 256>>>(A, big, 1); fill<<<1024, 256>>>(B, big, 2);
  double one = 1.0, zero = 0.0;
  struct Cfg { const char *name; cublasOperation_t ta, tb; int m, n, k; };
  std::vector<Cfg> cfgs = {
    {"C1 dgemm NN 1024^3", CUBLAS_OP_N, CUBLAS_OP_N, 1024, 1024, 1024},
    {"C2 dgemm TN 8192^3", CUBLAS_OP_T, CUBLAS_OP_N, 8192, 8192, 8192},
    {"C2' dgemm NN 8192^3", CUBLAS_OP_N, CUBLAS_OP_N, 8192, 8192, 8192},
    {"C2'' dgemm NT 8192^3", CUBLAS_OP_N, CUBLAS_OP_T, 8192, 8192, 8192},
    {"C2''' dgemm TT 8192^3", CUBLAS_OP_T, CUBLAS_OP_T, 8192, 8192, 8192},
    {"C3 dgemm NT 4097x3001x2049", CUBLAS_OP_N, CUBLAS_OP_T, 4097, 3001, 2049},
    {"dgemm NN 4096^3", CUBLAS_OP_N, CUBLAS_OP_N, 4096, 4096, 4096},
    {"dgemm NN 2048^3", CUBLAS_OP_N, CUBLAS_OP_N, 2048, 2048, 2048},
    {"C5/8 dgemm NN 8192x1024(n)x8192 (scaled column block)", CUBLAS_OP_N, CUBLAS_OP_N, 8192, 1024, 8192},
  };
  for (auto &c : cfgs) {
    int lda = c.ta == CUBLAS_OP_N ? c.m : c.k, ldb = c.tb == CUBLAS_OP_N ? c.k : c.n;
    float ms = best_ms([&] { cublasDgemm(h, c.ta, c.tb, c.m, c.n, c.k, &one, A, lda, B, ldb, &zero, C, c.m); }, reps);
    printf("%-58s %9.3f ms  %8.2f TFLOP/s\n", c.name, ms, 2.0 * c.m * c.n * c.k / ms * 1e-9);
  }
  // geam: transpose 8192^2, pad-copy C3 A / B
  {
    float ms = best_ms([&] { cublasDgeam(h, CUBLAS_OP_T, CUBLAS_OP_N, 8192, 8192, &one, A, 8192, &zero, C, 8192, C, 8192); }, reps);
    printf("%-58s %9.3f ms  %8.1f GB/s\n", "dgeam transpose 8192^2", ms, 2.0 * big * 8 / ms * 1e-6);
    ms = best_ms([&] { cublasDgeam(h, CUBLAS_OP_N, CUBLAS_OP_N, 8192, 8192, &one, A, 8192, &zero, C, 8192, C, 8192); }, reps);
    printf("%-58s %9.3f ms  %8.1f GB/s\n", "dgeam copy 8192^2", ms, 2.0 * big * 8 / ms * 1e-6);
    ms = best_ms([&] { cublasDgeam(h, CUBLAS_OP_N, CUBLAS_OP_N, 4097, 2049, &one, A, 4097, &zero, C, 4128, C, 4128); }, reps);
    printf("%-58s %9.3f ms  %8.1f GB/s\n", "dgeam pad-copy 4097x2049 ld4097->4128", ms, 2.0 * 4097 * 2049 * 8 / ms * 1e-6);
    ms = best_ms([&] { cublasDgeam(h, CUBLAS_OP_T, CUBLAS_OP_N, 2049, 4097, &one, A, 4097, &zero, C, 2080, C, 2080); }, reps);
    printf("%-58s %9.3f ms  %8.1f GB/s\n", "dgeam transpose 4097x2049 -> ld2080", ms, 2.0 * 4097 * 2049 * 8 / ms * 1e-6);
    cudaMemcpy(C, A, big * 8, cudaMemcpyDeviceToDevice);
    ms = best_ms([&] { cudaMemcpyAsync(C, A, big * 8, cudaMemcpyDeviceToDevice, 0); }, reps);
    printf("%-58s %9.3f ms  %8.1f GB/s\n", "cudaMemcpy D2D 512 MiB", ms, 2.0 * big * 8 / ms * 1e-6);
  }
  cudaFree(A); cudaFree(B); cudaFree(C);
  // SGEMM C4 (default math mode = true FP32) and with TF32 math for context
  {
    size_t m = 131072, n = 128, k = 4096;
    float *Af, *Bf, *Cf; cudaMalloc(&Af, m * k * 4); cudaMalloc(&Bf, k * n * 4); cudaMalloc(&Cf, m * n * 4);
    fillf<<<1024, 256>>>(Af, m * k, 3); fillf<<<64, 256>>>(Bf, k * n, 4);
    float onef = 1.f, zerof = 0.f;
    float ms = best_ms([&] { cublasSgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, m, n, k, &onef, Af, m, Bf, k, &zerof, Cf, m); }, reps);
    printf("%-58s %9.3f ms  %8.2f TFLOP/s\n", "C4 sgemm NN 131072x128x4096 (FP32 default math)", ms, 2.0 * m * n * k / ms * 1e-9);
    cublasSetMathMode(h, CUBLAS_TF32_TENSOR_OP_MATH);
    ms = best_ms([&] { cublasSgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, m, n, k, &onef, Af, m, Bf, k, &zerof, Cf, m); }, reps);
    printf("%-58s %9.3f ms  %8.2f TFLOP/s\n", "C4 sgemm NN (TF32 tensor-op math, context only)", ms, 2.0 * m * n * k / ms * 1e-9);
    cublasSetMathMode(h, CUBLAS_DEFAULT_MATH);
    // square sgemm for context
    ms = best_ms([&] { cublasSgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, 8192, 8192, 8192, &onef, Af, 8192, Af + 8192 * 8192, 8192, &zerof, Af + 2ull * 8192 * 8192, 8192); }, reps);
    printf("%-58s %9.3f ms  %8.2f TFLOP/s\n", "sgemm NN 8192^3 (FP32 default math)", ms, 2.0 * 8192.0 * 8192 * 8192 / ms * 1e-9);
  }
  return 0;
}
