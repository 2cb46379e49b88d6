// Probe: FP64 instruction issue rates on sm_100a.
// Decides whether the DGEMM inner loop is built on DMMA (mma.sync f64) or DFMA.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_rates fp64_rates.cu
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1684(double (&c)[4], const double (&a)[2], double b) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// MODE 0: DFMA, 32 independent chains per thread.  flops/thread/iter = 32*2
template <int NACC>
__global__ void k_dfma(double *out, int iters, double x, double y) {
  double acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) acc[i] = i + threadIdx.x;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) acc[i] = fma(acc[i], x, y);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_884(double *out, int iters, double x, double y) {
  double c0[NACC], c1[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) { c0[i] = i; c1[i] = threadIdx.x; }
  double a = x + threadIdx.x, b = y - threadIdx.x;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) dmma884(c0[i], c1[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_1684(double *out, int iters, double x, double y) {
  double c[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; i++) for (int j = 0; j < 4; j++) c[i][j] = i + j;
  double a[2] = {x + threadIdx.x, x - threadIdx.x}; double b = y - threadIdx.x;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) dmma1684(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) for (int j = 0; j < 4; j++) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_1688(double *out, int iters, double x, double y) {
  double c[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; i++) for (int j = 0; j < 4; j++) c[i][j] = i + j;
  double a[4] = {x + threadIdx.x, x - threadIdx.x, x, x * 2}; double b[2] = {y - threadIdx.x, y};
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) dmma1688(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) for (int j = 0; j < 4; j++) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_16816(double *out, int iters, double x, double y) {
  double c[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; i++) for (int j = 0; j < 4; j++) c[i][j] = i + j;
  double a[8], b[4];
#pragma unroll
  for (int j = 0; j < 8; j++) a[j] = x + j * threadIdx.x;
#pragma unroll
  for (int j = 0; j < 4; j++) b[j] = y - j * threadIdx.x;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) dmma16816(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) for (int j = 0; j < 4; j++) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  printf("device %s sms %d clock %d kHz\n", p.name, sms, p.clockRate);
  double *out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
  const int iters = 4096;
  for (int warps : {4, 8, 16, 32}) {
    for (int ctas : {1, 2}) {
      if (warps * ctas > 32) continue;
      dim3 g(sms * ctas), b(warps * 32);
      double nw = (double)sms * ctas * warps;
      float ms;
      ms = time_ms([&] { k_dfma<32><<<g, b>>>(out, iters, 1.0000001, 0.5); });
      printf("warps/cta %2d ctas/sm %d  DFMA x32        : %8.2f TFLOP/s\n", warps, ctas, nw * 32 * 32 * 2.0 * iters / ms * 1e-9);
      ms = time_ms([&] { k_884<8><<<g, b>>>(out, iters, 1.0000001, 0.5); });
      printf("warps/cta %2d ctas/sm %d  DMMA m8n8k4 x8  : %8.2f TFLOP/s\n", warps, ctas, nw * 8 * 512.0 * iters / ms * 1e-9);
      ms = time_ms([&] { k_884<16><<<g, b>>>(out, iters, 1.0000001, 0.5); });
      printf("warps/cta %2d ctas/sm %d  DMMA m8n8k4 x16 : %8.2f TFLOP/s\n", warps, ctas, nw * 16 * 512.0 * iters / ms * 1e-9);
      ms = time_ms([&] { k_1684<8><<<g, b>>>(out, iters, 1.0000001, 0.5); });
      printf("warps/cta %2d ctas/sm %d  DMMA m16n8k4 x8 : %8.2f TFLOP/s\n", warps, ctas, nw * 8 * 1024.0 * iters / ms * 1e-9);
      ms = time_ms([&] { k_1688<8><<<g, b>>>(out, iters, 1.0000001, 0.5); });
      printf("warps/cta %2d ctas/sm %d  DMMA m16n8k8 x8 : %8.2f TFLOP/s\n", warps, ctas, nw * 8 * 2048.0 * iters / ms * 1e-9);
      ms = time_ms([&] { k_16816<8><<<g, b>>>(out, iters, 1.0000001, 0.5); });
      printf("warps/cta %2d ctas/sm %d  DMMA m16n8k16 x8: %8.2f TFLOP/s\n", warps, ctas, nw * 8 * 4096.0 * iters / ms * 1e-9);
      ms = time_ms([&] { k_16816<16><<<g, b>>>(out, iters, 1.0000001, 0.5); });
      printf("warps/cta %2d ctas/sm %d  DMMA m16n8k16x16: %8.2f TFLOP/s\n", warps, ctas, nw * 16 * 4096.0 * iters / ms * 1e-9);
    }
  }
  // dependent-chain latency (1 warp, 1 accumulator)
  {
    float ms = time_ms([&] { k_884<1><<<1, 32>>>(out, iters * 16, 1.0000001, 0.5); });
    printf("DMMA m8n8k4 dependent chain: %.1f ns/op\n", ms * 1e6 / (iters * 16));
    ms = time_ms([&] { k_16816<1><<<1, 32>>>(out, iters * 16, 1.0000001, 0.5); });
    printf("DMMA m16n8k16 dependent chain: %.1f ns/op\n", ms * 1e6 / (iters * 16));
    ms = time_ms([&] { k_dfma<1><<<1, 32>>>(out, iters * 16, 1.0000001, 0.5); });
    printf("DFMA dependent chain: %.1f ns/op\n", ms * 1e6 / (iters * 16));
  }
  return 0;
}
