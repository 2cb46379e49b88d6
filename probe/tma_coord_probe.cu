// Does a 2-D tiled TMA load of 8-byte elements accept an inner start coordinate that is not a multiple of 16 bytes?
// X[i] = i (doubles); view: dims {rows, cols}, column stride 2*ld*8 bytes; box {16, 8}; load at (c0, 0) for c0 = 0..3,
// with SWIZZLE_NONE and SWIZZLE_128B.  Prints the first row of each landed box.  build: nvcc -arch=sm_100a -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

__global__ void probe(const __grid_constant__ CUtensorMap map, int c0, int c1, double *out) {
  __shared__ __align__(1024) double s[128];
  __shared__ __align__(8) unsigned long long bar;
  uint32_t sb = (uint32_t)__cvta_generic_to_shared(s), bb = (uint32_t)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(bb));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  for (int i = threadIdx.x; i < 128; i += 32) s[i] = -1.0;
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 1024;\n" ::"r"(bb) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(sb), "l"(&map),
                 "r"(c0), "r"(c1), "r"(bb)
                 : "memory");
  }
  asm volatile(
      "{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra.uni D;\nbra.uni W;\nD:\n}\n" ::"r"(bb)
      : "memory");
  for (int i = threadIdx.x; i < 128; i += 32) out[i] = s[i];
}

int main() {
  const int ld = 33, rows = 30, cols = 40;
  std::vector<double> h(ld * cols + 8);
  for (size_t i = 0; i < h.size(); i++) h[i] = (double)i;
  double *d, *out;
  cudaMalloc(&d, h.size() * 8);
  cudaMalloc(&out, 1024);
  cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  for (int swz = 0; swz < 2; swz++) {
    CUtensorMap map;
    cuuint64_t dims[2] = {(cuuint64_t)rows + 1, (cuuint64_t)cols / 2};
    cuuint64_t strides[1] = {2ull * ld * 8};
    cuuint32_t box[2] = {16, 8}, estr[2] = {1, 1};
    CUresult r = cuTensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        swz ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("swizzle %s encode -> %d\n", swz ? "128B" : "none", (int)r);
    for (int c0 = 0; c0 < 4; c0++) {
      probe<<<1, 32>>>(map, c0, 1, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("  c0=%d: %s\n", c0, cudaGetErrorString(e)); return 1; }
      double res[128];
      cudaMemcpy(res, out, 1024, cudaMemcpyDeviceToHost);
      printf("  c0=%d row0:", c0);
      for (int i = 0; i < 16; i++) printf(" %g", res[i]);
      printf(" | row1[0..3]: %g %g %g %g (expect first %d)\n", res[16], res[17], res[18], res[19], 2 * ld * 1 + c0);
    }
  }
  // rows near the end: coordinate 20 with 31 valid -> elements 20..30 then zero fill
  return 0;
}
