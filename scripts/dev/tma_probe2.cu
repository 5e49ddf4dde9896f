// canonical 2-D float32 TMA tile load (64x64 tensor, 32x32 box): does tensor TMA work on this box at all?
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <vector>
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void probe(const __grid_constant__ CUtensorMap map, float *out, int variant) {
  __shared__ __align__(1024) float buf[32 * 32];
  __shared__ __align__(8) unsigned long long mb;
  const unsigned mba = smem_u32(&mb);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mba), "r"(1) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mba), "r"(32 * 32 * 4) : "memory");
    if (variant == 0)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(buf)), "l"(&map),
                   "r"(mba), "r"(16), "r"(8)
                   : "memory");
    else
      asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(buf)), "l"(&map),
                   "r"(mba), "r"(16), "r"(8)
                   : "memory");
  }
  asm volatile("{\n.reg .pred P1;\nLW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra LD;\nbra LW;\nLD:\n}\n" ::"r"(mba), "r"(0) : "memory");
  for (int e = threadIdx.x; e < 1024; e += blockDim.x) out[e] = buf[e];
}
int main(int argc, char **argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  cudaFree(0);
  std::vector<float> h(64 * 64);
  for (int e = 0; e < 4096; e++) h[e] = (float)e;
  float *d, *out;
  cudaMalloc(&d, 4096 * 4);
  cudaMemcpy(d, h.data(), 4096 * 4, cudaMemcpyHostToDevice);
  cudaMalloc(&out, 1024 * 4);
  CUtensorMap m;
  cuuint64_t dims[2] = {64, 64}, strides[1] = {64 * 4};
  cuuint32_t box[2] = {32, 32}, es[2] = {1, 1};
  CUresult r = cuTensorMapEncodeTiled(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                      CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc=%d variant %d\n", (int)r, variant);
  probe<<<1, 128>>>(m, out, variant);
  cudaError_t e = cudaDeviceSynchronize();
  printf("run: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<float> o(1024);
  cudaMemcpy(o.data(), out, 4096, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int r2 = 0; r2 < 32; r2++)
    for (int c = 0; c < 32; c++)
      if (o[r2 * 32 + c] != (float)((8 + r2) * 64 + 16 + c)) bad++;
  printf("mismatches %d\n", bad);
  return 0;
}
