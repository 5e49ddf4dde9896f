// bisecting probe: rank-3 tensor (imt, km, jl), selectable element type / box / smem kind
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void probe(const __grid_constant__ CUtensorMap map, double *out, int c0, int c1, int c2, int bytes, int dyn) {
  __shared__ __align__(1024) double sbuf[34 * 24];
  extern __shared__ __align__(128) unsigned char raw[];
  __shared__ __align__(8) unsigned long long mb;
  double *buf = dyn ? reinterpret_cast<double *>(raw) : sbuf;
  const unsigned mba = smem_u32(&mb);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mba), "r"(1) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mba), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(buf)), "l"(&map),
                 "r"(mba), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
  }
  asm volatile("{\n.reg .pred P1;\nLW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra LD;\nbra LW;\nLD:\n}\n" ::"r"(mba), "r"(0) : "memory");
  for (int e = threadIdx.x; e < 34 * 24; e += blockDim.x) out[e] = buf[e];
}
int main(int argc, char **argv) {
  // args: dtype(0 f64, 1 f32-pairs) box0 box1 dyn c1
  const int dtype = atoi(argv[1]), box0 = atoi(argv[2]), box1 = atoi(argv[3]), dyn = atoi(argv[4]), c1 = atoi(argv[5]);
  const int imt = 102, km = 19, jl = 102;
  cudaFree(0);
  size_t n3 = (size_t)imt * km * jl;
  std::vector<double> h(n3);
  for (size_t e = 0; e < n3; e++) h[e] = (double)e;
  double *d, *out;
  cudaMalloc(&d, n3 * 8);
  cudaMemcpy(d, h.data(), n3 * 8, cudaMemcpyHostToDevice);
  cudaMalloc(&out, 34 * 24 * 8);
  const int f = dtype ? 2 : 1;
  CUtensorMap m;
  cuuint64_t dims[3] = {(cuuint64_t)imt * f, (cuuint64_t)km, (cuuint64_t)jl}, strides[2] = {(cuuint64_t)imt * 8, (cuuint64_t)imt * km * 8};
  cuuint32_t box[3] = {(cuuint32_t)box0 * f, (cuuint32_t)box1, 1}, es[3] = {1, 1, 1};
  CUresult r = cuTensorMapEncodeTiled(&m, dtype ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, d, dims, strides, box, es,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("dtype %d box %dx%d dyn %d c1 %d: encode rc=%d ", dtype, box0, box1, dyn, c1, (int)r);
  const int c0 = argc > 6 ? atoi(argv[6]) : 30, c2 = 5;
  if (dyn) cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 34 * 24 * 8);
  probe<<<1, 128, dyn ? 34 * 24 * 8 : 0>>>(m, out, c0 * f, c1, c2, box0 * box1 * 8, dyn);
  cudaError_t e = cudaDeviceSynchronize();
  printf("run: %s ", cudaGetErrorString(e));
  if (e != cudaSuccess) { printf("\n"); return 1; }
  std::vector<double> o(34 * 24);
  cudaMemcpy(o.data(), out, o.size() * 8, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int p = 0; p < box1; p++)
    for (int q = 0; q < box0; q++) {
      const int kk = c1 + p;
      const double want = (kk < 0 || kk >= km) ? 0.0 : (double)((size_t)(c0 + q) + (size_t)imt * (kk + (size_t)km * c2));
      if (o[p * box0 + q] != want) bad++;
    }
  printf("mismatches %d\n", bad);
  return 0;
}
