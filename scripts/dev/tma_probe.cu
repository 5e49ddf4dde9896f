// Stand-alone probe of the TMA usage of k_fct_march (FP64 tiled tensor maps, 3-D / 4-D boxes of 34 x nrows doubles, negative
// start coordinates, mbarrier completion).  Build + run on the GPU box:  nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/tma_probe scripts/dev/tma_probe.cu && /tmp/tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

struct Maps { CUtensorMap a, b; };

__global__ void probe(const __grid_constant__ Maps maps, const Maps *gmaps, const double *gptr, double *out, int c0, int c1, int c2, int c3, int nrows, int mode, int box0) {
  extern __shared__ __align__(128) unsigned char raw[];
  double *buf = reinterpret_cast<double *>(raw);
  unsigned long long *mb = reinterpret_cast<unsigned long long *>(raw + 2 * 34 * 32 * 8);
  const unsigned mba = smem_u32(mb);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mba), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (mode == 3) {
    if (threadIdx.x == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mba) : "memory");
  } else if (mode == 4) {
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mba), "r"(4096) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(buf)), "l"(gptr), "r"(4096), "r"(mba) : "memory");
    }
  } else
  if (threadIdx.x == 0) {
    const unsigned bytes = (mode == 2 ? 2u : 1u) * (unsigned)box0 * nrows * 8u;
    const CUtensorMap *ma = gmaps ? &gmaps->a : &maps.a, *mbp = gmaps ? &gmaps->b : &maps.b;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mba), "r"(bytes) : "memory");
    if (mode == 0 || mode == 2)
      asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(buf)),
                   "l"(ma), "r"(mba), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                   : "memory");
    if (mode == 1 || mode == 2)
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(buf + 34 * 32)),
                   "l"(mbp), "r"(mba), "r"(c0), "r"(c1), "r"(c2)
                   : "memory");
  }
  asm volatile(
      "{\n.reg .pred P1;\nLW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra LD;\nbra LW;\nLD:\n}\n" ::"r"(mba), "r"(0) : "memory");
  for (int e = threadIdx.x; e < 2 * 34 * 32; e += blockDim.x) out[e] = buf[e];
}

typedef CUresult (*Enc)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                        const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv) {
  const int amode = argc > 1 ? atoi(argv[1]) : 0, ac1 = argc > 2 ? atoi(argv[2]) : -1, box0 = argc > 3 ? atoi(argv[3]) : 34, dtype = argc > 4 ? atoi(argv[4]) : 0, gmem = argc > 5 ? atoi(argv[5]) : 0;
  const int imt = 102, km = 19, jl = 102, nt = 3, nrows = 23;
  void *p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) { printf("no entry point\n"); return 1; }
  Enc enc = (Enc)p;
  size_t n3 = (size_t)imt * km * jl;
  std::vector<double> h(n3 * nt);
  for (size_t e = 0; e < h.size(); e++) h[e] = (double)e;
  double *d, *out;
  cudaMalloc(&d, h.size() * 8);
  cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  cudaMalloc(&out, 2 * 34 * 32 * 8);
  Maps m;
  memset(&m, 0, sizeof m);
  const int es2 = (dtype == 2) ? 2 : 1;
  cuuint64_t dims[4] = {(cuuint64_t)imt * es2, (cuuint64_t)km, (cuuint64_t)jl, (cuuint64_t)nt};
  cuuint64_t strides[3] = {(cuuint64_t)imt * 8, (cuuint64_t)imt * km * 8, (cuuint64_t)n3 * 8};
  cuuint32_t box[4] = {(cuuint32_t)box0 * es2, (cuuint32_t)nrows, 1, 1}, es[4] = {1, 1, 1, 1};
  CUresult r4 = enc(&m.a, dtype == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : dtype ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CUresult r3 = enc(&m.b, dtype == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : dtype ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode 4d rc=%d 3d rc=%d\n", (int)r4, (int)r3);
  {
    CUtensorMap direct;
    memset(&direct, 0, sizeof direct);
    CUresult rd = cuTensorMapEncodeTiled(&direct, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("direct rc=%d same=%d\n", (int)rd, memcmp(&direct, &m.b, sizeof direct) == 0);
    const unsigned long long *w = (const unsigned long long *)&m.b;
    for (int q = 0; q < 16; q++) printf("%016llx%c", w[q], q % 4 == 3 ? '\n' : ' ');
    printf("device ptr %p\n", (void *)d);
  }
  const size_t shm = 2 * 34 * 32 * 8 + 64;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
  Maps *gm = nullptr;
  if (gmem) { cudaMalloc(&gm, sizeof(Maps)); cudaMemcpy(gm, &m, sizeof(Maps), cudaMemcpyHostToDevice); }
  for (int mode = amode; mode <= amode; mode++) {
    const int c0 = 29 * es2, c1 = ac1, c2 = 5, c3 = 1;
    printf("mode %d c1 %d box0 %d dtype %d gmem %d\n", mode, c1, box0, dtype, gmem);
    probe<<<1, 128, shm>>>(m, gm, d, out, c0, c1, c2, c3, nrows, mode, box0);
    cudaError_t e = cudaDeviceSynchronize();
    printf("mode %d: %s\n", mode, cudaGetErrorString(e));
    if (e != cudaSuccess) return 2;
    std::vector<double> o(2 * 34 * 32);
    cudaMemcpy(o.data(), out, o.size() * 8, cudaMemcpyDeviceToHost);
    // element (e, row p) <-> global (c0 + e, c1 + p, c2 [, c3])
    int bad = 0;
    if (mode == 4) { for (int e2 = 0; e2 < 512; e2++) if (o[e2] != (double)e2) bad++; printf("   bulk mismatches %d\n", bad); continue; }
    if (mode == 3) continue;
    for (int half = 0; half < 2; half++) {
      if ((mode == 0 && half == 1) || (mode == 1 && half == 0)) continue;
      for (int pr = 0; pr < nrows; pr++)
        for (int e2 = 0; e2 < box0; e2++) {
          const int kk = c1 + pr;
          double want = (kk < 0 || kk >= km) ? 0.0 : (double)((size_t)(c0 / es2 + e2) + (size_t)imt * (kk + (size_t)km * c2) + (half == 0 ? (size_t)c3 * n3 : 0));
          if (o[half * 34 * 32 + pr * box0 + e2] != want) bad++;
        }
    }
    printf("   mismatches %d\n", bad);
  }
  return 0;
}
