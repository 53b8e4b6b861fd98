// tools/tma_probe.cu -- isolates the TMA window load of k_int_search: u8 plane, 128x87 box, arbitrary (also negative) coordinates.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/tma_probe tools/tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <vector>
struct TMaps { CUtensorMap cur; CUtensorMap ref[16]; };
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int MODE>
__global__ void k(const __grid_constant__ TMaps tm, const int *idx, int x, int y, unsigned char *out) {
  __shared__ __align__(128) unsigned char win[87 * 128];
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ int sidx;
  if (threadIdx.x == 0) {
    sidx = idx[0];
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const CUtensorMap *m = MODE == 0 ? &tm.ref[0] : &tm.ref[sidx];
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(87 * 128) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(win)), "l"((unsigned long long)m), "r"(x), "r"(y), "r"(smem_u32(&mbar)) : "memory");
  }
  unsigned ok;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(&mbar)), "r"(0) : "memory");
  } while (!ok);
  for (int i = threadIdx.x; i < 87 * 128; i += blockDim.x) out[i] = win[i];
}
int main() {
  const int W = 1984, H = 1128, pitch = 2048;
  std::vector<unsigned char> h((size_t)pitch * H);
  for (int yy = 0; yy < H; yy++) for (int xx = 0; xx < pitch; xx++) h[(size_t)yy * pitch + xx] = (unsigned char)(xx * 7 + yy * 13);
  unsigned char *d, *o; int *di;
  cudaMalloc(&d, h.size() + 64); cudaMalloc(&o, 87 * 128); cudaMalloc(&di, 4); cudaMemset(di, 0, 4);
  cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
  void *fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  typedef CUresult (*enc_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                            CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  TMaps tm; memset(&tm, 0, sizeof(tm));
  cuuint64_t gd[2] = {W, H}, gs[1] = {pitch}; cuuint32_t box[2] = {128, 87}, es[2] = {1, 1};
  CUresult r = ((enc_t)fn)(&tm.ref[0], CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode -> %d (query %d)\n", (int)r, (int)q);
  std::vector<unsigned char> g(87 * 128);
  const int coords[][2] = {{0, 0}, {16, 8}, {58, 27}, {-12, -5}, {1900, 1100}, {3, 1}};
  for (int mode = 0; mode < 2; mode++)
    for (auto &c : coords) {
      if (mode == 0) k<0><<<1, 128>>>(tm, di, c[0], c[1], o); else k<1><<<1, 128>>>(tm, di, c[0], c[1], o);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d (%d,%d): %s\n", mode, c[0], c[1], cudaGetErrorString(e)); return 1; }
      cudaMemcpy(g.data(), o, g.size(), cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int rr = 0; rr < 87; rr++) for (int cc = 0; cc < 128; cc++) {
        int xx = c[0] + cc, yy = c[1] + rr;
        unsigned char want = (xx >= 0 && xx < W && yy >= 0 && yy < H) ? h[(size_t)yy * pitch + xx] : 0;
        bad += g[rr * 128 + cc] != want;
      }
      printf("mode %d (%d,%d): %d mismatches\n", mode, c[0], c[1], bad);
    }
  return 0;
}
