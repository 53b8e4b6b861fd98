// tools/ubench.cu -- instruction-throughput probes used to size the SAD kernel (not part of the product).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/ubench tools/ubench.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE> __global__ void k(unsigned *out, int iters, unsigned seed) {
  unsigned a0 = threadIdx.x * 0x01010101u + seed, a1 = a0 ^ 0x55aa55aa, a2 = a0 + 0x01234567, a3 = a1 * 3;
  unsigned c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0, c6 = 0, c7 = 0;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++) {
      if (MODE == 0) {   // vabsdiff4.add, 8 independent chains
        asm volatile("vabsdiff4.u32.u32.u32.add %0,%1,%2,%0;" : "+r"(c0) : "r"(a0), "r"(a1));
        asm volatile("vabsdiff4.u32.u32.u32.add %0,%1,%2,%0;" : "+r"(c1) : "r"(a1), "r"(a2));
        asm volatile("vabsdiff4.u32.u32.u32.add %0,%1,%2,%0;" : "+r"(c2) : "r"(a2), "r"(a3));
        asm volatile("vabsdiff4.u32.u32.u32.add %0,%1,%2,%0;" : "+r"(c3) : "r"(a3), "r"(a0));
        asm volatile("vabsdiff4.u32.u32.u32.add %0,%1,%2,%0;" : "+r"(c4) : "r"(a0), "r"(a2));
        asm volatile("vabsdiff4.u32.u32.u32.add %0,%1,%2,%0;" : "+r"(c5) : "r"(a1), "r"(a3));
        asm volatile("vabsdiff4.u32.u32.u32.add %0,%1,%2,%0;" : "+r"(c6) : "r"(a2), "r"(a0));
        asm volatile("vabsdiff4.u32.u32.u32.add %0,%1,%2,%0;" : "+r"(c7) : "r"(a3), "r"(a1));
      } else if (MODE == 1) {  // prmt
        asm volatile("prmt.b32 %0,%0,%1,0x4321;" : "+r"(c0) : "r"(a0)); asm volatile("prmt.b32 %0,%0,%1,0x4321;" : "+r"(c1) : "r"(a1));
        asm volatile("prmt.b32 %0,%0,%1,0x4321;" : "+r"(c2) : "r"(a2)); asm volatile("prmt.b32 %0,%0,%1,0x4321;" : "+r"(c3) : "r"(a3));
        asm volatile("prmt.b32 %0,%0,%1,0x5432;" : "+r"(c4) : "r"(a0)); asm volatile("prmt.b32 %0,%0,%1,0x5432;" : "+r"(c5) : "r"(a1));
        asm volatile("prmt.b32 %0,%0,%1,0x5432;" : "+r"(c6) : "r"(a2)); asm volatile("prmt.b32 %0,%0,%1,0x5432;" : "+r"(c7) : "r"(a3));
      } else if (MODE == 2) {  // iadd3 reference
        c0 += a0 + c1; c1 += a1 + c2; c2 += a2 + c3; c3 += a3 + c4; c4 += a0 + c5; c5 += a1 + c6; c6 += a2 + c7; c7 += a3 + c0;
      } else if (MODE == 4) {  // dp4a
        asm volatile("dp4a.u32.u32 %0,%1,%2,%0;" : "+r"(c0) : "r"(a0), "r"(a1)); asm volatile("dp4a.u32.u32 %0,%1,%2,%0;" : "+r"(c1) : "r"(a1), "r"(a2));
        asm volatile("dp4a.u32.u32 %0,%1,%2,%0;" : "+r"(c2) : "r"(a2), "r"(a3)); asm volatile("dp4a.u32.u32 %0,%1,%2,%0;" : "+r"(c3) : "r"(a3), "r"(a0));
        asm volatile("dp4a.u32.u32 %0,%1,%2,%0;" : "+r"(c4) : "r"(a0), "r"(a2)); asm volatile("dp4a.u32.u32 %0,%1,%2,%0;" : "+r"(c5) : "r"(a1), "r"(a3));
        asm volatile("dp4a.u32.u32 %0,%1,%2,%0;" : "+r"(c6) : "r"(a2), "r"(a0)); asm volatile("dp4a.u32.u32 %0,%1,%2,%0;" : "+r"(c7) : "r"(a3), "r"(a1));
      } else if (MODE == 5) {  // funnel shift
        asm volatile("shf.r.wrap.b32 %0,%0,%1,8;" : "+r"(c0) : "r"(a0)); asm volatile("shf.r.wrap.b32 %0,%0,%1,8;" : "+r"(c1) : "r"(a1));
        asm volatile("shf.r.wrap.b32 %0,%0,%1,8;" : "+r"(c2) : "r"(a2)); asm volatile("shf.r.wrap.b32 %0,%0,%1,8;" : "+r"(c3) : "r"(a3));
        asm volatile("shf.r.wrap.b32 %0,%0,%1,16;" : "+r"(c4) : "r"(a0)); asm volatile("shf.r.wrap.b32 %0,%0,%1,16;" : "+r"(c5) : "r"(a1));
        asm volatile("shf.r.wrap.b32 %0,%0,%1,16;" : "+r"(c6) : "r"(a2)); asm volatile("shf.r.wrap.b32 %0,%0,%1,16;" : "+r"(c7) : "r"(a3));
      } else if (MODE == 6) {  // vabsdiff4 (no add) + imad accumulate via dp4a
        unsigned t0, t1, t2, t3;
        asm volatile("vabsdiff4.u32.u32.u32 %0,%1,%2,%3;" : "=r"(t0) : "r"(a0), "r"(c4), "r"(0u));
        asm volatile("dp4a.u32.u32 %0,%1,%2,%0;" : "+r"(c0) : "r"(t0), "r"(0x01010101u));
        asm volatile("vabsdiff4.u32.u32.u32 %0,%1,%2,%3;" : "=r"(t1) : "r"(a1), "r"(c5), "r"(0u));
        asm volatile("dp4a.u32.u32 %0,%1,%2,%0;" : "+r"(c1) : "r"(t1), "r"(0x01010101u));
        asm volatile("vabsdiff4.u32.u32.u32 %0,%1,%2,%3;" : "=r"(t2) : "r"(a2), "r"(c6), "r"(0u));
        asm volatile("dp4a.u32.u32 %0,%1,%2,%0;" : "+r"(c2) : "r"(t2), "r"(0x01010101u));
        asm volatile("vabsdiff4.u32.u32.u32 %0,%1,%2,%3;" : "=r"(t3) : "r"(a3), "r"(c7), "r"(0u));
        asm volatile("dp4a.u32.u32 %0,%1,%2,%0;" : "+r"(c3) : "r"(t3), "r"(0x01010101u));
        c4 ^= c0; c5 ^= c1; c6 ^= c2; c7 ^= c3;
      } else if (MODE == 7) {  // isetp with predicate accumulation + select
        bool p = (c0 < a0) | (c1 < a1) | (c2 < a2) | (c3 < a3) | (c4 < a0) | (c5 < a1) | (c6 < a2) | (c7 < a3);
        c0 += p; c1 ^= c0; c2 ^= c1; c3 ^= c2; c4 ^= c3; c5 ^= c4; c6 ^= c5; c7 ^= c6;
      } else {  // mix: 4 vabsdiff4 + 3 prmt (the SAD inner loop ratio)
        asm volatile("vabsdiff4.u32.u32.u32.add %0,%1,%2,%0;" : "+r"(c0) : "r"(a0), "r"(c4));
        asm volatile("prmt.b32 %0,%1,%2,0x4321;" : "=r"(c4) : "r"(a1), "r"(c0));
        asm volatile("vabsdiff4.u32.u32.u32.add %0,%1,%2,%0;" : "+r"(c1) : "r"(a1), "r"(c5));
        asm volatile("prmt.b32 %0,%1,%2,0x5432;" : "=r"(c5) : "r"(a2), "r"(c1));
        asm volatile("vabsdiff4.u32.u32.u32.add %0,%1,%2,%0;" : "+r"(c2) : "r"(a2), "r"(c6));
        asm volatile("prmt.b32 %0,%1,%2,0x6543;" : "=r"(c6) : "r"(a3), "r"(c2));
        asm volatile("vabsdiff4.u32.u32.u32.add %0,%1,%2,%0;" : "+r"(c3) : "r"(a3), "r"(c7));
        c7 = c3 ^ a0;
      }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = c0 + c1 + c2 + c3 + c4 + c5 + c6 + c7;
}
template <int MODE> void run(const char *name, int per_iter) {
  unsigned *d; cudaMalloc(&d, 148 * 8 * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int iters = 4096;
  k<MODE><<<148 * 8, 256>>>(d, 64, 1); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<MODE><<<148 * 8, 256>>>(d, iters, 2); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = (double)148 * 8 * 256 * iters * 16 * per_iter;
  printf("%-28s %8.3f ms  %8.1f G thread-instr/s  = %6.1f lanes/clk/SM @1.9GHz\n", name, ms, ops / ms / 1e6, ops / ms / 1e6 / 148 / 1.9);
  cudaFree(d);
}
int main() {
  run<0>("vabsdiff4.add", 8); run<1>("prmt", 8); run<2>("iadd3", 8); run<3>("4 vabsdiff4 + 3 prmt + 1 lop", 8);
  run<4>("dp4a", 8); run<5>("shf.r.wrap", 8); run<6>("4x(vabsdiff4 + dp4a) + 4 lop", 12); run<7>("8 isetp.or + 8 alu", 16);
  return 0;
}
