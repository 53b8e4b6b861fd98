// k_tq.cu -- K7 + K8: H.264 forward integer transforms and quantisation, and the fused
// motion-compensation -> residual -> transform -> quantisation pass over a whole picture.
//
//  forward4x4 / forward8x8          lcommon/src/transform.c:20-68, :353-448
//  quant_4x4_normal / _around       lencod/src/quant4x4_normal.c:39-115, quant4x4_around.c:40-130
//  quant_8x8_normal / _around       lencod/src/quant8x8_normal.c:43-107, quant8x8_around.c:40-115
//  quant_8x8cavlc_normal / _around  lencod/src/quant8x8_normal.c:123-202, quant8x8_around.c:133-223
//  luma_prediction                  lencod/src/mc_prediction.c:117-236 (copy out of the quarter-pel
//                                   plane the mv selects, UMVLine4X origin clamp per predicted block)
//  luma_residual_coding(_8x8)       lencod/src/macroblock.c:919-1022, :1182-1257 (one 16x16 prediction unit for
//                                   mode 1, 8x8 units for modes 2..4, 4x4 units for modes 5..7; per-quadrant coeff_cost)
//
// All arithmetic is int32 exactly as in JM.  One thread owns one transform block; the work is a
// pure stream (residual in, levels out), i.e. HBM-bound.
#include "jmb_internal.h"

namespace {

__device__ __forceinline__ void fwd4(int *b) {
  int t[16];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    int *p = b + 4 * i;
    int t0 = p[0] + p[3], t1 = p[1] + p[2], t2 = p[1] - p[2], t3 = p[0] - p[3];
    t[4 * i] = t0 + t1; t[4 * i + 1] = (t3 << 1) + t2; t[4 * i + 2] = t0 - t1; t[4 * i + 3] = t3 - (t2 << 1);
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    int t0 = t[i] + t[12 + i], t1 = t[4 + i] + t[8 + i], t2 = t[4 + i] - t[8 + i], t3 = t[i] - t[12 + i];
    b[i] = t0 + t1; b[4 + i] = t2 + (t3 << 1); b[8 + i] = t0 - t1; b[12 + i] = t3 - (t2 << 1);
  }
}

__device__ __forceinline__ void fwd8_1d(const int *p, int s, int *o, int os) {
  int a0 = p[0] + p[7 * s], a1 = p[s] + p[6 * s], a2 = p[2 * s] + p[5 * s], a3 = p[3 * s] + p[4 * s];
  int b0 = a0 + a3, b1 = a1 + a2, b2 = a0 - a3, b3 = a1 - a2;
  a0 = p[0] - p[7 * s]; a1 = p[s] - p[6 * s]; a2 = p[2 * s] - p[5 * s]; a3 = p[3 * s] - p[4 * s];
  int b4 = a1 + a2 + ((a0 >> 1) + a0), b5 = a0 - a3 - ((a2 >> 1) + a2);
  int b6 = a0 + a3 - ((a1 >> 1) + a1), b7 = a1 - a2 + ((a3 >> 1) + a3);
  o[0] = b0 + b1; o[os] = b4 + (b7 >> 2); o[2 * os] = b2 + (b3 >> 1); o[3 * os] = b5 + (b6 >> 2);
  o[4 * os] = b0 - b1; o[5 * os] = b6 - (b5 >> 2); o[6 * os] = (b2 >> 1) - b3; o[7 * os] = (b4 >> 2) - b7;
}

__device__ void fwd8(int *b) {
  int t[64];
#pragma unroll
  for (int i = 0; i < 8; i++) fwd8_1d(b + 8 * i, 1, t + 8 * i, 1);
#pragma unroll
  for (int i = 0; i < 8; i++) fwd8_1d(t + i, 8, b + i, 8);
}

// inverse4x4 / inverse8x8, lcommon/src/transform.c:70-119, :450-547
__device__ __forceinline__ void inv4(int *b) {
  int t[16];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int *p = b + 4 * i;
    int p0 = p[0] + p[2], p1 = p[0] - p[2], p2 = (p[1] >> 1) - p[3], p3 = p[1] + (p[3] >> 1);
    t[4 * i] = p0 + p3; t[4 * i + 1] = p1 + p2; t[4 * i + 2] = p1 - p2; t[4 * i + 3] = p0 - p3;
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    int p0 = t[i] + t[8 + i], p1 = t[i] - t[8 + i], p2 = (t[4 + i] >> 1) - t[12 + i], p3 = t[4 + i] + (t[12 + i] >> 1);
    b[i] = p0 + p3; b[4 + i] = p1 + p2; b[8 + i] = p1 - p2; b[12 + i] = p0 - p3;
  }
}
__device__ __forceinline__ void inv8_1d(const int *p, int s, int *o, int os) {
  int a0 = p[0] + p[4 * s], a1 = p[0] - p[4 * s], a2 = p[6 * s] - (p[2 * s] >> 1), a3 = p[2 * s] + (p[6 * s] >> 1);
  int b0 = a0 + a3, b2 = a1 - a2, b4 = a1 + a2, b6 = a0 - a3;
  a0 = -p[3 * s] + p[5 * s] - p[7 * s] - (p[7 * s] >> 1);
  a1 = p[s] + p[7 * s] - p[3 * s] - (p[3 * s] >> 1);
  a2 = -p[s] + p[7 * s] + p[5 * s] + (p[5 * s] >> 1);
  a3 = p[3 * s] + p[5 * s] + p[s] + (p[s] >> 1);
  int b1 = a0 + (a3 >> 2), b3 = a1 + (a2 >> 2), b5 = a2 - (a1 >> 2), b7 = a3 - (a0 >> 2);
  o[0] = b0 + b7; o[os] = b2 - b5; o[2 * os] = b4 + b3; o[3 * os] = b6 + b1;
  o[4 * os] = b6 - b1; o[5 * os] = b4 - b3; o[6 * os] = b2 + b5; o[7 * os] = b0 - b7;
}
__device__ void inv8(int *b) {
  int t[64];
#pragma unroll
  for (int i = 0; i < 8; i++) inv8_1d(b + 8 * i, 1, t + 8 * i, 1);
#pragma unroll
  for (int i = 0; i < 8; i++) inv8_1d(t + i, 8, b + i, 8);
}

// H.264's frame zig-zag scans as {i (horizontal), j (vertical)} (the standard's Figure 8-8 / 8-9; JM: block.c:170,
// transform8x8.c:44,55).  When the caller's scan is one of these the quantiser runs with compile-time indices and the
// coefficient block stays in registers; any other scan (field scans) takes the table-driven path.
__device__ constexpr unsigned char STD_SCAN4[16][2] = {{0,0}, {1,0}, {0,1}, {0,2}, {1,1}, {2,0}, {3,0}, {2,1}, {1,2}, {0,3}, {1,3}, {2,2}, {3,1}, {3,2}, {2,3}, {3,3}};
__device__ constexpr unsigned char STD_SCAN8[64][2] = {{0,0}, {1,0}, {0,1}, {0,2}, {1,1}, {2,0}, {3,0}, {2,1}, {1,2}, {0,3}, {0,4}, {1,3}, {2,2}, {3,1}, {4,0}, {5,0}, {4,1}, {3,2}, {2,3}, {1,4}, {0,5}, {0,6}, {1,5}, {2,4}, {3,3}, {4,2}, {5,1}, {6,0}, {7,0}, {6,1}, {5,2}, {4,3}, {3,4}, {2,5}, {1,6}, {0,7}, {1,7}, {2,6}, {3,5}, {4,4}, {5,3}, {6,2}, {7,1}, {7,2}, {6,3}, {5,4}, {4,5}, {3,6}, {2,7}, {3,7}, {4,6}, {5,5}, {6,4}, {7,3}, {7,4}, {6,5}, {5,6}, {4,7}, {5,7}, {6,6}, {7,5}, {7,6}, {6,7}, {7,7}};
__device__ constexpr unsigned char STD_SCAN8_CAVLC[64][2] = {{0,0}, {1,1}, {1,2}, {2,2}, {4,1}, {0,5}, {3,3}, {7,0}, {3,4}, {1,7}, {5,3}, {6,3}, {2,7}, {6,4}, {5,6}, {7,5}, {1,0}, {2,0}, {0,3}, {3,1}, {3,2}, {0,6}, {4,2}, {6,1}, {2,5}, {2,6}, {6,2}, {5,4}, {3,7}, {7,3}, {4,7}, {7,6}, {0,1}, {3,0}, {0,4}, {4,0}, {2,3}, {1,5}, {5,1}, {5,2}, {1,6}, {3,5}, {7,1}, {4,5}, {4,6}, {7,4}, {5,7}, {6,7}, {0,2}, {2,1}, {1,3}, {5,0}, {1,4}, {2,4}, {6,0}, {4,3}, {0,7}, {4,4}, {7,2}, {3,6}, {5,5}, {6,5}, {6,6}, {7,7}};

// four consecutive samples at an arbitrary byte address: two aligned word loads + PRMT
__device__ __forceinline__ unsigned tq_ld4(const uint8_t *p) {
  const unsigned sh = (unsigned)(size_t)p & 3u;
  const unsigned *a = (const unsigned *)(p - sh);
  return __byte_perm(__ldg(a), __ldg(a + 1), 0x3210u + 0x1111u * sh);
}

struct QOut { int nonzero; int cost; };

// partition mode / reference of one quadrant of a jmb_mb_pred; 0 = fine (JMB_REQERR_* otherwise)
__device__ __forceinline__ int pred_check(int mode, int ref, int n, int nref) {
  int e = 0;
  if (mode < 1 || mode > 7 || (n == 8 && mode > 4)) e |= JMB_REQERR_BLOCKTYPE;
  if (ref >= nref) e |= JMB_REQERR_REF;
  return e;
}

// quantise one block in scan order.  coef: in = transformed, out = dequantised (JM leaves it in tblock).
// LISTS: write JM's (level, run) lists; otherwise write the level of every scan position to lv16.
// STD: 0 = scan from the descriptor (any order), 1 = the standard zig-zag, 2 = the standard 8x8 CAVLC interleave
template <int N, bool LISTS, int STD = 0>
__device__ __forceinline__ QOut quant_block(const jmb_quant_desc &q, int *coef, int *levels, int *runs, int *fadj, int16_t *lv16) {
  constexpr int NN = N * N;
  const int qp_per = q.qp / 6, q_bits = (N == 4 ? 15 : 16) + qp_per, dq = (N == 4) ? 4 : 6;
  const bool cavlc8 = (N == 8) && q.is_cavlc;
  const bool clip = (N == 4) ? (q.is_cavlc != 0) : cavlc8;
  int nl[4] = {0, 0, 0, 0}, run[4] = {0, 0, 0, 0};
  QOut o{0, 0};
#pragma unroll
  for (int k = 0; k < (STD ? NN : 0); k++) {       // compile-time scan: every index below is a constant after unrolling
    const int i = (N == 4) ? STD_SCAN4[k][0] : (STD == 2 ? STD_SCAN8_CAVLC[k][0] : STD_SCAN8[k][0]);
    const int j = (N == 4) ? STD_SCAN4[k][1] : (STD == 2 ? STD_SCAN8_CAVLC[k][1] : STD_SCAN8[k][1]);
    const int idx = j * N + i, s = (STD == 2) ? (k >> 4) : 0;
    const int c = coef[idx];
    int level = 0;
    if (c != 0) {
      const int scaled = abs(c) * q.qparams[idx][1];
      level = (scaled + q.qparams[idx][0]) >> q_bits;
      if (level != 0) {
        if (clip) level = min(level, 2063);
        o.cost += (level > 1) ? 999999 : q.c_cost[run[s]];
        if (c < 0) level = -level;
        coef[idx] = (((level * q.qparams[idx][2]) << qp_per) + (1 << (dq - 1))) >> dq;
        o.nonzero = 1;
      } else coef[idx] = 0;
    }
    lv16[k] = (int16_t)level;
    if (level != 0) run[s] = 0; else run[s]++;
  }
  if (STD) return o;
  for (int k = 0; k < NN; k++) {
    const int i = q.scan[k][0], j = q.scan[k][1], idx = j * N + i;
    const int s = cavlc8 ? (k >> 4) : 0;
    const int c = coef[idx];
    int level = 0, adj = 0;
    if (c != 0) {
      const int scaled = abs(c) * q.qparams[idx][1];
      level = (scaled + q.qparams[idx][0]) >> q_bits;
      if (level != 0) {
        if (clip) level = min(level, 2063);                      // CAVLC_LEVEL_LIMIT
        adj = (q.adapt_rnd_weight * (scaled - (level << q_bits)) + (1 << q_bits)) >> (q_bits + 1);
        o.cost += (level > 1) ? 999999 : q.c_cost[run[s]];        // MAX_VALUE
        if (c < 0) level = -level;
        coef[idx] = (((level * q.qparams[idx][2]) << qp_per) + (1 << (dq - 1))) >> dq;
        o.nonzero = 1;
      } else coef[idx] = 0;
    }
    if (fadj) fadj[idx] = adj;
    if (LISTS) {
      if (level != 0) {
        const int base = cavlc8 ? 17 * s : 0;
        levels[base + nl[s]] = level; runs[base + nl[s]] = run[s];
        nl[s]++; run[s] = 0;
      } else run[s]++;
    } else {
      lv16[k] = (int16_t)level;
      if (level != 0) run[s] = 0; else run[s]++;
    }
  }
  if (LISTS) {
    if (cavlc8) { for (int s = 0; s < 4; s++) levels[17 * s + nl[s]] = 0; }
    else levels[nl[0]] = 0;
  }
  return o;
}

template <int N>
__global__ void k_forward(int *blocks, int nblk) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nblk) return;
  int b[N * N];
  int *g = blocks + (size_t)i * N * N;
#pragma unroll
  for (int k = 0; k < N * N; k++) b[k] = g[k];
  if (N == 4) fwd4(b); else fwd8(b);
#pragma unroll
  for (int k = 0; k < N * N; k++) g[k] = b[k];
}

template <int N>
__global__ void k_quant_blocks(const jmb_quant_desc *__restrict__ qd, int do_transform, int *coef, int nblk, int lr_stride,
                               int *levels, int *runs, int *fadjust, int *coeff_cost, int *nonzero) {
  __shared__ jmb_quant_desc q;
  for (int i = threadIdx.x; i < (int)(sizeof(q) / 4); i += blockDim.x) ((int *)&q)[i] = ((const int *)qd)[i];
  __syncthreads();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nblk) return;
  int b[N * N];
  int *g = coef + (size_t)i * N * N;
  for (int k = 0; k < N * N; k++) b[k] = g[k];
  if (do_transform) { if (N == 4) fwd4(b); else fwd8(b); }
  QOut o = quant_block<N, true>(q, b, levels + (size_t)i * lr_stride, runs + (size_t)i * lr_stride,
                                (q.around && fadjust) ? fadjust + (size_t)i * N * N : nullptr, nullptr);
  for (int k = 0; k < N * N; k++) g[k] = b[k];
  coeff_cost[i] += o.cost;
  nonzero[i] = o.nonzero;
}

// one thread per transform block of the picture
template <int N>
__global__ void k_mc_tq(const jmb_mb_pred *__restrict__ pred, int n_mb, int mb_w, const jmb_quant_desc *__restrict__ qd,
                        const uint8_t *__restrict__ cur, int cur_pitch, const uint8_t *const *__restrict__ ref_planes,
                        size_t plane_bytes, int ref_pitch, int w, int h, int nref, int *__restrict__ err,
                        int16_t *__restrict__ levels, int *__restrict__ coeff_cost, unsigned *__restrict__ cbp_blk) {
  __shared__ jmb_quant_desc q;
  for (int i = threadIdx.x; i < (int)(sizeof(q) / 4); i += blockDim.x) ((int *)&q)[i] = ((const int *)qd)[i];
  __syncthreads();
  constexpr int PER_MB = (N == 4) ? 16 : 4;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_mb * PER_MB) return;
  const int mb = t / PER_MB, b = t - mb * PER_MB;
  const int mbx = (mb % mb_w) * 16, mby = (mb / mb_w) * 16;
  const int bx4 = (N == 4) ? (b & 3) : (b & 1) * 2, by4 = (N == 4) ? (b >> 2) : (b >> 1) * 2;   // 4x4 units in the MB
  const int b8 = (by4 >> 1) * 2 + (bx4 >> 1);
  const jmb_mb_pred *mp = pred + mb;
  const int mode = mp->b8mode[b8];
  {   // the table may live on the device: validated here, like the search requests (jmb_req_check)
    const int bad = pred_check(mode, mp->ref[b8], N, nref);
    if (bad) { jmb_req_report(err, bad, mb); return; }
  }
  // prediction unit: the 8x8 quadrant for modes 1..4, the 4x4 block for modes 5..7 (macroblock.c:946-971)
  int ux4 = bx4, uy4 = by4;
  if (mode < 5 || N == 8) { ux4 &= ~1; uy4 &= ~1; }
  if (mode == 1) { ux4 = 0; uy4 = 0; }                       // P16x16: one 16x16 prediction, one origin clamp (macroblock.c:1225)
  const int mvx = mp->mv[uy4 * 4 + ux4][0], mvy = mp->mv[uy4 * 4 + ux4][1];
  const int qx = ((mbx + ux4 * 4) << 2) + mvx, qy = ((mby + uy4 * 4) << 2) + mvy;
  const int iy = jmb_clip(-JMB_PAD_Y, h + JMB_PAD_Y - 1 - 16, qy >> 2), ix = jmb_clip(-JMB_PAD_X, w + JMB_PAD_X - 1 - 16, qx >> 2);
  const uint8_t *rp = ref_planes[mp->ref[b8]] + (size_t)((qy & 3) * 4 + (qx & 3)) * plane_bytes +
                      (size_t)(iy + JMB_PAD_Y + (by4 - uy4) * 4) * ref_pitch + (ix + JMB_PAD_X + (bx4 - ux4) * 4);
  const uint8_t *sp = cur + (size_t)(mby + by4 * 4) * cur_pitch + mbx + bx4 * 4;
  int r[N * N];
#pragma unroll
  for (int y = 0; y < N; y++)
#pragma unroll
    for (int x = 0; x < N; x++) r[y * N + x] = (int)sp[(size_t)y * cur_pitch + x] - (int)rp[(size_t)y * ref_pitch + x];
  if (N == 4) fwd4(r); else fwd8(r);
  QOut o = quant_block<N, false>(q, r, nullptr, nullptr, nullptr, levels + (size_t)mb * 256 + b * N * N);
  if (o.cost) atomicAdd(&coeff_cost[mb * 4 + b8], o.cost);
  if (o.nonzero) {
    unsigned bits = (N == 4) ? (1u << (by4 * 4 + bx4)) : (51u << (4 * b8 - 2 * (b8 & 1)));   // macroblock.c:1004
    atomicOr(&cbp_blk[mb], bits);
  }
}

// Every inter partition mode of every macroblock in ONE launch: the prediction comes straight from the 41 search
// results of the macroblock (the all_mv fill of mv_search.c:1005-1014 folded in), reference 0.  Thread = one transform
// block of one (mode, macroblock); outputs are mode-major.
// first request of partition mode m in a macroblock's 41 results, and the mode's partition size in 4x4 units
// (constant memory: as local arrays indexed by the mode they lived on the stack, 24 stores + 3 loads per thread)
__constant__ int c_mode_base[8] = {0, 0, 1, 3, 5, 9, 17, 25}, c_mode_w4[8] = {4, 4, 4, 2, 2, 2, 1, 1}, c_mode_h4[8] = {4, 4, 2, 4, 2, 1, 2, 1},
               c_mode_lw4[8] = {2, 2, 2, 1, 1, 1, 0, 0}, c_mode_lh4[8] = {2, 2, 1, 2, 1, 0, 1, 0};      // log2 of the two above

template <int N, int STD>
__global__ void k_mc_tq_modes(const jmb_me_res *__restrict__ res, int n_mb, int mb_w, unsigned mode_mask, const jmb_quant_desc *__restrict__ qd,
                              const uint8_t *__restrict__ cur, int cur_pitch, const uint8_t *__restrict__ ref_plane0,
                              size_t plane_bytes, int ref_pitch, int w, int h,
                              int16_t *__restrict__ levels, int *__restrict__ coeff_cost, unsigned *__restrict__ cbp_blk) {
  __shared__ jmb_quant_desc q;
  for (int i = threadIdx.x; i < (int)(sizeof(q) / 4); i += blockDim.x) ((int *)&q)[i] = ((const int *)qd)[i];
  __syncthreads();
  constexpr int PER_MB = (N == 4) ? 16 : 4;
  const int mode = blockIdx.y + 1;
  if (!((mode_mask >> blockIdx.y) & 1)) return;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_mb * PER_MB) return;
  const int mb = t / PER_MB, b = t - mb * PER_MB;
  const int mbx = (mb % mb_w) * 16, mby = (mb / mb_w) * 16;
  const int bx4 = (N == 4) ? (b & 3) : (b & 1) * 2, by4 = (N == 4) ? (b >> 2) : (b >> 1) * 2;
  const int b8 = (by4 >> 1) * 2 + (bx4 >> 1);
  int ux4 = bx4, uy4 = by4;                                  // prediction unit (macroblock.c:946-971)
  if (mode < 5 || N == 8) { ux4 &= ~1; uy4 &= ~1; }
  if (mode == 1) { ux4 = 0; uy4 = 0; }                       // P16x16: one 16x16 prediction, one origin clamp (macroblock.c:1225)
  const jmb_me_res r = res[mb * 41 + c_mode_base[mode] + (uy4 / c_mode_h4[mode]) * (4 / c_mode_w4[mode]) + ux4 / c_mode_w4[mode]];
  const int qx = ((mbx + ux4 * 4) << 2) + r.mv_x, qy = ((mby + uy4 * 4) << 2) + r.mv_y;
  const int iy = jmb_clip(-JMB_PAD_Y, h + JMB_PAD_Y - 1 - 16, qy >> 2), ix = jmb_clip(-JMB_PAD_X, w + JMB_PAD_X - 1 - 16, qx >> 2);
  const uint8_t *rp = ref_plane0 + (size_t)((qy & 3) * 4 + (qx & 3)) * plane_bytes +
                      (size_t)(iy + JMB_PAD_Y + (by4 - uy4) * 4) * ref_pitch + (ix + JMB_PAD_X + (bx4 - ux4) * 4);
  const uint8_t *sp = cur + (size_t)(mby + by4 * 4) * cur_pitch + mbx + bx4 * 4;
  int rr[N * N];
#pragma unroll
  for (int y = 0; y < N; y++)
#pragma unroll
    for (int x4 = 0; x4 < N; x4 += 4) {       // word loads: the source is 4-aligned, the prediction is not
      const unsigned sv = *(const unsigned *)(sp + (size_t)y * cur_pitch + x4), pv = tq_ld4(rp + (size_t)y * ref_pitch + x4);
#pragma unroll
      for (int x = 0; x < 4; x++) rr[y * N + x4 + x] = (int)((sv >> (8 * x)) & 255) - (int)((pv >> (8 * x)) & 255);
    }
  if (N == 4) fwd4(rr); else fwd8(rr);
  const size_t mo = (size_t)blockIdx.y * n_mb + mb;          // mode-major output index
  QOut o = quant_block<N, false, STD>(q, rr, nullptr, nullptr, nullptr, levels + mo * 256 + b * N * N);
  if (o.cost) atomicAdd(&coeff_cost[mo * 4 + b8], o.cost);
  if (o.nonzero) atomicOr(&cbp_blk[mo], (N == 4) ? (1u << (by4 * 4 + bx4)) : (51u << (4 * b8 - 2 * (b8 & 1))));
}


// jmb_mc_tq_modes_compact: the kernel above with JM's own output shape -- (level, run) tokens for the nonzero levels only and
// one 16-byte head per (mode, macroblock) -- and the quantiser description in the CONSTANT bank (a __grid_constant__
// parameter: with the compile-time scan every Scale / Offset is an immediate constant operand of the multiply-add, no load at
// all; the dense kernel spends most of its issue slots waiting for ~50 broadcast LDS of these per thread).
// Token space is handed out per macroblock: the threads of a macroblock (16 or 4 adjacent lanes) count their nonzero
// levels, scan the counts with shuffles, the first lane takes the macroblock's range with ONE atomicAdd.
// per position of a 4x4 block the smallest |coefficient| that quantises to a nonzero level (tq_thresholds)
struct TqThr { int ok; int t[16]; };

// bit mask of the scan positions of a block: 32 bits for a 4x4 block, 64 for an 8x8 one
template <bool WIDE> struct MaskOf { typedef unsigned type; };
template <> struct MaskOf<true> { typedef unsigned long long type; };
__device__ __forceinline__ int popc_of(unsigned m) { return __popc(m); }
__device__ __forceinline__ int popc_of(unsigned long long m) { return __popcll(m); }
__device__ __forceinline__ int msb_of(unsigned m) { return 31 - __clz(m); }
__device__ __forceinline__ int msb_of(unsigned long long m) { return 63 - __clzll(m); }

template <int N, int STD>
__global__ void __launch_bounds__(128)
k_mc_tq_modes_c(const jmb_me_res *__restrict__ res, int n_mb, int mb_w, unsigned mb_w_rcp, unsigned mode_mask, const __grid_constant__ jmb_quant_desc q,
                const __grid_constant__ TqThr thr,
                const uint8_t *__restrict__ cur, int cur_pitch, const uint8_t *__restrict__ ref_plane0, size_t plane_bytes, int ref_pitch,
                int w, int h, jmb_tq_head *__restrict__ heads, jmb_tq_token *__restrict__ tokens, unsigned token_cap, unsigned *__restrict__ tok_count) {
  constexpr int PER_MB = (N == 4) ? 16 : 4, NN = N * N;
  const int mode = blockIdx.y + 1;
  if (!((mode_mask >> blockIdx.y) & 1)) return;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = t < n_mb * PER_MB;                       // whole macroblocks are live or dead together
  const int mb = live ? t / PER_MB : 0, b = t % PER_MB;
  const int mbr = (int)__umulhi((unsigned)mb, mb_w_rcp), mbx = (mb - mbr * mb_w) * 16, mby = mbr * 16;      // mb / mb_w by its reciprocal (exact: see the launcher)
  const int bx4 = (N == 4) ? (b & 3) : (b & 1) * 2, by4 = (N == 4) ? (b >> 2) : (b >> 1) * 2;
  const int b8 = (by4 >> 1) * 2 + (bx4 >> 1);
  int ux4 = bx4, uy4 = by4;                                  // prediction unit (macroblock.c:946-971)
  if (mode < 5 || N == 8) { ux4 &= ~1; uy4 &= ~1; }
  if (mode == 1) { ux4 = 0; uy4 = 0; }
  const int lw = c_mode_lw4[mode], lh = c_mode_lh4[mode];     // partitions are 1, 2 or 4 blocks wide / high: shifts, not divisions
  const jmb_me_res r = res[mb * 41 + c_mode_base[mode] + ((uy4 >> lh) << (2 - lw)) + (ux4 >> lw)];
  const int qx = ((mbx + ux4 * 4) << 2) + r.mv_x, qy = ((mby + uy4 * 4) << 2) + r.mv_y;
  const int iy = jmb_clip(-JMB_PAD_Y, h + JMB_PAD_Y - 1 - 16, qy >> 2), ix = jmb_clip(-JMB_PAD_X, w + JMB_PAD_X - 1 - 16, qx >> 2);
  const uint8_t *rp = ref_plane0 + (size_t)((qy & 3) * 4 + (qx & 3)) * plane_bytes +
                      (size_t)(iy + JMB_PAD_Y + (by4 - uy4) * 4) * ref_pitch + (ix + JMB_PAD_X + (bx4 - ux4) * 4);
  const uint8_t *sp = cur + (size_t)(mby + by4 * 4) * cur_pitch + mbx + bx4 * 4;
  int rr[NN];
  const unsigned rsh = (unsigned)(size_t)rp & 3u, rsel = 0x3210u + 0x1111u * rsh;      // the pitch is a multiple of 4: one alignment for all rows
  const unsigned *const ra = (const unsigned *)(rp - rsh);
  const int rpw = ref_pitch >> 2;
#pragma unroll
  for (int y = 0; y < N; y++)
#pragma unroll
    for (int x4 = 0; x4 < N; x4 += 4) {
      const unsigned *const rq = ra + y * rpw + (x4 >> 2);
      const unsigned sv = *(const unsigned *)(sp + (size_t)y * cur_pitch + x4), pv = __byte_perm(__ldg(rq), __ldg(rq + 1), rsel);
#pragma unroll
      for (int x = 0; x < 4; x++) rr[y * N + x4 + x] = (int)((sv >> (8 * x)) & 255) - (int)((pv >> (8 * x)) & 255);
    }
  if (N == 4) fwd4(rr); else fwd8(rr);

  // quantisation in scan order (quant_4x4_normal, quant_8x8_normal, quant_8x8cavlc_normal): levels stay in rr[scan position]
  const int qp_per = q.qp / 6, q_bits = (N == 4 ? 15 : 16) + qp_per;
  const bool cavlc8 = (N == 8) && q.is_cavlc;
  const bool clip = (N == 4) ? (q.is_cavlc != 0) : cavlc8;
  // 4x4 blocks, first pass: which scan positions keep a nonzero level -- (|c| * Scale + Offset) >> q_bits != 0 is one multiply-add
  // and a compare; at the usual quantiser steps nearly all positions drop out here.  Second pass, only for the survivors: the level,
  // its clip, and JM's coefficient cost with the run since the previous survivor of the same list.
  typedef typename MaskOf<(NN > 32)>::type mask_t;
  mask_t nzm = 0;                                            // bit k: scan position k holds a nonzero level
  int lv[NN];
  int cost = 0;
  if constexpr (N == 4) {
    // thr.t[idx] = the smallest |c| >= 1 whose level is not 0 (worked out on the host); the multiply-add form for descriptors
    // whose parameters leave the range where that is the same thing
    const int q_one = 1 << q_bits;
    if (thr.ok) {
#pragma unroll
      for (int k = 0; k < NN; k++) {
        const int idx = STD ? STD_SCAN4[k][1] * N + STD_SCAN4[k][0] : q.scan[k][1] * N + q.scan[k][0];
        lv[k] = 0;
        if (abs(rr[idx]) >= thr.t[idx]) nzm |= (mask_t)1 << k;
      }
    } else {
#pragma unroll
      for (int k = 0; k < NN; k++) {
        const int idx = STD ? STD_SCAN4[k][1] * N + STD_SCAN4[k][0] : q.scan[k][1] * N + q.scan[k][0], c = rr[idx];
        lv[k] = 0;
        if (c != 0 && abs(c) * q.qparams[idx][1] + q.qparams[idx][0] >= q_one) nzm |= (mask_t)1 << k;
      }
    }
    if (nzm) {
#pragma unroll
      for (int k = 0; k < NN; k++) {
        if ((nzm >> k) & 1) {
          int i, j;
          if (STD) { i = (N == 4) ? STD_SCAN4[k][0] : (STD == 2 ? STD_SCAN8_CAVLC[k][0] : STD_SCAN8[k][0]);
                     j = (N == 4) ? STD_SCAN4[k][1] : (STD == 2 ? STD_SCAN8_CAVLC[k][1] : STD_SCAN8[k][1]); }
          else { i = q.scan[k][0]; j = q.scan[k][1]; }
          const int idx = j * N + i, c = rr[idx];
          int level = (abs(c) * q.qparams[idx][1] + q.qparams[idx][0]) >> q_bits;
          if (clip) level = min(level, 2063);
          const int k0 = cavlc8 ? (k & ~15) : 0;
          const mask_t before = nzm & (((mask_t)1 << k) - 1) & ~(((mask_t)1 << k0) - 1);      // earlier nonzero levels of the same list
          const int run = before ? (k - 1 - msb_of(before)) : (k - k0);
          cost += (level > 1) ? 999999 : q.c_cost[run];
          lv[k] = c < 0 ? -level : level;
        }
      }
    }
  } else {      // 8x8: one pass, levels as they come -- with 64 positions and 64-bit masks the two-pass form is slower (0.071 -> 0.082 ms at 4K)
    {
      int run[4] = {0, 0, 0, 0};
#pragma unroll
      for (int k = 0; k < NN; k++) {
        int i, j;
        if (STD) { i = (N == 4) ? STD_SCAN4[k][0] : (STD == 2 ? STD_SCAN8_CAVLC[k][0] : STD_SCAN8[k][0]);
                   j = (N == 4) ? STD_SCAN4[k][1] : (STD == 2 ? STD_SCAN8_CAVLC[k][1] : STD_SCAN8[k][1]); }
        else { i = q.scan[k][0]; j = q.scan[k][1]; }
        const int idx = j * N + i, s = cavlc8 ? (k >> 4) : 0;
        const int c = STD ? rr[idx] : rr[idx];
        int level = 0;
        if (c != 0) {
          level = (abs(c) * q.qparams[idx][1] + q.qparams[idx][0]) >> q_bits;
          if (level != 0) {
            if (clip) level = min(level, 2063);
            cost += (level > 1) ? 999999 : q.c_cost[run[s]];
            if (c < 0) level = -level;
            nzm |= (mask_t)1 << k;
          }
        }
        lv[k] = level;
        if (level != 0) run[s] = 0; else run[s]++;
      }
    }
  }
  const int cnt = popc_of(nzm);
  // macroblock-wide exchange: cost per quadrant, coded-block bits, token range
  int c8 = cost;
  if (N == 4) { c8 += __shfl_xor_sync(0xffffffffu, c8, 1); c8 += __shfl_xor_sync(0xffffffffu, c8, 4); }
  unsigned bits = cnt ? ((N == 4) ? (1u << (by4 * 4 + bx4)) : (51u << (4 * b8 - 2 * (b8 & 1)))) : 0u;
  int incl = cnt;
#pragma unroll
  for (int sh = 1; sh < PER_MB; sh <<= 1) {
    bits |= __shfl_xor_sync(0xffffffffu, bits, sh);
    const int up = __shfl_up_sync(0xffffffffu, incl, sh, PER_MB);
    if ((int)(threadIdx.x & (PER_MB - 1)) >= sh) incl += up;
  }
  const int lane = threadIdx.x & 31, leader = lane & ~(PER_MB - 1);
  const int total = __shfl_sync(0xffffffffu, incl, leader + PER_MB - 1);
  unsigned base = 0;
  if (lane == leader && live && total) base = atomicAdd(tok_count, (unsigned)total);
  base = __shfl_sync(0xffffffffu, base, leader);
  // cost8 of the four quadrants to the leader
  unsigned c8sat = (unsigned)min(c8, 255);
  unsigned cq[4];
#pragma unroll
  for (int k = 0; k < 4; k++) cq[k] = __shfl_sync(0xffffffffu, c8sat, leader + ((N == 4) ? ((k >> 1) * 8 + (k & 1) * 2) : k));
  if (!live) return;
  const size_t mo = (size_t)blockIdx.y * n_mb + mb;
  if (lane == leader) {
    uint4 hv;
    hv.x = bits; hv.y = base; hv.z = (unsigned)total | (cq[0] << 16) | (cq[1] << 24); hv.w = cq[2] | (cq[3] << 8);
    *(uint4 *)&heads[mo] = hv;
  }
  if (cnt && base + (unsigned)total <= token_cap) {
    jmb_tq_token *o = tokens + base + (incl - cnt);
#pragma unroll
    for (int k = 0; k < NN; k++) {
      if ((nzm >> k) & 1) {
        const int s = cavlc8 ? (k >> 4) : 0, k0 = cavlc8 ? (k & ~15) : 0;
        const mask_t before = nzm & (((mask_t)1 << k) - 1) & ~(((mask_t)1 << k0) - 1);     // earlier nonzero levels of the same list
        const int run = before ? (k - 1 - msb_of(before)) : (k - k0);
        jmb_tq_token tk;
        tk.level = (int16_t)lv[k]; tk.run = (uint8_t)run;
        tk.blk = (uint8_t)((N == 4) ? b : (cavlc8 ? b8 * 4 + s : b8));
        o[popc_of(nzm & (((mask_t)1 << k) - 1))] = tk;
      }
    }
  }
}

// List quantiser: the DC / AC members of JM's quantiser family (quant_ac4x4_*, quant_dc4x4_normal, quant_dc2x2_*,
// quant_dc4x2_*: lencod/src/quant4x4_normal.c:117,200, quant4x4_around.c:132, quantChroma_normal.c, quantChroma_around.c)
// are one loop over a list of m coefficients already in scan order, with per-position {Offset, Scale, InvScale}, one
// of three dequantisation forms and optional cost / adaptive-rounding outputs.  One thread per list.
__global__ void k_quant_list(const jmb_qlist_desc *__restrict__ qd, int *coef, int nlist, int *levels, int *runs, int *fadjust,
                             int *coeff_cost, int *nonzero) {
  __shared__ jmb_qlist_desc q;
  for (int i = threadIdx.x; i < (int)(sizeof(q) / 4); i += blockDim.x) ((int *)&q)[i] = ((const int *)qd)[i];
  __syncthreads();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nlist) return;
  int *c = coef + (size_t)t * q.m, *lv = levels + (size_t)t * 17, *rn = runs + (size_t)t * 17;
  int *fa = (fadjust && q.around) ? fadjust + (size_t)t * q.m : nullptr;
  int run = 0, n = 0, nz = 0, cost = coeff_cost ? coeff_cost[t] : 0;
  for (int k = 0; k < q.m; k++) {
    const int v = c[k];
    int adj = 0;
    if (v != 0) {
      const int scaled = abs(v) * q.params[k][1];
      int level = (scaled + q.params[k][0]) >> q.q_bits;
      if (level != 0) {
        if (q.clip) level = min(level, 2063);
        adj = (q.adapt_rnd_weight * (scaled - (level << q.q_bits)) + (1 << q.q_bits)) >> (q.q_bits + 1);
        if (q.use_cost) cost += (level > 1) ? 999999 : q.c_cost[run];
        if (v < 0) level = -level;
        const int dq = (level * q.params[k][2]) << q.qp_per;
        c[k] = q.dequant == JMB_DQ_LEVEL ? level : (q.dequant == JMB_DQ_SHIFT ? dq : ((dq + 8) >> 4));
        lv[n] = level; rn[n] = run; n++; run = 0; nz = 1;
      } else { c[k] = 0; run++; }
    } else run++;
    if (fa) fa[k] = adj;
  }
  lv[n] = 0;
  if (coeff_cost) coeff_cost[t] = cost;
  nonzero[t] = nz;
}

// hadamard4x4 / ihadamard4x4 / hadamard4x2 / ihadamard4x2 / hadamard2x2 / ihadamard2x2, lcommon/src/transform.c:121-330.
// One thread per block; flat layouts: 4x4 row-major [16]; 4x2: forward in/out = 2 rows x 4 [8], inverse in = 2 rows x 4,
// out = 4 rows x 2 [8]; 2x2: [4] = {b00, b04, b40, b44} in, {t0..t3} out (and the reverse).
__global__ void k_hadamard(int kind, int *vals, int nblk) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nblk) return;
  if (kind == JMB_HAD_4X4 || kind == JMB_IHAD_4X4) {
    int *b = vals + (size_t)t * 16, m[16];
    for (int i = 0; i < 4; i++) {
      const int p0 = b[4 * i], p1 = b[4 * i + 1], p2 = b[4 * i + 2], p3 = b[4 * i + 3];
      if (kind == JMB_HAD_4X4) { const int t0 = p0 + p3, t1 = p1 + p2, t2 = p1 - p2, t3 = p0 - p3; m[4 * i] = t0 + t1; m[4 * i + 1] = t3 + t2; m[4 * i + 2] = t0 - t1; m[4 * i + 3] = t3 - t2; }
      else { const int q0 = p0 + p2, q1 = p0 - p2, q2 = p1 - p3, q3 = p1 + p3; m[4 * i] = q0 + q3; m[4 * i + 1] = q1 + q2; m[4 * i + 2] = q1 - q2; m[4 * i + 3] = q0 - q3; }
    }
    for (int i = 0; i < 4; i++) {
      const int p0 = m[i], p1 = m[4 + i], p2 = m[8 + i], p3 = m[12 + i];
      if (kind == JMB_HAD_4X4) { const int t0 = p0 + p3, t1 = p1 + p2, t2 = p1 - p2, t3 = p0 - p3; b[i] = (t0 + t1) >> 1; b[4 + i] = (t2 + t3) >> 1; b[8 + i] = (t0 - t1) >> 1; b[12 + i] = (t3 - t2) >> 1; }
      else { const int q0 = p0 + p2, q1 = p0 - p2, q2 = p1 - p3, q3 = p1 + p3; b[i] = q0 + q3; b[4 + i] = q1 + q2; b[8 + i] = q1 - q2; b[12 + i] = q0 - q3; }
    }
  } else if (kind == JMB_HAD_4X2 || kind == JMB_IHAD_4X2) {
    int *b = vals + (size_t)t * 8, m[8], o[8];
    for (int i = 0; i < 4; i++) { m[i] = b[i] + b[4 + i]; m[4 + i] = b[i] - b[4 + i]; }
    for (int i = 0; i < 2; i++) {
      const int p0 = m[4 * i], p1 = m[4 * i + 1], p2 = m[4 * i + 2], p3 = m[4 * i + 3];
      if (kind == JMB_HAD_4X2) { const int t0 = p0 + p3, t1 = p1 + p2, t2 = p1 - p2, t3 = p0 - p3; o[4 * i] = t0 + t1; o[4 * i + 1] = t3 + t2; o[4 * i + 2] = t0 - t1; o[4 * i + 3] = t3 - t2; }
      else { const int t0 = p0 + p2, t1 = p0 - p2, t2 = p1 - p3, t3 = p1 + p3; o[i] = t0 + t3; o[2 + i] = t1 + t2; o[4 + i] = t1 - t2; o[6 + i] = t0 - t3; }   // block[r][i], 4 rows x 2
    }
    for (int i = 0; i < 8; i++) b[i] = o[i];
  } else {
    int *b = vals + (size_t)t * 4;
    const int a = b[0], c = b[1], d = b[2], e = b[3];
    if (kind == JMB_HAD_2X2) { const int p0 = a + c, p1 = a - c, p2 = d + e, p3 = d - e; b[0] = p0 + p2; b[1] = p1 + p3; b[2] = p0 - p2; b[3] = p1 - p3; }
    else { const int t0 = a + c, t1 = a - c, t2 = d + e, t3 = d - e; b[0] = t0 + t2; b[1] = t1 + t3; b[2] = t0 - t2; b[3] = t1 - t3; }
  }
}

template <int N>
__global__ void k_inverse(int *blocks, int nblk) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nblk) return;
  int b[N * N];
  int *g = blocks + (size_t)i * N * N;
#pragma unroll
  for (int k = 0; k < N * N; k++) b[k] = g[k];
  if (N == 4) inv4(b); else inv8(b);
#pragma unroll
  for (int k = 0; k < N * N; k++) g[k] = b[k];
}

// luma_residual_coding (lencod/src/macroblock.c:1182-1257) of a non-skipped inter macroblock of a P slice, for EVERY
// partition mode of every macroblock in one launch: prediction (from the 41 search results) -> residual -> forward
// transform -> quantisation -> per block inverse transform + sample_reconstruct (lcommon/src/blk_prediction.c:48) when a
// level is nonzero, else the prediction -> coefficient thresholding: quadrants with cost <= _LUMA_COEFF_COST_ (4) are
// reset (reset_block, macroblock.c:806), a macroblock whose summed cost is <= _LUMA_MB_COEFF_COST_ (5) drops its luma cbp
// and takes the prediction (:1248-1255) -> SSE of the reconstruction against the source (what RDCost_for_macroblocks
// charges as distortion).  Thread = one transform block; the 16 (4) threads of a macroblock exchange costs by shuffle.
template <int N, int STD>
__global__ void __launch_bounds__(128)
k_luma_rc_modes(const jmb_me_res *__restrict__ res, const jmb_mb_pred *__restrict__ pred, int first_mb, int n_mb, int mb_w, unsigned mode_mask,
                const jmb_quant_desc *__restrict__ qd, const uint8_t *__restrict__ cur, int cur_pitch, const uint8_t *const *__restrict__ ref_planes,
                size_t plane_bytes, int ref_pitch, int w, int h, int nref, int *__restrict__ err,
                int16_t *__restrict__ levels, int *__restrict__ cost8, unsigned *__restrict__ cbp_blk, unsigned *__restrict__ cbp,
                uint8_t *__restrict__ recon, int *__restrict__ sse) {
  __shared__ jmb_quant_desc q;
  for (int i = threadIdx.x; i < (int)(sizeof(q) / 4); i += blockDim.x) ((int *)&q)[i] = ((const int *)qd)[i];
  __syncthreads();
  constexpr int PER_MB = (N == 4) ? 16 : 4;
  if (!((mode_mask >> blockIdx.y) & 1)) return;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  bool live = t < n_mb * PER_MB;                               // dead threads still take part in the shuffles
  const int mb = live ? t / PER_MB : 0, b = t % PER_MB;        // mb: index into the outputs / pred; picture address = first_mb + mb
  const int mbx = ((first_mb + mb) % mb_w) * 16, mby = ((first_mb + mb) / mb_w) * 16;
  const int bx4 = (N == 4) ? (b & 3) : (b & 1) * 2, by4 = (N == 4) ? (b >> 2) : (b >> 1) * 2;
  const int b8 = (by4 >> 1) * 2 + (bx4 >> 1);
  // prediction source: an explicit table (mode / mvs / reference per quadrant) or mode blockIdx.y+1 of the search results
  int mode = pred ? pred[mb].b8mode[b8] : blockIdx.y + 1;
  if (pred) {   // a device-resident table is validated here; a rejected macroblock writes nothing (the whole macroblock: its
                // threads exchange costs by shuffle, so every one of them must see the same verdict)
    int bad = pred_check(mode, pred[mb].ref[b8], N, nref);
#pragma unroll
    for (int sh = 1; sh < PER_MB; sh <<= 1) bad |= __shfl_xor_sync(0xffffffffu, bad, sh);
    if (bad) { if (live && b == 0) jmb_req_report(err, bad, first_mb + mb); live = false; mode = 1; }
  }
  int ux4 = bx4, uy4 = by4;
  if (mode < 5 || N == 8) { ux4 &= ~1; uy4 &= ~1; }
  if (mode == 1) { ux4 = 0; uy4 = 0; }                       // P16x16: one 16x16 prediction, one origin clamp (macroblock.c:1225)
  int mvx, mvy, rf = 0;
  if (pred) { mvx = pred[mb].mv[uy4 * 4 + ux4][0]; mvy = pred[mb].mv[uy4 * 4 + ux4][1]; rf = min((int)pred[mb].ref[b8], nref - 1); }
  else {
    const jmb_me_res r = res[(first_mb + mb) * 41 + c_mode_base[mode] + (uy4 / c_mode_h4[mode]) * (4 / c_mode_w4[mode]) + ux4 / c_mode_w4[mode]];
    mvx = r.mv_x; mvy = r.mv_y;
  }
  const int qx = ((mbx + ux4 * 4) << 2) + mvx, qy = ((mby + uy4 * 4) << 2) + mvy;
  const int iy = jmb_clip(-JMB_PAD_Y, h + JMB_PAD_Y - 1 - 16, qy >> 2), ix = jmb_clip(-JMB_PAD_X, w + JMB_PAD_X - 1 - 16, qx >> 2);
  const uint8_t *rp = ref_planes[rf] + (size_t)((qy & 3) * 4 + (qx & 3)) * plane_bytes +
                      (size_t)(iy + JMB_PAD_Y + (by4 - uy4) * 4) * ref_pitch + (ix + JMB_PAD_X + (bx4 - ux4) * 4);
  const uint8_t *sp = cur + (size_t)(mby + by4 * 4) * cur_pitch + mbx + bx4 * 4;
  int rr[N * N];
  uint8_t pr[N * N], sr[N * N];
#pragma unroll
  for (int y = 0; y < N; y++)
#pragma unroll
    for (int x = 0; x < N; x++) {
      sr[y * N + x] = sp[(size_t)y * cur_pitch + x]; pr[y * N + x] = rp[(size_t)y * ref_pitch + x];
      rr[y * N + x] = (int)sr[y * N + x] - (int)pr[y * N + x];
    }
  if (N == 4) fwd4(rr); else fwd8(rr);
  const size_t mo = (size_t)blockIdx.y * n_mb + mb;          // mode-major output index
  int16_t lv[N * N];
  QOut o = quant_block<N, false, STD>(q, rr, nullptr, nullptr, nullptr, lv);
  // coefficient thresholding across the macroblock's threads
  int c8 = o.cost;
  if (N == 4) { c8 += __shfl_xor_sync(0xffffffffu, c8, 1); c8 += __shfl_xor_sync(0xffffffffu, c8, 4); }   // the quadrant's 4 blocks
  const bool reset8 = c8 <= 4;                               // _LUMA_COEFF_COST_
  if (reset8) c8 = 0;
  int tot = c8;
  if (N == 4) { tot += __shfl_xor_sync(0xffffffffu, tot, 2); tot += __shfl_xor_sync(0xffffffffu, tot, 8); }   // one lane per quadrant: b^2, b^8
  else { tot += __shfl_xor_sync(0xffffffffu, tot, 1); tot += __shfl_xor_sync(0xffffffffu, tot, 2); }
  const bool reset_mb = tot <= 5;                            // _LUMA_MB_COEFF_COST_
  const bool coded = o.nonzero && !reset8 && !reset_mb;
  int d2 = 0;
  if (o.nonzero && !reset8 && !reset_mb) { if (N == 4) inv4(rr); else inv8(rr); }
  uint8_t *rec = recon ? recon + mo * 256 + (by4 * 4) * 16 + bx4 * 4 : nullptr;
#pragma unroll
  for (int y = 0; y < N; y++)
#pragma unroll
    for (int x = 0; x < N; x++) {
      const int v = coded ? min(max(((rr[y * N + x] + 32) >> 6) + (int)pr[y * N + x], 0), 255) : (int)pr[y * N + x];
      const int d = (int)sr[y * N + x] - v;
      d2 += d * d;
      if (rec && live) rec[y * 16 + x] = (uint8_t)v;
    }
#pragma unroll
  for (int sh = 1; sh < PER_MB; sh <<= 1) d2 += __shfl_xor_sync(0xffffffffu, d2, sh);
  unsigned bits = coded ? ((N == 4) ? (1u << (by4 * 4 + bx4)) : (51u << (4 * b8 - 2 * (b8 & 1)))) : 0u;
  unsigned cb = coded ? (1u << b8) : 0u;
#pragma unroll
  for (int sh = 1; sh < PER_MB; sh <<= 1) { bits |= __shfl_xor_sync(0xffffffffu, bits, sh); cb |= __shfl_xor_sync(0xffffffffu, cb, sh); }
  if (!live) return;
  int16_t *lo = levels + mo * 256 + b * N * N;               // reset_block zeroes the quadrant's levels (cofAC memset)
#pragma unroll
  for (int k = 0; k < N * N; k++) lo[k] = reset8 ? (int16_t)0 : lv[k];
  if (N == 8 || (b & 5) == 0) cost8[mo * 4 + b8] = c8;       // one lane per quadrant (4x4: bx4, by4 even)
  if (b == 0) { cbp_blk[mo] = bits; cbp[mo] = cb; sse[mo] = d2; }
}

// all_mv fill of BlockMotionSearch (lencod/src/mv_search.c:1005-1014) for one partition mode of every
// macroblock: the mv of each partition is replicated over the 4x4 blocks it covers.
__global__ void k_pred_from_results(const jmb_me_res *__restrict__ res, int n_mb, int mode, jmb_mb_pred *__restrict__ pred) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_mb * 16) return;
  const int mb = t >> 4, b = t & 15, bx4 = b & 3, by4 = b >> 2;
  const int base[8] = {0, 0, 1, 3, 5, 9, 17, 25}, w4[8] = {4, 4, 4, 2, 2, 2, 1, 1}, h4[8] = {4, 4, 2, 4, 2, 1, 2, 1};
  const int slot = base[mode] + (by4 / h4[mode]) * (4 / w4[mode]) + bx4 / w4[mode];
  const jmb_me_res r = res[mb * 41 + slot];
  pred[mb].mv[b][0] = r.mv_x; pred[mb].mv[b][1] = r.mv_y;
  if (b < 4) { pred[mb].b8mode[b] = (uint8_t)mode; pred[mb].ref[b] = 0; }
}

}  // namespace

static int check_qdesc(jmb_ctx *ctx, const jmb_quant_desc *q) {
  if (!q || (q->n != 4 && q->n != 8)) return jmb_fail(ctx, JMB_ERR_ARG, "quant desc: n must be 4 or 8");
  if (q->qp < 0 || q->qp > 87) return jmb_fail(ctx, JMB_ERR_ARG, "quant desc: qp %d", q->qp);
  for (int k = 0; k < q->n * q->n; k++)
    if (q->scan[k][0] >= q->n || q->scan[k][1] >= q->n) return jmb_fail(ctx, JMB_ERR_ARG, "quant desc: scan[%d] outside the block", k);
  return 0;
}

// 1 / 2 when the descriptor's scan is the standard zig-zag / the standard 8x8 CAVLC interleave (compile-time scan
// kernels), else 0 (table-driven kernels)
// level = (|c| * Scale + Offset) >> q_bits is nonzero from some |c| on: that |c| per position of a 4x4 block.  Only claimed (ok)
// where the kernel's 32-bit arithmetic is exact for every coefficient a 4x4 residual can produce (|c| <= 255 * 36).
static TqThr tq_thresholds(const jmb_quant_desc *q) {
  TqThr o;
  memset(&o, 0, sizeof(o));
  if (q->n != 4) return o;
  const int q_bits = 15 + q->qp / 6;
  if (q_bits < 0 || q_bits > 30) return o;
  const long long one = 1ll << q_bits;
  for (int i = 0; i < 16; i++) {
    const long long off = q->qparams[i][0], sc = q->qparams[i][1];
    if (off < 0 || sc < 0 || 9180 * sc + off > 0x7fffffffll) return o;
    long long t;
    if (off >= one) t = 1;
    else if (sc == 0) t = 0x7fffffff;
    else t = (one - off + sc - 1) / sc;
    o.t[i] = (int)(t < 1 ? 1 : (t > 0x7fffffff ? 0x7fffffff : t));
  }
  o.ok = 1;
  return o;
}

static int std_scan_kind(const jmb_quant_desc *q) {
  static const unsigned char Z4[16][2] = {{0,0}, {1,0}, {0,1}, {0,2}, {1,1}, {2,0}, {3,0}, {2,1}, {1,2}, {0,3}, {1,3}, {2,2}, {3,1}, {3,2}, {2,3}, {3,3}};
  static const unsigned char Z8[64][2] = {{0,0}, {1,0}, {0,1}, {0,2}, {1,1}, {2,0}, {3,0}, {2,1}, {1,2}, {0,3}, {0,4}, {1,3}, {2,2}, {3,1}, {4,0}, {5,0}, {4,1}, {3,2}, {2,3}, {1,4}, {0,5}, {0,6}, {1,5}, {2,4}, {3,3}, {4,2}, {5,1}, {6,0}, {7,0}, {6,1}, {5,2}, {4,3}, {3,4}, {2,5}, {1,6}, {0,7}, {1,7}, {2,6}, {3,5}, {4,4}, {5,3}, {6,2}, {7,1}, {7,2}, {6,3}, {5,4}, {4,5}, {3,6}, {2,7}, {3,7}, {4,6}, {5,5}, {6,4}, {7,3}, {7,4}, {6,5}, {5,6}, {4,7}, {5,7}, {6,6}, {7,5}, {7,6}, {6,7}, {7,7}};
  static const unsigned char Z8C[64][2] = {{0,0}, {1,1}, {1,2}, {2,2}, {4,1}, {0,5}, {3,3}, {7,0}, {3,4}, {1,7}, {5,3}, {6,3}, {2,7}, {6,4}, {5,6}, {7,5}, {1,0}, {2,0}, {0,3}, {3,1}, {3,2}, {0,6}, {4,2}, {6,1}, {2,5}, {2,6}, {6,2}, {5,4}, {3,7}, {7,3}, {4,7}, {7,6}, {0,1}, {3,0}, {0,4}, {4,0}, {2,3}, {1,5}, {5,1}, {5,2}, {1,6}, {3,5}, {7,1}, {4,5}, {4,6}, {7,4}, {5,7}, {6,7}, {0,2}, {2,1}, {1,3}, {5,0}, {1,4}, {2,4}, {6,0}, {4,3}, {0,7}, {4,4}, {7,2}, {3,6}, {5,5}, {6,5}, {6,6}, {7,7}};
  if (q->n == 4) return memcmp(q->scan, Z4, sizeof(Z4)) ? 0 : 1;
  if (q->is_cavlc) return memcmp(q->scan, Z8C, sizeof(Z8C)) ? 0 : 2;
  return memcmp(q->scan, Z8, sizeof(Z8)) ? 0 : 1;
}

static int upload_ref_table(jmb_ctx *ctx, const uint8_t *const **d_tab) {
  const uint8_t *tab[JMB_MAX_REFS];
  for (int i = 0; i < JMB_MAX_REFS; i++) tab[i] = i < ctx->nref ? ctx->refs[ctx->ref_list[i]].planes : nullptr;
  int rc = jmb_reserve_dev(ctx, &ctx->d_reftab, &ctx->d_reftab_cap, sizeof(tab)); if (rc) return rc;
  JMB_CUDA(ctx, cudaMemcpyAsync(ctx->d_reftab, tab, sizeof(tab), cudaMemcpyHostToDevice, ctx->stream));
  *d_tab = (const uint8_t *const *)ctx->d_reftab;
  return 0;
}

static int upload_qdesc(jmb_ctx *ctx, const jmb_quant_desc *q, const jmb_quant_desc **d_q) {
  int rc = jmb_reserve_dev(ctx, &ctx->d_qdesc, &ctx->d_qdesc_cap, sizeof(*q)); if (rc) return rc;
  JMB_CUDA(ctx, cudaMemcpyAsync(ctx->d_qdesc, q, sizeof(*q), cudaMemcpyHostToDevice, ctx->stream));
  *d_q = (const jmb_quant_desc *)ctx->d_qdesc;
  return 0;
}

extern "C" {

int jmb_forward_transform(jmb_ctx *ctx, int32_t *blocks, int nblk, int n, int loc) {
  if (n != 4 && n != 8) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_forward_transform: n=%d", n);
  if (nblk <= 0) return JMB_OK;
  JMB_CUDA(ctx, cudaSetDevice(ctx->device));
  size_t bytes = (size_t)nblk * n * n * 4;
  int *d = blocks;
  if (loc == JMB_HOST) {
    int rc = jmb_reserve_dev(ctx, &ctx->d_stage, &ctx->d_stage_cap, bytes); if (rc) return rc;
    JMB_CUDA(ctx, cudaMemcpyAsync(ctx->d_stage, blocks, bytes, cudaMemcpyHostToDevice, ctx->stream));
    d = (int *)ctx->d_stage;
  }
  jmb_time_begin(ctx, JMB_K_FORWARD);
  if (n == 4) k_forward<4><<<(nblk + 127) / 128, 128, 0, ctx->stream>>>(d, nblk);
  else k_forward<8><<<(nblk + 63) / 64, 64, 0, ctx->stream>>>(d, nblk);
  jmb_time_end(ctx, JMB_K_FORWARD);
  JMB_LAUNCH_CHECK(ctx);
  if (loc == JMB_HOST) {
    JMB_CUDA(ctx, cudaMemcpyAsync(blocks, d, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    JMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return JMB_OK;
}

int jmb_quant_blocks(jmb_ctx *ctx, const jmb_quant_desc *q, int do_transform, int32_t *coef, int nblk,
                     int32_t *levels, int32_t *runs, int32_t *fadjust, int32_t *coeff_cost, int32_t *nonzero, int loc) {
  int rc = check_qdesc(ctx, q); if (rc) return rc;
  if (nblk <= 0) return JMB_OK;
  if (!coef || !levels || !runs || !coeff_cost || !nonzero) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_quant_blocks: NULL buffer");
  JMB_CUDA(ctx, cudaSetDevice(ctx->device));
  const int nn = q->n * q->n, lr = (q->n == 4) ? 17 : (q->is_cavlc ? 68 : 65);
  const jmb_quant_desc *d_q; rc = upload_qdesc(ctx, q, &d_q); if (rc) return rc;
  int *d_coef = coef, *d_lv = levels, *d_rn = runs, *d_fa = fadjust, *d_cc = coeff_cost, *d_nz = nonzero;
  if (loc == JMB_HOST) {
    // one device arena: coef | levels | runs | fadjust | cost | nonzero
    size_t o_lv = (size_t)nblk * nn, o_rn = o_lv + (size_t)nblk * lr, o_fa = o_rn + (size_t)nblk * lr, o_cc = o_fa + (size_t)nblk * nn,
           o_nz = o_cc + nblk, total = o_nz + nblk;
    rc = jmb_reserve_dev(ctx, &ctx->d_stage, &ctx->d_stage_cap, total * 4); if (rc) return rc;
    int *base = (int *)ctx->d_stage;
    d_coef = base; d_lv = base + o_lv; d_rn = base + o_rn; d_fa = base + o_fa; d_cc = base + o_cc; d_nz = base + o_nz;
    JMB_CUDA(ctx, cudaMemcpyAsync(d_coef, coef, (size_t)nblk * nn * 4, cudaMemcpyHostToDevice, ctx->stream));
    JMB_CUDA(ctx, cudaMemcpyAsync(d_cc, coeff_cost, (size_t)nblk * 4, cudaMemcpyHostToDevice, ctx->stream));
    JMB_CUDA(ctx, cudaMemsetAsync(d_lv, 0, (size_t)nblk * lr * 2 * 4, ctx->stream));
  }
  jmb_time_begin(ctx, JMB_K_QUANT);
  if (q->n == 4) k_quant_blocks<4><<<(nblk + 127) / 128, 128, 0, ctx->stream>>>(d_q, do_transform, d_coef, nblk, lr, d_lv, d_rn, d_fa, d_cc, d_nz);
  else k_quant_blocks<8><<<(nblk + 63) / 64, 64, 0, ctx->stream>>>(d_q, do_transform, d_coef, nblk, lr, d_lv, d_rn, d_fa, d_cc, d_nz);
  jmb_time_end(ctx, JMB_K_QUANT);
  JMB_LAUNCH_CHECK(ctx);
  if (loc == JMB_HOST) {
    JMB_CUDA(ctx, cudaMemcpyAsync(coef, d_coef, (size_t)nblk * nn * 4, cudaMemcpyDeviceToHost, ctx->stream));
    JMB_CUDA(ctx, cudaMemcpyAsync(levels, d_lv, (size_t)nblk * lr * 4, cudaMemcpyDeviceToHost, ctx->stream));
    JMB_CUDA(ctx, cudaMemcpyAsync(runs, d_rn, (size_t)nblk * lr * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (fadjust && q->around) JMB_CUDA(ctx, cudaMemcpyAsync(fadjust, d_fa, (size_t)nblk * nn * 4, cudaMemcpyDeviceToHost, ctx->stream));
    JMB_CUDA(ctx, cudaMemcpyAsync(coeff_cost, d_cc, (size_t)nblk * 4, cudaMemcpyDeviceToHost, ctx->stream));
    JMB_CUDA(ctx, cudaMemcpyAsync(nonzero, d_nz, (size_t)nblk * 4, cudaMemcpyDeviceToHost, ctx->stream));
    JMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return JMB_OK;
}

int jmb_pred_from_results(jmb_ctx *ctx, const jmb_me_res *res, int n_mb, int mode, jmb_mb_pred *pred, int loc) {
  if (mode < 1 || mode > 7 || n_mb <= 0) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_pred_from_results: mode %d n_mb %d", mode, n_mb);
  JMB_CUDA(ctx, cudaSetDevice(ctx->device));
  const jmb_me_res *d_res = res; jmb_mb_pred *d_pred = pred;
  if (!res) {          // the results of the last search call, still on the device
    if (!ctx->last_res || ctx->last_res_n < n_mb * 41) return jmb_fail(ctx, JMB_ERR_STATE, "jmb_pred_from_results: no resident search results for %d macroblocks", n_mb);
    d_res = ctx->last_res;
  } else if (loc == JMB_HOST) {
    int rc = jmb_reserve_dev(ctx, &ctx->d_stage4, &ctx->d_stage4_cap, (size_t)n_mb * 41 * sizeof(jmb_me_res)); if (rc) return rc;
    JMB_CUDA(ctx, cudaMemcpyAsync(ctx->d_stage4, res, (size_t)n_mb * 41 * sizeof(jmb_me_res), cudaMemcpyHostToDevice, ctx->stream));
    d_res = (const jmb_me_res *)ctx->d_stage4;
  }
  if (!pred || loc == JMB_HOST) {   // keep the prediction table in the context (jmb_mc_tq(pred = NULL) reads it)
    int rc = jmb_reserve_dev(ctx, &ctx->d_pred_keep, &ctx->d_pred_keep_cap, (size_t)n_mb * sizeof(jmb_mb_pred)); if (rc) return rc;
    d_pred = (jmb_mb_pred *)ctx->d_pred_keep; ctx->pred_keep_n = n_mb;
  }
  jmb_time_begin(ctx, JMB_K_PRED);
  k_pred_from_results<<<(n_mb * 16 + 255) / 256, 256, 0, ctx->stream>>>(d_res, n_mb, mode, d_pred);
  jmb_time_end(ctx, JMB_K_PRED);
  JMB_LAUNCH_CHECK(ctx);
  if (pred && loc == JMB_HOST) {
    JMB_CUDA(ctx, cudaMemcpyAsync(pred, d_pred, (size_t)n_mb * sizeof(jmb_mb_pred), cudaMemcpyDeviceToHost, ctx->stream));
    JMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return JMB_OK;
}

int jmb_mc_tq(jmb_ctx *ctx, const jmb_mb_pred *pred, int n_mb, const jmb_quant_desc *q,
              int16_t *levels, int32_t *coeff_cost, uint32_t *cbp_blk, int loc) {
  int rc = check_qdesc(ctx, q); if (rc) return rc;
  if (!ctx->cur || ctx->nref == 0) return jmb_fail(ctx, JMB_ERR_STATE, "jmb_mc_tq: call jmb_pic_begin first");
  const int mb_w = ctx->cur_w / 16, mb_total = mb_w * (ctx->cur_h / 16);
  if (n_mb <= 0 || n_mb > mb_total) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_mc_tq: n_mb %d (picture has %d)", n_mb, mb_total);
  JMB_CUDA(ctx, cudaSetDevice(ctx->device));
  const jmb_ref &r0 = ctx->refs[ctx->ref_list[0]];
  const uint8_t *tab[JMB_MAX_REFS];
  for (int i = 0; i < JMB_MAX_REFS; i++) tab[i] = i < ctx->nref ? ctx->refs[ctx->ref_list[i]].planes : nullptr;
  rc = jmb_reserve_dev(ctx, &ctx->d_reftab, &ctx->d_reftab_cap, sizeof(tab)); if (rc) return rc;
  JMB_CUDA(ctx, cudaMemcpyAsync(ctx->d_reftab, tab, sizeof(tab), cudaMemcpyHostToDevice, ctx->stream));
  const jmb_quant_desc *d_q; rc = upload_qdesc(ctx, q, &d_q); if (rc) return rc;
  const jmb_mb_pred *d_pred = pred; int16_t *d_lv = levels; int *d_cc = coeff_cost; unsigned *d_cbp = cbp_blk;
  if (!pred) {
    if (!ctx->d_pred_keep || ctx->pred_keep_n < n_mb) return jmb_fail(ctx, JMB_ERR_STATE, "jmb_mc_tq: no resident prediction table for %d macroblocks", n_mb);
    d_pred = (const jmb_mb_pred *)ctx->d_pred_keep;
  }
  if (loc == JMB_HOST) {
    if (pred) for (int i = 0; i < n_mb; i++)
      for (int k = 0; k < 4; k++)
        if (pred[i].b8mode[k] < 1 || pred[i].b8mode[k] > 7 || pred[i].ref[k] >= ctx->nref)
          return jmb_fail(ctx, JMB_ERR_ARG, "jmb_mc_tq: macroblock %d quadrant %d: mode %d ref %d", i, k, pred[i].b8mode[k], pred[i].ref[k]);
    rc = jmb_reserve_dev(ctx, &ctx->d_stage3, &ctx->d_stage3_cap, (size_t)n_mb * (512 + 16 + 4)); if (rc) return rc;
    if (pred) {
      rc = jmb_reserve_dev(ctx, &ctx->d_stage, &ctx->d_stage_cap, (size_t)n_mb * sizeof(jmb_mb_pred)); if (rc) return rc;
      JMB_CUDA(ctx, cudaMemcpyAsync(ctx->d_stage, pred, (size_t)n_mb * sizeof(jmb_mb_pred), cudaMemcpyHostToDevice, ctx->stream));
      d_pred = (const jmb_mb_pred *)ctx->d_stage;
    }
    d_lv = (int16_t *)ctx->d_stage3; d_cc = (int *)((char *)ctx->d_stage3 + (size_t)n_mb * 512); d_cbp = (unsigned *)(d_cc + (size_t)n_mb * 4);
  }
  JMB_CUDA(ctx, cudaMemsetAsync(d_cc, 0, (size_t)n_mb * 16, ctx->stream));
  JMB_CUDA(ctx, cudaMemsetAsync(d_cbp, 0, (size_t)n_mb * 4, ctx->stream));
  jmb_time_begin(ctx, JMB_K_MC_TQ);
  if (q->n == 4) k_mc_tq<4><<<(n_mb * 16 + 127) / 128, 128, 0, ctx->stream>>>(d_pred, n_mb, mb_w, d_q, ctx->cur, ctx->cur_pitch,
        (const uint8_t *const *)ctx->d_reftab, r0.plane_bytes, r0.pitch, ctx->cur_w, ctx->cur_h, ctx->nref, ctx->d_err, d_lv, d_cc, d_cbp);
  else k_mc_tq<8><<<(n_mb * 4 + 63) / 64, 64, 0, ctx->stream>>>(d_pred, n_mb, mb_w, d_q, ctx->cur, ctx->cur_pitch,
        (const uint8_t *const *)ctx->d_reftab, r0.plane_bytes, r0.pitch, ctx->cur_w, ctx->cur_h, ctx->nref, ctx->d_err, d_lv, d_cc, d_cbp);
  jmb_time_end(ctx, JMB_K_MC_TQ);
  JMB_LAUNCH_CHECK(ctx);
  if (loc == JMB_HOST) {
    JMB_CUDA(ctx, cudaMemcpyAsync(levels, d_lv, (size_t)n_mb * 512, cudaMemcpyDeviceToHost, ctx->stream));
    JMB_CUDA(ctx, cudaMemcpyAsync(coeff_cost, d_cc, (size_t)n_mb * 16, cudaMemcpyDeviceToHost, ctx->stream));
    JMB_CUDA(ctx, cudaMemcpyAsync(cbp_blk, d_cbp, (size_t)n_mb * 4, cudaMemcpyDeviceToHost, ctx->stream));
    JMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return JMB_OK;
}

int jmb_hadamard(jmb_ctx *ctx, int kind, int32_t *vals, int nblk, int loc) {
  if (kind < JMB_HAD_4X4 || kind > JMB_IHAD_2X2) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_hadamard: kind %d", kind);
  if (nblk <= 0) return JMB_OK;
  JMB_CUDA(ctx, cudaSetDevice(ctx->device));
  const int per = (kind <= JMB_IHAD_4X4) ? 16 : (kind <= JMB_IHAD_4X2 ? 8 : 4);
  const size_t bytes = (size_t)nblk * per * 4;
  int *d = vals;
  if (loc == JMB_HOST) {
    int rc = jmb_reserve_dev(ctx, &ctx->d_stage, &ctx->d_stage_cap, bytes); if (rc) return rc;
    JMB_CUDA(ctx, cudaMemcpyAsync(ctx->d_stage, vals, bytes, cudaMemcpyHostToDevice, ctx->stream));
    d = (int *)ctx->d_stage;
  }
  jmb_time_begin(ctx, JMB_K_FORWARD);
  k_hadamard<<<(nblk + 127) / 128, 128, 0, ctx->stream>>>(kind, d, nblk);
  jmb_time_end(ctx, JMB_K_FORWARD);
  JMB_LAUNCH_CHECK(ctx);
  if (loc == JMB_HOST) {
    JMB_CUDA(ctx, cudaMemcpyAsync(vals, d, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    JMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return JMB_OK;
}

int jmb_quant_list(jmb_ctx *ctx, const jmb_qlist_desc *q, int32_t *coef, int nlist, int32_t *levels, int32_t *runs, int32_t *fadjust,
                   int32_t *coeff_cost, int32_t *nonzero, int loc) {
  if (!q || q->m < 1 || q->m > 16 || q->q_bits < 1 || q->q_bits > 30 || q->qp_per < 0 || q->qp_per > 14 || q->dequant < JMB_DQ_LEVEL || q->dequant > JMB_DQ_SHIFT_RND4)
    return jmb_fail(ctx, JMB_ERR_ARG, "jmb_quant_list: bad descriptor");
  if (nlist <= 0) return JMB_OK;
  if (!coef || !levels || !runs || !nonzero) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_quant_list: NULL buffer");
  JMB_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = jmb_reserve_dev(ctx, &ctx->d_qdesc, &ctx->d_qdesc_cap, sizeof(*q)); if (rc) return rc;
  JMB_CUDA(ctx, cudaMemcpyAsync(ctx->d_qdesc, q, sizeof(*q), cudaMemcpyHostToDevice, ctx->stream));
  int *d_c = coef, *d_lv = levels, *d_rn = runs, *d_fa = fadjust, *d_cc = coeff_cost, *d_nz = nonzero;
  const size_t m = (size_t)q->m, n = (size_t)nlist;
  if (loc == JMB_HOST) {      // arena: coef | levels | runs | fadjust | cost | nonzero
    rc = jmb_reserve_dev(ctx, &ctx->d_stage, &ctx->d_stage_cap, n * (2 * m + 34 + 2) * 4); if (rc) return rc;
    int *base = (int *)ctx->d_stage;
    d_c = base; d_lv = base + n * m; d_rn = d_lv + n * 17; d_fa = fadjust ? d_rn + n * 17 : nullptr; d_cc = coeff_cost ? d_rn + n * 17 + n * m : nullptr; d_nz = d_rn + n * 17 + n * m + n;
    JMB_CUDA(ctx, cudaMemcpyAsync(d_c, coef, n * m * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (coeff_cost) JMB_CUDA(ctx, cudaMemcpyAsync(d_cc, coeff_cost, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    JMB_CUDA(ctx, cudaMemsetAsync(d_lv, 0, n * 34 * 4, ctx->stream));
  }
  jmb_time_begin(ctx, JMB_K_QUANT);
  k_quant_list<<<(nlist + 63) / 64, 64, 0, ctx->stream>>>((const jmb_qlist_desc *)ctx->d_qdesc, d_c, nlist, d_lv, d_rn, d_fa, d_cc, d_nz);
  jmb_time_end(ctx, JMB_K_QUANT);
  JMB_LAUNCH_CHECK(ctx);
  if (loc == JMB_HOST) {
    JMB_CUDA(ctx, cudaMemcpyAsync(coef, d_c, n * m * 4, cudaMemcpyDeviceToHost, ctx->stream));
    JMB_CUDA(ctx, cudaMemcpyAsync(levels, d_lv, n * 17 * 4, cudaMemcpyDeviceToHost, ctx->stream));
    JMB_CUDA(ctx, cudaMemcpyAsync(runs, d_rn, n * 17 * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (fadjust && q->around) JMB_CUDA(ctx, cudaMemcpyAsync(fadjust, d_fa, n * m * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (coeff_cost) JMB_CUDA(ctx, cudaMemcpyAsync(coeff_cost, d_cc, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    JMB_CUDA(ctx, cudaMemcpyAsync(nonzero, d_nz, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    JMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return JMB_OK;
}

int jmb_inverse_transform(jmb_ctx *ctx, int32_t *blocks, int nblk, int n, int loc) {
  if (n != 4 && n != 8) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_inverse_transform: n=%d", n);
  if (nblk <= 0) return JMB_OK;
  JMB_CUDA(ctx, cudaSetDevice(ctx->device));
  size_t bytes = (size_t)nblk * n * n * 4;
  int *d = blocks;
  if (loc == JMB_HOST) {
    int rc = jmb_reserve_dev(ctx, &ctx->d_stage, &ctx->d_stage_cap, bytes); if (rc) return rc;
    JMB_CUDA(ctx, cudaMemcpyAsync(ctx->d_stage, blocks, bytes, cudaMemcpyHostToDevice, ctx->stream));
    d = (int *)ctx->d_stage;
  }
  jmb_time_begin(ctx, JMB_K_FORWARD);
  if (n == 4) k_inverse<4><<<(nblk + 127) / 128, 128, 0, ctx->stream>>>(d, nblk);
  else k_inverse<8><<<(nblk + 63) / 64, 64, 0, ctx->stream>>>(d, nblk);
  jmb_time_end(ctx, JMB_K_FORWARD);
  JMB_LAUNCH_CHECK(ctx);
  if (loc == JMB_HOST) {
    JMB_CUDA(ctx, cudaMemcpyAsync(blocks, d, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    JMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return JMB_OK;
}

int jmb_luma_residual_coding_modes(jmb_ctx *ctx, const jmb_me_res *res, int n_mb, unsigned mode_mask, const jmb_quant_desc *q,
                                   int16_t *levels, int32_t *cost8, uint32_t *cbp_blk, uint32_t *cbp, uint8_t *recon, int32_t *sse, int loc) {
  int rc = check_qdesc(ctx, q); if (rc) return rc;
  if (!ctx->cur || ctx->nref == 0) return jmb_fail(ctx, JMB_ERR_STATE, "jmb_luma_residual_coding_modes: call jmb_pic_begin first");
  const int mb_w = ctx->cur_w / 16, mb_total = mb_w * (ctx->cur_h / 16);
  if (n_mb <= 0 || n_mb > mb_total) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_luma_residual_coding_modes: n_mb %d (picture has %d)", n_mb, mb_total);
  if (!mode_mask || (mode_mask >> 7)) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_luma_residual_coding_modes: mode_mask 0x%x (bits 0..6 = modes 1..7)", mode_mask);
  if (q->n == 8 && (mode_mask >> 4)) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_luma_residual_coding_modes: the 8x8 transform applies to modes 1..4 only");
  if (q->around) return jmb_fail(ctx, JMB_ERR_UNSUPPORTED, "jmb_luma_residual_coding_modes: adaptive rounding carries state from macroblock to macroblock (q_around.c); use the leaf form");
  if (!levels || !cost8 || !cbp_blk || !cbp || !sse) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_luma_residual_coding_modes: NULL output");
  JMB_CUDA(ctx, cudaSetDevice(ctx->device));
  const jmb_ref &r0 = ctx->refs[ctx->ref_list[0]];
  const jmb_me_res *d_res = res;
  if (!res) {
    if (!ctx->last_res || ctx->last_res_n < n_mb * 41) return jmb_fail(ctx, JMB_ERR_STATE, "jmb_luma_residual_coding_modes: no resident search results for %d macroblocks", n_mb);
    d_res = ctx->last_res;
  } else if (loc == JMB_HOST) {
    rc = jmb_reserve_dev(ctx, &ctx->d_stage4, &ctx->d_stage4_cap, (size_t)n_mb * 41 * sizeof(jmb_me_res)); if (rc) return rc;
    JMB_CUDA(ctx, cudaMemcpyAsync(ctx->d_stage4, res, (size_t)n_mb * 41 * sizeof(jmb_me_res), cudaMemcpyHostToDevice, ctx->stream));
    d_res = (const jmb_me_res *)ctx->d_stage4;
  }
  const jmb_quant_desc *d_q; rc = upload_qdesc(ctx, q, &d_q); if (rc) return rc;
  const size_t n7 = (size_t)7 * n_mb;
  int16_t *d_lv = levels; int *d_c8 = cost8; unsigned *d_cb = cbp_blk, *d_cbp = cbp; uint8_t *d_rec = recon; int *d_sse = sse;
  if (loc == JMB_HOST) {     // one arena: levels | cost8 | cbp_blk | cbp | sse | recon
    rc = jmb_reserve_dev(ctx, &ctx->d_stage3, &ctx->d_stage3_cap, n7 * (512 + 16 + 4 + 4 + 4 + 256)); if (rc) return rc;
    char *a = (char *)ctx->d_stage3;
    d_lv = (int16_t *)a; d_c8 = (int *)(a + n7 * 512); d_cb = (unsigned *)(a + n7 * 528); d_cbp = (unsigned *)(a + n7 * 532);
    d_sse = (int *)(a + n7 * 536); d_rec = recon ? (uint8_t *)(a + n7 * 540) : nullptr;
  }
  for (int m = 0; m < 7; m++)      // modes outside the mask are not computed: every output of theirs reads as zero
    if (!((mode_mask >> m) & 1)) {
      const size_t o = (size_t)m * n_mb;
      JMB_CUDA(ctx, cudaMemsetAsync(d_lv + o * 256, 0, (size_t)n_mb * 512, ctx->stream));
      JMB_CUDA(ctx, cudaMemsetAsync(d_c8 + o * 4, 0, (size_t)n_mb * 16, ctx->stream));
      JMB_CUDA(ctx, cudaMemsetAsync(d_cb + o, 0, (size_t)n_mb * 4, ctx->stream));
      JMB_CUDA(ctx, cudaMemsetAsync(d_cbp + o, 0, (size_t)n_mb * 4, ctx->stream));
      JMB_CUDA(ctx, cudaMemsetAsync(d_sse + o, 0, (size_t)n_mb * 4, ctx->stream));
      if (d_rec) JMB_CUDA(ctx, cudaMemsetAsync(d_rec + o * 256, 0, (size_t)n_mb * 256, ctx->stream));
    }
  const uint8_t *const *d_tab; rc = upload_ref_table(ctx, &d_tab); if (rc) return rc;
  jmb_time_begin(ctx, JMB_K_MC_TQ);
#define JMB_LRC(NN, STD, GRID) k_luma_rc_modes<NN, STD><<<GRID, 128, 0, ctx->stream>>>(d_res, nullptr, 0, n_mb, mb_w, mode_mask, d_q, ctx->cur, \
        ctx->cur_pitch, d_tab, r0.plane_bytes, r0.pitch, ctx->cur_w, ctx->cur_h, ctx->nref, ctx->d_err, d_lv, d_c8, d_cb, d_cbp, d_rec, d_sse)
  {
    const int kind = std_scan_kind(q);
    const dim3 g4((n_mb * 16 + 127) / 128, 7), g8((n_mb * 4 + 127) / 128, 7);
    if (q->n == 4) { if (kind == 1) JMB_LRC(4, 1, g4); else JMB_LRC(4, 0, g4); }
    else if (kind == 1) JMB_LRC(8, 1, g8); else if (kind == 2) JMB_LRC(8, 2, g8); else JMB_LRC(8, 0, g8);
  }
#undef JMB_LRC
  jmb_time_end(ctx, JMB_K_MC_TQ);
  JMB_LAUNCH_CHECK(ctx);
  if (loc == JMB_HOST) {
    JMB_CUDA(ctx, cudaMemcpyAsync(levels, d_lv, n7 * 512, cudaMemcpyDeviceToHost, ctx->stream));
    JMB_CUDA(ctx, cudaMemcpyAsync(cost8, d_c8, n7 * 16, cudaMemcpyDeviceToHost, ctx->stream));
    JMB_CUDA(ctx, cudaMemcpyAsync(cbp_blk, d_cb, n7 * 4, cudaMemcpyDeviceToHost, ctx->stream));
    JMB_CUDA(ctx, cudaMemcpyAsync(cbp, d_cbp, n7 * 4, cudaMemcpyDeviceToHost, ctx->stream));
    JMB_CUDA(ctx, cudaMemcpyAsync(sse, d_sse, n7 * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (recon) JMB_CUDA(ctx, cudaMemcpyAsync(recon, d_rec, n7 * 256, cudaMemcpyDeviceToHost, ctx->stream));
    JMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return JMB_OK;
}

int jmb_luma_residual_coding(jmb_ctx *ctx, const jmb_mb_pred *pred, int first_mb, int n_mb, const jmb_quant_desc *q,
                             int16_t *levels, int32_t *cost8, uint32_t *cbp_blk, uint32_t *cbp, uint8_t *recon, int32_t *sse, int loc) {
  int rc = check_qdesc(ctx, q); if (rc) return rc;
  if (!ctx->cur || ctx->nref == 0) return jmb_fail(ctx, JMB_ERR_STATE, "jmb_luma_residual_coding: call jmb_pic_begin first");
  const int mb_w = ctx->cur_w / 16, mb_total = mb_w * (ctx->cur_h / 16);
  if (n_mb <= 0 || first_mb < 0 || first_mb + n_mb > mb_total) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_luma_residual_coding: macroblocks %d..%d (picture has %d)", first_mb, first_mb + n_mb - 1, mb_total);
  if (q->around) return jmb_fail(ctx, JMB_ERR_UNSUPPORTED, "jmb_luma_residual_coding: adaptive rounding carries state from macroblock to macroblock (q_around.c); use the leaf form");
  if (!pred || !levels || !cost8 || !cbp_blk || !cbp || !sse) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_luma_residual_coding: NULL argument");
  JMB_CUDA(ctx, cudaSetDevice(ctx->device));
  const jmb_ref &r0 = ctx->refs[ctx->ref_list[0]];
  const jmb_mb_pred *d_pred = pred;
  const size_t nn = (size_t)n_mb;
  int16_t *d_lv = levels; int *d_c8 = cost8; unsigned *d_cb = cbp_blk, *d_cbp = cbp; uint8_t *d_rec = recon; int *d_sse = sse;
  if (loc == JMB_HOST) {
    for (int i = 0; i < n_mb; i++)
      for (int k = 0; k < 4; k++)
        if (pred[i].b8mode[k] < 1 || pred[i].b8mode[k] > 7 || pred[i].ref[k] >= ctx->nref || (q->n == 8 && pred[i].b8mode[k] > 4))
          return jmb_fail(ctx, JMB_ERR_ARG, "jmb_luma_residual_coding: macroblock %d quadrant %d: mode %d ref %d", i, k, pred[i].b8mode[k], pred[i].ref[k]);
    rc = jmb_reserve_dev(ctx, &ctx->d_stage, &ctx->d_stage_cap, nn * sizeof(jmb_mb_pred)); if (rc) return rc;
    JMB_CUDA(ctx, cudaMemcpyAsync(ctx->d_stage, pred, nn * sizeof(jmb_mb_pred), cudaMemcpyHostToDevice, ctx->stream));
    d_pred = (const jmb_mb_pred *)ctx->d_stage;
    rc = jmb_reserve_dev(ctx, &ctx->d_stage3, &ctx->d_stage3_cap, nn * (512 + 16 + 4 + 4 + 4 + 256)); if (rc) return rc;
    char *a = (char *)ctx->d_stage3;
    d_lv = (int16_t *)a; d_c8 = (int *)(a + nn * 512); d_cb = (unsigned *)(a + nn * 528); d_cbp = (unsigned *)(a + nn * 532);
    d_sse = (int *)(a + nn * 536); d_rec = recon ? (uint8_t *)(a + nn * 540) : nullptr;
  }
  const jmb_quant_desc *d_q; rc = upload_qdesc(ctx, q, &d_q); if (rc) return rc;
  const uint8_t *const *d_tab; rc = upload_ref_table(ctx, &d_tab); if (rc) return rc;
  jmb_time_begin(ctx, JMB_K_MC_TQ);
#define JMB_LRC(NN, STD, GRID) k_luma_rc_modes<NN, STD><<<GRID, 128, 0, ctx->stream>>>(nullptr, d_pred, first_mb, n_mb, mb_w, 1u, d_q, ctx->cur, \
        ctx->cur_pitch, d_tab, r0.plane_bytes, r0.pitch, ctx->cur_w, ctx->cur_h, ctx->nref, ctx->d_err, d_lv, d_c8, d_cb, d_cbp, d_rec, d_sse)
  {
    const int kind = std_scan_kind(q);
    const dim3 g4((n_mb * 16 + 127) / 128, 1), g8((n_mb * 4 + 127) / 128, 1);
    if (q->n == 4) { if (kind == 1) JMB_LRC(4, 1, g4); else JMB_LRC(4, 0, g4); }
    else if (kind == 1) JMB_LRC(8, 1, g8); else if (kind == 2) JMB_LRC(8, 2, g8); else JMB_LRC(8, 0, g8);
  }
#undef JMB_LRC
  jmb_time_end(ctx, JMB_K_MC_TQ);
  JMB_LAUNCH_CHECK(ctx);
  if (loc == JMB_HOST) {
    JMB_CUDA(ctx, cudaMemcpyAsync(levels, d_lv, nn * 512, cudaMemcpyDeviceToHost, ctx->stream));
    JMB_CUDA(ctx, cudaMemcpyAsync(cost8, d_c8, nn * 16, cudaMemcpyDeviceToHost, ctx->stream));
    JMB_CUDA(ctx, cudaMemcpyAsync(cbp_blk, d_cb, nn * 4, cudaMemcpyDeviceToHost, ctx->stream));
    JMB_CUDA(ctx, cudaMemcpyAsync(cbp, d_cbp, nn * 4, cudaMemcpyDeviceToHost, ctx->stream));
    JMB_CUDA(ctx, cudaMemcpyAsync(sse, d_sse, nn * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (recon) JMB_CUDA(ctx, cudaMemcpyAsync(recon, d_rec, nn * 256, cudaMemcpyDeviceToHost, ctx->stream));
    JMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return JMB_OK;
}

int jmb_mc_tq_modes(jmb_ctx *ctx, const jmb_me_res *res, int n_mb, unsigned mode_mask, const jmb_quant_desc *q,
                    int16_t *levels, int32_t *coeff_cost, uint32_t *cbp_blk, int loc) {
  int rc = check_qdesc(ctx, q); if (rc) return rc;
  if (!ctx->cur || ctx->nref == 0) return jmb_fail(ctx, JMB_ERR_STATE, "jmb_mc_tq_modes: call jmb_pic_begin first");
  const int mb_w = ctx->cur_w / 16, mb_total = mb_w * (ctx->cur_h / 16);
  if (n_mb <= 0 || n_mb > mb_total) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_mc_tq_modes: n_mb %d (picture has %d)", n_mb, mb_total);
  if (!mode_mask || (mode_mask >> 7)) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_mc_tq_modes: mode_mask 0x%x (bits 0..6 = modes 1..7)", mode_mask);
  if (q->n == 8 && (mode_mask >> 4)) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_mc_tq_modes: the 8x8 transform applies to modes 1..4 only");
  if (!levels || !coeff_cost || !cbp_blk) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_mc_tq_modes: NULL output");
  JMB_CUDA(ctx, cudaSetDevice(ctx->device));
  const jmb_ref &r0 = ctx->refs[ctx->ref_list[0]];
  const jmb_me_res *d_res = res;
  if (!res) {
    if (!ctx->last_res || ctx->last_res_n < n_mb * 41) return jmb_fail(ctx, JMB_ERR_STATE, "jmb_mc_tq_modes: no resident search results for %d macroblocks", n_mb);
    d_res = ctx->last_res;
  } else if (loc == JMB_HOST) {
    rc = jmb_reserve_dev(ctx, &ctx->d_stage4, &ctx->d_stage4_cap, (size_t)n_mb * 41 * sizeof(jmb_me_res)); if (rc) return rc;
    JMB_CUDA(ctx, cudaMemcpyAsync(ctx->d_stage4, res, (size_t)n_mb * 41 * sizeof(jmb_me_res), cudaMemcpyHostToDevice, ctx->stream));
    d_res = (const jmb_me_res *)ctx->d_stage4;
  }
  const jmb_quant_desc *d_q; rc = upload_qdesc(ctx, q, &d_q); if (rc) return rc;
  const size_t n7 = (size_t)7 * n_mb;
  int16_t *d_lv = levels; int *d_cc = coeff_cost; unsigned *d_cbp = cbp_blk;
  if (loc == JMB_HOST) {
    rc = jmb_reserve_dev(ctx, &ctx->d_stage3, &ctx->d_stage3_cap, n7 * (512 + 16 + 4)); if (rc) return rc;
    d_lv = (int16_t *)ctx->d_stage3; d_cc = (int *)((char *)ctx->d_stage3 + n7 * 512); d_cbp = (unsigned *)(d_cc + n7 * 4);
  }
  JMB_CUDA(ctx, cudaMemsetAsync(d_cc, 0, n7 * 16, ctx->stream));
  JMB_CUDA(ctx, cudaMemsetAsync(d_cbp, 0, n7 * 4, ctx->stream));
  for (int m = 0; m < 7; m++)      // modes outside the mask are not computed: their levels read as zero, never as stale memory
    if (!((mode_mask >> m) & 1)) JMB_CUDA(ctx, cudaMemsetAsync(d_lv + (size_t)m * n_mb * 256, 0, (size_t)n_mb * 512, ctx->stream));
  jmb_time_begin(ctx, JMB_K_MC_TQ);
#define JMB_MTQ(NN, STD, GRID, BLK) k_mc_tq_modes<NN, STD><<<GRID, BLK, 0, ctx->stream>>>(d_res, n_mb, mb_w, mode_mask, d_q, ctx->cur, ctx->cur_pitch, \
        r0.planes, r0.plane_bytes, r0.pitch, ctx->cur_w, ctx->cur_h, d_lv, d_cc, d_cbp)
  {
    const int kind = std_scan_kind(q);
    const dim3 g4((n_mb * 16 + 127) / 128, 7), g8((n_mb * 4 + 63) / 64, 7);
    if (q->n == 4) { if (kind == 1) JMB_MTQ(4, 1, g4, 128); else JMB_MTQ(4, 0, g4, 128); }
    else if (kind == 1) JMB_MTQ(8, 1, g8, 64); else if (kind == 2) JMB_MTQ(8, 2, g8, 64); else JMB_MTQ(8, 0, g8, 64);
  }
#undef JMB_MTQ
  jmb_time_end(ctx, JMB_K_MC_TQ);
  JMB_LAUNCH_CHECK(ctx);
  if (loc == JMB_HOST) {
    JMB_CUDA(ctx, cudaMemcpyAsync(levels, d_lv, n7 * 512, cudaMemcpyDeviceToHost, ctx->stream));
    JMB_CUDA(ctx, cudaMemcpyAsync(coeff_cost, d_cc, n7 * 16, cudaMemcpyDeviceToHost, ctx->stream));
    JMB_CUDA(ctx, cudaMemcpyAsync(cbp_blk, d_cbp, n7 * 4, cudaMemcpyDeviceToHost, ctx->stream));
    JMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return JMB_OK;
}

int jmb_mc_tq_modes_compact(jmb_ctx *ctx, const jmb_me_res *res, int n_mb, unsigned mode_mask, const jmb_quant_desc *q,
                            jmb_tq_head *heads, jmb_tq_token *tokens, uint32_t token_cap, uint32_t *n_tokens, int loc) {
  int rc = check_qdesc(ctx, q); if (rc) return rc;
  if (!ctx->cur || ctx->nref == 0) return jmb_fail(ctx, JMB_ERR_STATE, "jmb_mc_tq_modes_compact: call jmb_pic_begin first");
  const int mb_w = ctx->cur_w / 16, mb_total = mb_w * (ctx->cur_h / 16);
  if (n_mb <= 0 || n_mb > mb_total) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_mc_tq_modes_compact: n_mb %d (picture has %d)", n_mb, mb_total);
  if (!mode_mask || (mode_mask >> 7)) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_mc_tq_modes_compact: mode_mask 0x%x (bits 0..6 = modes 1..7)", mode_mask);
  if (q->n == 8 && (mode_mask >> 4)) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_mc_tq_modes_compact: the 8x8 transform applies to modes 1..4 only");
  if (q->around) return jmb_fail(ctx, JMB_ERR_UNSUPPORTED, "jmb_mc_tq_modes_compact: adaptive rounding is a leaf-form feature (jmb_quant_blocks)");
  if (!heads || !tokens || !n_tokens || !token_cap) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_mc_tq_modes_compact: NULL output");
  if (loc == JMB_HOST_ASYNC) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_mc_tq_modes_compact: the token count decides the second copy; use JMB_HOST or JMB_DEVICE");
  JMB_CUDA(ctx, cudaSetDevice(ctx->device));
  const jmb_ref &r0 = ctx->refs[ctx->ref_list[0]];
  const jmb_me_res *d_res = res;
  if (!res) {
    if (!ctx->last_res || ctx->last_res_n < n_mb * 41) return jmb_fail(ctx, JMB_ERR_STATE, "jmb_mc_tq_modes_compact: no resident search results for %d macroblocks", n_mb);
    d_res = ctx->last_res;
  } else if (loc == JMB_HOST) {
    rc = jmb_reserve_dev(ctx, &ctx->d_stage4, &ctx->d_stage4_cap, (size_t)n_mb * 41 * sizeof(jmb_me_res)); if (rc) return rc;
    JMB_CUDA(ctx, cudaMemcpyAsync(ctx->d_stage4, res, (size_t)n_mb * 41 * sizeof(jmb_me_res), cudaMemcpyHostToDevice, ctx->stream));
    d_res = (const jmb_me_res *)ctx->d_stage4;
  }
  const size_t n7 = (size_t)7 * n_mb;
  if (!ctx->d_tok_count) {
    JMB_CUDA(ctx, cudaMalloc(&ctx->d_tok_count, sizeof(unsigned)));
    JMB_CUDA(ctx, cudaHostAlloc(&ctx->h_tok_count, sizeof(unsigned), cudaHostAllocDefault));
  }
  jmb_tq_head *d_heads = heads; jmb_tq_token *d_tok = tokens; unsigned *d_cnt = ctx->d_tok_count;
  if (loc == JMB_HOST) {
    rc = jmb_reserve_dev(ctx, &ctx->d_heads, &ctx->d_heads_cap, n7 * sizeof(jmb_tq_head)); if (rc) return rc;
    rc = jmb_reserve_dev(ctx, &ctx->d_tokens, &ctx->d_tokens_cap, (size_t)token_cap * sizeof(jmb_tq_token)); if (rc) return rc;
    d_heads = (jmb_tq_head *)ctx->d_heads; d_tok = (jmb_tq_token *)ctx->d_tokens;
  }
  JMB_CUDA(ctx, cudaMemsetAsync(d_cnt, 0, sizeof(unsigned), ctx->stream));
  for (int m = 0; m < 7; m++)      // modes outside the mask: empty heads
    if (!((mode_mask >> m) & 1)) JMB_CUDA(ctx, cudaMemsetAsync(d_heads + (size_t)m * n_mb, 0, (size_t)n_mb * sizeof(jmb_tq_head), ctx->stream));
  jmb_time_begin(ctx, JMB_K_MC_TQ);
  // mb / mb_w inside the kernel is a multiply by ceil(2^32 / mb_w): exact while n_mb * mb_w < 2^32 (an 8K picture: 2^26)
  if ((unsigned long long)n_mb * (unsigned long long)mb_w >= (1ull << 32)) return jmb_fail(ctx, JMB_ERR_UNSUPPORTED, "jmb_mc_tq_modes_compact: %d macroblocks in rows of %d", n_mb, mb_w);
  const unsigned mb_w_rcp = (unsigned)(((1ull << 32) + (unsigned)mb_w - 1) / (unsigned)mb_w);
  const TqThr thr = tq_thresholds(q);
#define JMB_MTQC(NN, STD, GRID) k_mc_tq_modes_c<NN, STD><<<GRID, 128, 0, ctx->stream>>>(d_res, n_mb, mb_w, mb_w_rcp, mode_mask, *q, thr, ctx->cur, ctx->cur_pitch, \
        r0.planes, r0.plane_bytes, r0.pitch, ctx->cur_w, ctx->cur_h, d_heads, d_tok, token_cap, d_cnt)
  {
    const int kind = std_scan_kind(q);
    const dim3 g4((n_mb * 16 + 127) / 128, 7), g8((n_mb * 4 + 127) / 128, 7);
    if (q->n == 4) { if (kind == 1) JMB_MTQC(4, 1, g4); else JMB_MTQC(4, 0, g4); }
    else if (kind == 1) JMB_MTQC(8, 1, g8); else if (kind == 2) JMB_MTQC(8, 2, g8); else JMB_MTQC(8, 0, g8);
  }
#undef JMB_MTQC
  jmb_time_end(ctx, JMB_K_MC_TQ);
  JMB_LAUNCH_CHECK(ctx);
  if (loc == JMB_HOST) {
    JMB_CUDA(ctx, cudaMemcpyAsync(heads, d_heads, n7 * sizeof(jmb_tq_head), cudaMemcpyDeviceToHost, ctx->stream));
    JMB_CUDA(ctx, cudaMemcpyAsync(ctx->h_tok_count, d_cnt, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
    JMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *n_tokens = *ctx->h_tok_count;
    if (*n_tokens > token_cap) return jmb_fail(ctx, JMB_ERR_STATE, "jmb_mc_tq_modes_compact: %u tokens produced, room for %u", *n_tokens, token_cap);
    if (*n_tokens) {
      JMB_CUDA(ctx, cudaMemcpyAsync(tokens, d_tok, (size_t)*n_tokens * sizeof(jmb_tq_token), cudaMemcpyDeviceToHost, ctx->stream));
      JMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
  } else {
    JMB_CUDA(ctx, cudaMemcpyAsync(n_tokens, d_cnt, sizeof(unsigned), cudaMemcpyDeviceToDevice, ctx->stream));
  }
  return JMB_OK;
}

}  // extern "C"
