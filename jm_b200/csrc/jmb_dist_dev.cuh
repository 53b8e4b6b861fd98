// jmb_dist_dev.cuh -- device helpers shared by the distortion kernels (k_refine.cu, k_epzs.cu): JM's UMVLine4X addressing
// of the 16 quarter-pel planes, unaligned 4/8-sample loads, SAD / SSE / Hadamard SAD of 4x4 and 8x8 sub-blocks.
#pragma once
#include "jmb_internal.h"

namespace {


__constant__ signed char c_spiral9[9][2] = {{0,0},{0,-1},{0,1},{-1,-1},{1,-1},{-1,0},{1,0},{-1,1},{1,1}};
__constant__ unsigned char c_bsx[8] = {16, 16, 16, 8, 8, 8, 4, 4};
__constant__ unsigned char c_bsy[8] = {16, 16, 8, 16, 8, 4, 8, 4};

struct RefView { const uint8_t *planes; size_t plane_bytes; int pitch, w, h; };

// pointer to the sample at quarter-pel position (qx,qy) after UMVLine4X's origin clamp; the 16 planes of a reference span less
// than 4 GB (jmb_launch_subpel refuses more), so the offset is formed in 32 bits
__device__ __forceinline__ const uint8_t *umv(const RefView &rv, int qy, int qx) {
  const int iy = jmb_clip(-JMB_PAD_Y, rv.h + JMB_PAD_Y - 1 - 16, qy >> 2);
  const int ix = jmb_clip(-JMB_PAD_X, rv.w + JMB_PAD_X - 1 - 16, qx >> 2);
  const unsigned off = (unsigned)((qy & 3) * 4 + (qx & 3)) * (unsigned)rv.plane_bytes + (unsigned)((iy + JMB_PAD_Y) * rv.pitch + ix + JMB_PAD_X);
  return rv.planes + off;
}

__device__ __forceinline__ int hadamard4(const int *d) {   // d[16] row-major
  int m[16], s = 0;
#pragma unroll
  for (int c = 0; c < 4; c++) {
    int a0 = d[c] + d[12 + c], a1 = d[4 + c] + d[8 + c], a2 = d[4 + c] - d[8 + c], a3 = d[c] - d[12 + c];
    m[c] = a0 + a1; m[8 + c] = a0 - a1; m[4 + c] = a3 + a2; m[12 + c] = a3 - a2;
  }
#pragma unroll
  for (int r = 0; r < 4; r++) {
    int a0 = m[4 * r] + m[4 * r + 3], a1 = m[4 * r + 1] + m[4 * r + 2], a2 = m[4 * r + 1] - m[4 * r + 2], a3 = m[4 * r] - m[4 * r + 3];
    s += abs(a0 + a1) + abs(a0 - a1) + abs(a2 + a3) + abs(a3 - a2);
  }
  return (s + 1) >> 1;
}

__device__ int hadamard8(int *a) {   // a[64] row-major, destroyed
#pragma unroll
  for (int pass = 0; pass < 2; pass++) {
    const int step = pass ? 8 : 1, line = pass ? 1 : 8;
#pragma unroll
    for (int l = 0; l < 8; l++) {
      int *p = a + l * line;
#pragma unroll
      for (int len = 4; len >= 1; len >>= 1)
#pragma unroll
        for (int b = 0; b < 8; b += 2 * len)
#pragma unroll
          for (int k = 0; k < len; k++) {
            int u = p[(b + k) * step], v = p[(b + k + len) * step];
            p[(b + k) * step] = u + v; p[(b + k + len) * step] = u - v;
          }
    }
  }
  int s = 0;
#pragma unroll
  for (int i = 0; i < 64; i++) s += abs(a[i]);
  return (s + 2) >> 2;
}

// four (or eight) consecutive samples starting at an arbitrary byte address: aligned word loads + PRMT
__device__ __forceinline__ unsigned ld4(const uint8_t *p) {
  const unsigned sh = (unsigned)(size_t)p & 3u;
  const unsigned *a = (const unsigned *)(p - sh);
  return __byte_perm(__ldg(a), __ldg(a + 1), 0x3210u + 0x1111u * sh);
}
__device__ __forceinline__ void ld8(const uint8_t *p, unsigned &lo, unsigned &hi) {
  const unsigned sh = (unsigned)(size_t)p & 3u, sel = 0x3210u + 0x1111u * sh;
  const unsigned *a = (const unsigned *)(p - sh);
  const unsigned w0 = __ldg(a), w1 = __ldg(a + 1), w2 = __ldg(a + 2);
  lo = __byte_perm(w0, w1, sel); hi = __byte_perm(w1, w2, sel);
}

// The rows of a block inside a reference plane: the pitch is a multiple of 4 (jmb_ref_put rounds it up to 128), so every row
// has the word alignment of the first one -- one selector, one word stride.
struct RowsAt { const unsigned *a; unsigned sel; int pw; };
__device__ __forceinline__ RowsAt rows_at(const uint8_t *p, int pitch) {
  const unsigned sh = (unsigned)(size_t)p & 3u;
  return RowsAt{(const unsigned *)(p - sh), 0x3210u + 0x1111u * sh, pitch >> 2};
}
__device__ __forceinline__ unsigned row4(const RowsAt &R, int y) {
  const unsigned *q = R.a + y * R.pw;
  return __byte_perm(__ldg(q), __ldg(q + 1), R.sel);
}
__device__ __forceinline__ void row8(const RowsAt &R, int y, unsigned &lo, unsigned &hi) {
  const unsigned *q = R.a + y * R.pw;
  const unsigned w0 = __ldg(q), w1 = __ldg(q + 1), w2 = __ldg(q + 2);
  lo = __byte_perm(w0, w1, R.sel); hi = __byte_perm(w1, w2, R.sel);
}

// One sub-block of the source, kept in registers while the candidates of a refinement stage go by.
struct SrcBlk { unsigned w[16]; };   // n = 4: w[0..3] = rows; n = 8: w[2y], w[2y+1] = row y

__device__ __forceinline__ void load_src(SrcBlk &s, const uint8_t *cur, int cur_pitch, int x, int y, int n) {
  const uint8_t *p = cur + (size_t)y * cur_pitch + x;      // x is a multiple of 4: aligned
  if (n == 4) {
#pragma unroll
    for (int r = 0; r < 4; r++) s.w[r] = *(const unsigned *)(p + (size_t)r * cur_pitch);
  } else {
#pragma unroll
    for (int r = 0; r < 8; r++) { const unsigned *v = (const unsigned *)(p + (size_t)r * cur_pitch); s.w[2 * r] = v[0]; s.w[2 * r + 1] = v[1]; }
  }
}

// HadamardSAD8x8 (me_distortion.c:266-347) of an 8x8 sub-block in registers, two samples per register: a row's eight
// differences are four integers a + 65536 b (no field leaves +-16320 / 2: the last butterfly stage is never formed).  Two
// horizontal stages act on whole registers, three vertical ones on the rows, and the third horizontal stage is folded into
// |u + v| + |u - v| = 2 max(|u|, |v|): the maxima add up to half the coefficient sum S, and JM's (S + 2) >> 2 is (M + 1) >> 1.
__device__ __forceinline__ int hadamard8_packed(const SrcBlk &src, const uint8_t *ref, int pitch) {
  int p[8][4];
  const RowsAt R = rows_at(ref, pitch);
#pragma unroll
  for (int y = 0; y < 8; y++) {
    unsigned lo, hi;
    row8(R, y, lo, hi);
    const int d0 = (int)__byte_perm(src.w[2 * y], 0, 0x4140) - (int)__byte_perm(lo, 0, 0x4140), d1 = (int)__byte_perm(src.w[2 * y], 0, 0x4342) - (int)__byte_perm(lo, 0, 0x4342);
    const int d2 = (int)__byte_perm(src.w[2 * y + 1], 0, 0x4140) - (int)__byte_perm(hi, 0, 0x4140), d3 = (int)__byte_perm(src.w[2 * y + 1], 0, 0x4342) - (int)__byte_perm(hi, 0, 0x4342);
    const int a0 = d0 + d2, a1 = d1 + d3, a2 = d0 - d2, a3 = d1 - d3;
    p[y][0] = a0 + a1; p[y][1] = a0 - a1; p[y][2] = a2 + a3; p[y][3] = a2 - a3;
  }
  int s = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int b0 = p[0][k] + p[4][k], b1 = p[1][k] + p[5][k], b2 = p[2][k] + p[6][k], b3 = p[3][k] + p[7][k];
    const int b4 = p[0][k] - p[4][k], b5 = p[1][k] - p[5][k], b6 = p[2][k] - p[6][k], b7 = p[3][k] - p[7][k];
    const int c0 = b0 + b2, c1 = b1 + b3, c2 = b0 - b2, c3 = b1 - b3, c4 = b4 + b6, c5 = b5 + b7, c6 = b4 - b6, c7 = b5 - b7;
    const int e[8] = {c0 + c1, c0 - c1, c2 + c3, c2 - c3, c4 + c5, c4 - c5, c6 + c7, c6 - c7};
#pragma unroll
    for (int i = 0; i < 8; i++) { const int l = (int)(short)(e[i] & 0xffff), h = (e[i] - l) >> 16; s += max(abs(l), abs(h)); }
  }
  return (s + 1) >> 1;
}

// distortion contribution of sub-block (sbx, sby) [units of n pels] of a block at (pos_x,pos_y)
// against the candidate at absolute quarter-pel (cqx, cqy)
__device__ __forceinline__ int subblock_dist(const RefView &rv, const SrcBlk &src, int cqx, int cqy, int sbx, int sby, int n, int metric) {
  const uint8_t *ref;
  if (metric == JMB_SATD) ref = umv(rv, cqy + ((sby * n) << 2), cqx + ((sbx * n) << 2));   // per-sub-block clamp
  else ref = umv(rv, cqy, cqx) + (size_t)(sby * n) * rv.pitch + sbx * n;                     // partition clamp
  if (n == 4) {
    int d[16];
    const RowsAt R = rows_at(ref, rv.pitch);
#pragma unroll
    for (int y = 0; y < 4; y++) {
      const unsigned sv = src.w[y], rw = row4(R, y);
#pragma unroll
      for (int x = 0; x < 4; x++) d[y * 4 + x] = (int)((sv >> (8 * x)) & 255) - (int)((rw >> (8 * x)) & 255);
    }
    if (metric == JMB_SATD) return hadamard4(d);
    int s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += (metric == JMB_SAD) ? abs(d[i]) : d[i] * d[i];
    return s;
  }
  return hadamard8_packed(src, ref, rv.pitch);
}

// 4x4 sub-block, split in two so that the loads of several candidates can be in flight before the first Hadamard
__device__ __forceinline__ void load_ref4(const RefView &rv, int cqx, int cqy, int sbx, int sby, int metric, unsigned (&rw)[4]) {
  const uint8_t *ref;
  if (metric == JMB_SATD) ref = umv(rv, cqy + ((sby * 4) << 2), cqx + ((sbx * 4) << 2));   // per-sub-block clamp
  else ref = umv(rv, cqy, cqx) + (size_t)(sby * 4) * rv.pitch + sbx * 4;                     // partition clamp
  const RowsAt R = rows_at(ref, rv.pitch);
#pragma unroll
  for (int y = 0; y < 4; y++) rw[y] = row4(R, y);
}
// HadamardSAD4x4 (me_distortion.c:175-258) on two samples per register: a pair (a, b) is carried as the INTEGER
// a + 65536 * b, on which adds and subtracts act on both halves at once (no field ever overflows: |values| <= 4080).
// Vertical butterflies on the packed rows, one horizontal stage after swapping the halves of the right pair, and the
// last stage folded into |u + v| + |u - v| = 2 max(|u|, |v|), so the result is the sum of the eight maxima -- exactly
// JM's (sum |coefficient| + 1) >> 1, the sum being even.
__device__ __forceinline__ int hadamard4_packed(const unsigned (&sw)[4], const unsigned (&rw)[4]) {
  int lo[4], hi[4];
#pragma unroll
  for (int y = 0; y < 4; y++) {      // bytes 0,1 -> (b0, b1), bytes 2,3 -> (b2, b3) as 16-bit fields; difference as integers
    lo[y] = (int)__byte_perm(sw[y], 0, 0x4140) - (int)__byte_perm(rw[y], 0, 0x4140);
    hi[y] = (int)__byte_perm(sw[y], 0, 0x4342) - (int)__byte_perm(rw[y], 0, 0x4342);
  }
  int s = 0;
  int ml[4], mh[4];
  { const int a0 = lo[0] + lo[3], a1 = lo[1] + lo[2], a2 = lo[1] - lo[2], a3 = lo[0] - lo[3];
    ml[0] = a0 + a1; ml[2] = a0 - a1; ml[1] = a3 + a2; ml[3] = a3 - a2; }
  { const int a0 = hi[0] + hi[3], a1 = hi[1] + hi[2], a2 = hi[1] - hi[2], a3 = hi[0] - hi[3];
    mh[0] = a0 + a1; mh[2] = a0 - a1; mh[1] = a3 + a2; mh[3] = a3 - a2; }
#pragma unroll
  for (int r = 0; r < 4; r++) {
    // row r holds (m0, m1) in ml and (m2, m3) in mh; swapping mh's halves needs the integer form re-split first
    const int h0 = (int)(short)(mh[r] & 0xffff), h1 = (mh[r] - h0) >> 16;          // m2, m3
    const int l0 = (int)(short)(ml[r] & 0xffff), l1 = (ml[r] - l0) >> 16;          // m0, m1
    const int a0 = l0 + h1, a1 = l1 + h0, a2 = l1 - h0, a3 = l0 - h1;
    s += max(abs(a0), abs(a1)) + max(abs(a2), abs(a3));
  }
  return s;
}

__device__ __forceinline__ int dist4(const SrcBlk &src, const unsigned (&rw)[4], int metric) {
  if (metric == JMB_SATD) { const unsigned sw[4] = {src.w[0], src.w[1], src.w[2], src.w[3]}; return hadamard4_packed(sw, rw); }
  int d[16];
#pragma unroll
  for (int y = 0; y < 4; y++)
#pragma unroll
    for (int x = 0; x < 4; x++) d[y * 4 + x] = (int)((src.w[y] >> (8 * x)) & 255) - (int)((rw[y] >> (8 * x)) & 255);
  if (metric == JMB_SATD) return hadamard4(d);
  int s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += (metric == JMB_SAD) ? abs(d[i]) : d[i] * d[i];
  return s;
}


}  // namespace
