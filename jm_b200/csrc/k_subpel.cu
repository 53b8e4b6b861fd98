// k_subpel.cu -- K6: the 16 quarter-pel planes of a reference picture in ONE pass.
//
// Stands behind getSubImagesLuma (lencod/src/img_luma.c:611-680): integer plane with edge
// replication (:40-85), horizontal / vertical six-tap (20,-5,1) half-pel planes (:151-339), the
// centre half-pel plane from the UN-ROUNDED horizontal intermediates (:347-431) and the twelve
// bilinear quarter-pel planes (:440-596, source pairs :647-678).  JM makes 16 full passes over
// the frame through DRAM; here one CTA stages a (16+6) x (128+8) tile of the source in shared
// memory as 4-sample words, builds the un-rounded horizontal intermediates there (two samples per
// register), and emits all 16 planes of its 16 x 128 output tile with 128-byte row stores, four
// vertically adjacent rows per thread so that the staged words and their unpacking are shared.  Every tap index is clamped to the PADDED extent
// exactly as JM's edge branches do (img_luma.c:170-237, :272-331, :462-590).
//
// HBM traffic per reference: read w*h u16 once, write 16 * pitch * (h+40) bytes (u8 samples).
#include "jmb_internal.h"

namespace {

constexpr int TW = 128;  // output tile width  (32 lanes x 4 samples)
#ifndef JMB_SP_RPT
#define JMB_SP_RPT 4      // (warps per CTA, rows per thread) on the B200, 1080p / 4K: (8,1) 16.3 / 39.0 us, (8,2) 16.4 / 36.9, (4,2) 16.3 / 36.8, (4,4) 14.5 / 35.2
#endif
#ifndef JMB_SP_NW
#define JMB_SP_NW 4
#endif
constexpr int RPT = JMB_SP_RPT;  // rows per thread in the output stage
constexpr int SPT = 32 * JMB_SP_NW;   // threads per CTA
constexpr int TH = JMB_SP_NW * RPT;   // output tile height (RPT rows per warp)
constexpr int GWW = TW / 4 + 2;  // staged words per row: columns x0-4 .. x0+TW+3
constexpr int GH = TH + 6;       // staged rows    y0-2 .. y0+TH+3

// Two samples ride in one register as 16-bit fields; a word of four samples c0..c3 splits into its even columns (c0, c2) and its
// odd columns (c1, c3).  The six-tap sums are kept non-negative by a bias so that plain 32-bit adds never borrow between fields:
// 20*(c+d) - 5*(b+e) + (a+f) lies in [-2550, 10710]; + T_BIAS (a multiple of 32) it lies in [10, 13270].
constexpr uint32_t T_BIAS = 2560, T_BIAS2 = T_BIAS | (T_BIAS << 16);
__device__ __forceinline__ uint32_t even2(uint32_t v) { return v & 0x00ff00ffu; }
__device__ __forceinline__ uint32_t odd2(uint32_t v) { return __byte_perm(v, 0, 0x4341); }
// ONE_FOURTH_TAP {20,-5,1} on field pairs, biased: 20*(c+d) - 5*(b+e) + (a+f) + T_BIAS per field
__device__ __forceinline__ uint32_t tap6_biased(uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e, uint32_t f) {
  return (c + d) * 20u + (a + f + T_BIAS2) - (b + e) * 5u;
}
// clip255((v + 16) >> 5) per field of a biased six-tap pair: (v + T_BIAS + 16) >> 5 = ((v + 16) >> 5) + T_BIAS / 32
__device__ __forceinline__ uint32_t round5_clip(uint32_t biased) {
  const uint32_t q = ((biased + 0x00100010u) >> 5) & 0x07ff07ffu;
  constexpr uint32_t UNB = (uint32_t)(65536 - T_BIAS / 32);
  return __viaddmin_s16x2_relu(q, UNB | (UNB << 16), 0x00ff00ffu);     // max(min(q - 80, 255), 0) per field
}
__device__ __forceinline__ uint32_t bytes_of(uint32_t ev, uint32_t od) { return ev | (od << 8); }
// (a+b+1)>>1 per byte: (a | b) - ((a ^ b) >> 1), the shifted bit of each byte's neighbour masked off first
__device__ __forceinline__ uint32_t avg4(uint32_t a, uint32_t b) { return (a | b) - (((a ^ b) & 0xfefefefeu) >> 1); }

// four consecutive samples of one source row at padded coordinates (px .. px+3, py), each clamped to the padded extent and then
// to the picture (img_luma.c:40-85 and the edge branches of :170-237)
template <typename SRC>
__device__ __forceinline__ uint32_t stage_word(const SRC *__restrict__ src, int src_stride, int w, int h, int W, int H, int px, int py) {
  const int sy = jmb_clip(0, h - 1, jmb_clip(0, H - 1, py) - JMB_PAD_Y);
  const SRC *row = src + (size_t)sy * src_stride;
  const int sx = px - JMB_PAD_X;
  if (px >= 0 && px + 3 < W && sx >= 0 && sx + 3 < w && (((uintptr_t)(row + sx)) & (4 * sizeof(SRC) - 1)) == 0) {
    if (sizeof(SRC) == 1) return *(const uint32_t *)(row + sx);
    const uint2 v = *(const uint2 *)(row + sx);
    return __byte_perm(v.x, v.y, 0x6420);
  }
  uint32_t o = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) o |= (uint32_t)(uint8_t)row[jmb_clip(0, w - 1, jmb_clip(0, W - 1, px + k) - JMB_PAD_X)] << (8 * k);
  return o;
}

template <typename SRC>      // uint16_t (JM's imgpel) or uint8_t samples; the source may be a peer GPU's memory mapped over NVLink
__global__ void __launch_bounds__(SPT)
k_subpel_planes(const SRC *__restrict__ src, int src_stride, int w, int h, int W, int H,
                uint8_t *__restrict__ planes, int pitch, size_t plane_bytes) {
  __shared__ uint32_t sG[GH][GWW];        // integer samples, word j = columns x0-4+4j ..
  __shared__ uint2 sT[GH][TW / 4];        // biased un-rounded horizontal six-tap: .x = even columns, .y = odd columns

  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;   // padded coordinates of the tile origin
  const int tid = threadIdx.x;

  // stage 1: integer samples; a tile whose halo lies inside the picture needs no clamp at all
  const int ix = x0 - 4 - JMB_PAD_X, iy = y0 - 2 - JMB_PAD_Y;   // picture coordinates of the staged origin
  const bool interior = ix >= 0 && ix + 4 * GWW <= w && iy >= 0 && iy + GH <= h &&
                        ((((uintptr_t)src) | ((size_t)src_stride * sizeof(SRC))) & (4 * sizeof(SRC) - 1)) == 0;
  if (interior) {
    const SRC *base = src + (size_t)iy * src_stride + ix;
    for (int i = tid; i < GH * GWW; i += SPT) {
      const int r = i / GWW, j = i - r * GWW;
      const SRC *q = base + (size_t)r * src_stride + 4 * j;
      if (sizeof(SRC) == 1) sG[r][j] = *(const uint32_t *)q;
      else { const uint2 v = *(const uint2 *)q; sG[r][j] = __byte_perm(v.x, v.y, 0x6420); }
    }
  } else {
    for (int i = tid; i < GH * GWW; i += SPT) {
      const int r = i / GWW, j = i - r * GWW;
      sG[r][j] = stage_word(src, src_stride, w, h, W, H, x0 - 4 + 4 * j, y0 - 2 + r);
    }
  }
  __syncthreads();
  // stage 2: horizontal six-tap intermediates of every staged row, four columns per thread
  for (int i = tid; i < GH * (TW / 4); i += SPT) {
    const int r = i >> 5, j = i & 31;
    const uint32_t wl = sG[r][j], w0 = sG[r][j + 1], wr = sG[r][j + 2];
    const uint32_t fl = __funnelshift_r(wl, w0, 16), fr = __funnelshift_r(w0, wr, 16);     // columns -2..1 and 2..5
    const uint32_t pm2 = even2(fl), pm1 = odd2(fl), p0 = even2(w0), p1 = odd2(w0), p2 = even2(fr), p3 = odd2(fr), p4 = even2(wr);
    sT[r][j] = make_uint2(tap6_biased(pm2, pm1, p0, p1, p2, p3), tap6_biased(pm1, p0, p1, p2, p3, p4));
  }
  __syncthreads();

  // stage 3: a thread emits RPT vertically adjacent rows of four samples: the rows share their staged words and the unpacking
  const int ty = tid >> 5, lane = tid & 31;
  const int ya = y0 + ty * RPT, xb = x0 + lane * 4;
  const int r = ty * RPT + 2;                           // staged row of ya; row +1 is cy(y+1), column +1 is cx(x+1): the staging clamps

  uint32_t ge[5 + RPT], go[5 + RPT]; uint2 t[5 + RPT];
#pragma unroll
  for (int k = 0; k < 5 + RPT; k++) { const uint32_t g = sG[r - 2 + k][lane + 1]; ge[k] = even2(g); go[k] = odd2(g); t[k] = sT[r - 2 + k][lane]; }
  uint32_t Bn = bytes_of(round5_clip(t[2].x), round5_clip(t[2].y));      // the horizontal half-pel row of ya; each row hands its lower one on
#pragma unroll
  for (int i = 0; i < RPT; i++) {
    const int y = ya + i;
    const uint32_t G = bytes_of(ge[2 + i], go[2 + i]), Gd = bytes_of(ge[3 + i], go[3 + i]), Gr = __funnelshift_r(G, sG[r + i][lane + 2], 8);
    // vertical six-tap of the integer samples
    const uint32_t Hh = bytes_of(round5_clip(tap6_biased(ge[i], ge[i + 1], ge[i + 2], ge[i + 3], ge[i + 4], ge[i + 5])),
                                 round5_clip(tap6_biased(go[i], go[i + 1], go[i + 2], go[i + 3], go[i + 4], go[i + 5])));
    uint32_t Hn = __shfl_down_sync(0xffffffffu, Hh, 1);
    if (lane == 31) {                                   // column x+4 belongs to the next tile: its tap from the staged halo word
      uint32_t n[6];
#pragma unroll
      for (int k = 0; k < 6; k++) n[k] = sG[r + i - 2 + k][lane + 2] & 0xffu;
      Hn = round5_clip(tap6_biased(n[0], n[1], n[2], n[3], n[4], n[5])) & 0xffu;
    }
    const uint32_t Hr = __funnelshift_r(Hh, Hn, 8);
    const uint32_t B = Bn, Bd = bytes_of(round5_clip(t[3 + i].x), round5_clip(t[3 + i].y));
    Bn = Bd;
    // centre plane: vertical six-tap of the un-rounded horizontal intermediates in 32 bits; the six biases add up to 32 * T_BIAS
    uint32_t J;
    {
      const uint32_t s1e = t[i].x + t[i + 5].x, s5e = t[i + 1].x + t[i + 4].x, s20e = t[i + 2].x + t[i + 3].x;      // fields <= 26540: no carry, positive as s16
      const uint32_t s1o = t[i].y + t[i + 5].y, s5o = t[i + 1].y + t[i + 4].y, s20o = t[i + 2].y + t[i + 3].y;
      constexpr int C0 = 512 - 32 * (int)T_BIAS;
      constexpr int LO1 = 0x0001, LO5 = 0x00fb, LO20 = 0x0014, HI1 = 0x0100, HI5 = 0xfb00, HI20 = 0x1400;      // byte pairs (coef, 0) / (0, coef)
      const int j0 = __dp2a_lo((int)s20e, LO20, __dp2a_lo((int)s5e, LO5, __dp2a_lo((int)s1e, LO1, C0))) >> 10;
      const int j2 = __dp2a_lo((int)s20e, HI20, __dp2a_lo((int)s5e, HI5, __dp2a_lo((int)s1e, HI1, C0))) >> 10;
      const int j1 = __dp2a_lo((int)s20o, LO20, __dp2a_lo((int)s5o, LO5, __dp2a_lo((int)s1o, LO1, C0))) >> 10;
      const int j3 = __dp2a_lo((int)s20o, HI20, __dp2a_lo((int)s5o, HI5, __dp2a_lo((int)s1o, HI1, C0))) >> 10;
      J = (uint32_t)jmb_clip(0, 255, j0) | ((uint32_t)jmb_clip(0, 255, j1) << 8) | ((uint32_t)jmb_clip(0, 255, j2) << 16) | ((uint32_t)jmb_clip(0, 255, j3) << 24);
    }
    if (y >= H || xb >= W) continue;
    // the 16 planes of one reference span well under 4 GB: 32-bit offsets from the plane base
    const uint32_t o = (uint32_t)y * (uint32_t)pitch + (uint32_t)xb, pb = (uint32_t)plane_bytes;
#define PUT(fy, fx, v) *(uint32_t *)(planes + (o + (uint32_t)((fy) * 4 + (fx)) * pb)) = (v)
    PUT(0, 0, G);             PUT(0, 2, B);             PUT(2, 0, Hh);            PUT(2, 2, J);
    PUT(0, 1, avg4(G, B));    PUT(1, 0, avg4(G, Hh));   PUT(1, 1, avg4(B, Hh));   PUT(1, 2, avg4(B, J));
    PUT(2, 1, avg4(Hh, J));   PUT(0, 3, avg4(B, Gr));   PUT(1, 3, avg4(B, Hr));   PUT(2, 3, avg4(J, Hr));
    PUT(3, 0, avg4(Hh, Gd));  PUT(3, 1, avg4(Hh, Bd));  PUT(3, 2, avg4(J, Bd));   PUT(3, 3, avg4(Bd, Hr));
#undef PUT
  }
}

}  // namespace

int jmb_launch_subpel(jmb_ctx *ctx, const void *d_src, int sample_bytes, int src_stride, jmb_ref *r) {
  dim3 grid((r->W + TW - 1) / TW, (r->H + TH - 1) / TH);
  if (16 * r->plane_bytes > 0xffffffffull) return jmb_fail(ctx, JMB_ERR_ARG, "reference of %d x %d: the 16 planes exceed the 4 GB the plane kernel addresses", r->w, r->h);
  jmb_time_begin(ctx, JMB_K_SUBPEL);
  if (sample_bytes == 2) k_subpel_planes<uint16_t><<<grid, SPT, 0, ctx->stream>>>((const uint16_t *)d_src, src_stride, r->w, r->h, r->W, r->H, r->planes, r->pitch, r->plane_bytes);
  else k_subpel_planes<uint8_t><<<grid, SPT, 0, ctx->stream>>>((const uint8_t *)d_src, src_stride, r->w, r->h, r->W, r->H, r->planes, r->pitch, r->plane_bytes);
  jmb_time_end(ctx, JMB_K_SUBPEL);
  JMB_LAUNCH_CHECK(ctx);
  return JMB_OK;
}
