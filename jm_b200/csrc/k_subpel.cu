// k_subpel.cu -- K6: the 16 quarter-pel planes of a reference picture in ONE pass.
//
// Stands behind getSubImagesLuma (lencod/src/img_luma.c:611-680): integer plane with edge
// replication (:40-85), horizontal / vertical six-tap (20,-5,1) half-pel planes (:151-339), the
// centre half-pel plane from the UN-ROUNDED horizontal intermediates (:347-431) and the twelve
// bilinear quarter-pel planes (:440-596, source pairs :647-678).  JM makes 16 full passes over
// the frame through DRAM; here one CTA stages a (8+6) x (128+5) tile of the source in shared
// memory, builds the un-rounded horizontal intermediates there, and emits all 16 planes of its
// 8 x 128 output tile with 128-byte row stores.  Every tap index is clamped to the PADDED extent
// exactly as JM's edge branches do (img_luma.c:170-237, :272-331, :462-590).
//
// HBM traffic per reference: read w*h u16 once, write 16 * pitch * (h+40) bytes (u8 samples).
#include "jmb_internal.h"

namespace {

constexpr int TW = 128;  // output tile width  (32 lanes x 4 samples)
constexpr int TH = 8;    // output tile height (one warp per row)
constexpr int GW = TW + 5 + 3;  // staged columns x0-2 .. x0+TW+2, padded to a multiple of 4
constexpr int GH = TH + 6;      // staged rows    y0-2 .. y0+TH+3

__device__ __forceinline__ int tap6(int a, int b, int c, int d, int e, int f) {
  // ONE_FOURTH_TAP {20,-5,1}: 20*(c+d) - 5*(b+e) + (a+f)
  return 20 * (c + d) - 5 * (b + e) + (a + f);
}
__device__ __forceinline__ int clip255(int v) { return min(max(v, 0), 255); }
__device__ __forceinline__ uint32_t avg4(uint32_t a, uint32_t b) { return __vavgu4(a, b); }  // (a+b+1)>>1 per byte

template <typename SRC>      // uint16_t (JM's imgpel) or uint8_t samples; the source may be a peer GPU's memory mapped over NVLink
__global__ void __launch_bounds__(256)
k_subpel_planes(const SRC *__restrict__ src, int src_stride, int w, int h, int W, int H,
                uint8_t *__restrict__ planes, int pitch, size_t plane_bytes) {
  __shared__ uint8_t sG[GH][GW];
  __shared__ int16_t sT[GH][TW];   // un-rounded horizontal six-tap, range [-2550, 10710]

  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;   // padded coordinates of the tile origin
  const int tid = threadIdx.x;

  // stage 1: integer samples at padded coords (clamped to the padded extent, then to the picture)
  for (int i = tid; i < GH * GW; i += 256) {
    int r = i / GW, c = i - r * GW;
    int py = jmb_clip(0, H - 1, y0 - 2 + r), px = jmb_clip(0, W - 1, x0 - 2 + c);
    int sy = jmb_clip(0, h - 1, py - JMB_PAD_Y), sx = jmb_clip(0, w - 1, px - JMB_PAD_X);
    sG[r][c] = (uint8_t)src[(size_t)sy * src_stride + sx];
  }
  __syncthreads();
  // stage 2: horizontal six-tap intermediates for every staged row
  for (int i = tid; i < GH * TW; i += 256) {
    int r = i / TW, c = i - r * TW;              // tile column c <-> staged column c+2
    const uint8_t *g = &sG[r][c];
    sT[r][c] = (int16_t)tap6(g[0], g[1], g[2], g[3], g[4], g[5]);
  }
  __syncthreads();

  const int ty = tid >> 5, lane = tid & 31;
  const int y = y0 + ty, xb = x0 + lane * 4;
  if (y >= H || xb >= W) return;
  const int r = ty + 2;                                 // staged row of y
  const int r1 = min(y + 1, H - 1) - (y0 - 2);          // staged row of cy(y+1)

  uint32_t G = 0, B = 0, Hh = 0, J = 0, Gr = 0, Hr = 0, Gd = 0, Bd = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int c = lane * 4 + k;                          // tile column
    const int gc = c + 2;                                // staged column of x
    const int gc1 = min(xb + k + 1, W - 1) - (x0 - 2);   // staged column of cx(x+1)
    int g = sG[r][gc];
    int b = clip255((sT[r][c] + 16) >> 5);
    int bd = clip255((sT[r1][c] + 16) >> 5);
    int hv = clip255((tap6(sG[r - 2][gc], sG[r - 1][gc], sG[r][gc], sG[r + 1][gc], sG[r + 2][gc], sG[r + 3][gc]) + 16) >> 5);
    int hr = clip255((tap6(sG[r - 2][gc1], sG[r - 1][gc1], sG[r][gc1], sG[r + 1][gc1], sG[r + 2][gc1], sG[r + 3][gc1]) + 16) >> 5);
    int j = clip255((tap6(sT[r - 2][c], sT[r - 1][c], sT[r][c], sT[r + 1][c], sT[r + 2][c], sT[r + 3][c]) + 512) >> 10);
    int gr = sG[r][gc1], gd = sG[r1][gc];
    const int sh = 8 * k;
    G |= (uint32_t)g << sh;   B |= (uint32_t)b << sh;   Hh |= (uint32_t)hv << sh; J |= (uint32_t)j << sh;
    Gr |= (uint32_t)gr << sh; Hr |= (uint32_t)hr << sh; Gd |= (uint32_t)gd << sh; Bd |= (uint32_t)bd << sh;
  }
  uint8_t *o = planes + (size_t)y * pitch + xb;
#define PUT(fy, fx, v) *(uint32_t *)(o + (size_t)((fy) * 4 + (fx)) * plane_bytes) = (v)
  PUT(0, 0, G);             PUT(0, 2, B);             PUT(2, 0, Hh);            PUT(2, 2, J);
  PUT(0, 1, avg4(G, B));    PUT(1, 0, avg4(G, Hh));   PUT(1, 1, avg4(B, Hh));   PUT(1, 2, avg4(B, J));
  PUT(2, 1, avg4(Hh, J));   PUT(0, 3, avg4(B, Gr));   PUT(1, 3, avg4(B, Hr));   PUT(2, 3, avg4(J, Hr));
  PUT(3, 0, avg4(Hh, Gd));  PUT(3, 1, avg4(Hh, Bd));  PUT(3, 2, avg4(J, Bd));   PUT(3, 3, avg4(Bd, Hr));
#undef PUT
}

}  // namespace

int jmb_launch_subpel(jmb_ctx *ctx, const void *d_src, int sample_bytes, int src_stride, jmb_ref *r) {
  dim3 grid((r->W + TW - 1) / TW, (r->H + TH - 1) / TH);
  jmb_time_begin(ctx, JMB_K_SUBPEL);
  if (sample_bytes == 2) k_subpel_planes<uint16_t><<<grid, 256, 0, ctx->stream>>>((const uint16_t *)d_src, src_stride, r->w, r->h, r->W, r->H, r->planes, r->pitch, r->plane_bytes);
  else k_subpel_planes<uint8_t><<<grid, 256, 0, ctx->stream>>>((const uint8_t *)d_src, src_stride, r->w, r->h, r->W, r->H, r->planes, r->pitch, r->plane_bytes);
  jmb_time_end(ctx, JMB_K_SUBPEL);
  JMB_LAUNCH_CHECK(ctx);
  return JMB_OK;
}
