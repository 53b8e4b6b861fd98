// k_chroma.cu -- chroma of inter macroblocks: motion-compensated prediction and residual coding on the device.
//
//  prediction       OneComponentChromaPrediction4x4_regenerate (lencod/src/mc_prediction.c:292-352), called per 4x4 chroma block
//                   from chroma_prediction_4x4 / chroma_residual_coding (lencod/src/macroblock.c:1439-1500)
//  residual coding  residual_transform_quant_chroma_4x4 (lencod/src/block.c:954-1202) with hadamard2x2 / hadamard4x2
//                   (lcommon/src/transform.c:206-330), quant_dc2x2_normal / quant_dc4x2_normal (lencod/src/quantChroma_normal.c),
//                   quant_ac4x4_normal (lencod/src/quant4x4_normal.c:117), inverse4x4, sample_reconstruct
//
// One thread per 4x4 chroma block; the 4 (4:2:0) or 8 (4:2:2) threads of a component are adjacent lanes and exchange the DC
// coefficients, the running coefficient cost and the coded-block flags by warp shuffles; the first lane of the group runs
// the DC Hadamard + quantiser.  Streaming kernel: per macroblock 2 x (64 | 128) source + reference samples in, levels out.
#include "jmb_internal.h"

namespace {

__device__ __forceinline__ void c_fwd4(int *b) {      // forward4x4, lcommon/src/transform.c:20-68
  int t[16];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int *p = b + 4 * i;
    const int t0 = p[0] + p[3], t1 = p[1] + p[2], t2 = p[1] - p[2], t3 = p[0] - p[3];
    t[4 * i] = t0 + t1; t[4 * i + 1] = (t3 << 1) + t2; t[4 * i + 2] = t0 - t1; t[4 * i + 3] = t3 - (t2 << 1);
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int t0 = t[i] + t[12 + i], t1 = t[4 + i] + t[8 + i], t2 = t[4 + i] - t[8 + i], t3 = t[i] - t[12 + i];
    b[i] = t0 + t1; b[4 + i] = t2 + (t3 << 1); b[8 + i] = t0 - t1; b[12 + i] = t3 - (t2 << 1);
  }
}
__device__ __forceinline__ void c_inv4(int *b) {      // inverse4x4, lcommon/src/transform.c:70-119
  int t[16];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int *p = b + 4 * i;
    const int p0 = p[0] + p[2], p1 = p[0] - p[2], p2 = (p[1] >> 1) - p[3], p3 = p[1] + (p[3] >> 1);
    t[4 * i] = p0 + p3; t[4 * i + 1] = p1 + p2; t[4 * i + 2] = p1 - p2; t[4 * i + 3] = p0 - p3;
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int p0 = t[i] + t[8 + i], p1 = t[i] - t[8 + i], p2 = (t[4 + i] >> 1) - t[12 + i], p3 = t[4 + i] + (t[12 + i] >> 1);
    b[i] = p0 + p3; b[4 + i] = p1 + p2; b[8 + i] = p1 - p2; b[12 + i] = p0 - p3;
  }
}

__device__ constexpr unsigned char ZZ4[16][2] = {{0,0},{1,0},{0,1},{0,2},{1,1},{2,0},{3,0},{2,1},{1,2},{0,3},{1,3},{2,2},{3,1},{3,2},{2,3},{3,3}};
__constant__ int c_cmode_base[8] = {0, 0, 1, 3, 5, 9, 17, 25}, c_cmode_w4[8] = {4, 4, 4, 2, 2, 2, 1, 1}, c_cmode_h4[8] = {4, 4, 2, 4, 2, 1, 2, 1};

// one DC coefficient through quant_dc2x2_normal / quant_dc4x2_normal (doubled offset, one more bit of shift; dequantised value
// (level * InvScale) << qp_per): returns the level, *c becomes the dequantised coefficient
__device__ __forceinline__ int quant_dc1(int *c, const int *prm, int q_bits, int qp_per, bool clip) {
  const int v = *c;
  if (v == 0) return 0;
  int level = (abs(v) * prm[1] + (prm[0] << 1)) >> (q_bits + 1);
  if (level == 0) { *c = 0; return 0; }
  if (clip) level = min(level, 2063);
  if (v < 0) level = -level;
  *c = (level * prm[2]) << qp_per;
  return level;
}

template <int YUV>      // 1: 4:2:0 (2 x 2 blocks per component), 2: 4:2:2 (2 wide x 4 high)
__global__ void __launch_bounds__(128)
k_chroma_rc(const jmb_me_res *__restrict__ res, const jmb_mb_pred *__restrict__ pred, int mode, int first_mb, int n_mb, int mb_w,
            const __grid_constant__ jmb_chroma_desc d, const uint8_t *__restrict__ cur_c, int cur_pitch_c, size_t cur_plane,
            const uint8_t *const *__restrict__ ref_c, int ref_pitch_c, size_t ref_plane, int wc, int hc, int nref, int *__restrict__ err,
            int16_t *__restrict__ dc_levels, int16_t *__restrict__ ac_levels, unsigned *__restrict__ cbp_blk, unsigned *__restrict__ cr_cbp,
            uint8_t *__restrict__ recon) {
  constexpr int NB = (YUV == 1) ? 4 : 8, HMB = (YUV == 1) ? 8 : 16, F1Y = 64 / HMB, YDIV = HMB >> 2;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = t < n_mb * 2 * NB;
  const int item = live ? t / NB : 0, b = t % NB, mb = item >> 1, uv = item & 1;
  const int lane = threadIdx.x & 31, leader = lane & ~(NB - 1);
  const int bx = b & 1, by = b >> 1;                               // 4x4 block inside the component
  const int addr = first_mb + mb, mb_cx = (addr % mb_w) * 8, mb_cy = (addr / mb_w) * HMB;
  // ---- prediction + residual ----
  int rr[16], pr[16];
  int rf = 0;
  bool bad = false;
#pragma unroll
  for (int y = 0; y < 4; y++)
#pragma unroll
    for (int x = 0; x < 4; x++) {
      const int i = bx * 4 + x, j = by * 4 + y;
      const int lbx = i >> 1, lby = j / YDIV;                     // the luma 4x4 block above this sample
      int mvx, mvy;
      if (pred) {
        const jmb_mb_pred *mp = pred + mb;
        const int b8 = (lby >> 1) * 2 + (lbx >> 1);
        mvx = mp->mv[lby * 4 + lbx][0]; mvy = mp->mv[lby * 4 + lbx][1]; rf = mp->ref[b8];
        if (mp->b8mode[b8] < 1 || mp->b8mode[b8] > 7 || rf >= nref) { bad = true; rf = 0; }
      } else {
        const jmb_me_res r = res[addr * 41 + c_cmode_base[mode] + (lby / c_cmode_h4[mode]) * (4 / c_cmode_w4[mode]) + lbx / c_cmode_w4[mode]];
        mvx = r.mv_x; mvy = r.mv_y;
      }
      const uint8_t *rp = ref_c[rf] + (size_t)uv * ref_plane;
      const int ii = (i + mb_cx) * 8 + mvx, jj = (j + mb_cy) * F1Y + mvy;
      const int ii0 = jmb_clip(0, wc - 1, ii / 8), jj0 = jmb_clip(0, hc - 1, jj / F1Y);
      const int ii1 = jmb_clip(0, wc - 1, (ii + 7) / 8), jj1 = jmb_clip(0, hc - 1, (jj + F1Y - 1) / F1Y);
      const int if1 = ii & 7, if0 = 8 - if1, jf1 = jj & (F1Y - 1), jf0 = F1Y - jf1;
      const int p = (if0 * jf0 * rp[(size_t)jj0 * ref_pitch_c + ii0] + if1 * jf0 * rp[(size_t)jj0 * ref_pitch_c + ii1] +
                     if0 * jf1 * rp[(size_t)jj1 * ref_pitch_c + ii0] + if1 * jf1 * rp[(size_t)jj1 * ref_pitch_c + ii1] + (8 * F1Y >> 1)) / (8 * F1Y);
      pr[y * 4 + x] = p;
      rr[y * 4 + x] = (int)cur_c[(size_t)uv * cur_plane + (size_t)(mb_cy + j) * cur_pitch_c + mb_cx + i] - p;
    }
  c_fwd4(rr);
  // ---- DC of the component: gathered by the group's first lane (hadamard2x2 / hadamard4x2 + quant_dc_cr), handed back ----
  const bool clip = d.is_cavlc != 0;
  const int qp_dc = d.qp_dc[uv], qp_ac = d.qp_ac[uv];
  int dcv[NB];
#pragma unroll
  for (int k = 0; k < NB; k++) dcv[k] = __shfl_sync(0xffffffffu, rr[0], leader + k);
  int dc_lv[NB], dczero = 0, my_dc = 0;
  if (YUV == 1) {
    // hadamard2x2 (transform.c:298): inputs b00, b04, b40, b44 = blocks 0..3
    const int p0 = dcv[0] + dcv[1], p1 = dcv[0] - dcv[1], p2 = dcv[2] + dcv[3], p3 = dcv[2] - dcv[3];
    int m[4] = {p0 + p2, p1 + p3, p0 - p2, p1 - p3};
#pragma unroll
    for (int k = 0; k < 4; k++) { dc_lv[k] = quant_dc1(&m[k], d.params_dc[uv], 15 + qp_dc / 6, qp_dc / 6, clip); dczero |= dc_lv[k] != 0; }      // SCAN_YUV420: 0,1,2,3
    const int t0 = m[0] + m[1], t1 = m[0] - m[1], t2 = m[2] + m[3], t3 = m[2] - m[3];      // ihadamard2x2
    const int o[4] = {t0 + t2, t1 + t3, t0 - t2, t1 - t3};
#pragma unroll
    for (int k = 0; k < 4; k++) if (b == k) my_dc = o[k] >> 5;
  } else {
    // tblk[x][y] = DC of block (bx = x, by = y): 2 rows of 4 (block.c:1066-1070); hadamard4x2 (transform.c:206)
    int tin[8], m[8], o8[8];
#pragma unroll
    for (int x = 0; x < 2; x++)
#pragma unroll
      for (int y = 0; y < 4; y++) tin[x * 4 + y] = dcv[y * 2 + x];
#pragma unroll
    for (int i = 0; i < 4; i++) { m[i] = tin[i] + tin[4 + i]; m[4 + i] = tin[i] - tin[4 + i]; }
#pragma unroll
    for (int i = 0; i < 2; i++) {
      const int p0 = m[4 * i], p1 = m[4 * i + 1], p2 = m[4 * i + 2], p3 = m[4 * i + 3];
      const int t0 = p0 + p3, t1 = p1 + p2, t2 = p1 - p2, t3 = p0 - p3;
      o8[4 * i] = t0 + t1; o8[4 * i + 1] = t3 + t2; o8[4 * i + 2] = t0 - t1; o8[4 * i + 3] = t3 - t2;
    }
    // SCAN_YUV422 (block.c:87-94): {row, column} of the 2 x 4 array per list position
    constexpr int S422[8] = {0 * 4 + 0, 0 * 4 + 1, 1 * 4 + 0, 0 * 4 + 2, 0 * 4 + 3, 1 * 4 + 1, 1 * 4 + 2, 1 * 4 + 3};
#pragma unroll
    for (int k = 0; k < 8; k++) { dc_lv[k] = quant_dc1(&o8[S422[k]], d.params_dc[uv], 15 + qp_dc / 6, qp_dc / 6, clip); dczero |= dc_lv[k] != 0; }
    // ihadamard4x2 (transform.c:252): in 2 rows x 4, out 4 rows x 2
#pragma unroll
    for (int i = 0; i < 4; i++) { m[i] = o8[i] + o8[4 + i]; m[4 + i] = o8[i] - o8[4 + i]; }
    int out[8];
#pragma unroll
    for (int i = 0; i < 2; i++) {
      const int p0 = m[4 * i], p1 = m[4 * i + 1], p2 = m[4 * i + 2], p3 = m[4 * i + 3];
      const int t0 = p0 + p2, t1 = p0 - p2, t2 = p1 - p3, t3 = p1 + p3;
      out[i] = t0 + t3; out[2 + i] = t1 + t2; out[4 + i] = t1 - t2; out[6 + i] = t0 - t3;
    }
#pragma unroll
    for (int k = 0; k < 8; k++) if (b == k) my_dc = (out[(k >> 1) * 2 + (k & 1)] + 32) >> 6;      // mb_rres[j << 2][0 | 4] = rshift_rnd_sf(tblk[j][0 | 1], 6)
  }
  rr[0] = my_dc;
  // ---- AC: quant_ac4x4_normal, scan positions 1..15 ----
  int lv[15], cost = 0, nz = 0;
  {
    const int qp_per = qp_ac / 6, q_bits = 15 + qp_per;
    int run = 0;
#pragma unroll
    for (int k = 1; k < 16; k++) {
      const int i = ZZ4[k][0], j = ZZ4[k][1], idx = j * 4 + i;
      const int c = rr[idx];
      int level = 0;
      if (c != 0) {
        level = (abs(c) * d.params_ac[uv][idx][1] + d.params_ac[uv][idx][0]) >> q_bits;
        if (level != 0) {
          if (clip) level = min(level, 2063);
          cost += (level > 1) ? 999999 : d.c_cost[run];
          if (c < 0) level = -level;
          rr[idx] = (((level * d.params_ac[uv][idx][2]) << qp_per) + 8) >> 4;
          nz = 1;
        } else rr[idx] = 0;
      }
      lv[k - 1] = level;
      if (level != 0) run = 0; else run++;
    }
  }
  // ---- the component's running cost and flags (block.c:1134-1173) ----
  int tot = cost, any_ac = nz;
#pragma unroll
  for (int sh = 1; sh < NB; sh <<= 1) { tot += __shfl_xor_sync(0xffffffffu, tot, sh); any_ac |= __shfl_xor_sync(0xffffffffu, any_ac, sh); }
  const bool reset = any_ac && tot < 4;                            // _CHROMA_COEFF_COST_
  if (reset && nz) {
#pragma unroll
    for (int k = 1; k < 16; k++) rr[ZZ4[k][1] * 4 + ZZ4[k][0]] = 0;
#pragma unroll
    for (int k = 0; k < 15; k++) lv[k] = 0;
    nz = 0;
  }
  unsigned bits = dczero ? ((1u << NB) - 1) : 0u;
  unsigned acbits = (nz ? (1u << b) : 0u);
#pragma unroll
  for (int sh = 1; sh < NB; sh <<= 1) acbits |= __shfl_xor_sync(0xffffffffu, acbits, sh);
  bits |= acbits;
  const int ccbp = acbits ? 2 : (dczero ? 1 : 0);
  // ---- inverse transform + reconstruction ----
  int inv = (rr[0] != 0 || nz) ? 1 : 0;
  if (inv) c_inv4(rr);
  int any_inv = inv;
#pragma unroll
  for (int sh = 1; sh < NB; sh <<= 1) any_inv |= __shfl_xor_sync(0xffffffffu, any_inv, sh);
  if (pred) {
    int gb = bad;
#pragma unroll
    for (int sh = 1; sh < NB; sh <<= 1) gb |= __shfl_xor_sync(0xffffffffu, gb, sh);
    if (gb) { if (live && b == 0) jmb_req_report(err, JMB_REQERR_BLOCKTYPE | JMB_REQERR_REF, addr); return; }
  }
  if (!live) return;
  const size_t mo = (size_t)mb * 2 + uv;
  if (recon) {
    uint8_t *o = recon + mo * 128 + (by * 4) * 8 + bx * 4;
#pragma unroll
    for (int y = 0; y < 4; y++)
#pragma unroll
      for (int x = 0; x < 4; x++)
        o[y * 8 + x] = (uint8_t)(any_inv ? jmb_clip(0, 255, ((rr[y * 4 + x] + 32) >> 6) + pr[y * 4 + x]) : pr[y * 4 + x]);
  }
  int16_t *al = ac_levels + (mo * 8 + b) * 15;
#pragma unroll
  for (int k = 0; k < 15; k++) al[k] = (int16_t)lv[k];
  if (b == 0) {
#pragma unroll
    for (int k = 0; k < 8; k++) dc_levels[mo * 8 + k] = (int16_t)(k < NB ? dc_lv[k] : 0);
    if (YUV == 1) for (int k = NB; k < 8; k++) for (int e = 0; e < 15; e++) ac_levels[(mo * 8 + k) * 15 + e] = 0;
    cbp_blk[mo] = bits;
    cr_cbp[mo] = (unsigned)ccbp;
  }
}

// u8 or u16 samples -> the u8 chroma planes the kernel reads
template <typename SRC>
__global__ void k_pack_plane(const SRC *__restrict__ src, int stride, int w, int h, uint8_t *__restrict__ dst, int pitch) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x < w && y < h) dst[(size_t)y * pitch + x] = (uint8_t)src[(size_t)y * stride + x];
}

}  // namespace

static int put_planes(jmb_ctx *ctx, const void *u, const void *v, int sample_bytes, int wc, int hc, int stride, int loc, uint8_t *dst, int pitch) {
  const size_t bytes = (size_t)stride * hc * sample_bytes;
  const void *src[2] = {u, v};
  for (int c = 0; c < 2; c++) {
    const void *d_src = src[c];
    if (jmb_is_host(loc)) {
      int rc = jmb_reserve_dev(ctx, &ctx->d_stage, &ctx->d_stage_cap, 2 * bytes); if (rc) return rc;
      d_src = (char *)ctx->d_stage + c * bytes;
      JMB_CUDA(ctx, cudaMemcpyAsync((void *)d_src, src[c], bytes, cudaMemcpyHostToDevice, ctx->stream));
    }
    dim3 grid((wc + 255) / 256, hc);
    if (sample_bytes == 1) k_pack_plane<uint8_t><<<grid, 256, 0, ctx->stream>>>((const uint8_t *)d_src, stride, wc, hc, dst + (size_t)c * pitch * hc, pitch);
    else k_pack_plane<uint16_t><<<grid, 256, 0, ctx->stream>>>((const uint16_t *)d_src, stride, wc, hc, dst + (size_t)c * pitch * hc, pitch);
    JMB_LAUNCH_CHECK(ctx);
  }
  if (loc == JMB_HOST) JMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return JMB_OK;
}

extern "C" {

int jmb_ref_put_chroma(jmb_ctx *ctx, int slot, const void *u, const void *v, int sample_bytes, int width_c, int height_c, int stride, int loc) {
  if (slot < 0 || slot >= JMB_MAX_REFS || !ctx->refs[slot].valid) return jmb_fail(ctx, JMB_ERR_STATE, "jmb_ref_put_chroma: slot %d holds no luma picture", slot);
  jmb_ref *r = &ctx->refs[slot];
  if (!u || !v || (sample_bytes != 1 && sample_bytes != 2) || width_c != r->w / 2 || (height_c != r->h && height_c != r->h / 2) || stride < width_c)
    return jmb_fail(ctx, JMB_ERR_ARG, "jmb_ref_put_chroma: %dx%d stride %d for a %dx%d picture", width_c, height_c, stride, r->w, r->h);
  JMB_CUDA(ctx, cudaSetDevice(ctx->device));
  const int pitch = (width_c + 127) & ~127;
  if (!r->chroma || r->wc != width_c || r->hc != height_c) {
    if (r->chroma) { JMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); JMB_CUDA(ctx, cudaFree(r->chroma)); r->chroma = nullptr; }
    JMB_CUDA(ctx, cudaMalloc(&r->chroma, (size_t)2 * pitch * height_c));
    r->wc = width_c; r->hc = height_c; r->pitch_c = pitch;
  }
  return put_planes(ctx, u, v, sample_bytes, width_c, height_c, stride, loc, r->chroma, pitch);
}

int jmb_pic_chroma(jmb_ctx *ctx, const void *u, const void *v, int sample_bytes, int width_c, int height_c, int stride, int loc) {
  if (!ctx->cur) return jmb_fail(ctx, JMB_ERR_STATE, "jmb_pic_chroma: call jmb_pic_begin first");
  if (!u || !v || (sample_bytes != 1 && sample_bytes != 2) || width_c != ctx->cur_w / 2 || (height_c != ctx->cur_h && height_c != ctx->cur_h / 2) || stride < width_c)
    return jmb_fail(ctx, JMB_ERR_ARG, "jmb_pic_chroma: %dx%d stride %d for a %dx%d picture", width_c, height_c, stride, ctx->cur_w, ctx->cur_h);
  JMB_CUDA(ctx, cudaSetDevice(ctx->device));
  const int pitch = (width_c + 127) & ~127;
  const size_t bytes = (size_t)2 * pitch * height_c;
  if (bytes > ctx->cur_c_cap) {
    if (ctx->cur_c) { JMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); JMB_CUDA(ctx, cudaFree(ctx->cur_c)); ctx->cur_c = nullptr; }
    JMB_CUDA(ctx, cudaMalloc(&ctx->cur_c, bytes));
    ctx->cur_c_cap = bytes;
  }
  ctx->cur_wc = width_c; ctx->cur_hc = height_c; ctx->cur_pitch_c = pitch;
  return put_planes(ctx, u, v, sample_bytes, width_c, height_c, stride, loc, ctx->cur_c, pitch);
}

int jmb_chroma_residual_coding(jmb_ctx *ctx, const jmb_mb_pred *pred, int mode, int first_mb, int n_mb, const jmb_chroma_desc *d,
                               int16_t *dc_levels, int16_t *ac_levels, uint32_t *cbp_blk_chroma, uint32_t *cr_cbp, uint8_t *recon, int loc) {
  if (!d || !dc_levels || !ac_levels || !cbp_blk_chroma || !cr_cbp) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_chroma_residual_coding: NULL argument");
  if (!ctx->cur || !ctx->cur_c || ctx->nref == 0) return jmb_fail(ctx, JMB_ERR_STATE, "jmb_chroma_residual_coding: call jmb_pic_begin and jmb_pic_chroma first");
  const int mb_w = ctx->cur_w / 16, mb_total = mb_w * (ctx->cur_h / 16);
  if (n_mb <= 0 || first_mb < 0 || first_mb + n_mb > mb_total) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_chroma_residual_coding: macroblocks %d..%d (picture has %d)", first_mb, first_mb + n_mb - 1, mb_total);
  if (d->yuv_format != 1 && d->yuv_format != 2) return jmb_fail(ctx, JMB_ERR_UNSUPPORTED, "jmb_chroma_residual_coding: yuv_format %d (4:2:0 and 4:2:2 only)", d->yuv_format);
  if (ctx->cur_hc != (d->yuv_format == 1 ? ctx->cur_h / 2 : ctx->cur_h)) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_chroma_residual_coding: the chroma planes are %d rows, yuv_format %d", ctx->cur_hc, d->yuv_format);
  for (int c = 0; c < 2; c++)
    if (d->qp_ac[c] < 0 || d->qp_ac[c] > 87 || d->qp_dc[c] < 0 || d->qp_dc[c] > 90) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_chroma_residual_coding: chroma qp %d / %d", d->qp_ac[c], d->qp_dc[c]);
  if (!pred && (mode < 1 || mode > 7)) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_chroma_residual_coding: mode %d", mode);
  if (!pred && (!ctx->last_res || ctx->last_res_n < (first_mb + n_mb) * 41)) return jmb_fail(ctx, JMB_ERR_STATE, "jmb_chroma_residual_coding: no resident search results");
  JMB_CUDA(ctx, cudaSetDevice(ctx->device));
  const uint8_t *tab[JMB_MAX_REFS];
  for (int i = 0; i < JMB_MAX_REFS; i++) {
    tab[i] = nullptr;
    if (i < ctx->nref) {
      const jmb_ref &r = ctx->refs[ctx->ref_list[i]];
      if (!r.chroma || r.wc != ctx->cur_wc || r.hc != ctx->cur_hc) return jmb_fail(ctx, JMB_ERR_STATE, "jmb_chroma_residual_coding: reference %d has no chroma planes of this format (jmb_ref_put_chroma)", i);
      tab[i] = r.chroma;
    }
  }
  int rc = jmb_reserve_dev(ctx, &ctx->d_stage5, &ctx->d_stage5_cap, sizeof(tab)); if (rc) return rc;
  JMB_CUDA(ctx, cudaMemcpyAsync(ctx->d_stage5, tab, sizeof(tab), cudaMemcpyHostToDevice, ctx->stream));
  const jmb_ref &r0 = ctx->refs[ctx->ref_list[0]];
  const bool host = jmb_is_host(loc);
  const size_t n2 = (size_t)n_mb * 2;
  const jmb_mb_pred *d_pred = pred;
  int16_t *d_dc = dc_levels, *d_ac = ac_levels; unsigned *d_cb = cbp_blk_chroma, *d_cc = cr_cbp; uint8_t *d_rec = recon;
  if (host) {
    if (pred) {
      rc = jmb_reserve_dev(ctx, &ctx->d_stage, &ctx->d_stage_cap, (size_t)n_mb * sizeof(jmb_mb_pred)); if (rc) return rc;
      JMB_CUDA(ctx, cudaMemcpyAsync(ctx->d_stage, pred, (size_t)n_mb * sizeof(jmb_mb_pred), cudaMemcpyHostToDevice, ctx->stream));
      d_pred = (const jmb_mb_pred *)ctx->d_stage;
    }
    rc = jmb_reserve_dev(ctx, &ctx->d_stage3, &ctx->d_stage3_cap, n2 * (16 + 240 + 4 + 4 + 128)); if (rc) return rc;
    char *a = (char *)ctx->d_stage3;
    d_dc = (int16_t *)a; d_ac = (int16_t *)(a + n2 * 16); d_cb = (unsigned *)(a + n2 * 256); d_cc = (unsigned *)(a + n2 * 260);
    d_rec = recon ? (uint8_t *)(a + n2 * 264) : nullptr;
  }
  const int threads = n_mb * 2 * (d->yuv_format == 1 ? 4 : 8);
  jmb_time_begin(ctx, JMB_K_CHROMA);
  if (d->yuv_format == 1)
    k_chroma_rc<1><<<(threads + 127) / 128, 128, 0, ctx->stream>>>(ctx->last_res, d_pred, mode, first_mb, n_mb, mb_w, *d, ctx->cur_c, ctx->cur_pitch_c,
        (size_t)ctx->cur_pitch_c * ctx->cur_hc, (const uint8_t *const *)ctx->d_stage5, r0.pitch_c, (size_t)r0.pitch_c * r0.hc, ctx->cur_wc, ctx->cur_hc,
        ctx->nref, ctx->d_err, d_dc, d_ac, d_cb, d_cc, d_rec);
  else
    k_chroma_rc<2><<<(threads + 127) / 128, 128, 0, ctx->stream>>>(ctx->last_res, d_pred, mode, first_mb, n_mb, mb_w, *d, ctx->cur_c, ctx->cur_pitch_c,
        (size_t)ctx->cur_pitch_c * ctx->cur_hc, (const uint8_t *const *)ctx->d_stage5, r0.pitch_c, (size_t)r0.pitch_c * r0.hc, ctx->cur_wc, ctx->cur_hc,
        ctx->nref, ctx->d_err, d_dc, d_ac, d_cb, d_cc, d_rec);
  jmb_time_end(ctx, JMB_K_CHROMA);
  JMB_LAUNCH_CHECK(ctx);
  if (host) {
    JMB_CUDA(ctx, cudaMemcpyAsync(dc_levels, d_dc, n2 * 16, cudaMemcpyDeviceToHost, ctx->stream));
    JMB_CUDA(ctx, cudaMemcpyAsync(ac_levels, d_ac, n2 * 240, cudaMemcpyDeviceToHost, ctx->stream));
    JMB_CUDA(ctx, cudaMemcpyAsync(cbp_blk_chroma, d_cb, n2 * 4, cudaMemcpyDeviceToHost, ctx->stream));
    JMB_CUDA(ctx, cudaMemcpyAsync(cr_cbp, d_cc, n2 * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (recon) JMB_CUDA(ctx, cudaMemcpyAsync(recon, d_rec, n2 * 128, cudaMemcpyDeviceToHost, ctx->stream));
    if (loc == JMB_HOST) JMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return JMB_OK;
}

}  // extern "C"
