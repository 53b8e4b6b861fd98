// k_refine.cu -- K4 + K5: distortion of arbitrary (sub-pel) candidates and the half-/quarter-pel
// refinement that follows every integer search.
//
//  * block distortion = computeSAD / computeSSE / computeSATD (lencod/src/me_distortion.c:349,1190,745)
//    read from the 16 quarter-pel planes with JM's UMVLine4X addressing (lencod/inc/refbuf.h:22-26):
//    SAD/SSE clamp the PARTITION origin once (:367), SATD clamps EVERY 4x4 / 8x8 sub-block origin
//    (:771, :799).  HadamardSAD4x4 (:175-258) = (sum|H4 d H4| + 1) >> 1, HadamardSAD8x8 (:266-341)
//    = (sum|H8 d H8| + 2) >> 2.
//  * refinement = sub_pel_motion_estimation (lencod/src/me_fullsearch.c:186-289): nine half-pel
//    candidates around the integer mv, then eight (or nine) quarter-pel candidates around the best
//    half-pel one; J = lambda * mvbits + (D << 5); the first candidate in spiral order wins ties.
//
// One warp per search: the (candidate x sub-block) work items are spread over the 32 lanes, each
// lane does whole 4x4 (or 8x8) Hadamards in registers, per-candidate sums are combined with
// shared-memory atomics, lane 0 replays JM's sequential strict-'<' selection.
#include "jmb_internal.h"

namespace {

__constant__ signed char c_spiral9[9][2] = {{0,0},{0,-1},{0,1},{-1,-1},{1,-1},{-1,0},{1,0},{-1,1},{1,1}};
__constant__ unsigned char c_bsx[8] = {16, 16, 16, 8, 8, 8, 4, 4};
__constant__ unsigned char c_bsy[8] = {16, 16, 8, 16, 8, 4, 8, 4};

struct RefView { const uint8_t *planes; size_t plane_bytes; int pitch, w, h; };

// pointer to the sample at quarter-pel position (qx,qy) after UMVLine4X's origin clamp
__device__ __forceinline__ const uint8_t *umv(const RefView &rv, int qy, int qx) {
  int iy = jmb_clip(-JMB_PAD_Y, rv.h + JMB_PAD_Y - 1 - 16, qy >> 2);
  int ix = jmb_clip(-JMB_PAD_X, rv.w + JMB_PAD_X - 1 - 16, qx >> 2);
  return rv.planes + (size_t)((qy & 3) * 4 + (qx & 3)) * rv.plane_bytes + (size_t)(iy + JMB_PAD_Y) * rv.pitch + (ix + JMB_PAD_X);
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

// distortion contribution of sub-block (sbx, sby) [units of n pels] of a block at (pos_x,pos_y)
// against the candidate at absolute quarter-pel (cqx, cqy)
__device__ int subblock_dist(const RefView &rv, const uint8_t *cur, int cur_pitch, int pos_x, int pos_y,
                             int cqx, int cqy, int sbx, int sby, int n, int metric) {
  const uint8_t *src = cur + (size_t)(pos_y + sby * n) * cur_pitch + pos_x + sbx * n;
  const uint8_t *ref;
  if (metric == JMB_SATD) ref = umv(rv, cqy + ((sby * n) << 2), cqx + ((sbx * n) << 2));   // per-sub-block clamp
  else ref = umv(rv, cqy, cqx) + (size_t)(sby * n) * rv.pitch + sbx * n;                     // partition clamp
  if (n == 4) {
    int d[16];
#pragma unroll
    for (int y = 0; y < 4; y++) {
      unsigned sv = *(const unsigned *)(src + (size_t)y * cur_pitch);
      const uint8_t *rp = ref + (size_t)y * rv.pitch;
#pragma unroll
      for (int x = 0; x < 4; x++) d[y * 4 + x] = (int)((sv >> (8 * x)) & 255) - (int)rp[x];
    }
    if (metric == JMB_SATD) return hadamard4(d);
    int s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += (metric == JMB_SAD) ? abs(d[i]) : d[i] * d[i];
    return s;
  }
  int a[64];
#pragma unroll
  for (int y = 0; y < 8; y++) {
    const uint8_t *sp = src + (size_t)y * cur_pitch, *rp = ref + (size_t)y * rv.pitch;
#pragma unroll
    for (int x = 0; x < 8; x++) a[y * 8 + x] = (int)sp[x] - (int)rp[x];
  }
  return hadamard8(a);
}

// one refinement stage of sub_pel_motion_estimation: candidates mv + step*spiral[pos], pos0 <= pos < pos1
__device__ void refine_stage(const RefView &rv, const uint8_t *cur, int cur_pitch, const jmb_me_req &r, int lane,
                             int *sums /* shared, >= 9 per warp */, int step, int pos0, int pos1, int metric, int lam,
                             int &mvx, int &mvy, long long &min_mcost) {
  const int bsx = c_bsx[r.blocktype], bsy = c_bsy[r.blocktype];
  const int n = (metric == JMB_SATD && (r.flags & JMB_REQ_TEST8X8)) ? 8 : 4;
  const int nsx = bsx / n, nsub = nsx * (bsy / n);
  if (lane < 9) sums[lane] = 0;
  __syncwarp();
  const int ncand = pos1 - pos0;
  for (int it = lane; it < ncand * nsub; it += 32) {
    int c = it / nsub, sb = it - c * nsub;
    int pos = pos0 + c;
    int cqx = (r.pos_x << 2) + mvx + step * c_spiral9[pos][0], cqy = (r.pos_y << 2) + mvy + step * c_spiral9[pos][1];
    int d = subblock_dist(rv, cur, cur_pitch, r.pos_x, r.pos_y, cqx, cqy, sb % nsx, sb / nsx, n, metric);
    atomicAdd(&sums[pos], d);
  }
  __syncwarp();
  int best = 0;
  for (int pos = pos0; pos < pos1; pos++) {      // every lane replays JM's loop (uniform, tiny)
    int cx = mvx + step * c_spiral9[pos][0], cy = mvy + step * c_spiral9[pos][1];
    long long mc = (long long)lam * (jmb_mvbits(cx - r.pred_x) + jmb_mvbits(cy - r.pred_y));
    if (mc >= min_mcost) continue;
    mc += (long long)sums[pos] << 5;
    if (mc < min_mcost) { min_mcost = mc; best = pos; }
  }
  mvx += step * c_spiral9[best][0]; mvy += step * c_spiral9[best][1];
  __syncwarp();
}

__global__ void __launch_bounds__(256)
k_subpel_refine(const jmb_me_req *__restrict__ reqs, jmb_me_res *__restrict__ res, int n, const uint8_t *__restrict__ cur, int cur_pitch,
                const uint8_t *const *__restrict__ ref_planes, size_t plane_bytes, int ref_pitch, int w, int h, jmb_me_config me) {
  __shared__ int s_sums[8][12];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * 8 + warp;
  if (i >= n) return;
  const jmb_me_req r = reqs[i];
  if (!(r.flags & JMB_REQ_SUBPEL)) return;
  RefView rv{ref_planes[r.ref], plane_bytes, ref_pitch, w, h};
  int mvx, mvy; long long min_mcost;
  const long long DISTBLK_MAX = (long long)0x7fffffff << 5;     // lencod/inc/defines.h:136
  if (r.flags & JMB_REQ_SKIP_INT) { mvx = r.center_x; mvy = r.center_y; min_mcost = r.min_mcost; }
  else {
    mvx = res[i].imv_x; mvy = res[i].imv_y; min_mcost = res[i].icost;
    if (!me.start_hp) min_mcost = DISTBLK_MAX;                  // BlockMotionSearch, mv_search.c:971-974
  }
  const int max_pos2 = !me.start_hp ? max(1, me.search_pos2) : me.search_pos2;
  refine_stage(rv, cur, cur_pitch, r, lane, s_sums[warp], 2, me.start_hp, max_pos2, me.metric[1], r.lambda[1], mvx, mvy, min_mcost);
  if (!me.start_qp) min_mcost = DISTBLK_MAX;
  refine_stage(rv, cur, cur_pitch, r, lane, s_sums[warp], 1, me.start_qp, me.search_pos4, me.metric[2], r.lambda[2], mvx, mvy, min_mcost);
  if (lane == 0) { res[i].mv_x = (int16_t)mvx; res[i].mv_y = (int16_t)mvy; res[i].cost = min_mcost; }
}

__global__ void k_dist(const uint8_t *__restrict__ cur, int cur_pitch, RefView rv, int blocktype, int pos_x, int pos_y,
                       const int16_t *__restrict__ cand, int ncand, int metric, int test8x8, int *__restrict__ out) {
  const int bsx = c_bsx[blocktype], bsy = c_bsy[blocktype];
  const int n = (metric == JMB_SATD && test8x8) ? 8 : 4;
  const int nsx = bsx / n, nsub = nsx * (bsy / n);
  const int it = blockIdx.x * blockDim.x + threadIdx.x;
  if (it >= ncand * nsub) return;
  const int c = it / nsub, sb = it - c * nsub;
  int d = subblock_dist(rv, cur, cur_pitch, pos_x, pos_y, cand[2 * c], cand[2 * c + 1], sb % nsx, sb / nsx, n, metric);
  atomicAdd(&out[c], d);
}

}  // namespace

int jmb_launch_refine(jmb_ctx *ctx, const jmb_me_req *d_reqs, jmb_me_res *d_res, int n, const uint8_t *const *d_ref_planes) {
  const jmb_ref &r0 = ctx->refs[ctx->ref_list[0]];
  jmb_time_begin(ctx, JMB_K_REFINE);
  k_subpel_refine<<<(n + 7) / 8, 256, 0, ctx->stream>>>(d_reqs, d_res, n, ctx->cur, ctx->cur_pitch, d_ref_planes, r0.plane_bytes,
                                                        r0.pitch, ctx->cur_w, ctx->cur_h, ctx->me);
  jmb_time_end(ctx, JMB_K_REFINE);
  JMB_LAUNCH_CHECK(ctx);
  return JMB_OK;
}

extern "C" int jmb_dist(jmb_ctx *ctx, int ref, int metric, int blocktype, int pos_x, int pos_y,
                        const int16_t *cand_xy, int n, int test8x8, int32_t *out, int loc) {
  static const int bsx[8] = {0, 16, 16, 8, 8, 8, 4, 4}, bsy[8] = {0, 16, 8, 16, 8, 4, 8, 4};
  if (n <= 0) return JMB_OK;
  if (!ctx->cur || ref < 0 || ref >= ctx->nref) return jmb_fail(ctx, JMB_ERR_STATE, "jmb_dist: no picture / bad ref %d", ref);
  if (blocktype < 1 || blocktype > 7 || metric < 0 || metric > 2) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_dist: blocktype %d metric %d", blocktype, metric);
  if (pos_x < 0 || pos_y < 0 || (pos_x & 3) || (pos_y & 3) || pos_x + bsx[blocktype] > ctx->cur_w || pos_y + bsy[blocktype] > ctx->cur_h)
    return jmb_fail(ctx, JMB_ERR_ARG, "jmb_dist: block (%d,%d) type %d", pos_x, pos_y, blocktype);
  if (test8x8 && metric == JMB_SATD && blocktype > 4) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_dist: test8x8 needs blocktype <= 4");
  JMB_CUDA(ctx, cudaSetDevice(ctx->device));
  const jmb_ref &r = ctx->refs[ctx->ref_list[ref]];
  const int16_t *d_c = cand_xy; int *d_o = out;
  if (loc == JMB_HOST) {
    int rc = jmb_reserve_dev(ctx, &ctx->d_stage, &ctx->d_stage_cap, (size_t)n * 4); if (rc) return rc;
    rc = jmb_reserve_dev(ctx, &ctx->d_stage2, &ctx->d_stage2_cap, (size_t)n * 4); if (rc) return rc;
    JMB_CUDA(ctx, cudaMemcpyAsync(ctx->d_stage, cand_xy, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    d_c = (const int16_t *)ctx->d_stage; d_o = (int *)ctx->d_stage2;
  }
  JMB_CUDA(ctx, cudaMemsetAsync(d_o, 0, (size_t)n * 4, ctx->stream));
  const int nn = (metric == JMB_SATD && test8x8) ? 8 : 4;
  const int items = n * (bsx[blocktype] / nn) * (bsy[blocktype] / nn);
  RefView rv{r.planes, r.plane_bytes, r.pitch, r.w, r.h};
  jmb_time_begin(ctx, JMB_K_DIST);
  k_dist<<<(items + 127) / 128, 128, 0, ctx->stream>>>(ctx->cur, ctx->cur_pitch, rv, blocktype, pos_x, pos_y, d_c, n, metric, test8x8, d_o);
  jmb_time_end(ctx, JMB_K_DIST);
  JMB_LAUNCH_CHECK(ctx);
  if (loc == JMB_HOST) {
    JMB_CUDA(ctx, cudaMemcpyAsync(out, d_o, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    JMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return JMB_OK;
}
