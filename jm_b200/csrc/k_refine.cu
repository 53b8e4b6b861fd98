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
// One CTA per 41 consecutive searches (a macroblock in the picture layout).  Work item = one 4x4 (or 8x8) sub-block of one
// search; items whose distortions cannot differ (same sub-block, same mv, SATD's per-sub-block clamp) are found through a
// shared-memory hash table and evaluated once, each Hadamard whole in one thread's registers; the values are added up per
// search and one thread per search replays JM's sequential strict-'<' selection.
#include "jmb_internal.h"
#include "jmb_dist_dev.cuh"

namespace {

// Refinement of up to RQ consecutive requests per CTA (one macroblock's 41 searches in the frame layout).  Per stage
// (half-pel, quarter-pel): the requests name their sub-blocks, identical ones elect a leader (open addressing, atomicCAS), the
// distinct (sub-block, candidate group) units are spread over the threads -- the source sub-block in registers, one candidate
// after the other -- every item then adds its leader's values to sums[request][candidate], and one thread per request replays
// JM's sequential strict-'<' selection (me_fullsearch.c:221-289).
constexpr int RT = 128, RQ = 41;
static_assert(RQ == 41, "the picture form takes CTA b for macroblock b and thread t for its partition t");
// (round-1 A/B, kept for the record: batching the reference rows of several candidates in one thread lost to occupancy --
// (candidates in flight, CTAs/SM) (3,4) 0.243 ms, (2,6) 0.194, (2,8) 0.176, (1,8) 0.1745)
constexpr int DD_MAX = RT, DD_T = 256;     // one item per thread and chunk; the sub-blocks of one macroblock's 41 searches (112) are one chunk

struct RefineS {
  short pos_x, pos_y, pred_x, pred_y;
  int mvx, mvy;
  long long min_mcost;
  int lam_h, lam_q;
  unsigned char blocktype, ref, flags, nsub, n, nsx, nsx_sh;
  short first;                // index of the request's first work item
};

#ifndef JMB_RF_MINB
#define JMB_RF_MINB 10      // CTAs/SM (registers): 6 (80) 0.127 ms, 8 (64) 0.113, 10 (48) 0.106, 12 (40) 0.107 -- the phases are short and
#endif                      // separated by barriers, so resident CTAs count for more than the 216 bytes of spills at 48 registers
__global__ void __launch_bounds__(RT, JMB_RF_MINB)
k_subpel_refine(const jmb_me_req *__restrict__ reqs, jmb_me_res *__restrict__ res, int n, const uint8_t *__restrict__ cur, int cur_pitch,
                const uint8_t *const *__restrict__ ref_planes, size_t plane_bytes, int ref_pitch, int w, int h, jmb_me_config me, int nref, int *__restrict__ err,
                const jmb_frame_gen gen, const jmb_pack_out pack) {
  __shared__ RefineS sr[RQ];
  __shared__ int sums[RQ][9];
  __shared__ int s_total;
  __shared__ int4 dd_key[DD_MAX];                  // (sub-block position, mv, reference | size | request for SAD/SSE)
  __shared__ int dd_val[DD_MAX][9];                // distortion of distinct sub-block x candidate
  __shared__ int dd_table[DD_T], dd_nlead;
  __shared__ short dd_info[DD_MAX], dd_lead[DD_MAX], dd_slot[DD_MAX], dd_list[DD_MAX];
  const int tid = threadIdx.x, base = blockIdx.x * RQ, cnt = min(RQ, n - base);
  const long long DISTBLK_MAX = (long long)0x7fffffff << 5;     // lencod/inc/defines.h:136

  if (tid < cnt) {
    const jmb_me_req r = gen.pred ? jmb_frame_request(gen, blockIdx.x, tid) : reqs[base + tid];      // picture form: CTA = macroblock (RQ = 41)
    RefineS q;
    q.pos_x = r.pos_x; q.pos_y = r.pos_y; q.pred_x = r.pred_x; q.pred_y = r.pred_y;
    q.lam_h = r.lambda[1]; q.lam_q = r.lambda[2];
    q.blocktype = r.blocktype; q.ref = r.ref; q.flags = r.flags;
    q.nsub = 0; q.n = 4; q.nsx = 1; q.nsx_sh = 0; q.first = 0; q.mvx = q.mvy = 0; q.min_mcost = 0;
    const int bad = jmb_req_check(r, w, h, nref);
    jmb_req_report(err, bad, base + tid);
    if (bad) q.flags = 0;
    else if (r.flags & JMB_REQ_SUBPEL) {
      if (r.flags & JMB_REQ_SKIP_INT) { q.mvx = r.center_x; q.mvy = r.center_y; q.min_mcost = r.min_mcost; }
      else {
        const jmb_me_res o = res[base + tid];
        q.mvx = o.imv_x; q.mvy = o.imv_y; q.min_mcost = o.icost;
        if (!me.start_hp) q.min_mcost = DISTBLK_MAX;                  // BlockMotionSearch, mv_search.c:971-974
      }
    }
    sr[tid] = q;
  }
  __syncthreads();

#pragma unroll 1
  for (int stage = 0; stage < 2; stage++) {
    const int metric = me.metric[1 + stage], step = stage ? 1 : 2;
    const int pos0 = stage ? me.start_qp : me.start_hp;
    const int pos1 = stage ? me.search_pos4 : (!me.start_hp ? max(1, me.search_pos2) : me.search_pos2);
    if (tid < cnt) {      // work items of this stage (the sub-block size depends on the stage's metric)
      RefineS &q = sr[tid];
      q.nsub = 0;
      if (q.flags & JMB_REQ_SUBPEL) {
        const int nn = (metric == JMB_SATD && (q.flags & JMB_REQ_TEST8X8)) ? 8 : 4;
        const int nsx = c_bsx[q.blocktype] / nn;                     // 1, 2 or 4 sub-blocks across
        q.n = (unsigned char)nn; q.nsx = (unsigned char)nsx; q.nsx_sh = (unsigned char)(nsx >> 1);
        q.nsub = (unsigned char)(nsx * (c_bsy[q.blocktype] / nn));
        if (stage == 1 && !me.start_qp) q.min_mcost = DISTBLK_MAX;
      }
    }
    for (int i = tid; i < cnt * 9; i += RT) (&sums[0][0])[i] = 0;
    __syncthreads();
    if (tid < 32) {       // exclusive prefix sum of the item counts: warp 0 scans requests tid and tid + 32
      static_assert(RQ <= 64, "the scan covers two requests per lane");
      const int n0 = tid < cnt ? sr[tid].nsub : 0, n1 = tid + 32 < cnt ? sr[tid + 32].nsub : 0;
      int s0 = n0, s1 = n1;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t0 = __shfl_up_sync(0xffffffffu, s0, d), t1 = __shfl_up_sync(0xffffffffu, s1, d);
        if (tid >= d) { s0 += t0; s1 += t1; }
      }
      const int half = __shfl_sync(0xffffffffu, s0, 31);
      if (tid < cnt) sr[tid].first = s0 - n0;
      if (tid + 32 < cnt) sr[tid + 32].first = half + s1 - n1;
      if (tid == 31) s_total = half + s1;
    }
    __syncthreads();
    const int total = s_total;
    // With SATD every 4x4 / 8x8 sub-block clamps its own origin (me_distortion.c:771,799), so the distortion of a sub-block at a
    // candidate depends only on where the sub-block lies, the reference and the candidate's ABSOLUTE mv -- not on which partition
    // it belongs to.  Partitions of one macroblock that left the integer search with the same mv (the common case in coherent
    // motion) walk the same candidates: their sub-blocks share one evaluation.  (SAD / SSE clamp the partition origin: the
    // request index joins the key and nothing is shared.)  The items go by in chunks of DD_MAX: a macroblock's 112 are one chunk.
#pragma unroll 1
    for (int c0 = 0; c0 < total; c0 += DD_MAX) {
      const int nit = min(DD_MAX, total - c0);
      for (int i = tid; i < DD_T; i += RT) dd_table[i] = -1;
      if (tid == 0) dd_nlead = 0;
      if (tid < cnt) {                                   // every request names its sub-blocks of this chunk
        const RefineS &q = sr[tid];
        const int nn = q.n, tail = q.ref | (nn << 8) | (metric != JMB_SATD ? (tid + 1) << 16 : 0);
        for (int item = max((int)q.first, c0); item < min(q.first + q.nsub, c0 + DD_MAX); item++) {
          const int sb = item - q.first, sbx = sb & (q.nsx - 1), sby = sb >> q.nsx_sh;
          dd_key[item - c0] = make_int4((q.pos_x + sbx * nn) | ((q.pos_y + sby * nn) << 16), q.mvx, q.mvy, tail);
          dd_info[item - c0] = (short)(tid | (sb << 8));
        }
      }
      __syncthreads();
      if (tid < nit) {                                   // open addressing: the first claimant of a key evaluates it
        const int4 k = dd_key[tid];
        const unsigned hsh = ((unsigned)k.x * 0x9e3779b1u) ^ ((unsigned)k.y * 0x85ebca6bu) ^ ((unsigned)k.z * 0xc2b2ae35u) ^ ((unsigned)k.w * 0x27d4eb2fu);
        int slot = (int)(hsh >> 16) & (DD_T - 1), lead;
        for (;;) {
          const int o = atomicCAS(&dd_table[slot], -1, tid);
          if (o == -1) { lead = tid; break; }
          const int4 ko = dd_key[o];
          if (ko.x == k.x && ko.y == k.y && ko.z == k.z && ko.w == k.w) { lead = o; break; }
          slot = (slot + 1) & (DD_T - 1);
        }
        dd_lead[tid] = (short)lead;
        if (lead == tid) { const int idx = atomicAdd(&dd_nlead, 1); dd_list[idx] = (short)tid; dd_slot[tid] = (short)idx; }
      }
      __syncthreads();
      // (distinct sub-block, candidate group) units: the fewer distinct sub-blocks, the finer the split over the threads
      const int nl = dd_nlead, ncand = pos1 - pos0;
      int G = ncand;
      if (nl * ncand <= RT) G = 1; else if (nl * ((ncand + 1) >> 1) <= RT) G = 2; else if (nl * ((ncand + 3) >> 2) <= RT) G = 4;
      const int groups = ncand > 0 ? (ncand + G - 1) / G : 0;
      for (int u = tid; u < nl * groups; u += RT) {
        const int idx = u / groups, g = u - idx * groups, info = dd_info[dd_list[idx]], lo = info & 0xff, sb = info >> 8;
        const RefineS &q = sr[lo];
        const int nn = q.n, sbx = sb & (q.nsx - 1), sby = sb >> q.nsx_sh;
        RefView rv{ref_planes[q.ref], plane_bytes, ref_pitch, w, h};
        SrcBlk src;
        load_src(src, cur, cur_pitch, q.pos_x + sbx * nn, q.pos_y + sby * nn, nn);
        const int bqx = (q.pos_x << 2) + q.mvx, bqy = (q.pos_y << 2) + q.mvy;
        const int pa = pos0 + g * G, pb = min(pos1, pa + G);
        if (nn == 4) {
          for (int pos = pa; pos < pb; pos++) {
            unsigned rw[4];
            load_ref4(rv, bqx + step * c_spiral9[pos][0], bqy + step * c_spiral9[pos][1], sbx, sby, metric, rw);
            dd_val[idx][pos] = dist4(src, rw, metric);
          }
        } else {
          for (int pos = pa; pos < pb; pos++)
            dd_val[idx][pos] = subblock_dist(rv, src, bqx + step * c_spiral9[pos][0], bqy + step * c_spiral9[pos][1], sbx, sby, 8, metric);
        }
      }
      __syncthreads();
      if (tid < nit) {
        const int lo = dd_info[tid] & 0xff, idx = dd_slot[dd_lead[tid]];
        for (int pos = pos0; pos < pos1; pos++) atomicAdd(&sums[lo][pos], dd_val[idx][pos]);
      }
      __syncthreads();
    }
    if (tid < cnt && (sr[tid].flags & JMB_REQ_SUBPEL)) {
      // JM's sequential strict-'<' walk (me_fullsearch.c:221-289).  A candidate's cost fits 32 bits (lambda <= 65535 times at most
      // 130 bits, plus a block distortion < 2^24 shifted by 5); an incumbent beyond that loses to any candidate.
      RefineS &q = sr[tid];
      const unsigned lam = (unsigned)(stage ? q.lam_q : q.lam_h);
      unsigned cur = q.min_mcost > 0xffffffffll ? 0xffffffffu : (unsigned)q.min_mcost;
      int best = -1;
      for (int pos = pos0; pos < pos1; pos++) {
        const int cx = q.mvx + step * c_spiral9[pos][0], cy = q.mvy + step * c_spiral9[pos][1];
        const unsigned mc = lam * (unsigned)(jmb_mvbits(cx - q.pred_x) + jmb_mvbits(cy - q.pred_y)) + ((unsigned)sums[tid][pos] << 5);
        if (mc < cur) { cur = mc; best = pos; }
      }
      if (best >= 0) { q.min_mcost = cur; q.mvx += step * c_spiral9[best][0]; q.mvy += step * c_spiral9[best][1]; }
    }
    __syncthreads();
  }
  if (tid < cnt) {
    jmb_me_res *const o = res + base + tid;
    int mvx, mvy; long long cost;
    if (sr[tid].flags & JMB_REQ_SUBPEL) { mvx = sr[tid].mvx; mvy = sr[tid].mvy; cost = sr[tid].min_mcost; o->cost = cost; }
    else { if (!pack.on) return; mvx = o->mv_x; mvy = o->mv_y; cost = o->cost; }      // not refined: the integer stage's answer stands
    if (pack.on) {      // picture form: the final clip of the mv (mv_search.c:981) and the 8-byte result, as k_pack_results does
      mvx = jmb_clip(gen.fp.mv_min_x, gen.fp.mv_max_x, mvx); mvy = jmb_clip(gen.fp.mv_min_y, gen.fp.mv_max_y, mvy);
      if (pack.out) {
        jmb_me_res8 o8;
        o8.mv_x = (int16_t)mvx; o8.mv_y = (int16_t)mvy; o8.cost = cost > 0x7fffffffLL ? 0x7fffffff : (int32_t)cost;
        pack.out[base + tid] = o8;
      }
    }
    o->mv_x = (int16_t)mvx; o->mv_y = (int16_t)mvy;
  }
}

// prediction sample of jmb_dist_ex: one or two references, optionally weighted (me_distortion.c:434-1520)
struct DistPred { int form, c2x, c2y, w1, w2, off, shift, round; };
__device__ __forceinline__ unsigned pred_word(unsigned a, unsigned b, const DistPred &P) {
  unsigned o = 0;
#pragma unroll
  for (int x = 0; x < 4; x++) {
    const int ra = (int)((a >> (8 * x)) & 255), rb = (int)((b >> (8 * x)) & 255);
    int v;
    if (P.form == JMB_PRED_WEIGHTED) v = jmb_clip(0, 255, ((P.w1 * ra + P.round) >> P.shift) + P.off);
    else if (P.form == JMB_PRED_AVERAGE) v = (ra + rb + 1) >> 1;
    else v = jmb_clip(0, 255, ((P.w1 * ra + P.w2 * rb + P.round) >> P.shift) + P.off);
    o |= (unsigned)v << (8 * x);
  }
  return o;
}

// one thread per (candidate, sub-block); the sub-block's prediction is formed in registers and handed to the same
// SAD / SSE / Hadamard code the single-reference path uses
__global__ void k_dist(const uint8_t *__restrict__ cur, int cur_pitch, RefView rv, RefView rv2, DistPred P, int blocktype, int pos_x, int pos_y,
                       const int16_t *__restrict__ cand, int ncand, int metric, int test8x8, int *__restrict__ out) {
  const int bsx = c_bsx[blocktype], bsy = c_bsy[blocktype];
  const int n = (metric == JMB_SATD && test8x8) ? 8 : 4;
  const int nsx = bsx / n, nsub = nsx * (bsy / n);
  const int it = blockIdx.x * blockDim.x + threadIdx.x;
  if (it >= ncand * nsub) return;
  const int c = it / nsub, sb = it - c * nsub, sbx = sb % nsx, sby = sb / nsx;
  SrcBlk src;
  load_src(src, cur, cur_pitch, pos_x + sbx * n, pos_y + sby * n, n);
  const int c1x = cand[2 * c], c1y = cand[2 * c + 1];
  const bool two = P.form >= JMB_PRED_AVERAGE;
  const uint8_t *r1, *r2 = nullptr;
  if (metric == JMB_SATD) {      // every sub-block origin is clamped, in each reference
    r1 = umv(rv, c1y + ((sby * n) << 2), c1x + ((sbx * n) << 2));
    if (two) r2 = umv(rv2, P.c2y + ((sby * n) << 2), P.c2x + ((sbx * n) << 2));
  } else {                       // the partition origin is clamped
    r1 = umv(rv, c1y, c1x) + (size_t)(sby * n) * rv.pitch + sbx * n;
    if (two) r2 = umv(rv2, P.c2y, P.c2x) + (size_t)(sby * n) * rv2.pitch + sbx * n;
  }
  int d;
  if (n == 4) {
    unsigned rw[4];
#pragma unroll
    for (int y = 0; y < 4; y++) {
      rw[y] = ld4(r1 + (size_t)y * rv.pitch);
      if (P.form != JMB_PRED_PLAIN) rw[y] = pred_word(rw[y], two ? ld4(r2 + (size_t)y * rv2.pitch) : 0u, P);
    }
    d = dist4(src, rw, metric);
  } else {
    int a[64];
#pragma unroll
    for (int y = 0; y < 8; y++) {
      unsigned lo, hi, lo2 = 0, hi2 = 0;
      ld8(r1 + (size_t)y * rv.pitch, lo, hi);
      if (two) ld8(r2 + (size_t)y * rv2.pitch, lo2, hi2);
      if (P.form != JMB_PRED_PLAIN) { lo = pred_word(lo, lo2, P); hi = pred_word(hi, hi2, P); }
      unsigned s0 = src.w[2 * y], s1 = src.w[2 * y + 1];
      if (P.form == JMB_PRED_WEIGHTED_AVERAGE && y) {
        // computeBiPredSATD2's 8x8 branch never advances the source pointer past the eighth sample of a row
        // (me_distortion.c:1166), so row y reads JM's block-compact source copy y samples early -- reproduced
        s0 = s1 = 0;
        const int L0 = (sby * 8 + y) * bsx + sbx * 8 - y;
#pragma unroll
        for (int x = 0; x < 8; x++) {
          const int L = L0 + x, row = L / bsx, col = L - row * bsx;
          const unsigned v = cur[(size_t)(pos_y + row) * cur_pitch + pos_x + col];
          if (x < 4) s0 |= v << (8 * x); else s1 |= v << (8 * (x - 4));
        }
      }
#pragma unroll
      for (int x = 0; x < 4; x++) {
        a[y * 8 + x] = (int)((s0 >> (8 * x)) & 255) - (int)((lo >> (8 * x)) & 255);
        a[y * 8 + 4 + x] = (int)((s1 >> (8 * x)) & 255) - (int)((hi >> (8 * x)) & 255);
      }
    }
    d = hadamard8(a);
  }
  atomicAdd(&out[c], d);
}

// Mode-decision distortion back-ends (lencod/src/me_distortion.c:38-146): distortion4x4SAD / SSE / SATD and distortion8x8SAD /
// SADthres / SSE / SATD of DIFFERENCE blocks the caller formed (skip / direct / bi-predictive candidates: mv_search.c:589-675,
// :1159-1325, macroblock.c:1413, intra_chroma.c:443).  One thread per block.  thres > 0 reproduces distortion8x8SADthres:
// the row loop stops once the running sum exceeds thres and the partial sum is what JM returns.
__global__ void k_block_dist(const int16_t *__restrict__ diff, int nblk, int n, int metric, const int *__restrict__ thres, int *__restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nblk) return;
  const int16_t *d = diff + (size_t)t * n * n;
  int s = 0;
  if (metric == JMB_SATD) {
    if (n == 4) { int v[16]; for (int k = 0; k < 16; k++) v[k] = d[k]; s = hadamard4(v); }
    else { int v[64]; for (int k = 0; k < 64; k++) v[k] = d[k]; s = hadamard8(v); }
  } else if (metric == JMB_SSE) {
    for (int k = 0; k < n * n; k++) s += (int)d[k] * (int)d[k];
  } else {
    const int lim = thres ? thres[t] : 0x7fffffff;
    for (int j = 0; j < n; j++) {
      for (int i = 0; i < n; i++) s += abs((int)d[j * n + i]);
      if (n == 8 && thres && s > lim) break;
    }
  }
  out[t] = s;
}
}  // namespace

int jmb_launch_refine(jmb_ctx *ctx, const jmb_me_req *d_reqs, jmb_me_res *d_res, int n, const uint8_t *const *d_ref_planes) {
  const jmb_ref &r0 = ctx->refs[ctx->ref_list[0]];
  jmb_time_begin(ctx, JMB_K_REFINE);
  jmb_frame_gen gen; memset(&gen, 0, sizeof(gen));
  if (ctx->gen_pred) { gen.pred = ctx->gen_pred; gen.fp = ctx->gen_fp; gen.R = ctx->gen_R; gen.mb_w = ctx->gen_mb_w; }
  jmb_pack_out pack; pack.out = ctx->pack_out; pack.on = ctx->pack_on ? 1 : 0;
  k_subpel_refine<<<(n + RQ - 1) / RQ, RT, 0, ctx->stream>>>(d_reqs, d_res, n, ctx->cur, ctx->cur_pitch, d_ref_planes, r0.plane_bytes,
                                                        r0.pitch, ctx->cur_w, ctx->cur_h, ctx->me, ctx->nref, ctx->d_err, gen, pack);
  jmb_time_end(ctx, JMB_K_REFINE);
  JMB_LAUNCH_CHECK(ctx);
  return JMB_OK;
}

extern "C" int jmb_dist_ex(jmb_ctx *ctx, int ref, const jmb_dist_pred *pred, int metric, int blocktype, int pos_x, int pos_y,
                           const int16_t *cand_xy, int n, int test8x8, int32_t *out, int loc) {
  static const int bsx[8] = {0, 16, 16, 8, 8, 8, 4, 4}, bsy[8] = {0, 16, 8, 16, 8, 4, 8, 4};
  if (n <= 0) return JMB_OK;
  if (!ctx->cur || ref < 0 || ref >= ctx->nref) return jmb_fail(ctx, JMB_ERR_STATE, "jmb_dist: no picture / bad ref %d", ref);
  if (blocktype < 1 || blocktype > 7 || metric < 0 || metric > 2) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_dist: blocktype %d metric %d", blocktype, metric);
  if (pos_x < 0 || pos_y < 0 || (pos_x & 3) || (pos_y & 3) || pos_x + bsx[blocktype] > ctx->cur_w || pos_y + bsy[blocktype] > ctx->cur_h)
    return jmb_fail(ctx, JMB_ERR_ARG, "jmb_dist: block (%d,%d) type %d", pos_x, pos_y, blocktype);
  if (test8x8 && metric == JMB_SATD && blocktype > 4) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_dist: test8x8 needs blocktype <= 4");
  DistPred P{JMB_PRED_PLAIN, 0, 0, 0, 0, 0, 0, 0};
  int ref2 = ref;
  if (pred) {
    if (pred->form < JMB_PRED_PLAIN || pred->form > JMB_PRED_WEIGHTED_AVERAGE) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_dist_ex: prediction form %d", pred->form);
    const bool two = pred->form >= JMB_PRED_AVERAGE, wp = pred->form == JMB_PRED_WEIGHTED || pred->form == JMB_PRED_WEIGHTED_AVERAGE;
    if (two && (pred->ref2 < 0 || pred->ref2 >= ctx->nref)) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_dist_ex: bad ref2 %d", pred->ref2);
    if (two && (pred->cand2_x < -32768 || pred->cand2_x > 32767 || pred->cand2_y < -32768 || pred->cand2_y > 32767))
      return jmb_fail(ctx, JMB_ERR_ARG, "jmb_dist_ex: cand2 (%d,%d)", pred->cand2_x, pred->cand2_y);
    if (wp && (pred->log_weight_denom < 0 || pred->log_weight_denom > 7 || pred->weight1 < -32768 || pred->weight1 > 32767 ||
               pred->weight2 < -32768 || pred->weight2 > 32767 || pred->offset < -32768 || pred->offset > 32767 || pred->wp_round < 0 || pred->wp_round > 64))
      return jmb_fail(ctx, JMB_ERR_ARG, "jmb_dist_ex: weights (%d,%d) offset %d denom %d round %d", pred->weight1, pred->weight2, pred->offset,
                      pred->log_weight_denom, pred->wp_round);
    P.form = pred->form; P.c2x = pred->cand2_x; P.c2y = pred->cand2_y; P.w1 = pred->weight1; P.w2 = pred->weight2; P.off = pred->offset;
    P.shift = pred->log_weight_denom + (two ? 1 : 0); P.round = pred->wp_round * (two ? 2 : 1);
    if (two) ref2 = pred->ref2;
  }
  JMB_CUDA(ctx, cudaSetDevice(ctx->device));
  const jmb_ref &r = ctx->refs[ctx->ref_list[ref]], &rb = ctx->refs[ctx->ref_list[ref2]];
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
  RefView rv{r.planes, r.plane_bytes, r.pitch, r.w, r.h}, rv2{rb.planes, rb.plane_bytes, rb.pitch, rb.w, rb.h};
  jmb_time_begin(ctx, JMB_K_DIST);
  k_dist<<<(items + 127) / 128, 128, 0, ctx->stream>>>(ctx->cur, ctx->cur_pitch, rv, rv2, P, blocktype, pos_x, pos_y, d_c, n, metric, test8x8, d_o);
  jmb_time_end(ctx, JMB_K_DIST);
  JMB_LAUNCH_CHECK(ctx);
  if (loc == JMB_HOST) {
    JMB_CUDA(ctx, cudaMemcpyAsync(out, d_o, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    JMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return JMB_OK;
}

extern "C" int jmb_dist(jmb_ctx *ctx, int ref, int metric, int blocktype, int pos_x, int pos_y,
                        const int16_t *cand_xy, int n, int test8x8, int32_t *out, int loc) {
  return jmb_dist_ex(ctx, ref, nullptr, metric, blocktype, pos_x, pos_y, cand_xy, n, test8x8, out, loc);
}

extern "C" int jmb_block_distortion(jmb_ctx *ctx, int metric, int n, const int16_t *diff, int nblk, const int32_t *thres, int32_t *out, int loc) {
  if (nblk <= 0) return JMB_OK;
  if ((n != 4 && n != 8) || metric < JMB_SAD || metric > JMB_SATD || !diff || !out) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_block_distortion: n %d metric %d", n, metric);
  if (thres && (n != 8 || metric != JMB_SAD)) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_block_distortion: thresholds belong to the 8x8 SAD (distortion8x8SADthres)");
  JMB_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t db = (size_t)nblk * n * n * sizeof(int16_t), ob = (size_t)nblk * sizeof(int32_t);
  const int16_t *d_d = diff; const int *d_t = thres; int *d_o = out;
  if (jmb_is_host(loc)) {
    int rc = jmb_reserve_dev(ctx, &ctx->d_stage, &ctx->d_stage_cap, db + 2 * ob + 16); if (rc) return rc;
    char *a = (char *)ctx->d_stage;
    JMB_CUDA(ctx, cudaMemcpyAsync(a, diff, db, cudaMemcpyHostToDevice, ctx->stream));
    d_d = (const int16_t *)a; d_o = (int *)(a + ((db + 15) & ~(size_t)15));
    if (thres) { d_t = d_o + nblk; JMB_CUDA(ctx, cudaMemcpyAsync((void *)d_t, thres, ob, cudaMemcpyHostToDevice, ctx->stream)); }
  }
  jmb_time_begin(ctx, JMB_K_DIST);
  k_block_dist<<<(nblk + 127) / 128, 128, 0, ctx->stream>>>(d_d, nblk, n, metric, d_t, d_o);
  jmb_time_end(ctx, JMB_K_DIST);
  JMB_LAUNCH_CHECK(ctx);
  if (jmb_is_host(loc)) {
    JMB_CUDA(ctx, cudaMemcpyAsync(out, d_o, ob, cudaMemcpyDeviceToHost, ctx->stream));
    if (loc == JMB_HOST) JMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return JMB_OK;
}
