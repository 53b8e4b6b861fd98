// jmb_internal.h -- shared declarations of libjmb200 (not part of the C ABI).
#pragma once
#include <cuda.h>            // CUtensorMap (types only: the encoder is fetched with cudaGetDriverEntryPoint)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "jmb200.h"

#define JMB_PAD_X 32   // IMG_PAD_SIZE_X, lencod/inc/defines.h:121
#define JMB_PAD_Y 20   // IMG_PAD_SIZE_Y, lencod/inc/defines.h:122
#define JMB_MAX_SEARCH_RANGE 64   // largest SearchRange the search kernels index (k_search.cu IDX_BITS)

// One reference picture in HBM: 16 quarter-pel planes of u8 samples (bit depth 8), plane-major,
// each (h+40) rows of `pitch` bytes (pitch = (w+64) rounded up to 128 B so every row starts on a
// 128-byte line).  plane(fy,fx) = base + (fy*4+fx)*plane_bytes.
struct jmb_ref {
  uint8_t *planes = nullptr;
  int w = 0, h = 0, W = 0, H = 0, pitch = 0;
  size_t plane_bytes = 0;
  bool valid = false;
  uint8_t *chroma = nullptr; int wc = 0, hc = 0, pitch_c = 0;      // U plane then V plane, u8, hc rows of pitch_c bytes each
  CUtensorMap tmap_int;      // TMA descriptor of the integer plane [0][0]: u8, W x H, box = search-window tile
};

struct jmb_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  char err[512] = {0};
  uint64_t launches = 0;
  jmb_ref refs[JMB_MAX_REFS];
  // current picture (u8) and the reference list of this picture
  uint8_t *cur = nullptr; int cur_w = 0, cur_h = 0, cur_pitch = 0; size_t cur_cap = 0;
  uint8_t *cur_c = nullptr; int cur_wc = 0, cur_hc = 0, cur_pitch_c = 0; size_t cur_c_cap = 0;      // current picture's U, V planes
  CUtensorMap tmap_cur;      // TMA descriptor of the current picture: box = one 16x16 macroblock
  int ref_list[JMB_MAX_REFS]; int nref = 0;
  jmb_me_config me;
  bool me_configured = false;
  // staging: pinned host buffer + device scratch, grown on demand
  void *h_stage = nullptr; size_t h_stage_cap = 0;
  void *d_stage = nullptr; size_t d_stage_cap = 0;
  void *d_stage2 = nullptr; size_t d_stage2_cap = 0;
  void *d_groups = nullptr; size_t d_groups_cap = 0;
  void *h_groups = nullptr; size_t h_groups_cap = 0;
  // per-kernel event timing
  bool timing = false;
  struct EvPair { cudaEvent_t a, b; };
  EvPair *ev[16] = {nullptr}; int ev_n[16] = {0}; int ev_cap[16] = {0};
  void *d_reftab = nullptr; size_t d_reftab_cap = 0;
  void *d_qdesc = nullptr; size_t d_qdesc_cap = 0;
  void *d_stage3 = nullptr; size_t d_stage3_cap = 0;
  void *d_stage4 = nullptr; size_t d_stage4_cap = 0;
  void *d_stage5 = nullptr; size_t d_stage5_cap = 0;
  // context-resident intermediates (NULL arguments of jmb_pred_from_results / jmb_mc_tq refer to them)
  void *d_res_keep = nullptr; size_t d_res_keep_cap = 0; const jmb_me_res *last_res = nullptr; int last_res_n = 0;
  void *d_pred_keep = nullptr; size_t d_pred_keep_cap = 0; int pred_keep_n = 0;
  // request validation happens on the device (the requests may live there); the kernels OR a JMB_REQERR_* code and the
  // index of one offending request into d_err[0..1], which every synchronising call reads back
  int *d_err = nullptr; int *h_err = nullptr;
  bool smem_opt_in = false;   // k_int_search's dynamic shared memory opt-in done on this context's device
  // macroblock-resident SAD surfaces (jmb_mb_surfaces) per reference of the picture's list, and the host-mapped mailbox the
  // per-partition searches answer through (jmb_mb_search)
  struct Surf { void *buf = nullptr; size_t cap = 0; bool valid = false; int mb_x = 0, mb_y = 0, x0 = 0, y0 = 0, n = 0; unsigned long pic_serial = 0; };
  Surf surf[JMB_MAX_REFS];
  unsigned long pic_serial = 0, reftab_serial = 0, mb_calls = 0;      // pic_serial: bumped by every jmb_pic_begin / jmb_ref_put
  void *mbox = nullptr, *d_mbox = nullptr, *d_one = nullptr; int mbox_seq = 0;
  // picture form with device-generated requests / compact outputs
  void *d_mvpred = nullptr; size_t d_mvpred_cap = 0;
  void *d_res8 = nullptr; size_t d_res8_cap = 0;
  void *d_heads = nullptr; size_t d_heads_cap = 0;
  void *d_tokens = nullptr; size_t d_tokens_cap = 0;
  unsigned *d_tok_count = nullptr; unsigned *h_tok_count = nullptr;
  // picture form in flight (jmb_me_search_frame_pred): what the search / refinement kernels form their requests from, and where
  // the refinement kernel leaves the 8-byte results (k_search.cu sets them around its launches)
  const jmb_mb_mvpred *gen_pred = nullptr; jmb_frame_params gen_fp; int gen_R = 0, gen_mb_w = 0;
  jmb_me_res8 *pack_out = nullptr; bool pack_on = false;
  // deblocking: ticket counter + per-macroblock completion flags (k_deblock), stamped with db_serial
  void *d_db = nullptr; size_t d_db_cap = 0; int db_serial = 0;
  // peer buffers opened with jmb_peer_open (cudaIpcOpenMemHandle is expensive: one mapping per handle)
  struct Peer { unsigned char handle[JMB_IPC_HANDLE_BYTES]; void *mapped; };
  Peer peers[32]; int n_peers = 0;
};
int jmb_check_device_errors(jmb_ctx *ctx);   // after a stream synchronisation

int jmb_fail(jmb_ctx *ctx, int code, const char *fmt, ...);
// 2-D u8 tensor map (width x height samples, `pitch` bytes per row, box_w x box_h tile, zero fill outside)
int jmb_make_tmap_u8(jmb_ctx *ctx, CUtensorMap *out, const void *base, int width, int height, int pitch, int box_w, int box_h);
#define JMB_WIN_BOX_W 128   // search-window tile staged by TMA in k_int_search: 128 x 87 samples
#define JMB_WIN_BOX_H 87
int jmb_reserve_host(jmb_ctx *ctx, void **p, size_t *cap, size_t bytes);
int jmb_reserve_dev(jmb_ctx *ctx, void **p, size_t *cap, size_t bytes);

#define JMB_CUDA(ctx, call)                                                                     \
  do {                                                                                          \
    cudaError_t e_ = (call);                                                                    \
    if (e_ != cudaSuccess)                                                                      \
      return jmb_fail((ctx), JMB_ERR_CUDA, "%s:%d: %s -> %s", __FILE__, __LINE__, #call,        \
                      cudaGetErrorString(e_));                                                  \
  } while (0)

#define JMB_LAUNCH_CHECK(ctx)                                                                   \
  do {                                                                                          \
    (ctx)->launches++;                                                                          \
    JMB_CUDA((ctx), cudaGetLastError());                                                        \
  } while (0)

// kernel classes for the optional per-kernel CUDA-event timing (jmb_timing_enable / jmb_timing_get)
enum { JMB_K_SUBPEL = 0, JMB_K_PACK, JMB_K_INT_SEARCH, JMB_K_REFINE, JMB_K_DIST, JMB_K_FFS_SURF, JMB_K_FORWARD,
       JMB_K_QUANT, JMB_K_MC_TQ, JMB_K_PRED, JMB_K_GEN, JMB_K_EPZS, JMB_K_CHROMA, JMB_K_DEBLOCK, JMB_K_ARGMIN, JMB_K_COUNT };
static_assert(JMB_K_COUNT <= 16, "jmb_ctx::ev holds 16 kernel classes");
void jmb_time_begin(jmb_ctx *ctx, int kid);
void jmb_time_end(jmb_ctx *ctx, int kid);

// ---- device helpers shared by the kernels -----------------------------------------------------
__device__ __forceinline__ int jmb_clip(int lo, int hi, int v) { return min(max(v, lo), hi); }

// mvbits[] of lencod/src/mv_search.c:366-374: 1 for 0, else 2*floor(log2|v|)+3
__device__ __forceinline__ int jmb_mvbits(int v) {
  return 65 - 2 * __clz(abs(v));      // clz(0) = 32 -> 1; |v| in [2^k, 2^(k+1)) -> 2k + 3
}

// index of displacement (dx,dy) in JM's square spiral (lencod/src/mv_search.c:406-442)
__device__ __forceinline__ int jmb_spiral_index(int dx, int dy) {
  int l = max(abs(dx), abs(dy));
  if (l == 0) return 0;
  int base = (2 * l - 1) * (2 * l - 1);
  if (abs(dy) == l && abs(dx) < l) return base + 2 * (dx + l - 1) + (dy > 0);
  return base + 2 * (2 * l - 1) + 2 * (dy + l) + (dx > 0);
}

// request validation shared by the search and refinement kernels; 0 = fine
enum { JMB_REQERR_BLOCKTYPE = 1, JMB_REQERR_REF = 2, JMB_REQERR_POS = 4, JMB_REQERR_CENTER = 8, JMB_REQERR_MODE = 16,
       JMB_REQERR_LAMBDA = 32, JMB_REQERR_MINCOST = 64, JMB_REQERR_LAYOUT = 128, JMB_REQERR_FPEL_METRIC = 256 };
__device__ __forceinline__ int jmb_req_check(const jmb_me_req &r, int w, int h, int nref) {
  if (r.blocktype < 1 || r.blocktype > 7) return JMB_REQERR_BLOCKTYPE;
  const int bsx = (r.blocktype <= 2) ? 16 : (r.blocktype <= 5 ? 8 : 4);
  const int bsy = (r.blocktype == 1 || r.blocktype == 3) ? 16 : ((r.blocktype == 2 || r.blocktype == 4 || r.blocktype == 6) ? 8 : 4);
  int e = 0;
  if (r.ref >= nref) e |= JMB_REQERR_REF;
  if (r.pos_x < 0 || r.pos_y < 0 || r.pos_x + bsx > w || r.pos_y + bsy > h || (r.pos_x % bsx) || (r.pos_y % bsy)) e |= JMB_REQERR_POS;
  if (!(r.flags & JMB_REQ_SKIP_INT) && ((r.center_x | r.center_y) & 3)) e |= JMB_REQERR_CENTER;
  if (r.mode > JMB_SEARCH_FAST_FULL) e |= JMB_REQERR_MODE;
  if ((unsigned)r.lambda[0] > 65535u || (unsigned)r.lambda[1] > 65535u || (unsigned)r.lambda[2] > 65535u) e |= JMB_REQERR_LAMBDA;
  if (r.min_mcost < 0 || r.min_mcost > ((long long)1 << 48)) e |= JMB_REQERR_MINCOST;
  return e;
}
__device__ __forceinline__ void jmb_req_report(int *err, int code, int index) {
  if (code) { atomicOr(&err[0], code); err[1] = index; }
}

// ---- picture form: the requests of jmb_me_search_frame_pred are never written to memory ------------------------------------
// The search and refinement kernels form the request of (macroblock, partition in canonical order) from the predictor
// table and the slice's parameters where they would read it; gen.pred == nullptr means "requests come from memory".
struct jmb_frame_gen { const jmb_mb_mvpred *pred; jmb_frame_params fp; int R, mb_w; };
// The refinement kernel's last step for the picture form: final clip of the mv (mv_search.c:981) and the 8-byte results.
struct jmb_pack_out { jmb_me_res8 *out; int on; };

__device__ __forceinline__ jmb_me_req jmb_frame_request(const jmb_frame_gen &g, int mb, int p) {      // partition p (0..40) of macroblock mb
  // canonical partition order: by type, then raster order of the partitions inside the macroblock
  const int type = p < 1 ? 1 : p < 3 ? 2 : p < 5 ? 3 : p < 9 ? 4 : p < 17 ? 5 : p < 25 ? 6 : 7;
  const int first = type == 1 ? 0 : type == 2 ? 1 : type == 3 ? 3 : type == 4 ? 5 : type == 5 ? 9 : type == 6 ? 17 : 25;
  const int lx = (type <= 2) ? 0 : (type <= 5 ? 1 : 2);      // 1, 2 or 4 partitions across: 16, 8 or 4 samples wide
  const int bsy = (type == 1 || type == 3) ? 16 : ((type == 2 || type == 4 || type == 6) ? 8 : 4);
  const int k = p - first, row = k >> lx, col = k & ((1 << lx) - 1);
  const int mby = mb / g.mb_w, mbx = mb - mby * g.mb_w;
  const jmb_frame_params &fp = g.fp;
  const int px = g.pred[mb].pred[p][0], py = g.pred[mb].pred[p][1];
  jmb_me_req q;
  q.pos_x = (int16_t)(mbx * 16 + col * (16 >> lx)); q.pos_y = (int16_t)(mby * 16 + row * bsy);
  q.pred_x = (int16_t)px; q.pred_y = (int16_t)py;
  if (fp.mode == JMB_SEARCH_FAST_FULL) {      // one centre per macroblock: the rounded 16x16 predictor (me_fullfast.c:309-327)
    const int bx = g.pred[mb].pred[0][0], by = g.pred[mb].pred[0][1];
    q.center_x = (int16_t)jmb_clip(fp.mv_min_x + 4 * g.R, fp.mv_max_x - 4 * g.R, ((bx + 2) >> 2) * 4);
    q.center_y = (int16_t)jmb_clip(fp.mv_min_y + 4 * g.R, fp.mv_max_y - 4 * g.R, ((by + 2) >> 2) * 4);
  } else {                                    // mv_search.c:931-932, clip_mv_range :957
    q.center_x = (int16_t)jmb_clip(fp.mv_min_x, fp.mv_max_x, ((px + 2) >> 2) * 4);
    q.center_y = (int16_t)jmb_clip(fp.mv_min_y, fp.mv_max_y, ((py + 2) >> 2) * 4);
  }
  q.blocktype = (uint8_t)type; q.ref = (uint8_t)fp.ref; q.mode = (uint8_t)fp.mode;
  q.flags = (uint8_t)(fp.flags & (JMB_REQ_SUBPEL | (type <= 4 ? JMB_REQ_TEST8X8 : 0)));
  q.lambda[0] = fp.lambda[0]; q.lambda[1] = fp.lambda[1]; q.lambda[2] = fp.lambda[2];
  q.reserved_ = 0;
  q.min_mcost = (int64_t)0x7fffffff << 5;      // DISTBLK_MAX, lencod/inc/defines.h:136
  return q;
}

// kernels (defined in k_*.cu), launched through these host wrappers
int jmb_launch_subpel(jmb_ctx *ctx, const void *d_src, int sample_bytes, int src_stride, jmb_ref *r);
static inline bool jmb_is_host(int loc) { return loc == JMB_HOST || loc == JMB_HOST_ASYNC; }
