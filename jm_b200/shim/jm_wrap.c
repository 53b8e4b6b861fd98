/*
 * jm_wrap.c -- the reference-side binding of libjmb200: GNU ld --wrap stubs that sit behind JM 19.0
 * lencod's own call sites and forward them to the C ABI in include/jmb200.h.
 *
 * Compiled with JM's own headers and flags (gcc -std=gnu99 -fsigned-char ...) and linked with JM's
 * UNMODIFIED objects:
 *
 *   gcc -o lencod_jmb  <all lencod/lcommon objects>  jm_wrap.o  -ljmb200 -lm \
 *       -Wl,--wrap=getSubImagesLuma -Wl,--wrap=encode_one_slice \
 *       -Wl,--wrap=full_search_motion_estimation -Wl,--wrap=fast_full_search_motion_estimation \
 *       -Wl,--wrap=setup_fast_full_search -Wl,--wrap=sub_pel_motion_estimation \
 *       -Wl,--wrap=computeSAD -Wl,--wrap=computeSSE -Wl,--wrap=computeSATD \
 *       -Wl,--wrap=forward4x4 -Wl,--wrap=forward8x8 -Wl,--wrap=inverse4x4 -Wl,--wrap=inverse8x8 \
 *       -Wl,--wrap=quant_4x4_normal -Wl,--wrap=quant_4x4_around \
 *       -Wl,--wrap=quant_8x8_normal -Wl,--wrap=quant_8x8_around \
 *       -Wl,--wrap=quant_8x8cavlc_normal -Wl,--wrap=quant_8x8cavlc_around -Wl,--wrap=luma_residual_coding \
 *       -Wl,--wrap=quant_ac4x4_normal -Wl,--wrap=quant_ac4x4_around -Wl,--wrap=quant_dc4x4_normal \
 *       -Wl,--wrap=quant_dc2x2_normal -Wl,--wrap=quant_dc2x2_around -Wl,--wrap=quant_dc4x2_normal -Wl,--wrap=quant_dc4x2_around \
 *       -Wl,--wrap=hadamard4x4 -Wl,--wrap=ihadamard4x4 -Wl,--wrap=hadamard4x2 -Wl,--wrap=ihadamard4x2 \
 *       -Wl,--wrap=hadamard2x2 -Wl,--wrap=ihadamard2x2 \
 *       -Wl,--wrap=computeSADWP -Wl,--wrap=computeSSEWP -Wl,--wrap=computeSATDWP -Wl,--wrap=computeBiPredSAD1 -Wl,--wrap=computeBiPredSAD2 \
 *       -Wl,--wrap=computeBiPredSSE1 -Wl,--wrap=computeBiPredSSE2 -Wl,--wrap=computeBiPredSATD1 -Wl,--wrap=computeBiPredSATD2 \
 *       -Wl,--wrap=full_search_bipred_motion_estimation -Wl,--wrap=sub_pel_bipred_motion_estimation
 *
 * No JM source file is edited: every symbol above is defined in one translation unit and referenced
 * from another, so the linker redirects the reference to __wrap_<sym>.
 *
 * What runs where: the arithmetic of every wrapped function runs on the GPU; this file only converts
 * JM's row-pointer arrays and structs to the flat arguments of the ABI and back.  There is no CPU
 * path behind the ABI.  The ONLY use of __real_<sym> is the explicit plumbing switch
 * JMB_SHIM=passthrough (config 0 of BASELINE.json: proves the re-link itself is neutral) and the
 * per-family bisect switches JMB_SHIM_OFF=planes,me,subpel,tq -- both are debugging aids, off by default.
 * Unsupported configurations (field/MBAFF pictures, weighted-prediction ME, RDOptimization=0's (0,0)
 * bias, chroma ME, on-the-fly interpolation) stop the encoder the way JM's own error() does: message on stderr, exit.
 *
 * Granularity: JM decides one block at a time (each search's predictor depends on the previous
 * decisions), so this drop-in issues one small launch per leaf call.  It is the
 * bit-exactness proof on JM's real call sequence; the throughput path is the whole-picture form of the
 * same entry points (jmb_me_search_frame, jmb_mc_tq) that bench.py drives.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "global.h"
#include "mbuffer.h"
#include "mv_search.h"
#include "me_fullsearch.h"
#include "me_fullfast.h"
#include "img_luma.h"
#include "transform.h"
#include "quant4x4.h"
#include "quant8x8.h"
#include "quantChroma.h"
#include "me_distortion.h"
#include "mv_prediction.h"
#include "me_epzs.h"
#include "me_epzs_int.h"
#include "me_epzs_common.h"

#include "jmb200.h"

/* ---- the symbols the linker leaves reachable under their real names ------------------------------ */
void    __real_getSubImagesLuma(VideoParameters *p_Vid, StorablePicture *s);
int     __real_encode_one_slice(VideoParameters *p_Vid, int SliceGroupId, int TotalCodedMBs);
distblk __real_full_search_motion_estimation(Macroblock *, MotionVector *, MEBlock *, distblk, int);
distblk __real_fast_full_search_motion_estimation(Macroblock *, MotionVector *, MEBlock *, distblk, int);
void    __real_setup_fast_full_search(Macroblock *, MEBlock *, int);
distblk __real_sub_pel_motion_estimation(Macroblock *, MotionVector *, MEBlock *, distblk, int *);
void    __real_forward4x4(int **block, int **tblock, int pos_y, int pos_x);
void    __real_forward8x8(int **block, int **tblock, int pos_y, int pos_x);
void    __real_inverse4x4(int **tblock, int **block, int pos_y, int pos_x);
void    __real_inverse8x8(int **tblock, int **block, int pos_x);
int     __real_quant_4x4_normal(Macroblock *, int **, struct quant_methods *);
int     __real_quant_4x4_around(Macroblock *, int **, struct quant_methods *);
int     __real_quant_8x8_normal(Macroblock *, int **, struct quant_methods *);
int     __real_quant_8x8_around(Macroblock *, int **, struct quant_methods *);
int     __real_quant_8x8cavlc_normal(Macroblock *, int **, struct quant_methods *, int ***);
int     __real_quant_8x8cavlc_around(Macroblock *, int **, struct quant_methods *, int ***);
void    __real_luma_residual_coding(Macroblock *currMB);
void    __real_chroma_residual_coding(Macroblock *currMB);
distblk __real_computeSADWP(StorablePicture *, MEBlock *, distblk, MotionVector *);
distblk __real_computeSSEWP(StorablePicture *, MEBlock *, distblk, MotionVector *);
distblk __real_computeSATDWP(StorablePicture *, MEBlock *, distblk, MotionVector *);
distblk __real_computeBiPredSAD1(StorablePicture *, StorablePicture *, MEBlock *, distblk, MotionVector *, MotionVector *);
distblk __real_computeBiPredSAD2(StorablePicture *, StorablePicture *, MEBlock *, distblk, MotionVector *, MotionVector *);
distblk __real_computeBiPredSSE1(StorablePicture *, StorablePicture *, MEBlock *, distblk, MotionVector *, MotionVector *);
distblk __real_computeBiPredSSE2(StorablePicture *, StorablePicture *, MEBlock *, distblk, MotionVector *, MotionVector *);
distblk __real_computeBiPredSATD1(StorablePicture *, StorablePicture *, MEBlock *, distblk, MotionVector *, MotionVector *);
distblk __real_computeBiPredSATD2(StorablePicture *, StorablePicture *, MEBlock *, distblk, MotionVector *, MotionVector *);
distblk __real_full_search_bipred_motion_estimation(Macroblock *, int, MotionVector *, MotionVector *, MotionVector *, MotionVector *, MEBlock *, int, distblk, int);
distblk __real_sub_pel_bipred_motion_estimation(Macroblock *, MEBlock *, int, MotionVector *, MotionVector *, MotionVector *, MotionVector *, distblk, int *);
int     __real_quant_ac4x4_normal(Macroblock *, int **, struct quant_methods *);
int     __real_quant_ac4x4_around(Macroblock *, int **, struct quant_methods *);
int     __real_quant_dc4x4_normal(Macroblock *, int **, int, int *, int *, LevelQuantParams *, const byte (*)[2]);
int     __real_quant_dc2x2_normal(Macroblock *, int **, int, int *, int *, LevelQuantParams *, int **, const byte (*)[2]);
int     __real_quant_dc2x2_around(Macroblock *, int **, int, int *, int *, LevelQuantParams *, int **, const byte (*)[2]);
int     __real_quant_dc4x2_normal(Macroblock *, int **, int, int *, int *, LevelQuantParams *, int **, const byte (*)[2]);
int     __real_quant_dc4x2_around(Macroblock *, int **, int, int *, int *, LevelQuantParams *, int **, const byte (*)[2]);
void    __real_hadamard4x4(int **, int **);
void    __real_ihadamard4x4(int **, int **);
void    __real_hadamard4x2(int **, int **);
void    __real_ihadamard4x2(int **, int **);
void    __real_hadamard2x2(int **, int *);
void    __real_ihadamard2x2(int *, int *);
distblk __real_EPZS_integer_motion_estimation(Macroblock *, MotionVector *, MEBlock *, distblk, int);
distblk __real_EPZS_sub_pel_motion_estimation(Macroblock *, MotionVector *, MEBlock *, distblk, int *);
void    __real_select_distortion(VideoParameters *p_Vid, InputParameters *p_Inp);
void    __real_DeblockFrame(VideoParameters *p_Vid, imgpel **imgY, imgpel ***imgUV);
distblk __real_computeSAD(StorablePicture *, MEBlock *, distblk, MotionVector *);
distblk __real_computeSSE(StorablePicture *, MEBlock *, distblk, MotionVector *);
distblk __real_computeSATD(StorablePicture *, MEBlock *, distblk, MotionVector *);


enum { FAM_PLANES = 1, FAM_ME = 2, FAM_SUBPEL = 4, FAM_TQ = 8, FAM_DIST = 16, FAM_DEBLOCK = 32 };
enum { AHEAD_N = 8 * JMB_CHAIN_MAX };      /* answers kept ahead of JM: a call's worth for each of several references */

static struct
{
  int              init;              /* 0 = not yet, 1 = GPU, 2 = passthrough */
  int              off;               /* FAM_* families handed back to JM (bisecting only) */
  jmb_ctx         *ctx;
  StorablePicture *slot_pic[JMB_MAX_REFS];
  unsigned long    slot_stamp[JMB_MAX_REFS];
  unsigned long    stamp;
  int              pic_valid;         /* jmb_pic_begin done for the current slice and reference set */
  int              list[JMB_MAX_REFS], nlist;
  jmb_me_config    cfg;
  int              cfg_valid;
  unsigned long    calls[10];
  int              verify;            /* JMB_SHIM_VERIFY=1: differential check of the device luma_residual_coding */
  unsigned long    verified;
  /* surfaces resident on the device for the full search (jmb_mb_surfaces): macroblock and centre per device reference */
  struct { int valid, mb_x, mb_y, cx, cy, radius; unsigned long pic; } surf[JMB_MAX_REFS];
  unsigned long    pic_count;         /* bumped whenever the current picture / reference set is (re)sent */
  /* the sub-pel refinement computed speculatively with the integer search of the same block (one device call instead of two) */
  struct { int valid, pos_x, pos_y, blocktype, list, ref_idx, test8x8, lambda_h, lambda_q; MotionVector pred, imv, mv; distblk min_in, cost; } spec;
  unsigned long    spec_hits, surf_builds;
  /* answers the device gave AHEAD of JM's call sequence (jmb_mb_chain): each is handed out only when JM arrives at that block
   * with exactly the predictor, centre, bound and lambda the device assumed */
  struct { int valid, pos_x, pos_y, blocktype, list, ref_idx, test8x8, lambda, mode, R; unsigned long pic; MotionVector pred, center; jmb_me_res res; } ahead[AHEAD_N];
  int              chains_off;        /* JMB_SHIM_CHAIN=0: one device call per search (A/B of the run-ahead) */
  unsigned long    chain_calls, chain_hits, chain_stale;
  unsigned long    deblocked, deblock_verified;
  int              rc_device;         /* JMB_SHIM_RC=device: luma_residual_coding of eligible inter macroblocks IS the device call (results written into JM's state) */
  unsigned long    rc_written;
} S;

/* JM's convention for fatal conditions is error(text, code) -> message on stderr, exit(code) (lencod/src/lencod.c).
 * JM's error() also flushes the DPB first, which re-enters the wrapped functions; the shim therefore prints the same
 * way and exits itself. */
static void fatal(int code)
{
  fprintf(stderr, "%s\n", errortext);
  exit(code);
}

static void jmb_die(const char *what, int rc)
{
  snprintf(errortext, ET_SIZE, "libjmb200 shim: %s failed (%d): %s", what, rc, jmb_last_error(S.ctx));
  fatal(700);
}

static void unsupported(const char *what)
{
  snprintf(errortext, ET_SIZE, "libjmb200 shim: %s is not supported by the GPU path (and there is no CPU fallback)", what);
  fatal(701);
}

static unsigned long chroma_verified;
static void report(void)
{
  if (S.init == 1 && S.verify)
    fprintf(stderr, "[jmb shim] luma_residual_coding verified on %lu macroblock codings (levels, cbp, cbp_blk, reconstruction); chroma_residual_coding verified on %lu; DeblockFrame verified on %lu pictures (every sample)\n", S.verified, chroma_verified, S.deblock_verified);
  if (S.init == 1 && getenv("JMB_SHIM_VERBOSE"))
    fprintf(stderr, "[jmb shim] planes %lu  full %lu  fastfull %lu  subpel %lu  fwd4 %lu  fwd8 %lu  quant4 %lu  quant8 %lu  dist %lu  kernel launches %llu  bipred/weighted distortions %lu  subpel served by the integer search's call %lu  surface builds %lu  chain calls %lu  searches answered ahead %lu  (discarded %lu)  pictures deblocked %lu (verified %lu)  luma_residual_coding answered by the device %lu\n",
            S.calls[0], S.calls[1], S.calls[2], S.calls[3], S.calls[4], S.calls[5], S.calls[6], S.calls[7], S.calls[8],
            (unsigned long long)jmb_launch_count(S.ctx), S.calls[9], S.spec_hits, S.surf_builds, S.chain_calls, S.chain_hits, S.chain_stale, S.deblocked, S.deblock_verified, S.rc_written);
}

static int shim_on(int family)
{
  if (!S.init)
  {
    const char *m = getenv("JMB_SHIM"), *off = getenv("JMB_SHIM_OFF");
    if (m && !strcmp(m, "passthrough"))
      S.init = 2;
    else
    {
      int dev = getenv("JMB_DEVICE") ? atoi(getenv("JMB_DEVICE")) : 0;
      int rc = jmb_create(dev, &S.ctx);
      if (rc)
      {
        snprintf(errortext, ET_SIZE, "libjmb200 shim: jmb_create(%d) failed (%d): %s", dev, rc, jmb_last_error(NULL));
        fatal(700);
      }
      S.init = 1;
      S.verify = getenv("JMB_SHIM_VERIFY") != NULL;
      S.chains_off = getenv("JMB_SHIM_CHAIN") && !strcmp(getenv("JMB_SHIM_CHAIN"), "0");
      S.rc_device = getenv("JMB_SHIM_RC") && !strcmp(getenv("JMB_SHIM_RC"), "device");
      if (off)
      {
        if (strstr(off, "planes")) S.off |= FAM_PLANES;
        if (strstr(off, "me"))     S.off |= FAM_ME;
        if (strstr(off, "subpel")) S.off |= FAM_SUBPEL;
        if (strstr(off, "tq"))     S.off |= FAM_TQ;
        if (strstr(off, "dist"))   S.off |= FAM_DIST;
        if (strstr(off, "deblock")) S.off |= FAM_DEBLOCK;
      }
      atexit(report);
    }
  }
  return S.init == 1 && !(S.off & family);
}

/* ---- reference pictures ------------------------------------------------------------------------------
 * key = StorablePicture*; a slot is recycled least-recently-registered first, and a picture whose slot
 * was recycled is simply uploaded again when a search names it (JM owns the samples, we only cache). */
static int put_ref(VideoParameters *p_Vid, StorablePicture *s)
{
  int i, slot = -1;
  for (i = 0; i < JMB_MAX_REFS; i++)
    if (S.slot_pic[i] == s) { slot = i; break; }
  if (slot < 0)
    for (i = 0; i < JMB_MAX_REFS; i++)
      if (!S.slot_pic[i]) { slot = i; break; }
  if (slot < 0)
  {
    slot = 0;
    for (i = 1; i < JMB_MAX_REFS; i++)
      if (S.slot_stamp[i] < S.slot_stamp[slot]) slot = i;
  }
  {
    imgpel **img = s->p_curr_img;
    int stride = (int)(img[1] - img[0]);
    int rc = jmb_ref_put(S.ctx, slot, (const uint16_t *)&img[0][0], s->size_x, s->size_y, stride, p_Vid->bitdepth_luma, JMB_HOST);
    if (rc) jmb_die("jmb_ref_put", rc);
    if (S.verify && (p_Vid->yuv_format == YUV420 || p_Vid->yuv_format == YUV422) && s->imgUV && p_Vid->bitdepth_chroma == 8)
    { /* the chroma path is checked differentially (JMB_SHIM_VERIFY): it needs the reference's chroma planes too */
      rc = jmb_ref_put_chroma(S.ctx, slot, &s->imgUV[0][0][0], &s->imgUV[1][0][0], (int)sizeof(imgpel), s->size_x_cr, s->size_y_cr,
                              (int)(s->imgUV[0][1] - s->imgUV[0][0]), JMB_HOST);
      if (rc) jmb_die("jmb_ref_put_chroma", rc);
    }
  }
  S.slot_pic[slot] = s;
  S.slot_stamp[slot] = ++S.stamp;
  S.pic_valid = 0;
  return slot;
}

/* stands behind getSubImagesLuma (lencod/src/img_luma.c:611), called from UnifiedOneForthPix (image.c:2187):
 * the 16 quarter-pel planes are built on the device, kept there for the searches, and copied into JM's own
 * imgY_sub planes because JM's untouched host code (luma_prediction, mode decision) reads them. */
void __wrap_getSubImagesLuma(VideoParameters *p_Vid, StorablePicture *s)
{
  int fy, fx, slot;
  if (!shim_on(FAM_PLANES)) { __real_getSubImagesLuma(p_Vid, s); return; }
  if (p_Vid->p_Inp->OnTheFlyFractMCP) unsupported("OnTheFlyFractMCP");
  if (s->p_curr_img != s->imgY || s->p_curr_img_sub != s->imgY_sub)
  { /* 4:4:4 joint coding interpolates U and V through this function too: not part of the luma ME path */
    unsupported("getSubImagesLuma on a chroma plane (4:4:4 joint)");
  }
  S.calls[0]++;
  slot = put_ref(p_Vid, s);
  for (fy = 0; fy < 4; fy++)
    for (fx = 0; fx < 4; fx++)
    { /* each plane is one calloc of (h+40) x (w+64) imgpel with biased row pointers (memalloc.c:881-904) */
      imgpel *flat = &s->imgY_sub[fy][fx][-IMG_PAD_SIZE_Y][-IMG_PAD_SIZE_X];
      int rc = jmb_ref_get_plane(S.ctx, slot, fy, fx, (uint16_t *)flat, JMB_HOST);
      if (rc) jmb_die("jmb_ref_get_plane", rc);
    }
}

/* a new slice: the current picture must be (re)sent before its first search */
int __wrap_encode_one_slice(VideoParameters *p_Vid, int SliceGroupId, int TotalCodedMBs)
{
  S.pic_valid = 0;
  return __real_encode_one_slice(p_Vid, SliceGroupId, TotalCodedMBs);
}

static int ref_index_of(VideoParameters *p_Vid, Slice *currSlice, MEBlock *mv_block, StorablePicture *ref_picture)
{
  int i, slot = -1;

  if (p_Vid->structure != FRAME || currSlice->mb_aff_frame_flag) unsupported("field / MBAFF picture");
  if (mv_block->ChromaMEEnable) unsupported("ChromaMEEnable");
  if (!ref_picture) unsupported("a search without a reference picture");

  for (i = 0; i < JMB_MAX_REFS; i++)
    if (S.slot_pic[i] == ref_picture) { slot = i; break; }
  if (slot < 0) slot = put_ref(p_Vid, ref_picture);        /* its slot was recycled: upload again */
  else S.slot_stamp[slot] = ++S.stamp;

  if (!S.pic_valid)
  {
    imgpel **cur = p_Vid->pCurImg;
    int rc;
    S.nlist = 0;
    for (i = 0; i < JMB_MAX_REFS; i++)
      if (S.slot_pic[i]) S.list[S.nlist++] = i;
    rc = jmb_pic_begin(S.ctx, (const uint16_t *)&cur[0][0], p_Vid->width, p_Vid->height, (int)(cur[1] - cur[0]), JMB_HOST, S.list, S.nlist);
    if (rc) jmb_die("jmb_pic_begin", rc);
    S.pic_valid = 1;
    S.pic_count++;
    S.spec.valid = 0;
    if (S.verify && (p_Vid->yuv_format == YUV420 || p_Vid->yuv_format == YUV422) && p_Vid->bitdepth_chroma == 8)
    {
      imgpel **cu = p_Vid->pImgOrg[1], **cv = p_Vid->pImgOrg[2];
      rc = jmb_pic_chroma(S.ctx, &cu[0][0], &cv[0][0], (int)sizeof(imgpel), p_Vid->width_cr, p_Vid->height_cr, (int)(cu[1] - cu[0]), JMB_HOST);
      if (rc) jmb_die("jmb_pic_chroma", rc);
    }
  }
  for (i = 0; i < S.nlist; i++)
    if (S.list[i] == slot) return i;
  snprintf(errortext, ET_SIZE, "libjmb200 shim: reference picture not in the device list");
  fatal(702);
  return -1;
}

static int ref_index(Macroblock *currMB, MEBlock *mv_block)
{
  Slice *currSlice = currMB->p_Slice;
  return ref_index_of(currMB->p_Vid, currSlice, mv_block, currSlice->listX[mv_block->list + currMB->list_offset][(int)mv_block->ref_idx]);
}

static void configure(Macroblock *currMB, MEBlock *mv_block, int search_range)
{
  VideoParameters *p_Vid = currMB->p_Vid;
  InputParameters *p_Inp = currMB->p_Inp;
  jmb_me_config c;
  memset(&c, 0, sizeof(c));
  c.search_range = search_range > 0 ? search_range : (S.cfg_valid ? S.cfg.search_range : p_Inp->search_range[0]);
  c.max_mvd = p_Vid->max_mvd;
  c.metric[0] = p_Inp->MEErrorMetric[F_PEL];
  c.metric[1] = p_Inp->MEErrorMetric[H_PEL];
  c.metric[2] = p_Inp->MEErrorMetric[Q_PEL];
  c.start_hp = p_Vid->start_me_refinement_hp;
  c.start_qp = p_Vid->start_me_refinement_qp;
  c.search_pos2 = mv_block->search_pos2;
  c.search_pos4 = mv_block->search_pos4;
  if (!S.cfg_valid || memcmp(&c, &S.cfg, sizeof(c)))
  {
    int rc = jmb_me_configure(S.ctx, &c);
    if (rc) jmb_die("jmb_me_configure", rc);
    S.cfg = c;
    S.cfg_valid = 1;
  }
}

static void fill_request(jmb_me_req *q, MEBlock *mv_block, MotionVector *pred_mv, int ref, distblk min_mcost)
{
  memset(q, 0, sizeof(*q));
  q->pos_x = mv_block->pos_x;
  q->pos_y = mv_block->pos_y;
  q->pred_x = pred_mv->mv_x;
  q->pred_y = pred_mv->mv_y;
  q->blocktype = (uint8_t)mv_block->blocktype;
  q->ref = (uint8_t)ref;
  q->min_mcost = min_mcost;
}

/* the sub-pel refinement that came with an integer search's answer, kept for the SubPelME call that follows */
static void note_spec(Macroblock *currMB, MotionVector *pred_mv, MEBlock *mv_block, int lambda_factor, int spec, const jmb_me_res *r)
{
  VideoParameters *p_Vid = currMB->p_Vid;
  S.spec.valid = spec;
  S.spec.pos_x = mv_block->pos_x; S.spec.pos_y = mv_block->pos_y; S.spec.blocktype = mv_block->blocktype;
  S.spec.list = mv_block->list; S.spec.ref_idx = mv_block->ref_idx; S.spec.test8x8 = mv_block->test8x8;
  S.spec.lambda_h = S.spec.lambda_q = lambda_factor;
  S.spec.pred = *pred_mv;
  S.spec.imv.mv_x = r->imv_x; S.spec.imv.mv_y = r->imv_y;
  S.spec.mv.mv_x = r->mv_x;   S.spec.mv.mv_y = r->mv_y;
  S.spec.min_in = p_Vid->start_me_refinement_hp ? (distblk)r->icost : DISTBLK_MAX;      /* what BlockMotionSearch hands to SubPelME */
  S.spec.cost = (distblk)r->cost;
}

/* One partition's search + (speculatively) its sub-pel refinement in ONE device call over the macroblock's resident surfaces.
 * BlockMotionSearch calls SubPelME right after IntPelME with the same block, predictor and -- in every configuration JM
 * ships -- the same lambda at all three levels (mv_search.c:960-976); the refinement is therefore computed along with the
 * integer search and handed out when SubPelME asks for exactly that (anything else: a device call of its own). */
static int mb_search(Macroblock *currMB, MotionVector *pred_mv, MEBlock *mv_block, distblk min_mcost, int lambda_factor, int mode,
                     int ri, int center_x, int center_y, jmb_me_res *r)
{
  VideoParameters *p_Vid = currMB->p_Vid;
  jmb_me_req q;
  int rc, spec = !currMB->p_Inp->DisableSubpelME[p_Vid->view_id];
  fill_request(&q, mv_block, pred_mv, ri, min_mcost);
  q.center_x = (int16_t)center_x;
  q.center_y = (int16_t)center_y;
  q.mode = (uint8_t)mode;
  q.flags = (uint8_t)(spec ? (JMB_REQ_SUBPEL | (mv_block->test8x8 ? JMB_REQ_TEST8X8 : 0)) : 0);
  q.lambda[0] = q.lambda[1] = q.lambda[2] = lambda_factor;
  rc = jmb_mb_search(S.ctx, &q, r);
  if (rc) return rc;
  note_spec(currMB, pred_mv, mv_block, lambda_factor, spec, r);
  return 0;
}

/* ---- running ahead of JM's call sequence --------------------------------------------------------------------------------
 * JM searches the partitions of a macroblock one BlockMotionSearch at a time because each block's predictor is the median of
 * its neighbours' vectors, and some neighbours are blocks searched just before (mv_search.c:1560-1850).  Nothing else ties
 * them together, so at the FIRST search of a region -- the 16x16 block (modes 1-3 follow), or the 8x8 block of a quadrant
 * (its sub-modes 5-7 follow) -- the shim describes all searches of the region to the device in ONE call (jmb_mb_chain):
 * geometry, and per block the three neighbours as JM's own get_neighbors finds them, either by value (mv_info as it stands)
 * or by naming the earlier search of the call that will have written that position.  The device derives predictor and centre
 * as BlockMotionSearch does and searches block after block.  When JM arrives at a later block, the stored answer is used only
 * if JM's own predictor, centre, bound, lambda and range are what the device assumed (the assumption that can fail: with
 * several references JM may settle the first block of a 16x8 / 8x16 pair on another reference); otherwise it is discarded
 * and the block is searched by a call of its own.  Either way every number JM sees is the one it would have computed. */
static const struct { int type, x, y, chain; } CHAIN_WHOLE[5] = {{1, 0, 0, 0}, {2, 0, 0, 1}, {2, 0, 8, 1}, {3, 0, 0, 2}, {3, 8, 0, 2}};
static const struct { int type, x, y, chain; } CHAIN_QUAD[9] = {{4, 0, 0, 0}, {5, 0, 0, 1}, {5, 0, 4, 1}, {6, 0, 0, 2}, {6, 4, 0, 2},
                                                                {7, 0, 0, 3}, {7, 4, 0, 3}, {7, 0, 4, 3}, {7, 4, 4, 3}};
static const int CHAIN_BSX[8] = {0, 16, 16, 8, 8, 8, 4, 4}, CHAIN_BSY[8] = {0, 16, 8, 16, 8, 4, 8, 4};

static int ahead_take(Macroblock *currMB, MotionVector *pred_mv, MEBlock *mv_block, distblk min_mcost, int lambda_factor, int mode, int R,
                      int center_x, int center_y, jmb_me_res *r)
{
  int i;
  for (i = 0; i < AHEAD_N; i++)
    if (S.ahead[i].valid && S.ahead[i].pos_x == mv_block->pos_x && S.ahead[i].pos_y == mv_block->pos_y && S.ahead[i].blocktype == mv_block->blocktype &&
        S.ahead[i].list == mv_block->list && S.ahead[i].ref_idx == mv_block->ref_idx)
    {
      S.ahead[i].valid = 0;
      if (S.ahead[i].pic == S.pic_count && S.ahead[i].test8x8 == mv_block->test8x8 && S.ahead[i].lambda == lambda_factor && S.ahead[i].mode == mode &&
          S.ahead[i].R == R && min_mcost == DISTBLK_MAX && S.ahead[i].pred.mv_x == pred_mv->mv_x && S.ahead[i].pred.mv_y == pred_mv->mv_y &&
          S.ahead[i].center.mv_x == center_x && S.ahead[i].center.mv_y == center_y)
      {
        *r = S.ahead[i].res;
        S.chain_hits++;
        note_spec(currMB, pred_mv, mv_block, lambda_factor, !currMB->p_Inp->DisableSubpelME[currMB->p_Vid->view_id], r);
        return 1;
      }
      S.chain_stale++;
    }
  return 0;
}

/* the region that starts with this block, in one device call; 1 = the block's own answer is in *r */
static int ahead_run(Macroblock *currMB, MotionVector *pred_mv, MEBlock *mv_block, int lambda_factor, int mode, int R, int ri,
                     int center_x, int center_y, jmb_me_res *r)
{
  VideoParameters *p_Vid = currMB->p_Vid;
  InputParameters *p_Inp = currMB->p_Inp;
  Slice *currSlice = currMB->p_Slice;
  PicMotionParams **mv_info = p_Vid->enc_picture->mv_info;
  jmb_chain_req q[JMB_CHAIN_MAX];
  jmb_chain_res o[JMB_CHAIN_MAX];
  int32_t lim[4];
  int n, i, j, k, rc, list = mv_block->list, spec = !p_Inp->DisableSubpelME[p_Vid->view_id];
  int mbx = mv_block->pos_x & ~15, mby = mv_block->pos_y & ~15, qx = mv_block->pos_x & 15, qy = mv_block->pos_y & 15;
  if (S.chains_off || currSlice->mb_aff_frame_flag || p_Inp->DisableMEPrediction || !p_Inp->rdopt || currSlice->rdoq_motion_copy == 1) return 0;
  if (mv_block->blocktype == 1) n = 5;
  else if (mv_block->blocktype == 4) n = 9;
  else return 0;
  memset(q, 0, sizeof(q));
  for (i = 0; i < n; i++)
  {
    int type = n == 5 ? CHAIN_WHOLE[i].type : CHAIN_QUAD[i].type;
    int x = n == 5 ? CHAIN_WHOLE[i].x : qx + CHAIN_QUAD[i].x, y = n == 5 ? CHAIN_WHOLE[i].y : qy + CHAIN_QUAD[i].y;
    PixelPos block[4];
    q[i].req.pos_x = (int16_t)(mbx + x); q[i].req.pos_y = (int16_t)(mby + y);
    q[i].req.blocktype = (uint8_t)type; q[i].req.ref = (uint8_t)ri; q[i].req.mode = (uint8_t)mode;
    q[i].req.flags = (uint8_t)(spec ? (JMB_REQ_SUBPEL | ((p_Inp->Transform8x8Mode && type <= 4) ? JMB_REQ_TEST8X8 : 0)) : 0);      /* mv_search.c:1630,1769 */
    q[i].req.lambda[0] = q[i].req.lambda[1] = q[i].req.lambda[2] = lambda_factor;
    q[i].req.min_mcost = DISTBLK_MAX;
    q[i].req.center_x = (int16_t)center_x; q[i].req.center_y = (int16_t)center_y;      /* FAST_FULL: one centre per macroblock and reference; FULL: derived */
    q[i].jm_ref = (int8_t)mv_block->ref_idx;
    q[i].chain = (int8_t)(n == 5 ? CHAIN_WHOLE[i].chain : CHAIN_QUAD[i].chain);
    get_neighbors(currMB, block, x, y, CHAIN_BSX[type]);
    for (k = 0; k < 3; k++)
    {
      jmb_chain_nb *nb = &q[i].nb[k];
      nb->available = (int8_t)(block[k].available != 0);
      nb->dep = -1;
      if (!block[k].available) continue;
      for (j = 0; j < i; j++)      /* a position inside an earlier block of the same chain: that search's vector will stand there */
        if (q[j].chain == q[i].chain && block[k].pos_x * 4 >= q[j].req.pos_x && block[k].pos_x * 4 < q[j].req.pos_x + CHAIN_BSX[q[j].req.blocktype] &&
            block[k].pos_y * 4 >= q[j].req.pos_y && block[k].pos_y * 4 < q[j].req.pos_y + CHAIN_BSY[q[j].req.blocktype])
          nb->dep = (int8_t)j;
      if (nb->dep >= 0) nb->ref_idx = (int8_t)mv_block->ref_idx;
      else
      {
        nb->ref_idx = (int8_t)mv_info[block[k].pos_y][block[k].pos_x].ref_idx[list];
        nb->mv_x = mv_info[block[k].pos_y][block[k].pos_x].mv[list].mv_x;
        nb->mv_y = mv_info[block[k].pos_y][block[k].pos_x].mv[list].mv_y;
      }
    }
  }
  lim[0] = p_Vid->MaxHmvR[4]; lim[1] = p_Vid->MaxHmvR[5]; lim[2] = p_Vid->MaxVmvR[4]; lim[3] = p_Vid->MaxVmvR[5];      /* clip_mv_range(.., Q_PEL), conformance.c:640 */
  rc = jmb_mb_chain(S.ctx, q, n, lim, JM_INT_DIVIDE, o);
  if (rc) jmb_die("jmb_mb_chain", rc);
  S.chain_calls++;
  /* what is still waiting for this list and reference (or for another macroblock / picture) will not be asked for any more */
  for (i = 0; i < AHEAD_N; i++)
    if (S.ahead[i].valid && ((S.ahead[i].list == list && S.ahead[i].ref_idx == mv_block->ref_idx) || S.ahead[i].pic != S.pic_count ||
                             (S.ahead[i].pos_x & ~15) != mbx || (S.ahead[i].pos_y & ~15) != mby))
    { S.ahead[i].valid = 0; S.chain_stale++; }
  for (i = 1, j = 0; i < n; i++)
    if (o[i].status == JMB_CHAIN_DONE)
    {
      while (j < AHEAD_N && S.ahead[j].valid) j++;
      if (j == AHEAD_N) break;
      S.ahead[j].valid = 1; S.ahead[j].pos_x = q[i].req.pos_x; S.ahead[j].pos_y = q[i].req.pos_y; S.ahead[j].blocktype = q[i].req.blocktype;
      S.ahead[j].list = list; S.ahead[j].ref_idx = mv_block->ref_idx; S.ahead[j].test8x8 = (q[i].req.flags & JMB_REQ_TEST8X8) != 0;
      S.ahead[j].lambda = lambda_factor; S.ahead[j].mode = mode; S.ahead[j].R = R; S.ahead[j].pic = S.pic_count;
      S.ahead[j].pred.mv_x = o[i].pred_x; S.ahead[j].pred.mv_y = o[i].pred_y;
      S.ahead[j].center.mv_x = o[i].center_x; S.ahead[j].center.mv_y = o[i].center_y;
      S.ahead[j].res = o[i].res;
    }
  /* the block JM asked for: the device must have derived JM's own predictor and centre for it */
  if (o[0].status != JMB_CHAIN_DONE || o[0].pred_x != pred_mv->mv_x || o[0].pred_y != pred_mv->mv_y || o[0].center_x != center_x || o[0].center_y != center_y ||
      ((q[0].req.flags & JMB_REQ_TEST8X8) != 0) != (mv_block->test8x8 != 0))
    return 0;
  *r = o[0].res;
  note_spec(currMB, pred_mv, mv_block, lambda_factor, spec, r);
  return 1;
}

/* stands behind full_search_motion_estimation (lencod/src/me_fullsearch.c:39) = currMB->IntPelME for SearchMode -1.
 * The 4x4 SADs of the macroblock are computed once (jmb_mb_surfaces, a window E pels wider than the search window around the
 * first partition's centre) and every partition's arg-min runs over them; a partition whose own window leaves that area gets
 * fresh surfaces around its centre. */
distblk __wrap_full_search_motion_estimation(Macroblock *currMB, MotionVector *pred_mv, MEBlock *mv_block, distblk min_mcost, int lambda_factor)
{
  jmb_me_req q;
  jmb_me_res r;
  MotionVector *mv = &mv_block->mv[(short)mv_block->list];
  int rc, ri, R, E, mbx, mby, pass;
  if (!shim_on(FAM_ME)) return __real_full_search_motion_estimation(currMB, pred_mv, mv_block, min_mcost, lambda_factor);
  /* weighted-reference ME (UseWeightedReferenceME): JM's own loop runs and takes every distortion from the device
   * through mv_block->computePredFPel = computeSADWP (wrapped below) */
  if (mv_block->apply_weights || currMB->p_Inp->MEErrorMetric[F_PEL] != ERROR_SAD)      /* likewise a full-pel metric other than SAD */
    return __real_full_search_motion_estimation(currMB, pred_mv, mv_block, min_mcost, lambda_factor);
  if (!currMB->p_Inp->rdopt) unsupported("RDOptimization=0 (the (0,0) bias of full_search_motion_estimation)");
  S.calls[1]++;
  R = imin(mv_block->searchRange.max_x, mv_block->searchRange.max_y) >> 2;
  configure(currMB, mv_block, R);
  ri = ref_index(currMB, mv_block);
  mbx = mv_block->pos_x & ~15; mby = mv_block->pos_y & ~15;
  E = imin(8, (91 - (2 * R + 1)) / 2);
  if (E >= 0 && ahead_take(currMB, pred_mv, mv_block, min_mcost, lambda_factor, JMB_SEARCH_FULL, R, mv->mv_x, mv->mv_y, &r))
  { /* searched ahead of this call, with exactly this predictor and centre */
    mv->mv_x = r.imv_x; mv->mv_y = r.imv_y;
    return (distblk)r.icost;
  }
  if (E >= 0)
    for (pass = 0; pass < 2; pass++)
    {
      if (pass || !S.surf[ri].valid || S.surf[ri].pic != S.pic_count || S.surf[ri].mb_x != mbx || S.surf[ri].mb_y != mby)
      { /* (re)build the surfaces around this partition's centre (mv_search.c:931-957 left it in mv) */
        rc = jmb_mb_surfaces(S.ctx, ri, mbx, mby, mv->mv_x, mv->mv_y, R + E);
        if (rc) jmb_die("jmb_mb_surfaces", rc);
        S.surf[ri].valid = 1; S.surf[ri].pic = S.pic_count; S.surf[ri].mb_x = mbx; S.surf[ri].mb_y = mby;
        S.surf_builds++;
      }
      if (min_mcost == DISTBLK_MAX && ahead_run(currMB, pred_mv, mv_block, lambda_factor, JMB_SEARCH_FULL, R, ri, mv->mv_x, mv->mv_y, &r))
      { mv->mv_x = r.imv_x; mv->mv_y = r.imv_y; return (distblk)r.icost; }
      rc = mb_search(currMB, pred_mv, mv_block, min_mcost, lambda_factor, JMB_SEARCH_FULL, ri, mv->mv_x, mv->mv_y, &r);
      if (!rc) { mv->mv_x = r.imv_x; mv->mv_y = r.imv_y; return (distblk)r.icost; }
      if (rc != JMB_ERR_STATE) jmb_die("jmb_mb_search(full)", rc);      /* JMB_ERR_STATE: window not covered -> fresh surfaces */
    }
  /* search ranges the surface kernel does not take (2R+1 > 91 displacements): one launch per request */
  fill_request(&q, mv_block, pred_mv, ri, min_mcost);
  q.center_x = mv->mv_x;           /* the search centre BlockMotionSearch left in mv (mv_search.c:931-957) */
  q.center_y = mv->mv_y;
  q.mode = JMB_SEARCH_FULL;
  q.lambda[0] = lambda_factor;
  S.spec.valid = 0;
  rc = jmb_me_search(S.ctx, &q, 1, &r, JMB_HOST);
  if (rc) jmb_die("jmb_me_search(full)", rc);
  mv->mv_x = r.imv_x;
  mv->mv_y = r.imv_y;
  return (distblk)r.icost;
}

/* setup_fast_full_search (lencod/src/me_fullfast.c:269): the search-centre rule (:305-329) on the host; the BlockSAD surfaces
 * (:492-556) are built on the device and STAY there (jmb_mb_surfaces) -- JM's own BlockSAD arrays are never filled. */
void __wrap_setup_fast_full_search(Macroblock *currMB, MEBlock *mv_block, int list)
{
  VideoParameters *p_Vid = currMB->p_Vid;
  InputParameters *p_Inp = currMB->p_Inp;
  MEFullFast *ff = p_Vid->p_ffast_me;
  Slice *currSlice = currMB->p_Slice;
  short ref = mv_block->ref_idx;
  int search_range = ff->max_search_range[list][ref] << 2, sr = ff->max_search_range[list][ref], ri, rc;
  int list_offset = p_Vid->mb_data[currMB->mbAddrX].list_offset;
  MotionVector pmv, *c = &ff->search_center[list][ref];
  PixelPos block[4];
  if (!shim_on(FAM_ME)) { __real_setup_fast_full_search(currMB, mv_block, list); return; }
  if (!p_Inp->rdopt) unsupported("RDOptimization=0 with fast full search");
  /* setup_fast_full_search builds SQUARED-error surfaces for any full-pel metric but SAD (dist_method, me_fullfast.c:274) */
  if (p_Inp->MEErrorMetric[F_PEL] != ERROR_SAD) unsupported("fast full search with MEDistortionFPel other than SAD (squared-error BlockSAD surfaces)");
  if (mv_block->apply_weights) unsupported("fast full search with UseWeightedReferenceME (its weighted SAD surfaces are inline host code; use SearchMode -1 or 3)");
  if (2 * sr + 1 > 91) unsupported("fast full search with a SearchRange above 45");
  get_neighbors(currMB, block, 0, 0, 16);
  currMB->GetMVPredictor(currMB, block, &pmv, ref, p_Vid->enc_picture->mv_info, list, 0, 0, 16, 16);
#if (JM_INT_DIVIDE)
  c->mv_x = ((pmv.mv_x + 2) >> 2) * 4;
  c->mv_y = ((pmv.mv_y + 2) >> 2) * 4;
#else
  c->mv_x = (pmv.mv_x / 4) * 4;
  c->mv_y = (pmv.mv_y / 4) * 4;
#endif
  c->mv_x = (short)iClip3(p_Vid->MaxHmvR[4] + search_range, p_Vid->MaxHmvR[5] - search_range, c->mv_x);
  c->mv_y = (short)iClip3(p_Vid->MaxVmvR[4] + search_range, p_Vid->MaxVmvR[5] - search_range, c->mv_y);
  ff->search_center_padded[list][ref] = pad_MVs(*c, mv_block);
  if (mv_block->pos_x != currMB->pix_x || mv_block->pos_y != currMB->opix_y)
    unsupported("setup_fast_full_search entered with a block that is not at the macroblock origin");
  configure(currMB, mv_block, sr);
  ri = ref_index_of(p_Vid, currSlice, mv_block, currSlice->listX[list + list_offset][ref]);
  rc = jmb_mb_surfaces(S.ctx, ri, currMB->pix_x, currMB->opix_y, c->mv_x, c->mv_y, sr);
  if (rc) jmb_die("jmb_mb_surfaces", rc);
  S.surf_builds++;
  S.calls[2]++;
  ff->search_setup_done[list][ref] = 1;
}

/* stands behind fast_full_search_motion_estimation (lencod/src/me_fullfast.c:618) = currMB->IntPelME for SearchMode 0:
 * the arg-min of one partition over the resident surfaces, on the device */
distblk __wrap_fast_full_search_motion_estimation(Macroblock *currMB, MotionVector *pred_mv, MEBlock *mv_block, distblk min_mcost, int lambda_factor)
{
  VideoParameters *p_Vid = currMB->p_Vid;
  MEFullFast *ff = p_Vid->p_ffast_me;
  int list = mv_block->list, rc;
  short ref = mv_block->ref_idx;
  jmb_me_res r;
  if (!shim_on(FAM_ME)) return __real_fast_full_search_motion_estimation(currMB, pred_mv, mv_block, min_mcost, lambda_factor);
  if (!currMB->p_Inp->rdopt) unsupported("RDOptimization=0 with fast full search");
  S.calls[2]++;
  if (!ff->search_setup_done[list][ref]) currMB->p_SetupFastFullPelSearch(currMB, mv_block, list);
  configure(currMB, mv_block, imax(mv_block->searchRange.max_x, mv_block->searchRange.max_y) >> 2);
  if (ahead_take(currMB, pred_mv, mv_block, min_mcost, lambda_factor, JMB_SEARCH_FAST_FULL, S.cfg.search_range,
                 ff->search_center[list][ref].mv_x, ff->search_center[list][ref].mv_y, &r) ||
      (min_mcost == DISTBLK_MAX && ahead_run(currMB, pred_mv, mv_block, lambda_factor, JMB_SEARCH_FAST_FULL, S.cfg.search_range, ref_index(currMB, mv_block),
                                             ff->search_center[list][ref].mv_x, ff->search_center[list][ref].mv_y, &r)))
  {
    mv_block->mv[list].mv_x = r.imv_x;
    mv_block->mv[list].mv_y = r.imv_y;
    return (distblk)r.icost;
  }
  rc = mb_search(currMB, pred_mv, mv_block, min_mcost, lambda_factor, JMB_SEARCH_FAST_FULL, ref_index(currMB, mv_block),
                 ff->search_center[list][ref].mv_x, ff->search_center[list][ref].mv_y, &r);
  if (rc == JMB_ERR_STATE)
  { /* the device's reference set changed since the setup (a recycled slot was uploaded again): build the surfaces again */
    MEBlock at_mb = *mv_block;
    at_mb.pos_x = currMB->pix_x; at_mb.pos_y = currMB->opix_y;
    currMB->p_SetupFastFullPelSearch(currMB, &at_mb, list);
    rc = mb_search(currMB, pred_mv, mv_block, min_mcost, lambda_factor, JMB_SEARCH_FAST_FULL, ref_index(currMB, mv_block),
                   ff->search_center[list][ref].mv_x, ff->search_center[list][ref].mv_y, &r);
  }
  if (rc) jmb_die("jmb_mb_search(fast full)", rc);
  mv_block->mv[list].mv_x = r.imv_x;
  mv_block->mv[list].mv_y = r.imv_y;
  return (distblk)r.icost;
}

/* stands behind sub_pel_motion_estimation (lencod/src/me_fullsearch.c:186) = currMB->SubPelME */
distblk __wrap_sub_pel_motion_estimation(Macroblock *currMB, MotionVector *pred, MEBlock *mv_block, distblk min_mcost, int *lambda)
{
  jmb_me_req q;
  jmb_me_res r;
  MotionVector *mv = &mv_block->mv[(int)mv_block->list];
  int rc;
  if (!shim_on(FAM_SUBPEL)) return __real_sub_pel_motion_estimation(currMB, pred, mv_block, min_mcost, lambda);
  if (mv_block->apply_weights) return __real_sub_pel_motion_estimation(currMB, pred, mv_block, min_mcost, lambda);   /* -> compute*WP on the device */
  if (!currMB->p_Inp->rdopt) unsupported("RDOptimization=0 (the (0,0) bias of sub_pel_motion_estimation)");
  S.calls[3]++;
  if (S.spec.valid && S.spec.pos_x == mv_block->pos_x && S.spec.pos_y == mv_block->pos_y && S.spec.blocktype == mv_block->blocktype &&
      S.spec.list == mv_block->list && S.spec.ref_idx == mv_block->ref_idx && S.spec.test8x8 == mv_block->test8x8 &&
      S.spec.lambda_h == lambda[H_PEL] && S.spec.lambda_q == lambda[Q_PEL] && S.spec.pred.mv_x == pred->mv_x && S.spec.pred.mv_y == pred->mv_y &&
      S.spec.imv.mv_x == mv->mv_x && S.spec.imv.mv_y == mv->mv_y && S.spec.min_in == min_mcost)
  { /* exactly the refinement the integer search's device call already did */
    S.spec.valid = 0;
    S.spec_hits++;
    *mv = S.spec.mv;
    return S.spec.cost;
  }
  S.spec.valid = 0;
  configure(currMB, mv_block, 0);
  fill_request(&q, mv_block, pred, ref_index(currMB, mv_block), min_mcost);
  q.center_x = mv->mv_x;
  q.center_y = mv->mv_y;
  q.mode = JMB_SEARCH_FULL;
  q.flags = JMB_REQ_SUBPEL | JMB_REQ_SKIP_INT | (mv_block->test8x8 ? JMB_REQ_TEST8X8 : 0);
  q.lambda[0] = lambda[F_PEL];
  q.lambda[1] = lambda[H_PEL];
  q.lambda[2] = lambda[Q_PEL];
  rc = jmb_mb_search(S.ctx, &q, &r);
  if (rc) jmb_die("jmb_mb_search(sub-pel)", rc);
  mv->mv_x = r.mv_x;
  mv->mv_y = r.mv_y;
  return (distblk)r.cost;
}

/* ---- EPZS (SearchMode 3 with EPZSSubPelGrid: currMB->IntPelME = EPZS_integer_motion_estimation, me_epzs_common.c:155;
 *      currMB->SubPelME = EPZS_sub_pel_motion_estimation with EPZSSubPelME = 1) --------------------------------------------
 * The state machine of a search runs on the device in ONE call (jmb_epzs_search).  What stays here is JM's host state:
 * the predictor generators (JM's own functions, called in JM's order, me_epzs_int.c:152-212) fill p_EPZS->predictor->point;
 * the wrapper only notes where each generator's output starts and which cost gate JM puts in front of it, so that the
 * device can apply the gate once it knows the cost of the start mv.  The two generators that take the start cost as an
 * argument are called with both outcomes (EPZS_temporal_predictors with a cost below and above the stop criterion: the
 * co-located predictor comes first either way, its neighbours only in the second case).
 * JMB_EPZS_CAPTURE=<file> (with JMB_SHIM=passthrough, CPU only): every call is recorded -- request, predictor list, the
 * pictures, and what JM's REAL function returned -- for tests/golden/make_epzs_golden.py, which pins the oracle. */
static FILE *epzs_cap;
static int   epzs_cap_init;
static struct { StorablePicture *pic; int poc; } epzs_cap_ref[64];
static int   epzs_cap_nref, epzs_cap_cur_id = -1, epzs_cap_cur_poc = -0x7fffffff;
static unsigned long epzs_cap_calls;

static int epzs_capture_on(void)
{
  if (!epzs_cap_init)
  {
    const char *f = getenv("JMB_EPZS_CAPTURE");
    epzs_cap_init = 1;
    if (f) { epzs_cap = fopen(f, "wb"); if (!epzs_cap) { snprintf(errortext, ET_SIZE, "cannot open %s", f); fatal(704); } }
  }
  return epzs_cap != NULL;
}

static void epzs_cap_picture(int id, imgpel **img, int w, int h)
{
  int32_t hdr[4] = {1, id, w, h};
  int x, y;
  fwrite(hdr, 4, 4, epzs_cap);
  for (y = 0; y < h; y++) for (x = 0; x < w; x++) fputc((int)img[y][x], epzs_cap);
}

static void epzs_cap_ids(VideoParameters *p_Vid, StorablePicture *ref, int *ref_id, int *cur_id)
{
  int i;
  for (i = 0; i < epzs_cap_nref; i++) if (epzs_cap_ref[i].pic == ref && epzs_cap_ref[i].poc == ref->poc) break;
  if (i == epzs_cap_nref)
  {
    if (epzs_cap_nref == 64) epzs_cap_nref = 0, i = 0;
    epzs_cap_ref[i].pic = ref; epzs_cap_ref[i].poc = ref->poc; epzs_cap_nref++;
    epzs_cap_picture(1000 + ref->poc * 64 + i, ref->imgY, ref->size_x, ref->size_y);
  }
  *ref_id = 1000 + ref->poc * 64 + i;
  if (epzs_cap_cur_poc != p_Vid->enc_picture->poc)
  {
    epzs_cap_cur_poc = p_Vid->enc_picture->poc;
    epzs_cap_cur_id = 500000 + epzs_cap_cur_poc;
    epzs_cap_picture(epzs_cap_cur_id, p_Vid->pCurImg, p_Vid->width, p_Vid->height);
  }
  *cur_id = epzs_cap_cur_id;
}

static int epzs_pattern_id(VideoParameters *p_Vid, EPZSStructure *p)
{
  if (p == p_Vid->sdiamond)  return JMB_EPZS_PAT_SDIAMOND;
  if (p == p_Vid->square)    return JMB_EPZS_PAT_SQUARE;
  if (p == p_Vid->ediamond)  return JMB_EPZS_PAT_EDIAMOND;
  if (p == p_Vid->ldiamond)  return JMB_EPZS_PAT_LDIAMOND;
  if (p == p_Vid->sbdiamond) return JMB_EPZS_PAT_SBDIAMOND;
  if (p == p_Vid->pmvfast)   return JMB_EPZS_PAT_PMVFAST;
  unsupported("an EPZS refinement pattern that is not one of JM's six");
  return 0;
}

#define EPZS_MAX_CANDS 512
/* the request of one EPZS_integer_motion_estimation call; returns the number of candidates written to cands (x,y pairs) */
static int epzs_build_request(Macroblock *currMB, MotionVector *pred_mv, MEBlock *mv_block, int lambda_factor, jmb_epzs_req *q, int16_t *cands)
{
  Slice *currSlice = currMB->p_Slice;
  VideoParameters *p_Vid = currMB->p_Vid;
  InputParameters *p_Inp = currMB->p_Inp;
  EPZSParameters *p_EPZS = currSlice->p_EPZS;
  int blocktype = mv_block->blocktype, list = mv_block->list, cur_list = list + currMB->list_offset, ref = mv_block->ref_idx;
  MotionVector *mv = &mv_block->mv[list];
  StorablePicture *ref_picture = currSlice->listX[cur_list][ref];
  distblk lambda_dist = weighted_cost(lambda_factor, 2);
  distblk *prevSad = &p_EPZS->distortion[cur_list][blocktype - 1][mv_block->pos_x2];
  SPoint *point = p_EPZS->predictor->point;
  int prednum = 5, seg_end[4], s, i, k = 0, n0, frame_gt0 = (ref > 0 && currSlice->structure == FRAME), fixed3_border, static_ok;
  short invalid_refs;

  memset(q, 0, sizeof(*q));
  q->pos_x = mv_block->pos_x;  q->pos_y = mv_block->pos_y;
  q->pred_x = pred_mv->mv_x;   q->pred_y = pred_mv->mv_y;
  q->start_x = mv->mv_x;       q->start_y = mv->mv_y;
  q->blocktype = (uint8_t)blocktype;
  q->jm_ref = (uint8_t)ref;
  q->lambda[0] = lambda_factor;
  q->range_x = (int16_t)mv_block->searchRange.max_x;  q->range_y = (int16_t)mv_block->searchRange.max_y;
  q->pattern = (uint8_t)epzs_pattern_id(p_Vid, p_EPZS->searchPattern);
  q->pattern_dual = (uint8_t)epzs_pattern_id(p_Vid, p_EPZS->searchPatternD);
  q->flags = (uint8_t)((frame_gt0 ? JMB_EPZS_REF_GT0_FRAME : 0) | (p_Inp->EPZSPattern != 0 ? JMB_EPZS_ADAPT_PATTERN : 0) |
                       ((ref > 0 && blocktype != 1) ? JMB_EPZS_SQUARE_HINT : 0) | (p_Inp->EPZSDual > 0 ? JMB_EPZS_DUAL : 0));
  q->medthres = p_EPZS->medthres[blocktype];
  q->subthres = p_EPZS->subthres[blocktype];
  q->prev_sad = *prevSad;
  q->stop = EPZSDetermineStopCriterion(p_EPZS, prevSad, mv_block, lambda_dist);
  q->min_mcost = DISTBLK_MAX;

  /* segment 0: generators JM always runs (me_epzs_int.c:152-170) + the co-located predictor */
  invalid_refs = EPZS_spatial_predictors(p_EPZS, mv_block, list, currMB->list_offset, ref, p_Vid->enc_picture->mv_info);
  if (p_Inp->EPZSSpatialMem) EPZS_spatial_memory_predictors(p_EPZS, mv_block, cur_list, &prednum, ref_picture->size_x >> 2);
  if (p_Inp->HMEEnable == 1 && p_Inp->EPZSUseHMEPredictors == 1) EPZS_hierarchical_predictors(p_EPZS, mv_block, &prednum, ref_picture, currSlice);
  n0 = prednum;
#if (MVC_EXTENSION_ENABLE)
  if (p_Inp->EPZSTemporal[currSlice->view_id])
#else
  if (p_Inp->EPZSTemporal)
#endif
  {
    EPZS_temporal_predictors(currMB, ref_picture, p_EPZS, mv_block, &prednum, q->stop, 0);                 /* start cost <= stop */
    seg_end[0] = prednum;
    prednum = n0;
    EPZS_temporal_predictors(currMB, ref_picture, p_EPZS, mv_block, &prednum, q->stop, DISTBLK_MAX);       /* start cost > stop */
    seg_end[1] = prednum;                                                                                   /* segment 1: gate 1 */
    q->gate[1] = 1;
  }
  else seg_end[0] = seg_end[1] = prednum;
  /* segment 2: window predictors (:193-203) */
  fixed3_border = (p_Inp->EPZSFixed == 3 && (currMB->mb_x == 0 || currMB->mb_y == 0));
  static_ok = ((ref < 2 && blocktype < 4) || (ref < 1 && blocktype == 4) || ((currSlice->structure != FRAME || currMB->list_offset) && ref < 3))
              && (p_Inp->EPZSFixed > 1 || (p_Inp->EPZSFixed && currSlice->slice_type == P_SLICE));
  if (fixed3_border || static_ok)
  {
    EPZSWindowPredictors(mv, p_EPZS->predictor, &prednum,
                         (p_Inp->EPZSAggressiveWindow != 0) || ((invalid_refs > 2) && (ref < 1 + (currSlice->structure != FRAME || currMB->list_offset)))
                         ? p_EPZS->window_predictor_ext : p_EPZS->window_predictor);
    q->gate[2] = fixed3_border ? 0 : 3;
  }
  seg_end[2] = prednum;
  /* segment 3: block-type / reference predictors (:211-214) */
  if (currMB->mbAddrX != 0 && p_Inp->EPZSBlockType)
  {
    EPZSBlockTypePredictorsMB(currSlice, mv_block, point, &prednum);
    q->gate[3] = (ref == 0) ? 0 : 2;
  }
  seg_end[3] = prednum;
  if (prednum > EPZS_MAX_CANDS) unsupported("more than 512 EPZS predictors for one block");
  for (s = 0, i = 0; s < 4; s++)
  {
    int cnt = seg_end[s] - i;
    if (cnt > 255) unsupported("more than 255 EPZS predictors in one generator group");
    q->n_cand[s] = (uint8_t)cnt;
    for (; i < seg_end[s]; i++) { cands[2 * k] = point[i].motion.mv_x; cands[2 * k + 1] = point[i].motion.mv_y; k++; }
  }
  return k;
}

static int epzs_device_ok(Macroblock *currMB, MEBlock *mv_block)
{
  VideoParameters *p_Vid = currMB->p_Vid;
  return !mv_block->apply_weights && !mv_block->ChromaMEEnable && p_Vid->structure == FRAME && !currMB->p_Slice->mb_aff_frame_flag &&
         currMB->p_Inp->MEErrorMetric[F_PEL] == ERROR_SAD && p_Vid->start_me_refinement_qp == 1 && p_Vid->bitdepth_luma == 8;
}

/* stands behind EPZS_integer_motion_estimation (lencod/src/me_epzs_int.c:42) */
distblk __wrap_EPZS_integer_motion_estimation(Macroblock *currMB, MotionVector *pred_mv, MEBlock *mv_block, distblk min_mcost, int lambda_factor)
{
  Slice *currSlice = currMB->p_Slice;
  InputParameters *p_Inp = currMB->p_Inp;
  EPZSParameters *p_EPZS = currSlice->p_EPZS;
  int list = mv_block->list, cur_list = list + currMB->list_offset, ref = mv_block->ref_idx, blocktype = mv_block->blocktype, ncand, rc;
  MotionVector *mv = &mv_block->mv[list];
  distblk *prevSad = &p_EPZS->distortion[cur_list][blocktype - 1][mv_block->pos_x2];
  jmb_epzs_req q;
  jmb_epzs_res r;
  int16_t cands[2 * EPZS_MAX_CANDS];

  if (epzs_capture_on())
  {
    int32_t hdr[8];
    distblk ret;
    ncand = epzs_build_request(currMB, pred_mv, mv_block, lambda_factor, &q, cands);
    epzs_cap_ids(currMB->p_Vid, currSlice->listX[cur_list][ref], &hdr[1], &hdr[2]);
    ret = __real_EPZS_integer_motion_estimation(currMB, pred_mv, mv_block, min_mcost, lambda_factor);
    hdr[0] = 2; hdr[3] = ncand; hdr[4] = mv->mv_x; hdr[5] = mv->mv_y; hdr[6] = 0; hdr[7] = 0;
    fwrite(hdr, 4, 8, epzs_cap); fwrite(&q, sizeof(q), 1, epzs_cap); fwrite(cands, 4, (size_t)ncand, epzs_cap);
    { int64_t o[2] = {(int64_t)ret, (int64_t)*prevSad}; fwrite(o, 8, 2, epzs_cap); }
    epzs_cap_calls++;
    return ret;
  }
  if (!shim_on(FAM_ME) || !epzs_device_ok(currMB, mv_block))      /* JM's own walk; its distortions still come from the device (computeSAD*) */
    return __real_EPZS_integer_motion_estimation(currMB, pred_mv, mv_block, min_mcost, lambda_factor);
  S.calls[1]++;
  configure(currMB, mv_block, 0);
  ncand = epzs_build_request(currMB, pred_mv, mv_block, lambda_factor, &q, cands);
  q.ref = (uint8_t)ref_index(currMB, mv_block);
  ++p_EPZS->BlkCount;      /* JM's visit stamp advances once per search (me_epzs_int.c:77-79); its map itself is not used */
  if (p_EPZS->BlkCount == 0) ++p_EPZS->BlkCount;
  rc = jmb_epzs_search(S.ctx, &q, 1, cands, ncand, &r, JMB_HOST);
  if (rc) jmb_die("jmb_epzs_search", rc);
  *prevSad = (distblk)r.prev_sad;
#if EPZSREF
  if (p_Inp->EPZSSpatialMem)
    p_EPZS->p_motion[cur_list][ref][blocktype - 1][mv_block->block_y][mv_block->pos_x2].mv_x = r.imv_x,
    p_EPZS->p_motion[cur_list][ref][blocktype - 1][mv_block->block_y][mv_block->pos_x2].mv_y = r.imv_y;
#else
  if (p_Inp->EPZSSpatialMem && ref == 0)
    p_EPZS->p_motion[cur_list][blocktype - 1][mv_block->block_y][mv_block->pos_x2].mv_x = r.imv_x,
    p_EPZS->p_motion[cur_list][blocktype - 1][mv_block->block_y][mv_block->pos_x2].mv_y = r.imv_y;
#endif
  mv->mv_x = r.imv_x;
  mv->mv_y = r.imv_y;
  return (distblk)r.icost;
}

/* stands behind EPZS_sub_pel_motion_estimation (lencod/src/me_epzs_sub.c:30) */
distblk __wrap_EPZS_sub_pel_motion_estimation(Macroblock *currMB, MotionVector *pred, MEBlock *mv_block, distblk min_mcost, int *lambda)
{
  Slice *currSlice = currMB->p_Slice;
  EPZSParameters *p_EPZS = currSlice->p_EPZS;
  MotionVector *mv = &mv_block->mv[(int)mv_block->list];
  jmb_epzs_req q;
  jmb_epzs_res r;
  int16_t none[2] = {0, 0};
  int rc, capture = epzs_capture_on();

  if (!capture && (!shim_on(FAM_SUBPEL) || !epzs_device_ok(currMB, mv_block)))
    return __real_EPZS_sub_pel_motion_estimation(currMB, pred, mv_block, min_mcost, lambda);
  memset(&q, 0, sizeof(q));
  q.pos_x = mv_block->pos_x;  q.pos_y = mv_block->pos_y;
  q.pred_x = pred->mv_x;      q.pred_y = pred->mv_y;
  q.start_x = mv->mv_x;       q.start_y = mv->mv_y;
  q.blocktype = (uint8_t)mv_block->blocktype;
  q.flags = (uint8_t)(JMB_EPZS_SUBPEL | JMB_EPZS_SKIP_INT | (mv_block->test8x8 ? JMB_EPZS_TEST8X8 : 0));
  q.lambda[0] = lambda[F_PEL]; q.lambda[1] = lambda[H_PEL]; q.lambda[2] = lambda[Q_PEL];
  q.range_x = q.range_y = 1;
  q.subthres = p_EPZS->subthres[mv_block->blocktype];
  q.prev_sad = DISTBLK_MAX;
  q.min_mcost = min_mcost;
  if (capture)
  {
    int32_t hdr[8];
    distblk ret;
    epzs_cap_ids(currMB->p_Vid, currSlice->listX[mv_block->list + currMB->list_offset][(int)mv_block->ref_idx], &hdr[1], &hdr[2]);
    ret = __real_EPZS_sub_pel_motion_estimation(currMB, pred, mv_block, min_mcost, lambda);
    hdr[0] = 3; hdr[3] = 0; hdr[4] = mv->mv_x; hdr[5] = mv->mv_y;
    hdr[6] = currMB->p_Vid->start_me_refinement_hp | (currMB->p_Vid->start_me_refinement_qp << 1) | (mv_block->search_pos2 << 2);
    hdr[7] = currMB->p_Inp->MEErrorMetric[H_PEL] | (currMB->p_Inp->MEErrorMetric[Q_PEL] << 4);
    fwrite(hdr, 4, 8, epzs_cap); fwrite(&q, sizeof(q), 1, epzs_cap);
    { int64_t o[2] = {(int64_t)ret, 0}; fwrite(o, 8, 2, epzs_cap); }
    return ret;
  }
  S.calls[3]++;
  configure(currMB, mv_block, 0);
  q.ref = (uint8_t)ref_index(currMB, mv_block);
  rc = jmb_epzs_search(S.ctx, &q, 1, none, 0, &r, JMB_HOST);
  if (rc) jmb_die("jmb_epzs_search(sub-pel)", rc);
  mv->mv_x = r.mv_x;
  mv->mv_y = r.mv_y;
  return (distblk)r.cost;
}

/* ---- block distortion (lencod/src/me_distortion.c:349 computeSAD, :1190 computeSSE, :745 computeSATD) --------------
 * The finest-grain hook: p_Vid->computeUniPred[] (mv_search.c:486-500) -> mv_block->computePred{F,H,Q}Pel, i.e. every
 * search engine that is NOT wrapped as a whole (EPZS, UMHex, the bi-predictive searches' single-list parts) becomes a
 * client of the device distortion oracle.  JM's early exit returns dist_scale_f(x) = min_mcost (mv_search.h:19-23) as soon
 * as a running sum exceeds min_mcost >> 5; the sums only grow and the check also follows the last row / sub-block, so the
 * exit happens iff the TOTAL exceeds it -- the wrapper reproduces it from the full distortion. */
static distblk block_dist(int metric, StorablePicture *ref1, MEBlock *mv_block, distblk min_mcost, MotionVector *cand)
{
  VideoParameters *p_Vid = mv_block->p_Vid;
  int16_t xy[2];
  int32_t d = 0;
  int rc, ref = ref_index_of(p_Vid, mv_block->p_Slice, mv_block, ref1);
  S.calls[8]++;
  xy[0] = cand->mv_x;
  xy[1] = cand->mv_y;
  rc = jmb_dist(S.ctx, ref, metric, mv_block->blocktype, mv_block->pos_x, mv_block->pos_y, xy, 1,
                metric == JMB_SATD && mv_block->test8x8, &d, JMB_HOST);
  if (rc) jmb_die("jmb_dist", rc);
  return d > dist_down(min_mcost) ? min_mcost : dist_scale((distblk)d);
}

distblk __wrap_computeSAD(StorablePicture *ref1, MEBlock *mv_block, distblk min_mcost, MotionVector *cand)
{
  if (!shim_on(FAM_DIST)) return __real_computeSAD(ref1, mv_block, min_mcost, cand);
  return block_dist(JMB_SAD, ref1, mv_block, min_mcost, cand);
}
distblk __wrap_computeSSE(StorablePicture *ref1, MEBlock *mv_block, distblk min_mcost, MotionVector *cand)
{
  if (!shim_on(FAM_DIST)) return __real_computeSSE(ref1, mv_block, min_mcost, cand);
  return block_dist(JMB_SSE, ref1, mv_block, min_mcost, cand);
}
distblk __wrap_computeSATD(StorablePicture *ref1, MEBlock *mv_block, distblk min_mcost, MotionVector *cand)
{
  if (!shim_on(FAM_DIST)) return __real_computeSATD(ref1, mv_block, min_mcost, cand);
  return block_dist(JMB_SATD, ref1, mv_block, min_mcost, cand);
}


/* ---- mode-decision distortion back-ends: p_Vid->distortion4x4 / distortion8x8 (lencod/src/me_distortion.c:38-170) ------------
 * select_distortion installs them from inside their own translation unit, so the wrapper lets it run and then points the two
 * function pointers at device-backed versions (jmb_block_distortion).  Callers: GetSkipCostMB, BPredPartitionCost,
 * BIDPartitionCost (mv_search.c:589-675, :1159-1325), the transform-size decision (macroblock.c:1413), intra chroma mode cost. */
static int md_metric;
static distblk md_block(short *diff, int n)
{
  int32_t d = 0;
  int rc;
  if (!shim_on(FAM_DIST))
  { /* bisecting: JM's own arithmetic */
    if (n == 4) return md_metric == ERROR_SAD ? distortion4x4SAD(diff, DISTBLK_MAX) : md_metric == ERROR_SSE ? distortion4x4SSE(diff, DISTBLK_MAX) : distortion4x4SATD(diff, DISTBLK_MAX);
    return md_metric == ERROR_SAD ? distortion8x8SAD(diff, DISTBLK_MAX) : md_metric == ERROR_SSE ? distortion8x8SSE(diff, DISTBLK_MAX) : distortion8x8SATD(diff, DISTBLK_MAX);
  }
  rc = jmb_block_distortion(S.ctx, md_metric == ERROR_SAD ? JMB_SAD : md_metric == ERROR_SSE ? JMB_SSE : JMB_SATD, n, diff, 1, NULL, &d, JMB_HOST);
  if (rc) jmb_die("jmb_block_distortion", rc);
  S.calls[8]++;
  return dist_scale((distblk)d);
}
static distblk md_distortion4x4(short *diff, distblk min_dist) { (void)min_dist; return md_block(diff, 4); }
static distblk md_distortion8x8(short *diff, distblk min_dist) { (void)min_dist; return md_block(diff, 8); }

void __wrap_select_distortion(VideoParameters *p_Vid, InputParameters *p_Inp)
{
  __real_select_distortion(p_Vid, p_Inp);
  md_metric = (p_Inp->ModeDecisionMetric == ERROR_SAD || p_Inp->ModeDecisionMetric == ERROR_SSE) ? p_Inp->ModeDecisionMetric : ERROR_SATD;
  if (getenv("JMB_SHIM") && !strcmp(getenv("JMB_SHIM"), "passthrough")) return;
  p_Vid->distortion4x4 = md_distortion4x4;
  p_Vid->distortion8x8 = md_distortion8x8;
}

/* ---- the other nine members of the distortion table (mv_search.c:486-506): weighted and bi-predictive forms -------
 * computeSADWP/SSEWP/SATDWP (me_distortion.c:434,1261,833), computeBiPredSAD1/SSE1/SATD1 (:525,:1353,:943),
 * computeBiPredSAD2/SSE2/SATD2 (:624,:1438,:1038) -> jmb_dist_ex.  Weights as PrepareMEParams / PrepareBiPredMEParams
 * (mv_search.c:183,206) left them in the MEBlock. */
static void dist_refs(MEBlock *mv_block, StorablePicture *ref1, StorablePicture *ref2, int *i1, int *i2)
{
  VideoParameters *p_Vid = mv_block->p_Vid;
  *i1 = ref_index_of(p_Vid, mv_block->p_Slice, mv_block, ref1);
  *i2 = ref2 ? ref_index_of(p_Vid, mv_block->p_Slice, mv_block, ref2) : *i1;
  if (ref2) *i1 = ref_index_of(p_Vid, mv_block->p_Slice, mv_block, ref1);      /* registering ref2 may have re-ordered the list */
}

static void dist_pred_of(jmb_dist_pred *P, int form, MEBlock *mv_block, int ref2, const MotionVector *cand2)
{
  Slice *currSlice = mv_block->p_Slice;
  memset(P, 0, sizeof(*P));
  P->form = form;
  P->ref2 = ref2;
  if (cand2) { P->cand2_x = cand2->mv_x; P->cand2_y = cand2->mv_y; }
  if (mv_block->p_Vid->max_imgpel_value != 255) unsupported("weighted / bi-predictive distortion at a bit depth other than 8");
  if (form == JMB_PRED_WEIGHTED) { P->weight1 = mv_block->weight_luma; P->offset = mv_block->offset_luma; }
  if (form == JMB_PRED_WEIGHTED_AVERAGE) { P->weight1 = mv_block->weight1; P->weight2 = mv_block->weight2; P->offset = mv_block->offsetBi; }
  P->log_weight_denom = currSlice->luma_log_weight_denom;
  P->wp_round = currSlice->wp_luma_round;
}

static distblk block_dist_ex(int metric, int form, StorablePicture *ref1, StorablePicture *ref2, MEBlock *mv_block, distblk min_mcost,
                             MotionVector *cand1, MotionVector *cand2)
{
  jmb_dist_pred P;
  int16_t xy[2];
  int32_t d = 0;
  int rc, i1, i2;
  S.calls[8]++;
  S.calls[9]++;
  dist_refs(mv_block, ref1, ref2, &i1, &i2);
  dist_pred_of(&P, form, mv_block, i2, cand2);
  xy[0] = cand1->mv_x;
  xy[1] = cand1->mv_y;
  rc = jmb_dist_ex(S.ctx, i1, &P, metric, mv_block->blocktype, mv_block->pos_x, mv_block->pos_y, xy, 1,
                   metric == JMB_SATD && mv_block->test8x8, &d, JMB_HOST);
  if (rc) jmb_die("jmb_dist_ex", rc);
  return d > dist_down(min_mcost) ? min_mcost : dist_scale((distblk)d);
}

#define WRAP_UNI_WP(NAME, METRIC) \
distblk __wrap_##NAME(StorablePicture *ref1, MEBlock *mv_block, distblk min_mcost, MotionVector *cand) \
{ \
  if (!shim_on(FAM_DIST)) return __real_##NAME(ref1, mv_block, min_mcost, cand); \
  return block_dist_ex(METRIC, JMB_PRED_WEIGHTED, ref1, NULL, mv_block, min_mcost, cand, NULL); \
}
#define WRAP_BIPRED(NAME, METRIC, FORM) \
distblk __wrap_##NAME(StorablePicture *ref1, StorablePicture *ref2, MEBlock *mv_block, distblk min_mcost, MotionVector *cand1, MotionVector *cand2) \
{ \
  if (!shim_on(FAM_DIST)) return __real_##NAME(ref1, ref2, mv_block, min_mcost, cand1, cand2); \
  return block_dist_ex(METRIC, FORM, ref1, ref2, mv_block, min_mcost, cand1, cand2); \
}
WRAP_UNI_WP(computeSADWP, JMB_SAD)
WRAP_UNI_WP(computeSSEWP, JMB_SSE)
WRAP_UNI_WP(computeSATDWP, JMB_SATD)
WRAP_BIPRED(computeBiPredSAD1, JMB_SAD, JMB_PRED_AVERAGE)
WRAP_BIPRED(computeBiPredSSE1, JMB_SSE, JMB_PRED_AVERAGE)
WRAP_BIPRED(computeBiPredSATD1, JMB_SATD, JMB_PRED_AVERAGE)
WRAP_BIPRED(computeBiPredSAD2, JMB_SAD, JMB_PRED_WEIGHTED_AVERAGE)
WRAP_BIPRED(computeBiPredSSE2, JMB_SSE, JMB_PRED_WEIGHTED_AVERAGE)
WRAP_BIPRED(computeBiPredSATD2, JMB_SATD, JMB_PRED_WEIGHTED_AVERAGE)

/* ---- bi-predictive searches (currMB->BiPredME / SubPelBiPredME for SearchMode -1 and 0, mv_search.c:162-170) ----------
 * One list moves, the other stands still.  Every candidate's distortion comes from ONE jmb_dist_ex call; the host then
 * replays JM's sequential selection (strict '<', candidates whose mv cost alone reaches min_mcost skipped, the early exit
 * of the distortion = "return the threshold") on the complete distortions. */
static int bipred_metric(MEBlock *mv_block, int level)
{
  int m = mv_block->p_Vid->p_Inp->MEErrorMetric[level];
  if (m != ERROR_SAD && m != ERROR_SSE && m != ERROR_SATD) unsupported("a motion-estimation error metric other than SAD / SSE / SATD");
  return m == ERROR_SAD ? JMB_SAD : m == ERROR_SSE ? JMB_SSE : JMB_SATD;
}

static void bipred_dists(Macroblock *currMB, MEBlock *mv_block, int list, int level, const int16_t *cands, int n,
                         const MotionVector *cand2, int32_t *d)
{
  Slice *currSlice = currMB->p_Slice;
  int list_offset = currMB->p_Vid->mb_data[currMB->mbAddrX].list_offset;
  StorablePicture *ref1 = currSlice->listX[list + list_offset][(int)mv_block->ref_idx];
  StorablePicture *ref2 = currSlice->listX[(list ^ 1) + list_offset][0];
  int metric = bipred_metric(mv_block, level), i1, i2, rc;
  jmb_dist_pred P;
  dist_refs(mv_block, ref1, ref2, &i1, &i2);
  dist_pred_of(&P, mv_block->apply_weights ? JMB_PRED_WEIGHTED_AVERAGE : JMB_PRED_AVERAGE, mv_block, i2, cand2);
  rc = jmb_dist_ex(S.ctx, i1, &P, metric, mv_block->blocktype, mv_block->pos_x, mv_block->pos_y, cands, n,
                   metric == JMB_SATD && mv_block->test8x8, d, JMB_HOST);
  if (rc) jmb_die("jmb_dist_ex(bipred)", rc);
  S.calls[8] += (unsigned long)n;
  S.calls[9] += (unsigned long)n;
}

distblk __wrap_full_search_bipred_motion_estimation(Macroblock *currMB, int list, MotionVector *pred_mv1, MotionVector *pred_mv2,
                                                    MotionVector *mv1, MotionVector *mv2, MEBlock *mv_block, int search_range,
                                                    distblk min_mcost, int lambda_factor)
{
  VideoParameters *p_Vid = currMB->p_Vid;
  int pos, best_pos = 0, max_pos = (2 * (search_range >> 2) + 1) * (2 * (search_range >> 2) + 1);
  MotionVector center1, center2, cand, pred1, pred2;
  int16_t *xy;
  int32_t *d;
  distblk mcost, mc2;
  if (!shim_on(FAM_ME)) return __real_full_search_bipred_motion_estimation(currMB, list, pred_mv1, pred_mv2, mv1, mv2, mv_block, search_range, min_mcost, lambda_factor);
  pred1.mv_x = mv_block->pos_x_padded + pred_mv1->mv_x;  pred1.mv_y = mv_block->pos_y_padded + pred_mv1->mv_y;
  pred2.mv_x = mv_block->pos_x_padded + pred_mv2->mv_x;  pred2.mv_y = mv_block->pos_y_padded + pred_mv2->mv_y;
  center1.mv_x = mv_block->pos_x_padded + mv1->mv_x;     center1.mv_y = mv_block->pos_y_padded + mv1->mv_y;
  center2.mv_x = mv_block->pos_x_padded + mv2->mv_x;     center2.mv_y = mv_block->pos_y_padded + mv2->mv_y;
  xy = (int16_t *)malloc((size_t)max_pos * (2 * sizeof(int16_t) + sizeof(int32_t)));
  if (!xy) no_mem_exit("jmb shim: bipred candidates");
  d = (int32_t *)(xy + 2 * (size_t)max_pos);
  for (pos = 0; pos < max_pos; pos++)
  {
    xy[2 * pos]     = (int16_t)(center1.mv_x + p_Vid->spiral_qpel_search[pos].mv_x);
    xy[2 * pos + 1] = (int16_t)(center1.mv_y + p_Vid->spiral_qpel_search[pos].mv_y);
  }
  bipred_dists(currMB, mv_block, list, F_PEL, xy, max_pos, &center2, d);
  mc2 = mv_cost(p_Vid, lambda_factor, &center2, &pred2);
  for (pos = 0; pos < max_pos; pos++)
  {
    distblk thr;
    cand.mv_x = xy[2 * pos];  cand.mv_y = xy[2 * pos + 1];
    mcost = mv_cost(p_Vid, lambda_factor, &cand, &pred1) + mc2;
    if (mcost >= min_mcost) continue;
    thr = min_mcost - mcost;
    mcost += d[pos] > dist_down(thr) ? thr : dist_scale((distblk)d[pos]);
    if (mcost < min_mcost) { best_pos = pos; min_mcost = mcost; }
  }
  free(xy);
  if (best_pos) add_mvs(mv1, &p_Vid->spiral_qpel_search[best_pos]);
  S.calls[1]++;
  return min_mcost;
}

distblk __wrap_sub_pel_bipred_motion_estimation(Macroblock *currMB, MEBlock *mv_block, int list, MotionVector *pred_mv1, MotionVector *pred_mv2,
                                                MotionVector *mv1, MotionVector *mv2, distblk min_mcost, int *lambda)
{
  VideoParameters *p_Vid = currMB->p_Vid;
  int stage;
  MotionVector smv;
  if (!shim_on(FAM_SUBPEL)) return __real_sub_pel_bipred_motion_estimation(currMB, mv_block, list, pred_mv1, pred_mv2, mv1, mv2, min_mcost, lambda);
  smv = pad_MVs(*mv2, mv_block);
  for (stage = 0; stage < 2; stage++)      /* half-pel (me_fullsearch.c:330-358), then quarter-pel (:365-392) */
  {
    const MotionVector *spiral = stage ? p_Vid->spiral_search : p_Vid->spiral_hpel_search;
    int lambda_factor = lambda[stage ? Q_PEL : H_PEL];
    int pos0, pos1, pos, best_pos = 0, n;
    int16_t xy[2 * 9];
    int32_t d[9];
    distblk mc2, mcost;
    MotionVector cand;
    if (!stage)
    {
      pos0 = (min_mcost == DISTBLK_MAX) ? 0 : p_Vid->start_me_refinement_hp;
      pos1 = !p_Vid->start_me_refinement_hp ? imax(1, mv_block->search_pos2) : mv_block->search_pos2;
    }
    else
    {
      if (!p_Vid->start_me_refinement_qp) min_mcost = DISTBLK_MAX;
      pos0 = p_Vid->start_me_refinement_qp;
      pos1 = mv_block->search_pos4;
    }
    if (pos1 > 9) unsupported("more than nine sub-pel refinement positions");
    n = pos1 - pos0;
    if (n <= 0) continue;
    for (pos = pos0; pos < pos1; pos++)
    {
      xy[2 * (pos - pos0)]     = (int16_t)(mv1->mv_x + spiral[pos].mv_x + mv_block->pos_x_padded);      /* pad_MVs */
      xy[2 * (pos - pos0) + 1] = (int16_t)(mv1->mv_y + spiral[pos].mv_y + mv_block->pos_y_padded);
    }
    bipred_dists(currMB, mv_block, list, stage ? Q_PEL : H_PEL, xy, n, &smv, d);
    mc2 = mv_cost(p_Vid, lambda_factor, mv2, pred_mv2);
    for (pos = pos0; pos < pos1; pos++)
    {
      distblk thr;
      cand.mv_x = mv1->mv_x + spiral[pos].mv_x;  cand.mv_y = mv1->mv_y + spiral[pos].mv_y;
      mcost = mv_cost(p_Vid, lambda_factor, &cand, pred_mv1) + mc2;
      if (mcost >= min_mcost) continue;
      thr = min_mcost - mcost;
      mcost += d[pos - pos0] > dist_down(thr) ? thr : dist_scale((distblk)d[pos - pos0]);
      if (mcost < min_mcost) { min_mcost = mcost; best_pos = pos; }
    }
    if (best_pos) add_mvs(mv1, &spiral[best_pos]);
  }
  S.calls[3]++;
  return min_mcost;
}

/* ---- the DC / AC members of the quantiser family and the Hadamard transforms (Intra16x16 and chroma DC paths) ----
 * quant_ac4x4_normal/_around (quant4x4_normal.c:117, quant4x4_around.c:132), quant_dc4x4_normal (:200), quant_dc2x2_* and
 * quant_dc4x2_* (quantChroma_normal.c, quantChroma_around.c) all map onto jmb_quant_list: the wrapper gathers the
 * coefficients in the function's scan order with the per-position parameters and scatters the results back. */
static int quant_list(Macroblock *currMB, int m, int **cptr /* m pointers into tblock */, int32_t params[][3], int q_bits, int qp_per,
                      int dequant, int use_cost, int around, const byte *c_cost, int *coeff_cost, int *L, int *R, int **fptr)
{
  jmb_qlist_desc d;
  int32_t coef[16], levels[17], runs[17], fadj[16], cost = coeff_cost ? *coeff_cost : 0, nz = 0;
  int k, rc;
  memset(&d, 0, sizeof(d));
  d.m = m; d.q_bits = q_bits; d.qp_per = qp_per; d.dequant = dequant; d.use_cost = use_cost; d.around = around;
  d.clip = (currMB->p_Slice->symbol_mode == CAVLC);
  d.adapt_rnd_weight = currMB->p_Vid->AdaptRndWeight;
  for (k = 0; k < m; k++) { d.params[k][0] = params[k][0]; d.params[k][1] = params[k][1]; d.params[k][2] = params[k][2]; coef[k] = *cptr[k]; }
  if (c_cost) for (k = 0; k < 16; k++) d.c_cost[k] = c_cost[k];
  rc = jmb_quant_list(S.ctx, &d, coef, 1, levels, runs, around ? fadj : NULL, &cost, &nz, JMB_HOST);
  if (rc) jmb_die("jmb_quant_list", rc);
  for (k = 0; k < m; k++) { *cptr[k] = coef[k]; if (around && fptr) *fptr[k] = fadj[k]; }
  for (k = 0; levels[k] != 0; k++) { L[k] = levels[k]; R[k] = runs[k]; }
  L[k] = 0;
  if (coeff_cost) *coeff_cost = cost;
  S.calls[6]++;
  return nz;
}

static int quant_ac4x4(Macroblock *currMB, int **tblock, struct quant_methods *qm, int around)
{
  int *cptr[15], *fptr[15];
  int32_t params[15][3];
  int k, qp_per = currMB->p_Vid->p_Quant->qp_per_matrix[qm->qp];
  for (k = 1; k < 16; k++)
  {
    int i = qm->pos_scan[k][0], j = qm->pos_scan[k][1];
    cptr[k - 1] = &tblock[j][qm->block_x + i];
    fptr[k - 1] = (around && qm->fadjust) ? &qm->fadjust[j][qm->block_x + i] : NULL;
    params[k - 1][0] = qm->q_params[j][i].OffsetComp; params[k - 1][1] = qm->q_params[j][i].ScaleComp; params[k - 1][2] = qm->q_params[j][i].InvScaleComp;
  }
  return quant_list(currMB, 15, cptr, params, Q_BITS + qp_per, qp_per, JMB_DQ_SHIFT_RND4, 1, around, qm->c_cost, qm->coeff_cost,
                    qm->ACLevel, qm->ACRun, (around && qm->fadjust) ? fptr : NULL);
}
int __wrap_quant_ac4x4_normal(Macroblock *currMB, int **tblock, struct quant_methods *q_method)
{
  if (!shim_on(FAM_TQ)) return __real_quant_ac4x4_normal(currMB, tblock, q_method);
  return quant_ac4x4(currMB, tblock, q_method, 0);
}
int __wrap_quant_ac4x4_around(Macroblock *currMB, int **tblock, struct quant_methods *q_method)
{
  if (!shim_on(FAM_TQ)) return __real_quant_ac4x4_around(currMB, tblock, q_method);
  return quant_ac4x4(currMB, tblock, q_method, 1);
}

/* the DC forms use ONE parameter triple, a doubled offset and one more bit of shift */
static int quant_dc(Macroblock *currMB, int m, int **cptr, int qp, LevelQuantParams *qp1, int dequant, int *L, int *R)
{
  int32_t params[16][3];
  int k, qp_per = currMB->p_Vid->p_Quant->qp_per_matrix[qp];
  for (k = 0; k < m; k++) { params[k][0] = qp1->OffsetComp << 1; params[k][1] = qp1->ScaleComp; params[k][2] = qp1->InvScaleComp; }
  return quant_list(currMB, m, cptr, params, Q_BITS + qp_per + 1, qp_per, dequant, 0, 0, NULL, NULL, L, R, NULL);
}
int __wrap_quant_dc4x4_normal(Macroblock *currMB, int **tblock, int qp, int *DCLevel, int *DCRun, LevelQuantParams *q_params_4x4, const byte (*pos_scan)[2])
{
  int *cptr[16], k;
  if (!shim_on(FAM_TQ)) return __real_quant_dc4x4_normal(currMB, tblock, qp, DCLevel, DCRun, q_params_4x4, pos_scan);
  for (k = 0; k < 16; k++) cptr[k] = &tblock[pos_scan[k][1]][pos_scan[k][0]];
  return quant_dc(currMB, 16, cptr, qp, q_params_4x4, JMB_DQ_LEVEL, DCLevel, DCRun);
}
static int quant_dc2x2(Macroblock *currMB, int **tblock, int qp, int *DCLevel, int *DCRun, LevelQuantParams *qp1)
{
  int *cptr[4], k;
  for (k = 0; k < 4; k++) cptr[k] = &(*tblock)[k];
  return quant_dc(currMB, 4, cptr, qp, qp1, JMB_DQ_SHIFT, DCLevel, DCRun);
}
int __wrap_quant_dc2x2_normal(Macroblock *currMB, int **tblock, int qp, int *DCLevel, int *DCRun, LevelQuantParams *q, int **fadjust, const byte (*pos_scan)[2])
{
  if (!shim_on(FAM_TQ)) return __real_quant_dc2x2_normal(currMB, tblock, qp, DCLevel, DCRun, q, fadjust, pos_scan);
  return quant_dc2x2(currMB, tblock, qp, DCLevel, DCRun, q);
}
int __wrap_quant_dc2x2_around(Macroblock *currMB, int **tblock, int qp, int *DCLevel, int *DCRun, LevelQuantParams *q, int **fadjust, const byte (*pos_scan)[2])
{
  if (!shim_on(FAM_TQ)) return __real_quant_dc2x2_around(currMB, tblock, qp, DCLevel, DCRun, q, fadjust, pos_scan);
  return quant_dc2x2(currMB, tblock, qp, DCLevel, DCRun, q);
}
static int quant_dc4x2(Macroblock *currMB, int **tblock, int qp, int *DCLevel, int *DCRun, LevelQuantParams *qp1, const byte (*pos_scan)[2])
{
  int *cptr[8], k;
  for (k = 0; k < 8; k++) cptr[k] = &tblock[pos_scan[k][0]][pos_scan[k][1]];     /* j first: quantChroma_normal.c:128 */
  return quant_dc(currMB, 8, cptr, qp, qp1, JMB_DQ_SHIFT, DCLevel, DCRun);
}
int __wrap_quant_dc4x2_normal(Macroblock *currMB, int **tblock, int qp, int *DCLevel, int *DCRun, LevelQuantParams *q, int **fadjust, const byte (*pos_scan)[2])
{
  if (!shim_on(FAM_TQ)) return __real_quant_dc4x2_normal(currMB, tblock, qp, DCLevel, DCRun, q, fadjust, pos_scan);
  return quant_dc4x2(currMB, tblock, qp, DCLevel, DCRun, q, pos_scan);
}
int __wrap_quant_dc4x2_around(Macroblock *currMB, int **tblock, int qp, int *DCLevel, int *DCRun, LevelQuantParams *q, int **fadjust, const byte (*pos_scan)[2])
{
  if (!shim_on(FAM_TQ)) return __real_quant_dc4x2_around(currMB, tblock, qp, DCLevel, DCRun, q, fadjust, pos_scan);
  return quant_dc4x2(currMB, tblock, qp, DCLevel, DCRun, q, pos_scan);
}

/* hadamard4x4 / ihadamard4x4 / hadamard4x2 / ihadamard4x2 / hadamard2x2 / ihadamard2x2 (lcommon/src/transform.c:121-330) */
static void hadamard_call(int kind, int32_t *v)
{
  int rc = jmb_hadamard(S.ctx, kind, v, 1, JMB_HOST);
  if (rc) jmb_die("jmb_hadamard", rc);
  S.calls[4]++;
}
void __wrap_hadamard4x4(int **block, int **tblock)
{
  int32_t v[16]; int i, j;
  if (!shim_on(FAM_TQ)) { __real_hadamard4x4(block, tblock); return; }
  for (j = 0; j < 4; j++) for (i = 0; i < 4; i++) v[4 * j + i] = block[j][i];
  hadamard_call(JMB_HAD_4X4, v);
  for (j = 0; j < 4; j++) for (i = 0; i < 4; i++) tblock[j][i] = v[4 * j + i];
}
void __wrap_ihadamard4x4(int **tblock, int **block)
{
  int32_t v[16]; int i, j;
  if (!shim_on(FAM_TQ)) { __real_ihadamard4x4(tblock, block); return; }
  for (j = 0; j < 4; j++) for (i = 0; i < 4; i++) v[4 * j + i] = tblock[j][i];
  hadamard_call(JMB_IHAD_4X4, v);
  for (j = 0; j < 4; j++) for (i = 0; i < 4; i++) block[j][i] = v[4 * j + i];
}
void __wrap_hadamard4x2(int **block, int **tblock)
{
  int32_t v[8]; int i, j;
  if (!shim_on(FAM_TQ)) { __real_hadamard4x2(block, tblock); return; }
  for (j = 0; j < 2; j++) for (i = 0; i < 4; i++) v[4 * j + i] = block[j][i];
  hadamard_call(JMB_HAD_4X2, v);
  for (j = 0; j < 2; j++) for (i = 0; i < 4; i++) tblock[j][i] = v[4 * j + i];
}
void __wrap_ihadamard4x2(int **tblock, int **block)
{
  int32_t v[8]; int i, j;
  if (!shim_on(FAM_TQ)) { __real_ihadamard4x2(tblock, block); return; }
  for (j = 0; j < 2; j++) for (i = 0; i < 4; i++) v[4 * j + i] = tblock[j][i];
  hadamard_call(JMB_IHAD_4X2, v);
  for (j = 0; j < 4; j++) for (i = 0; i < 2; i++) block[j][i] = v[2 * j + i];
}
void __wrap_hadamard2x2(int **block, int tblock[4])
{
  int32_t v[4]; int k;
  if (!shim_on(FAM_TQ)) { __real_hadamard2x2(block, tblock); return; }
  v[0] = block[0][0]; v[1] = block[0][4]; v[2] = block[4][0]; v[3] = block[4][4];
  hadamard_call(JMB_HAD_2X2, v);
  for (k = 0; k < 4; k++) tblock[k] = v[k];
}
void __wrap_ihadamard2x2(int tblock[4], int block[4])
{
  int32_t v[4]; int k;
  if (!shim_on(FAM_TQ)) { __real_ihadamard2x2(tblock, block); return; }
  for (k = 0; k < 4; k++) v[k] = tblock[k];
  hadamard_call(JMB_IHAD_2X2, v);
  for (k = 0; k < 4; k++) block[k] = v[k];
}

/* ---- differential check of the whole-macroblock device path against JM itself (JMB_SHIM_VERIFY=1) -----------------
 * luma_residual_coding (lencod/src/macroblock.c:1182) runs as usual (its leaves are already on the GPU); afterwards the
 * same macroblock is coded once more by jmb_luma_residual_coding from JM's own state (partition modes in currMB->b8x8,
 * mvs in currSlice->all_mv, references in enc_picture->mv_info) and the device's levels, cbp, cbp_blk and reconstruction
 * must equal what JM left in cofAC / currMB / enc_picture.  Any difference stops the encoder.  H.264's zig-zag scans and
 * JM's coefficient-cost rows are file-static in JM (block.c:72-77,170-176, transform8x8.c:44-90), so the wrapper carries them. */
static const uint8_t V_SCAN4[16][2] = {{0,0},{1,0},{0,1},{0,2},{1,1},{2,0},{3,0},{2,1},{1,2},{0,3},{1,3},{2,2},{3,1},{3,2},{2,3},{3,3}};
static const uint8_t V_COST4[3][16] = {{3,2,2,1,1,1,0,0,0,0,0,0,0,0,0,0},{9,9,9,9,9,9,9,9,9,9,9,9,9,9,9,9},{3,2,2,1,1,1,0,0,0,0,0,0,0,0,0,0}};

static void zigzag8(uint8_t scan[64][2])
{
  int i = 0, j = 0, up = 1, k;
  for (k = 0; k < 64; k++)
  {
    scan[k][0] = (uint8_t)i; scan[k][1] = (uint8_t)j;
    if (up) { if (i == 7) { j++; up = 0; } else if (j == 0) { i++; up = 0; } else { i++; j--; } }
    else    { if (j == 7) { i++; up = 1; } else if (i == 0) { j++; up = 1; } else { i--; j++; } }
  }
}

static void verify_mismatch(Macroblock *currMB, const char *what, int a, int b)
{
  snprintf(errortext, ET_SIZE, "libjmb200 shim: VERIFY luma_residual_coding: macroblock %d (type %d): %s differs (JM %d, device %d)",
           currMB->mbAddrX, currMB->mb_type, what, a, b);
  fatal(703);
}

void __wrap_luma_residual_coding(Macroblock *currMB)
{
  Slice *currSlice = currMB->p_Slice;
  VideoParameters *p_Vid = currMB->p_Vid;
  jmb_mb_pred pred;
  jmb_quant_desc d;
  static int16_t levels[256];
  static uint8_t recon[256];
  int32_t cost8[4], sse;
  uint32_t cbp_blk, cbp;
  int k, bx, by, i, j, rc, n, cavlc8, qp;
  uint8_t scan8[64][2];

  /* JMB_SHIM_RC=device: for the macroblocks the device path covers (inter macroblocks of P slices, list-0 prediction, no
   * adaptive rounding / weighted prediction / error-robust RDO) the device call REPLACES JM's function: its levels, cbp, cbp_blk and
   * reconstruction are written into JM's state (cofAC, currMB, enc_picture) and JM's own code is not run.  Otherwise JM's function
   * runs, and with JMB_SHIM_VERIFY the device result is compared with what it left. */
  int write_back = 0, eligible = shim_on(FAM_TQ) || (S.init == 1 && S.rc_device);
  if (eligible && (currSlice->slice_type != P_SLICE || p_Vid->AdaptiveRounding || p_Vid->structure != FRAME || currSlice->mb_aff_frame_flag)) eligible = 0;
  if (eligible && !(currMB->mb_type == P16x16 || currMB->mb_type == P16x8 || currMB->mb_type == P8x16 || currMB->mb_type == P8x8)) eligible = 0;
  if (eligible && (currSlice->weighted_prediction || currSlice->NoResidueDirect == 1 || p_Vid->bitdepth_luma != 8 || currMB->p_Inp->rdopt == 3)) eligible = 0;
  if (eligible && currMB->p_Inp->UseRDOQuant) eligible = 0;
  for (k = 0; eligible && k < 4; k++)
  {
    int mode = currMB->b8x8[k].mode;
    if (currMB->b8x8[k].pdir != 0 || mode < 1 || mode > 7 || (currMB->luma_transform_size_8x8_flag && mode > 4) ||
        p_Vid->enc_picture->mv_info[currMB->block_y + 2 * (k >> 1)][currMB->block_x + 2 * (k & 1)].ref_idx[LIST_0] < 0) eligible = 0;
  }
  write_back = eligible && S.rc_device;
  if (!write_back) __real_luma_residual_coding(currMB);
  if (S.init != 1 || !eligible || (!S.verify && !write_back)) return;
  n = currMB->luma_transform_size_8x8_flag ? 8 : 4;
  cavlc8 = (n == 8 && currSlice->symbol_mode == CAVLC);
  memset(&pred, 0, sizeof(pred));
  for (k = 0; k < 4; k++)
  {
    PicMotionParams *m = &p_Vid->enc_picture->mv_info[currMB->block_y + 2 * (k >> 1)][currMB->block_x + 2 * (k & 1)];
    int mode = currMB->b8x8[k].mode, ref = m->ref_idx[LIST_0], slot = -1;
    MEBlock probe;
    if (currMB->b8x8[k].pdir != 0 || mode < 1 || mode > 7 || ref < 0 || (n == 8 && mode > 4)) return;
    memset(&probe, 0, sizeof(probe));
    slot = ref_index_of(p_Vid, currSlice, &probe, currSlice->listX[LIST_0 + currMB->list_offset][ref]);
    pred.b8mode[k] = (uint8_t)mode;
    pred.ref[k] = (uint8_t)slot;
    for (by = 2 * (k >> 1); by < 2 * (k >> 1) + 2; by++)
      for (bx = 2 * (k & 1); bx < 2 * (k & 1) + 2; bx++)
      {
        MotionVector *mv = &currSlice->all_mv[LIST_0][ref][mode][by][bx];
        pred.mv[by * 4 + bx][0] = mv->mv_x;
        pred.mv[by * 4 + bx][1] = mv->mv_y;
      }
  }
  qp = currMB->qp_scaled[0];
  memset(&d, 0, sizeof(d));
  d.n = n; d.qp = qp; d.is_cavlc = (n == 4) ? (currSlice->symbol_mode == CAVLC) : cavlc8;
  {
    LevelQuantParams **qpar = (n == 4) ? p_Vid->p_Quant->q_params_4x4[0][0][qp] : p_Vid->p_Quant->q_params_8x8[0][0][qp];
    for (j = 0; j < n; j++)
      for (i = 0; i < n; i++)
      {
        d.qparams[j * n + i][0] = qpar[j][i].OffsetComp;
        d.qparams[j * n + i][1] = qpar[j][i].ScaleComp;
        d.qparams[j * n + i][2] = qpar[j][i].InvScaleComp;
      }
  }
  if (n == 4)
  {
    memcpy(d.scan, V_SCAN4, sizeof(V_SCAN4));
    memcpy(d.c_cost, V_COST4[currSlice->disthres], 16);
  }
  else
  {
    zigzag8(scan8);
    for (k = 0; k < 64; k++)       /* CAVLC: four interleaved lists, list s takes zig-zag entries 4m+s (transform8x8.c:55) */
    {
      int src = cavlc8 ? 4 * (k & 15) + (k >> 4) : k;
      d.scan[k][0] = scan8[src][0]; d.scan[k][1] = scan8[src][1];
    }
    for (k = 0; k < 64; k++) d.c_cost[k] = (currSlice->disthres == 1) ? 9 : (k < 4 ? 3 : k < 12 ? 2 : k < 24 ? 1 : 0);
  }
  rc = jmb_luma_residual_coding(S.ctx, &pred, currMB->mbAddrX, 1, &d, levels, cost8, &cbp_blk, &cbp, recon, &sse, JMB_HOST);
  if (rc) jmb_die("jmb_luma_residual_coding", rc);

  if (write_back)
  { /* what luma_residual_coding leaves behind (macroblock.c:1182-1257, block.c:661-725): cbp / cbp_blk, the reconstruction, and per
       transform block the (level, run) list in cofAC -- zero lists where thresholding cleared a quadrant */
    currMB->cbp = (int)cbp;
    currMB->cbp_blk = (int64)cbp_blk;
    currSlice->cmp_cbp[1] = currSlice->cmp_cbp[2] = 0;
    currSlice->cur_cbp_blk[1] = currSlice->cur_cbp_blk[2] = 0;
    for (j = 0; j < 16; j++)
      for (i = 0; i < 16; i++) p_Vid->enc_picture->imgY[currMB->pix_y + j][currMB->pix_x + i] = recon[j * 16 + i];
    for (k = 0; k < 4; k++)
    {
      int lists = (n == 4 || cavlc8) ? 4 : 1, len = (n == 4 || cavlc8) ? 16 : 64, s, e;
      if (cost8[k] == 0)
      { /* the quadrant fell to reset_block (cost <= _LUMA_COEFF_COST_): JM clears its whole cofAC slab (macroblock.c:818) */
        memset(currSlice->cofAC[k][0][0], 0, 4 * 2 * 65 * sizeof(int));
        continue;
      }
      /* otherwise the lists stand as the quantiser wrote them -- also when the macroblock-level threshold then drops the cbp
         (macroblock.c:1248-1255 leaves cofAC alone); like JM's quantisers, nothing is written behind the terminating level */
      for (s = 0; s < lists; s++)
      {
        int *ACL = currSlice->cofAC[k][s][0], *ACR = currSlice->cofAC[k][s][1], run = 0, cnt = 0;
        const int16_t *lv = (n == 4) ? &levels[((2 * (k >> 1) + (s >> 1)) * 4 + 2 * (k & 1) + (s & 1)) * 16] : &levels[k * 64 + s * 16];
        for (e = 0; e < len; e++)
        {
          if (lv[e]) { ACL[cnt] = lv[e]; ACR[cnt] = run; cnt++; run = 0; }
          else run++;
        }
        ACL[cnt] = 0;
      }
    }
    S.rc_written++;
    return;
  }
  if ((currMB->cbp & 15) != (int)cbp) verify_mismatch(currMB, "cbp", currMB->cbp & 15, (int)cbp);
  if ((int)(currMB->cbp_blk & 0xffff) != (int)cbp_blk) verify_mismatch(currMB, "cbp_blk", (int)(currMB->cbp_blk & 0xffff), (int)cbp_blk);
  for (j = 0; j < 16; j++)
    for (i = 0; i < 16; i++)
      if (p_Vid->enc_picture->imgY[currMB->pix_y + j][currMB->pix_x + i] != recon[j * 16 + i])
        verify_mismatch(currMB, "reconstruction", p_Vid->enc_picture->imgY[currMB->pix_y + j][currMB->pix_x + i], recon[j * 16 + i]);
  /* levels of every block JM says is coded: expand JM's (level, run) lists to scan order */
  for (k = 0; k < 4; k++)
  {
    if (!(cbp & (1u << k))) continue;
    if (n == 4)
    {
      int b4;
      for (b4 = 0; b4 < 4; b4++)
      {
        int blk = (2 * (k >> 1) + (b4 >> 1)) * 4 + 2 * (k & 1) + (b4 & 1), pos = 0, e;
        int *ACL = currSlice->cofAC[k][b4][0], *ACR = currSlice->cofAC[k][b4][1];
        int16_t want[16];
        memset(want, 0, sizeof(want));
        for (e = 0; e < 16 && ACL[e] != 0; e++) { pos += ACR[e]; want[pos++] = (int16_t)ACL[e]; }
        for (e = 0; e < 16; e++)
          if (want[e] != levels[blk * 16 + e]) verify_mismatch(currMB, "a 4x4 level", want[e], levels[blk * 16 + e]);
      }
    }
    else
    {
      int s, e;
      for (s = 0; s < (cavlc8 ? 4 : 1); s++)
      {
        int *ACL = cavlc8 ? currSlice->cofAC[k][s][0] : currSlice->cofAC[k][0][0], *ACR = cavlc8 ? currSlice->cofAC[k][s][1] : currSlice->cofAC[k][0][1];
        int16_t want[64];
        int pos = 0, len = cavlc8 ? 16 : 64;
        memset(want, 0, sizeof(want));
        for (e = 0; e < len && ACL[e] != 0; e++) { pos += ACR[e]; want[pos++] = (int16_t)ACL[e]; }
        for (e = 0; e < len; e++)
          if (want[e] != levels[k * 64 + s * 16 + e]) verify_mismatch(currMB, "an 8x8 level", want[e], levels[k * 64 + s * 16 + e]);
      }
    }
  }
  S.verified++;
}


/* ---- differential check of the device chroma path against JM itself (JMB_SHIM_VERIFY=1) ------------------------------------
 * chroma_residual_coding (lencod/src/macroblock.c:1439) runs as usual; afterwards the same inter macroblock's two chroma
 * components are predicted and residual-coded by jmb_chroma_residual_coding from JM's own state and the DC / AC levels, the
 * chroma bits of cbp_blk, the chroma cbp and the reconstruction must equal what JM left.  Any difference stops the encoder. */
static void chroma_mismatch(Macroblock *currMB, const char *what, int uv, int a, int b)
{
  snprintf(errortext, ET_SIZE, "libjmb200 shim: VERIFY chroma_residual_coding: macroblock %d (type %d) component %d: %s differs (JM %d, device %d)",
           currMB->mbAddrX, currMB->mb_type, uv, what, a, b);
  fatal(705);
}

void __wrap_chroma_residual_coding(Macroblock *currMB)
{
  Slice *currSlice = currMB->p_Slice;
  VideoParameters *p_Vid = currMB->p_Vid;
  QuantParameters *p_Quant = p_Vid->p_Quant;
  jmb_mb_pred pred;
  jmb_chroma_desc d;
  static int16_t dc[2][8], ac[2][8][15];
  static uint8_t recon[2][16][8];
  uint32_t cb[2], cc[2];
  int k, bx, by, uv, i, j, rc, yuv = p_Vid->yuv_format, nb, hmb, cbp_before = currMB->cbp;

  __real_chroma_residual_coding(currMB);
  if (S.init != 1 || !S.verify || (S.off & FAM_TQ)) return;
  if (currSlice->slice_type != P_SLICE || p_Vid->AdaptiveRounding || p_Vid->structure != FRAME || currSlice->mb_aff_frame_flag) return;
  if (!(currMB->mb_type == P16x16 || currMB->mb_type == P16x8 || currMB->mb_type == P8x16 || currMB->mb_type == P8x8)) return;
  if (currSlice->weighted_prediction || currSlice->NoResidueDirect || p_Vid->bitdepth_chroma != 8 || (yuv != YUV420 && yuv != YUV422)) return;
  if (currMB->is_field_mode) return;
  nb = (yuv == YUV420) ? 4 : 8; hmb = 2 * nb;
  memset(&pred, 0, sizeof(pred));
  for (k = 0; k < 4; k++)
  {
    PicMotionParams *m = &p_Vid->enc_picture->mv_info[currMB->block_y + 2 * (k >> 1)][currMB->block_x + 2 * (k & 1)];
    int mode = currMB->b8x8[k].mode, ref = m->ref_idx[LIST_0];
    MEBlock probe;
    if (currMB->b8x8[k].pdir != 0 || mode < 1 || mode > 7 || ref < 0) return;
    memset(&probe, 0, sizeof(probe));
    pred.b8mode[k] = (uint8_t)mode;
    pred.ref[k] = (uint8_t)ref_index_of(p_Vid, currSlice, &probe, currSlice->listX[LIST_0 + currMB->list_offset][ref]);
    for (by = 2 * (k >> 1); by < 2 * (k >> 1) + 2; by++)
      for (bx = 2 * (k & 1); bx < 2 * (k & 1) + 2; bx++)
      {
        MotionVector *mv = &currSlice->all_mv[LIST_0][ref][mode][by][bx];
        pred.mv[by * 4 + bx][0] = mv->mv_x;
        pred.mv[by * 4 + bx][1] = mv->mv_y;
      }
  }
  memset(&d, 0, sizeof(d));
  d.yuv_format = yuv;
  d.is_cavlc = (currSlice->symbol_mode == CAVLC);
  for (uv = 0; uv < 2; uv++)
  {
    int qp = currMB->qpc[uv] + currSlice->bitdepth_chroma_qp_scale, qp_dc = qp + (yuv == YUV422 ? 3 : 0);
    LevelQuantParams **qa = p_Quant->q_params_4x4[uv + 1][0][qp], **qd = p_Quant->q_params_4x4[uv + 1][0][qp_dc];
    d.qp_ac[uv] = qp; d.qp_dc[uv] = qp_dc;
    for (j = 0; j < 4; j++)
      for (i = 0; i < 4; i++)
      {
        d.params_ac[uv][j * 4 + i][0] = qa[j][i].OffsetComp; d.params_ac[uv][j * 4 + i][1] = qa[j][i].ScaleComp; d.params_ac[uv][j * 4 + i][2] = qa[j][i].InvScaleComp;
      }
    d.params_dc[uv][0] = qd[0][0].OffsetComp; d.params_dc[uv][1] = qd[0][0].ScaleComp; d.params_dc[uv][2] = qd[0][0].InvScaleComp;
  }
  memcpy(d.c_cost, V_COST4[currSlice->disthres], 16);
  rc = jmb_chroma_residual_coding(S.ctx, &pred, 0, currMB->mbAddrX, 1, &d, &dc[0][0], &ac[0][0][0], cb, cc, &recon[0][0][0], JMB_HOST);
  if (rc) jmb_die("jmb_chroma_residual_coding", rc);

  if ((int)(cc[0] > cc[1] ? cc[0] : cc[1]) != ((currMB->cbp - cbp_before) >> 4)) chroma_mismatch(currMB, "chroma cbp", 0, (currMB->cbp - cbp_before) >> 4, (int)(cc[0] > cc[1] ? cc[0] : cc[1]));
  for (uv = 0; uv < 2; uv++)
  {
    int uv_scale = uv * (p_Vid->num_blk8x8_uv >> 1), e, pos, b;
    int16_t want[16];
    int jm_bits = (int)((currMB->cbp_blk >> (16 + uv * nb)) & ((1 << nb) - 1));
    if (jm_bits != (int)cb[uv]) chroma_mismatch(currMB, "cbp_blk chroma bits", uv, jm_bits, (int)cb[uv]);
    for (j = 0; j < hmb; j++)
      for (i = 0; i < 8; i++)
        if (p_Vid->enc_picture->imgUV[uv][currMB->pix_c_y + j][currMB->pix_c_x + i] != recon[uv][j][i])
          chroma_mismatch(currMB, "reconstruction", uv, p_Vid->enc_picture->imgUV[uv][currMB->pix_c_y + j][currMB->pix_c_x + i], recon[uv][j][i]);
    memset(want, 0, sizeof(want));
    for (e = 0, pos = 0; e < nb && currSlice->cofDC[uv + 1][0][e] != 0; e++) { pos += currSlice->cofDC[uv + 1][1][e]; want[pos++] = (int16_t)currSlice->cofDC[uv + 1][0][e]; }
    for (e = 0; e < nb; e++) if (want[e] != dc[uv][e]) chroma_mismatch(currMB, "a DC level", uv, want[e], dc[uv][e]);
    for (b = 0; b < nb; b++)
    {
      int *ACL = currSlice->cofAC[4 + (b >> 2) + uv_scale][b & 3][0], *ACR = currSlice->cofAC[4 + (b >> 2) + uv_scale][b & 3][1];
      memset(want, 0, sizeof(want));
      for (e = 0, pos = 0; e < 15 && ACL[e] != 0; e++) { pos += ACR[e]; want[pos++] = (int16_t)ACL[e]; }
      for (e = 0; e < 15; e++) if (want[e] != ac[uv][b][e]) chroma_mismatch(currMB, "an AC level", uv, want[e], ac[uv][b][e]);
    }
  }
  chroma_verified++;
}

/* ---- transforms (lcommon/src/transform.c:20, :353) ---------------------------------------------------- */
static void forward_nxn(int **block, int **tblock, int pos_y, int pos_x, int n)
{
  int32_t buf[64];
  int i, j, rc;
  for (j = 0; j < n; j++)
    for (i = 0; i < n; i++) buf[j * n + i] = block[pos_y + j][pos_x + i];
  rc = jmb_forward_transform(S.ctx, buf, 1, n, JMB_HOST);
  if (rc) jmb_die("jmb_forward_transform", rc);
  for (j = 0; j < n; j++)
    for (i = 0; i < n; i++) tblock[pos_y + j][pos_x + i] = buf[j * n + i];
}

void __wrap_forward4x4(int **block, int **tblock, int pos_y, int pos_x)
{
  if (!shim_on(FAM_TQ)) { __real_forward4x4(block, tblock, pos_y, pos_x); return; }
  S.calls[4]++;
  forward_nxn(block, tblock, pos_y, pos_x, 4);
}

void __wrap_forward8x8(int **block, int **tblock, int pos_y, int pos_x)
{
  if (!shim_on(FAM_TQ)) { __real_forward8x8(block, tblock, pos_y, pos_x); return; }
  S.calls[5]++;
  forward_nxn(block, tblock, pos_y, pos_x, 8);
}

/* inverse4x4 / inverse8x8 (lcommon/src/transform.c:70, :450): dequantised coefficients -> residual, used by the
 * reconstruction half of residual_transform_quant_luma_* (block.c:708, transform8x8.c:572) and the chroma / intra paths */
static void inverse_nxn(int **tblock, int **block, int pos_y, int pos_x, int n)
{
  int32_t buf[64];
  int i, j, rc;
  for (j = 0; j < n; j++)
    for (i = 0; i < n; i++) buf[j * n + i] = tblock[pos_y + j][pos_x + i];
  rc = jmb_inverse_transform(S.ctx, buf, 1, n, JMB_HOST);
  if (rc) jmb_die("jmb_inverse_transform", rc);
  for (j = 0; j < n; j++)
    for (i = 0; i < n; i++) block[pos_y + j][pos_x + i] = buf[j * n + i];
}

void __wrap_inverse4x4(int **tblock, int **block, int pos_y, int pos_x)
{
  if (!shim_on(FAM_TQ)) { __real_inverse4x4(tblock, block, pos_y, pos_x); return; }
  S.calls[4]++;
  inverse_nxn(tblock, block, pos_y, pos_x, 4);
}

void __wrap_inverse8x8(int **tblock, int **block, int pos_x)
{
  if (!shim_on(FAM_TQ)) { __real_inverse8x8(tblock, block, pos_x); return; }
  S.calls[5]++;
  inverse_nxn(tblock, block, 0, pos_x, 8);
}

/* ---- quantisation (lencod/src/quant4x4_normal.c:39, quant4x4_around.c:40, quant8x8_normal.c:43,:123,
 *      quant8x8_around.c) --------------------------------------------------------------------------------
 * tblock is JM's row-pointer view of the coefficient block (rows already offset by the caller, columns by
 * q_method->block_x); on return it holds the dequantised coefficients, ACLevel/ACRun (or cofAC for the 8x8
 * CAVLC form) the level/run lists, *coeff_cost the running cost, fadjust the adaptive-rounding terms. */
static int quant_nxn(Macroblock *currMB, int **tblock, struct quant_methods *qm, int n, int around, int cavlc8, int ***cofAC)
{
  VideoParameters *p_Vid = currMB->p_Vid;
  Slice *currSlice = currMB->p_Slice;
  jmb_quant_desc d;
  int32_t coef[64], levels[68], runs[68], fadj[64], cost, nz;
  int i, j, k, rc, bx = qm->block_x;
  const int nn = n * n;

  memset(&d, 0, sizeof(d));
  d.n = n;
  d.qp = qm->qp;
  d.is_cavlc = (n == 4) ? (currSlice->symbol_mode == CAVLC) : cavlc8;
  d.around = around;
  d.adapt_rnd_weight = p_Vid->AdaptRndWeight;
  for (j = 0; j < n; j++)
    for (i = 0; i < n; i++)
    {
      d.qparams[j * n + i][0] = qm->q_params[j][i].OffsetComp;
      d.qparams[j * n + i][1] = qm->q_params[j][i].ScaleComp;
      d.qparams[j * n + i][2] = qm->q_params[j][i].InvScaleComp;
      coef[j * n + i] = tblock[j][bx + i];
    }
  for (k = 0; k < nn; k++)
  {
    d.scan[k][0] = qm->pos_scan[k][0];
    d.scan[k][1] = qm->pos_scan[k][1];
  }
  /* c_cost is indexed by the run (< 16 per list for 4x4 / 8x8 CAVLC, < 64 for 8x8); COEFF_COST rows are 16 / 64 long */
  for (k = 0; k < ((n == 4 || cavlc8) ? 16 : 64); k++) d.c_cost[k] = qm->c_cost[k];
  cost = *qm->coeff_cost;
  rc = jmb_quant_blocks(S.ctx, &d, 0, coef, 1, levels, runs, around ? fadj : NULL, &cost, &nz, JMB_HOST);
  if (rc) jmb_die("jmb_quant_blocks", rc);
  *qm->coeff_cost = cost;
  for (j = 0; j < n; j++)
    for (i = 0; i < n; i++)
    {
      tblock[j][bx + i] = coef[j * n + i];
      if (around) qm->fadjust[j][bx + i] = fadj[j * n + i];
    }
  if (cavlc8)
  {
    for (k = 0; k < 4; k++)
    {
      int *ACL = &cofAC[k][0][0], *ACR = &cofAC[k][1][0];
      for (i = 0; levels[17 * k + i] != 0; i++) { ACL[i] = levels[17 * k + i]; ACR[i] = runs[17 * k + i]; }
      ACL[i] = 0;
    }
  }
  else
  {
    for (i = 0; levels[i] != 0; i++) { qm->ACLevel[i] = levels[i]; qm->ACRun[i] = runs[i]; }
    qm->ACLevel[i] = 0;
  }
  return nz;
}

int __wrap_quant_4x4_normal(Macroblock *currMB, int **tblock, struct quant_methods *q_method)
{
  if (!shim_on(FAM_TQ)) return __real_quant_4x4_normal(currMB, tblock, q_method);
  S.calls[6]++;
  return quant_nxn(currMB, tblock, q_method, 4, 0, 0, NULL);
}
int __wrap_quant_4x4_around(Macroblock *currMB, int **tblock, struct quant_methods *q_method)
{
  if (!shim_on(FAM_TQ)) return __real_quant_4x4_around(currMB, tblock, q_method);
  S.calls[6]++;
  return quant_nxn(currMB, tblock, q_method, 4, 1, 0, NULL);
}
int __wrap_quant_8x8_normal(Macroblock *currMB, int **tblock, struct quant_methods *q_method)
{
  if (!shim_on(FAM_TQ)) return __real_quant_8x8_normal(currMB, tblock, q_method);
  S.calls[7]++;
  return quant_nxn(currMB, tblock, q_method, 8, 0, 0, NULL);
}
int __wrap_quant_8x8_around(Macroblock *currMB, int **tblock, struct quant_methods *q_method)
{
  if (!shim_on(FAM_TQ)) return __real_quant_8x8_around(currMB, tblock, q_method);
  S.calls[7]++;
  return quant_nxn(currMB, tblock, q_method, 8, 1, 0, NULL);
}
int __wrap_quant_8x8cavlc_normal(Macroblock *currMB, int **tblock, struct quant_methods *q_method, int ***cofAC)
{
  if (!shim_on(FAM_TQ)) return __real_quant_8x8cavlc_normal(currMB, tblock, q_method, cofAC);
  S.calls[7]++;
  return quant_nxn(currMB, tblock, q_method, 8, 0, 1, cofAC);
}
int __wrap_quant_8x8cavlc_around(Macroblock *currMB, int **tblock, struct quant_methods *q_method, int ***cofAC)
{
  if (!shim_on(FAM_TQ)) return __real_quant_8x8cavlc_around(currMB, tblock, q_method, cofAC);
  S.calls[7]++;
  return quant_nxn(currMB, tblock, q_method, 8, 1, 1, cofAC);
}


/* ---- deblocking: DeblockFrame (lencod/src/loopFilter.c:63), called once per coded picture (image.c:236) -----------------------
 * The wrapper describes every macroblock as DeblockMb and the strength functions read it (jmb_db_mb), hands the reconstructed
 * planes over as bytes, and takes them back filtered (jmb_deblock_picture: a wavefront of macroblocks on the device).  With
 * JMB_SHIM_VERIFY the real DeblockFrame runs as well, on a copy, and every sample is compared. */
void __wrap_DeblockFrame(VideoParameters *p_Vid, imgpel **imgY, imgpel ***imgUV)
{
  const int w = p_Vid->width, h = p_Vid->height, mbw = w / 16, mbh = h / 16, n = mbw * mbh;
  const int yuv = imgUV ? p_Vid->yuv_format : 0, wc = w / 2, hc = p_Vid->yuv_format == YUV420 ? h / 2 : h;
  PicMotionParams **mv_info = p_Vid->enc_picture->mv_info;
  StorablePicture *pics[64];
  int npics = 0, i, k, l, x, y, rc, slice_type;
  jmb_db_mb *mbs;
  uint8_t *pl[3];
  if (!shim_on(FAM_DEBLOCK)) { __real_DeblockFrame(p_Vid, imgY, imgUV); return; }
  if (p_Vid->mb_aff_frame_flag || p_Vid->structure != FRAME) unsupported("deblocking of field / MBAFF pictures");
  if (p_Vid->P444_joined || p_Vid->yuv_format == YUV444) unsupported("deblocking of 4:4:4 pictures");
  if (p_Vid->bitdepth_luma != 8 || (p_Vid->yuv_format != YUV400 && p_Vid->bitdepth_chroma != 8)) unsupported("deblocking at bit depths above 8");
  if ((int)p_Vid->PicSizeInMbs != n) unsupported("deblocking: picture size is not a whole number of macroblocks");
  slice_type = p_Vid->mb_data[0].p_Slice->slice_type;
  if (slice_type != P_SLICE && slice_type != B_SLICE && slice_type != I_SLICE) unsupported("deblocking of SP / SI slices");
  mbs = (jmb_db_mb *)calloc(n, sizeof(jmb_db_mb));
  pl[0] = (uint8_t *)malloc((size_t)w * h); pl[1] = (uint8_t *)malloc((size_t)wc * hc + 1); pl[2] = (uint8_t *)malloc((size_t)wc * hc + 1);
  if (!mbs || !pl[0] || !pl[1] || !pl[2]) no_mem_exit("libjmb200 shim: deblocking buffers");
  for (i = 0; i < n; i++)
  {
    Macroblock *m = &p_Vid->mb_data[i];
    jmb_db_mb *d = &mbs[i];
    int ipcm = m->mb_type == IPCM;      /* init_Deblock (loopFilter.c:40-47) zeroes the QPs of IPCM macroblocks, for good */
    if (ipcm) { m->qp = 0; m->qpc[0] = 0; m->qpc[1] = 0; }
    if (m->p_Slice->slice_type != slice_type) unsupported("deblocking of a picture whose slices differ in type");
    d->mb_type = (uint8_t)m->mb_type;
    d->flags = (uint8_t)((m->luma_transform_size_8x8_flag ? JMB_DB_T8X8 : 0) | (m->cbp ? JMB_DB_CBP : 0) | (m->mbAvailA ? JMB_DB_AVAIL_A : 0) | (m->mbAvailB ? JMB_DB_AVAIL_B : 0));
    d->qp = (int8_t)(ipcm ? 0 : m->qp); d->qpc[0] = (int8_t)(ipcm ? 0 : m->qpc[0]); d->qpc[1] = (int8_t)(ipcm ? 0 : m->qpc[1]);
    d->df_disable_idc = (int8_t)m->DFDisableIdc; d->df_alpha_c0_offset = (int8_t)m->DFAlphaC0Offset; d->df_beta_offset = (int8_t)m->DFBetaOffset;
    d->cbp_blk = (uint32_t)(m->cbp_blk & 0xffff);
    for (k = 0; k < 16; k++)
    {
      PicMotionParams *p = &mv_info[(i / mbw) * 4 + k / 4][(i % mbw) * 4 + k % 4];
      for (l = 0; l < 2; l++)
      {
        int id = -1;
        if (p->ref_idx[l] != -1)
        {
          for (id = 0; id < npics && pics[id] != p->ref_pic[l]; id++) ;
          if (id == npics) { if (npics == 64) unsupported("deblocking: more than 64 distinct reference pictures in one picture"); pics[npics++] = p->ref_pic[l]; }
        }
        d->ref_id[l][k] = (int8_t)id;
        d->mv[l][k][0] = p->mv[l].mv_x; d->mv[l][k][1] = p->mv[l].mv_y;
      }
    }
  }
  for (y = 0; y < h; y++) for (x = 0; x < w; x++) pl[0][(size_t)y * w + x] = (uint8_t)imgY[y][x];
  if (yuv) for (k = 0; k < 2; k++) for (y = 0; y < hc; y++) for (x = 0; x < wc; x++) pl[1 + k][(size_t)y * wc + x] = (uint8_t)imgUV[k][y][x];
  rc = jmb_deblock_picture(S.ctx, pl[0], w, yuv ? pl[1] : NULL, yuv ? pl[2] : NULL, wc, w, h, yuv, slice_type, p_Vid->active_sps->direct_8x8_inference_flag, mbs, JMB_HOST);
  if (rc) jmb_die("jmb_deblock_picture", rc);
  if (S.verify)
  { /* JM's own filter on JM's own planes; the device result must be the same sample for sample */
    __real_DeblockFrame(p_Vid, imgY, imgUV);
    for (y = 0; y < h; y++) for (x = 0; x < w; x++)
      if (imgY[y][x] != pl[0][(size_t)y * w + x]) { snprintf(errortext, ET_SIZE, "libjmb200 shim: deblocked luma differs from JM's at (%d,%d): %d vs %d", x, y, pl[0][(size_t)y * w + x], imgY[y][x]); fatal(703); }
    if (yuv) for (k = 0; k < 2; k++) for (y = 0; y < hc; y++) for (x = 0; x < wc; x++)
      if (imgUV[k][y][x] != pl[1 + k][(size_t)y * wc + x]) { snprintf(errortext, ET_SIZE, "libjmb200 shim: deblocked chroma %d differs from JM's at (%d,%d)", k, x, y); fatal(703); }
    S.deblock_verified++;
  }
  else
  {
    for (y = 0; y < h; y++) for (x = 0; x < w; x++) imgY[y][x] = pl[0][(size_t)y * w + x];
    if (yuv) for (k = 0; k < 2; k++) for (y = 0; y < hc; y++) for (x = 0; x < wc; x++) imgUV[k][y][x] = pl[1 + k][(size_t)y * wc + x];
  }
  S.deblocked++;
  free(mbs); free(pl[0]); free(pl[1]); free(pl[2]);
}
