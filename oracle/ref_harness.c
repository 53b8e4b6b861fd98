/*
 * oracle/ref_harness.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Flat C entry points around the REAL JM 19.0 leaf functions of the ME + transform/quant hot
 * path.  It is compiled against JM's own headers and linked with JM's own (unmodified) objects
 * into oracle/_ref/libjmref.so by oracle/Makefile ("make -C oracle ref").  Nothing of JM is
 * re-implemented here: this file only builds the minimal VideoParameters / InputParameters /
 * Slice / Macroblock / MEBlock / StorablePicture state each leaf dereferences, then calls it.
 *
 * Leaves called (all in /root/reference):
 *   getSubImagesLuma                     lencod/src/img_luma.c:611
 *   init_motion_search_module            lencod/src/mv_search.c:315   (spiral + mvbits tables)
 *   full_search_motion_estimation        lencod/src/me_fullsearch.c:39
 *   sub_pel_motion_estimation            lencod/src/me_fullsearch.c:186
 *   setup_fast_full_search               lencod/src/me_fullfast.c:269
 *   fast_full_search_motion_estimation   lencod/src/me_fullfast.c:618
 *   computeSAD / computeSATD / computeSSE  lencod/src/me_distortion.c:349,745,1190
 *   forward4x4 / forward8x8              lcommon/src/transform.c:20,353
 *   quant_4x4_normal/_around             lencod/src/quant4x4_normal.c:39, quant4x4_around.c:40
 *   quant_8x8_normal/_around, quant_8x8cavlc_normal/_around   lencod/src/quant8x8_*.c
 *   GetMotionVectorPredictorNormal       lcommon/src/mv_prediction.c:192 (through init_motion_vector_prediction, :306)
 */
#include <stdint.h>
#include <string.h>
#include <stdlib.h>

#include "global.h"
#include "mbuffer.h"
#include "memalloc.h"
#include "img_luma.h"
#include "mv_search.h"
#include "me_fullsearch.h"
#include "me_fullfast.h"
#include "me_distortion.h"
#include "transform.h"
#include "quant4x4.h"
#include "quant8x8.h"
#include "quantChroma.h"
#include "refbuf.h"
#include "mv_prediction.h"

typedef struct jmref_ctx
{
  VideoParameters *p_Vid;
  InputParameters *p_Inp;
  Slice           *slice;
  Macroblock      *mb;
  StorablePicture *ref;
  StorablePicture *ref_list[2];
  StorablePicture *ref2;       /* second reference of the bi-predictive distortions (jmref_set_ref2) */
  imgpel         **cur;        /* current (source) luma, row pointers */
  int              w, h;
  MotionVector     stub_pmv;   /* what the stubbed GetMVPredictor returns */
  QuantParameters *quant;
  int             *qp_per, *qp_rem;
} jmref_ctx;

static jmref_ctx *g_ctx; /* the stubs below need it; single-threaded like JM */

static void stub_getNeighbour(Macroblock *currMB, int xN, int yN, int mb_size[2], PixelPos *pix)
{
  (void)currMB; (void)xN; (void)yN; (void)mb_size;
  memset(pix, 0, sizeof(*pix));
}

static void stub_GetMVPredictor(Macroblock *currMB, PixelPos *block, MotionVector *pmv, short ref_frame,
                                struct pic_motion_params **mv_info, int list, int mb_x, int mb_y, int bsx, int bsy)
{
  (void)currMB; (void)block; (void)ref_frame; (void)mv_info; (void)list; (void)mb_x; (void)mb_y; (void)bsx; (void)bsy;
  *pmv = g_ctx->stub_pmv;
}

/* metric codes as JM's: 0 = SAD, 1 = SSE, 2 = SATD (lencod/inc/defines.h ERROR_*) */
void *jmref_open(int width, int height, int search_range, int metric_f, int metric_h, int metric_q,
                 int fast_full, int rdopt, int bitdepth, int level_vmv_qpel /* e.g. 2048 for level >= 3.1 */)
{
  jmref_ctx *c = (jmref_ctx *)calloc(1, sizeof(*c));
  VideoParameters *p_Vid = (VideoParameters *)calloc(1, sizeof(VideoParameters));
  InputParameters *p_Inp = (InputParameters *)calloc(1, sizeof(InputParameters));
  c->p_Vid = p_Vid; c->p_Inp = p_Inp; c->w = width; c->h = height;
  p_Vid->p_Inp = p_Inp;

  p_Inp->search_range[0] = p_Inp->search_range[1] = search_range;
  p_Inp->MEErrorMetric[F_PEL] = metric_f;
  p_Inp->MEErrorMetric[H_PEL] = metric_h;
  p_Inp->MEErrorMetric[Q_PEL] = metric_q;
  p_Inp->ModeDecisionMetric = ERROR_SATD;
  p_Inp->SearchMode[0] = p_Inp->SearchMode[1] = fast_full ? FAST_FULL_SEARCH : FULL_SEARCH;
  p_Inp->full_search = 2;              /* RestrictSearchRange = 2: no restriction (all bundled cfgs) */
  p_Inp->rdopt = rdopt;
  p_Inp->ChromaMEEnable = 0;
  p_Inp->OnTheFlyFractMCP = 0;

  p_Vid->max_num_references = 1;
  p_Vid->bitdepth_luma = (short)bitdepth;
  p_Vid->max_pel_value_comp[0] = p_Vid->max_pel_value_comp[1] = p_Vid->max_pel_value_comp[2] = (1 << bitdepth) - 1;
  p_Vid->max_imgpel_value = (short)((1 << bitdepth) - 1);
  p_Vid->width = width; p_Vid->height = height;
  p_Vid->padded_size_x      = width + 2 * IMG_PAD_SIZE_X;
  p_Vid->padded_size_x_m8x8 = p_Vid->padded_size_x - BLOCK_SIZE_8x8;
  p_Vid->padded_size_x_m4x4 = p_Vid->padded_size_x - BLOCK_SIZE;
  p_Vid->mb_size[0][0] = p_Vid->mb_size[0][1] = MB_BLOCK_SIZE;
  p_Vid->getNeighbour = stub_getNeighbour;
  p_Vid->searchRange.min_x = p_Vid->searchRange.min_y = -(search_range << 2);
  p_Vid->searchRange.max_x = p_Vid->searchRange.max_y =  (search_range << 2);
  p_Vid->MaxHmvR[0] = -2047; p_Vid->MaxHmvR[1] = 2047; p_Vid->MaxHmvR[2] = -4096;
  p_Vid->MaxHmvR[3] = 4095;  p_Vid->MaxHmvR[4] = -8192; p_Vid->MaxHmvR[5] = 8191;
  p_Vid->MaxVmvR[4] = -level_vmv_qpel;       p_Vid->MaxVmvR[5] = level_vmv_qpel - 1;
  p_Vid->MaxVmvR[2] = -(level_vmv_qpel / 2); p_Vid->MaxVmvR[3] = level_vmv_qpel / 2 - 1;
  p_Vid->MaxVmvR[0] = -(level_vmv_qpel / 4) + 1; p_Vid->MaxVmvR[1] = level_vmv_qpel / 4 - 1;
  p_Vid->active_pps = (pic_parameter_set_rbsp_t *)calloc(1, sizeof(pic_parameter_set_rbsp_t));
  p_Vid->mb_data = (Macroblock *)calloc(1, sizeof(Macroblock));
  p_Vid->enc_picture = (StorablePicture *)calloc(1, sizeof(StorablePicture)); /* only ->mv_info is read, by the stub */
  get_mem2Dint_pad(&p_Vid->imgY_sub_tmp, height, width, IMG_PAD_SIZE_Y, IMG_PAD_SIZE_X);

  init_motion_search_module(p_Vid, p_Inp);

  /* reference picture with its 16 quarter-pel planes (layout of mbuffer.c:438,562-565) */
  StorablePicture *s = (StorablePicture *)calloc(1, sizeof(StorablePicture));
  s->size_x = width; s->size_y = height;
  s->size_x_padded = width + 2 * IMG_PAD_SIZE_X;
  s->size_y_padded = height + 2 * IMG_PAD_SIZE_Y;
  s->size_x_pad = width + 2 * IMG_PAD_SIZE_X - 1 - MB_BLOCK_SIZE - IMG_PAD_SIZE_X;
  s->size_y_pad = height + 2 * IMG_PAD_SIZE_Y - 1 - MB_BLOCK_SIZE - IMG_PAD_SIZE_Y;
  get_mem2Dpel(&s->imgY, height, width);
  get_mem4Dpel_pad(&s->imgY_sub, 4, 4, height, width, IMG_PAD_SIZE_Y, IMG_PAD_SIZE_X);
  s->p_img_sub[0] = s->imgY_sub;
  s->p_curr_img = s->imgY;
  s->p_curr_img_sub = s->imgY_sub;
  c->ref = s;
  c->ref_list[0] = s; c->ref_list[1] = NULL;

  get_mem2Dpel(&c->cur, height, width);
  p_Vid->pCurImg = c->cur;

  c->slice = (Slice *)calloc(1, sizeof(Slice));
  c->slice->p_Vid = p_Vid; c->slice->p_Inp = p_Inp;
  c->slice->slice_type = P_SLICE;
  c->slice->listX[0] = c->ref_list;
  c->slice->listXsize[0] = 1;
  c->slice->symbol_mode = CAVLC;

  c->mb = (Macroblock *)calloc(1, sizeof(Macroblock));
  c->mb->p_Vid = p_Vid; c->mb->p_Inp = p_Inp; c->mb->p_Slice = c->slice;
  c->mb->GetMVPredictor = stub_GetMVPredictor;
  c->mb->p_SetupFastFullPelSearch = setup_fast_full_search;

  /* quantiser state dereferenced by quant_*: p_Vid->p_Quant->qp_per_matrix */
  c->quant = (QuantParameters *)calloc(1, sizeof(QuantParameters));
  c->qp_per = (int *)calloc(128, sizeof(int));
  c->qp_rem = (int *)calloc(128, sizeof(int));
  for (int i = 0; i < 128; i++) { c->qp_per[i] = i / 6; c->qp_rem[i] = i % 6; }
  c->quant->qp_per_matrix = c->qp_per; c->quant->qp_rem_matrix = c->qp_rem;
  p_Vid->p_Quant = c->quant;

  g_ctx = c;
  return c;
}

static StorablePicture *new_ref_picture(int width, int height)
{
  StorablePicture *s = (StorablePicture *)calloc(1, sizeof(StorablePicture));
  s->size_x = width; s->size_y = height;
  s->size_x_padded = width + 2 * IMG_PAD_SIZE_X;
  s->size_y_padded = height + 2 * IMG_PAD_SIZE_Y;
  s->size_x_pad = width + 2 * IMG_PAD_SIZE_X - 1 - MB_BLOCK_SIZE - IMG_PAD_SIZE_X;
  s->size_y_pad = height + 2 * IMG_PAD_SIZE_Y - 1 - MB_BLOCK_SIZE - IMG_PAD_SIZE_Y;
  get_mem2Dpel(&s->imgY, height, width);
  get_mem4Dpel_pad(&s->imgY_sub, 4, 4, height, width, IMG_PAD_SIZE_Y, IMG_PAD_SIZE_X);
  s->p_img_sub[0] = s->imgY_sub;
  s->p_curr_img = s->imgY;
  s->p_curr_img_sub = s->imgY_sub;
  return s;
}

void jmref_set_ref2(void *h, const uint16_t *luma, int stride)
{
  jmref_ctx *c = (jmref_ctx *)h;
  if (!c->ref2) c->ref2 = new_ref_picture(c->w, c->h);
  for (int y = 0; y < c->h; y++)
    memcpy(c->ref2->imgY[y], luma + (size_t)y * stride, c->w * sizeof(imgpel));
  getSubImagesLuma(c->p_Vid, c->ref2);
}

void jmref_set_ref(void *h, const uint16_t *luma, int stride)
{
  jmref_ctx *c = (jmref_ctx *)h;
  for (int y = 0; y < c->h; y++)
    memcpy(c->ref->imgY[y], luma + (size_t)y * stride, c->w * sizeof(imgpel));
  getSubImagesLuma(c->p_Vid, c->ref);
}

/* out: (h+2*PADY) x (w+2*PADX) samples, row-major, origin = padded top-left */
void jmref_get_subplane(void *h, int fy, int fx, uint16_t *out)
{
  jmref_ctx *c = (jmref_ctx *)h;
  int W = c->w + 2 * IMG_PAD_SIZE_X, H = c->h + 2 * IMG_PAD_SIZE_Y;
  for (int y = 0; y < H; y++)
    memcpy(out + (size_t)y * W, c->ref->imgY_sub[fy][fx][y - IMG_PAD_SIZE_Y] - IMG_PAD_SIZE_X, W * sizeof(imgpel));
}

void jmref_set_cur(void *h, const uint16_t *luma, int stride)
{
  jmref_ctx *c = (jmref_ctx *)h;
  for (int y = 0; y < c->h; y++)
    memcpy(c->cur[y], luma + (size_t)y * stride, c->w * sizeof(imgpel));
}

static const short k_bsize[8][2] = {{16,16},{16,16},{16,8},{8,16},{8,8},{8,4},{4,8},{4,4}};

static void fill_mv_block(jmref_ctx *c, MEBlock *b, int blocktype, int pos_x, int pos_y, int test8x8)
{
  VideoParameters *p_Vid = c->p_Vid;
  memset(b, 0, sizeof(*b));
  b->p_Vid = p_Vid; b->p_Slice = c->slice;
  b->blocktype = (short)blocktype;
  b->blocksize_x = k_bsize[blocktype][0];
  b->blocksize_y = k_bsize[blocktype][1];
  b->pos_x = (short)pos_x; b->pos_y = (short)pos_y;
  b->pos_x2 = (short)(pos_x >> 2); b->pos_y2 = (short)(pos_y >> 2);
  b->pos_x_padded = (short)(pos_x << 2); b->pos_y_padded = (short)(pos_y << 2);
  b->block_x = (short)((pos_x & 15) >> 2); b->block_y = (short)((pos_y & 15) >> 2);
  b->list = 0; b->ref_idx = 0;
  b->searchRange = p_Vid->searchRange;
  b->search_pos2 = 9; b->search_pos4 = 9;
  b->test8x8 = test8x8;
  b->computePredFPel = p_Vid->computeUniPred[F_PEL];
  b->computePredHPel = p_Vid->computeUniPred[H_PEL];
  b->computePredQPel = p_Vid->computeUniPred[Q_PEL];
  get_mem2Dpel(&b->orig_pic, 1, b->blocksize_x * b->blocksize_y);
  get_original_block(p_Vid, b);
  c->mb->pix_x = (short)(pos_x & ~15); c->mb->pix_y = (short)(pos_y & ~15);
  c->mb->opix_y = (short)(pos_y & ~15);
}

int64_t jmref_full_search(void *h, int blocktype, int pos_x, int pos_y, int pred_x, int pred_y,
                          int center_x, int center_y, int lambda, int64_t min_mcost, int16_t *mv_out)
{
  jmref_ctx *c = (jmref_ctx *)h; g_ctx = c;
  MEBlock b; MotionVector pred;
  fill_mv_block(c, &b, blocktype, pos_x, pos_y, 0);
  pred.mv_x = (short)pred_x; pred.mv_y = (short)pred_y;
  b.mv[0].mv_x = (short)center_x; b.mv[0].mv_y = (short)center_y;
  distblk r = full_search_motion_estimation(c->mb, &pred, &b, (distblk)min_mcost, lambda);
  mv_out[0] = b.mv[0].mv_x; mv_out[1] = b.mv[0].mv_y;
  free_mem2Dpel(b.orig_pic);
  return (int64_t)r;
}

int64_t jmref_sub_pel(void *h, int blocktype, int pos_x, int pos_y, int pred_x, int pred_y,
                      int mv_x, int mv_y, const int *lambda3, int64_t min_mcost, int test8x8, int16_t *mv_out)
{
  jmref_ctx *c = (jmref_ctx *)h; g_ctx = c;
  MEBlock b; MotionVector pred; int lam[3] = { lambda3[0], lambda3[1], lambda3[2] };
  fill_mv_block(c, &b, blocktype, pos_x, pos_y, test8x8);
  pred.mv_x = (short)pred_x; pred.mv_y = (short)pred_y;
  b.mv[0].mv_x = (short)mv_x; b.mv[0].mv_y = (short)mv_y;
  distblk r = sub_pel_motion_estimation(c->mb, &pred, &b, (distblk)min_mcost, lam);
  mv_out[0] = b.mv[0].mv_x; mv_out[1] = b.mv[0].mv_y;
  free_mem2Dpel(b.orig_pic);
  return (int64_t)r;
}

/* metric: 0 SAD, 1 SSE, 2 SATD; cand = absolute quarter-pel position (pos<<2 + mv) */
int64_t jmref_dist(void *h, int metric, int blocktype, int pos_x, int pos_y, int cand_x, int cand_y,
                   int test8x8, int64_t min_mcost)
{
  jmref_ctx *c = (jmref_ctx *)h; g_ctx = c;
  MEBlock b; MotionVector cand;
  fill_mv_block(c, &b, blocktype, pos_x, pos_y, test8x8);
  cand.mv_x = (short)cand_x; cand.mv_y = (short)cand_y;
  distblk r;
  if (metric == 0)      r = computeSAD (c->ref, &b, (distblk)min_mcost, &cand);
  else if (metric == 1) r = computeSSE (c->ref, &b, (distblk)min_mcost, &cand);
  else                  r = computeSATD(c->ref, &b, (distblk)min_mcost, &cand);
  free_mem2Dpel(b.orig_pic);
  return (int64_t)r;
}

/* the twelve members of the distortion table (mv_search.c:486-506): form 0 computeSAD/SSE/SATD, 1 compute*WP,
 * 2 computeBiPred*1, 3 computeBiPred*2; wp = {weight1, weight2, offset, luma_log_weight_denom, wp_luma_round} */
int64_t jmref_dist_ex(void *h, int metric, int form, int blocktype, int pos_x, int pos_y, int c1x, int c1y, int c2x, int c2y,
                      int test8x8, int64_t min_mcost, const int *wp)
{
  jmref_ctx *c = (jmref_ctx *)h; g_ctx = c;
  MEBlock b; MotionVector cand1, cand2;
  fill_mv_block(c, &b, blocktype, pos_x, pos_y, test8x8);
  cand1.mv_x = (short)c1x; cand1.mv_y = (short)c1y; cand2.mv_x = (short)c2x; cand2.mv_y = (short)c2y;
  b.weight_luma = (short)wp[0]; b.offset_luma = (short)wp[2];
  b.weight1 = (short)wp[0]; b.weight2 = (short)wp[1]; b.offsetBi = (short)wp[2];
  c->slice->luma_log_weight_denom = (short)wp[3]; c->slice->wp_luma_round = wp[4];
  distblk r = 0, m = (distblk)min_mcost;
  switch (form * 3 + metric) {
  case 0:  r = computeSAD (c->ref, &b, m, &cand1); break;
  case 1:  r = computeSSE (c->ref, &b, m, &cand1); break;
  case 2:  r = computeSATD(c->ref, &b, m, &cand1); break;
  case 3:  r = computeSADWP (c->ref, &b, m, &cand1); break;
  case 4:  r = computeSSEWP (c->ref, &b, m, &cand1); break;
  case 5:  r = computeSATDWP(c->ref, &b, m, &cand1); break;
  case 6:  r = computeBiPredSAD1 (c->ref, c->ref2, &b, m, &cand1, &cand2); break;
  case 7:  r = computeBiPredSSE1 (c->ref, c->ref2, &b, m, &cand1, &cand2); break;
  case 8:  r = computeBiPredSATD1(c->ref, c->ref2, &b, m, &cand1, &cand2); break;
  case 9:  r = computeBiPredSAD2 (c->ref, c->ref2, &b, m, &cand1, &cand2); break;
  case 10: r = computeBiPredSSE2 (c->ref, c->ref2, &b, m, &cand1, &cand2); break;
  case 11: r = computeBiPredSATD2(c->ref, c->ref2, &b, m, &cand1, &cand2); break;
  }
  free_mem2Dpel(b.orig_pic);
  return (int64_t)r;
}

/* The real bi-predictive searches (me_fullsearch.c:112, :299): list 0 moves (reference = c->ref), list 1 stands still at mv2
 * (reference = c->ref2).  form 2: computeBiPred*1, form 3: computeBiPred*2 with wp = {weight1, weight2, offsetBi, denom, round}.
 * which = 0: full_search_bipred_motion_estimation (search_range in quarter-pels), 1: sub_pel_bipred_motion_estimation. */
int64_t jmref_bipred_search(void *h, int which, int form, int blocktype, int pos_x, int pos_y, const int *pred1, const int *pred2,
                            const int *mv1_in, const int *mv2_in, int search_range_qpel, const int *lambda3, int64_t min_mcost,
                            int test8x8, const int *wp, int16_t *mv_out)
{
  jmref_ctx *c = (jmref_ctx *)h; g_ctx = c;
  VideoParameters *p_Vid = c->p_Vid;
  MEBlock b; MotionVector p1, p2, m1, m2; int lam[3] = { lambda3[0], lambda3[1], lambda3[2] };
  StorablePicture *l1[1] = { c->ref2 };
  fill_mv_block(c, &b, blocktype, pos_x, pos_y, test8x8);
  c->slice->listX[1] = l1; c->slice->listXsize[1] = 1;
  b.weight1 = (short)wp[0]; b.weight2 = (short)wp[1]; b.offsetBi = (short)wp[2];
  c->slice->luma_log_weight_denom = (short)wp[3]; c->slice->wp_luma_round = wp[4];
  b.computeBiPredFPel = form == 3 ? p_Vid->computeBiPred2[F_PEL] : p_Vid->computeBiPred1[F_PEL];
  b.computeBiPredHPel = form == 3 ? p_Vid->computeBiPred2[H_PEL] : p_Vid->computeBiPred1[H_PEL];
  b.computeBiPredQPel = form == 3 ? p_Vid->computeBiPred2[Q_PEL] : p_Vid->computeBiPred1[Q_PEL];
  p1.mv_x = (short)pred1[0]; p1.mv_y = (short)pred1[1]; p2.mv_x = (short)pred2[0]; p2.mv_y = (short)pred2[1];
  m1.mv_x = (short)mv1_in[0]; m1.mv_y = (short)mv1_in[1]; m2.mv_x = (short)mv2_in[0]; m2.mv_y = (short)mv2_in[1];
  distblk r;
  if (which == 0) r = full_search_bipred_motion_estimation(c->mb, 0, &p1, &p2, &m1, &m2, &b, search_range_qpel, (distblk)min_mcost, lam[0]);
  else            r = sub_pel_bipred_motion_estimation(c->mb, &b, 0, &p1, &p2, &m1, &m2, (distblk)min_mcost, lam);
  mv_out[0] = m1.mv_x; mv_out[1] = m1.mv_y;
  free_mem2Dpel(b.orig_pic);
  return (int64_t)r;
}

/* Fast full search: one setup per macroblock (pmv = 16x16 predictor), then per-partition arg-min. */
void jmref_ffs_setup(void *h, int mb_pix_x, int mb_pix_y, int pmv_x, int pmv_y, int16_t *center_out)
{
  jmref_ctx *c = (jmref_ctx *)h; g_ctx = c;
  MEBlock b;
  fill_mv_block(c, &b, 1, mb_pix_x, mb_pix_y, 0);
  c->stub_pmv.mv_x = (short)pmv_x; c->stub_pmv.mv_y = (short)pmv_y;
  reset_fast_full_search(c->p_Vid);
  setup_fast_full_search(c->mb, &b, 0);
  center_out[0] = c->p_Vid->p_ffast_me->search_center[0][0].mv_x;
  center_out[1] = c->p_Vid->p_ffast_me->search_center[0][0].mv_y;
  free_mem2Dpel(b.orig_pic);
}

void jmref_ffs_get_sad(void *h, int blocktype, int block_index, uint32_t *out, int max_pos)
{
  jmref_ctx *c = (jmref_ctx *)h;
  distpel *s = c->p_Vid->p_ffast_me->BlockSAD[0][0][blocktype][block_index];
  for (int i = 0; i < max_pos; i++) out[i] = (uint32_t)s[i];
}

int64_t jmref_ffs_search(void *h, int blocktype, int pos_x, int pos_y, int pred_x, int pred_y,
                         int lambda, int64_t min_mcost, int16_t *mv_out)
{
  jmref_ctx *c = (jmref_ctx *)h; g_ctx = c;
  MEBlock b; MotionVector pred;
  fill_mv_block(c, &b, blocktype, pos_x, pos_y, 0);
  pred.mv_x = (short)pred_x; pred.mv_y = (short)pred_y;
  distblk r = fast_full_search_motion_estimation(c->mb, &pred, &b, (distblk)min_mcost, lambda);
  mv_out[0] = b.mv[0].mv_x; mv_out[1] = b.mv[0].mv_y;
  free_mem2Dpel(b.orig_pic);
  return (int64_t)r;
}

int jmref_spiral(void *h, int16_t *out_xy, int max_pos)
{
  jmref_ctx *c = (jmref_ctx *)h;
  for (int i = 0; i < max_pos; i++) {
    out_xy[2 * i] = c->p_Vid->spiral_search[i].mv_x;
    out_xy[2 * i + 1] = c->p_Vid->spiral_search[i].mv_y;
  }
  return c->p_Vid->max_mvd;
}

int jmref_mvbits(void *h, int v) { return ((jmref_ctx *)h)->p_Vid->mvbits[v]; }

/* ---- transforms (stateless) ---- */
static void call_fwd(void (*f)(int **, int **, int, int), int *blk, int n)
{
  int *rows[8];
  for (int i = 0; i < n; i++) rows[i] = blk + i * n;
  f(rows, rows, 0, 0);
}
void jmref_forward4x4(int *blk16) { call_fwd(forward4x4, blk16, 4); }
void jmref_forward8x8(int *blk64) { call_fwd(forward8x8, blk64, 8); }
void jmref_inverse4x4(int *blk16)
{
  int *rows[4];
  for (int i = 0; i < 4; i++) rows[i] = blk16 + i * 4;
  inverse4x4(rows, rows, 0, 0);
}
void jmref_inverse8x8(int *blk64)
{
  int *rows[8];
  for (int i = 0; i < 8; i++) rows[i] = blk64 + i * 8;
  inverse8x8(rows, rows, 0);
}
int  jmref_hadamard_sad4x4(short *d) { return HadamardSAD4x4(d); }
int  jmref_hadamard_sad8x8(short *d) { return HadamardSAD8x8(d); }

/* ---- quantisation ----
 * variant: 0 quant_4x4_normal, 1 quant_4x4_around, 2 quant_8x8_normal, 3 quant_8x8_around,
 *          4 quant_8x8cavlc_normal, 5 quant_8x8cavlc_around
 * coef   : n*n transformed coefficients, row-major (in) -> dequantised coefficients (out)
 * qparams: n*n triples {OffsetComp, ScaleComp, InvScaleComp}, row-major [j][i]
 * scan   : n*n pairs {i (horizontal), j (vertical)}
 * levels/runs: 4x4: [17]; 8x8: [65]; 8x8cavlc: [4][17] each
 * fadjust: n*n (around variants only)
 * returns nonzero; *coeff_cost is accumulated into.
 */
int jmref_quant(void *h, int variant, int *coef, int qp, const int *qparams, const uint8_t *scan,
                const uint8_t *c_cost, int is_cavlc, int adapt_rnd_weight,
                int *levels, int *runs, int *fadjust, int *coeff_cost)
{
  jmref_ctx *c = (jmref_ctx *)h; g_ctx = c;
  int n = (variant < 2) ? 4 : 8;
  int *rows[8], *frow[8];
  LevelQuantParams qp_store[64], *qrows[8];
  int fadj_dummy[64];
  for (int i = 0; i < n; i++) {
    rows[i] = coef + i * n;
    frow[i] = (fadjust ? fadjust : fadj_dummy) + i * n;
    qrows[i] = qp_store + i * n;
  }
  for (int i = 0; i < n * n; i++) {
    qp_store[i].OffsetComp = qparams[3 * i];
    qp_store[i].ScaleComp = qparams[3 * i + 1];
    qp_store[i].InvScaleComp = qparams[3 * i + 2];
  }
  c->slice->symbol_mode = is_cavlc ? CAVLC : CABAC;
  c->p_Vid->AdaptRndWeight = adapt_rnd_weight;
  QuantMethods q; memset(&q, 0, sizeof(q));
  q.block_x = 0; q.block_y = 0; q.qp = qp;
  q.ACLevel = levels; q.ACRun = runs; q.fadjust = frow;
  q.q_params = qrows; q.coeff_cost = coeff_cost;
  q.pos_scan = (const byte (*)[2])scan; q.c_cost = c_cost;
  if (variant >= 4) {
    /* cofAC[k][0|1][..] for the four interleaved CAVLC sub-blocks */
    int *lv[4][2]; int **cof[4];
    for (int k = 0; k < 4; k++) { lv[k][0] = levels + 17 * k; lv[k][1] = runs + 17 * k; cof[k] = lv[k]; }
    return (variant == 4) ? quant_8x8cavlc_normal(c->mb, rows, &q, cof) : quant_8x8cavlc_around(c->mb, rows, &q, cof);
  }
  switch (variant) {
    case 0: return quant_4x4_normal(c->mb, rows, &q);
    case 1: return quant_4x4_around(c->mb, rows, &q);
    case 2: return quant_8x8_normal(c->mb, rows, &q);
    default: return quant_8x8_around(c->mb, rows, &q);
  }
}

/* ---- the DC / AC members of the quantiser family, and the Hadamard transforms ----
 * variant: 6 quant_ac4x4_normal, 7 quant_ac4x4_around, 8 quant_dc4x4_normal, 9 quant_dc2x2_normal, 10 quant_dc2x2_around,
 *          11 quant_dc4x2_normal, 12 quant_dc4x2_around
 * coef: 6,7,8: 4x4 row-major [16]; 9,10: [4]; 11,12: 2 rows x 4 [8] (tblock[j][i], j < 2, i < 4)
 * qparams: 6,7: 16 triples [j][i]; others: ONE triple.  scan: pairs as JM reads them for that function. */
int jmref_quant_misc(void *h, int variant, int *coef, int qp, const int *qparams, const uint8_t *scan, const uint8_t *c_cost,
                     int is_cavlc, int adapt_rnd_weight, int *levels, int *runs, int *fadjust, int *coeff_cost)
{
  jmref_ctx *c = (jmref_ctx *)h; g_ctx = c;
  c->slice->symbol_mode = is_cavlc ? CAVLC : CABAC;
  c->p_Vid->AdaptRndWeight = adapt_rnd_weight;
  if (variant <= 7) {
    int *rows[4], *frow[4], fadj_dummy[16];
    LevelQuantParams qp_store[16], *qrows[4];
    for (int i = 0; i < 4; i++) { rows[i] = coef + 4 * i; frow[i] = (fadjust ? fadjust : fadj_dummy) + 4 * i; qrows[i] = qp_store + 4 * i; }
    for (int i = 0; i < 16; i++) { qp_store[i].OffsetComp = qparams[3 * i]; qp_store[i].ScaleComp = qparams[3 * i + 1]; qp_store[i].InvScaleComp = qparams[3 * i + 2]; }
    QuantMethods q; memset(&q, 0, sizeof(q));
    q.qp = qp; q.ACLevel = levels; q.ACRun = runs; q.fadjust = frow; q.q_params = qrows; q.coeff_cost = coeff_cost;
    q.pos_scan = (const byte (*)[2])scan; q.c_cost = c_cost;
    return variant == 6 ? quant_ac4x4_normal(c->mb, rows, &q) : quant_ac4x4_around(c->mb, rows, &q);
  }
  LevelQuantParams one = { qparams[0], qparams[1], qparams[2] };
  if (variant == 8) {
    int *rows[4];
    for (int i = 0; i < 4; i++) rows[i] = coef + 4 * i;
    return quant_dc4x4_normal(c->mb, rows, qp, levels, runs, &one, (const byte (*)[2])scan);
  }
  if (variant <= 10) {
    int *rows[1] = { coef };
    return variant == 9 ? quant_dc2x2_normal(c->mb, rows, qp, levels, runs, &one, NULL, (const byte (*)[2])scan)
                        : quant_dc2x2_around(c->mb, rows, qp, levels, runs, &one, NULL, (const byte (*)[2])scan);
  }
  {
    int *rows[2] = { coef, coef + 4 };        /* tblock[j][i], j = scan[k][0] < 2, i = scan[k][1] < 4 (block.c:88-94, :1076) */
    return variant == 11 ? quant_dc4x2_normal(c->mb, rows, qp, levels, runs, &one, NULL, (const byte (*)[2])scan)
                         : quant_dc4x2_around(c->mb, rows, qp, levels, runs, &one, NULL, (const byte (*)[2])scan);
  }
}

/* kind 0..5: hadamard4x4, ihadamard4x4, hadamard4x2, ihadamard4x2, hadamard2x2, ihadamard2x2; flat layouts of jmb_hadamard */
void jmref_hadamard(int kind, int *v)
{
  if (kind <= 1) {
    int *rows[4], out[16], *orows[4];
    for (int i = 0; i < 4; i++) { rows[i] = v + 4 * i; orows[i] = out + 4 * i; }
    if (kind == 0) hadamard4x4(rows, orows); else ihadamard4x4(rows, orows);
    memcpy(v, out, sizeof(out));
  } else if (kind == 2) {
    int *rows[2] = { v, v + 4 }, out[8], *orows[2] = { out, out + 4 };
    hadamard4x2(rows, orows);
    memcpy(v, out, sizeof(out));
  } else if (kind == 3) {
    int *rows[2] = { v, v + 4 }, out[8], *orows[4] = { out, out + 2, out + 4, out + 6 };
    ihadamard4x2(rows, orows);
    memcpy(v, out, sizeof(out));
  } else if (kind == 4) {
    int grid[5][5], *rows[5], out[4];
    memset(grid, 0, sizeof(grid));
    for (int i = 0; i < 5; i++) rows[i] = grid[i];
    grid[0][0] = v[0]; grid[0][4] = v[1]; grid[4][0] = v[2]; grid[4][4] = v[3];
    hadamard2x2(rows, out);
    memcpy(v, out, sizeof(out));
  } else {
    int out[4];
    ihadamard2x2(v, out);
    memcpy(v, out, sizeof(out));
  }
}

/* ------------------------------------------------------------------------------------------------
 * CPU baseline leg of bench.py: the same per-macroblock work the GPU step does, executed by JM's own
 * functions: 41 x (full_search_motion_estimation + sub_pel_motion_estimation) and, for each of the 7
 * partition modes, prediction (UMVLine4X copy, as OneComponentLumaPrediction mc_prediction.c:117) ->
 * residual -> forward4x4 -> quant_4x4_normal for the 16 luma blocks.
 *   preds  [n_mb][41][2] predictors (qpel), canonical partition order (include/jmb200.h)
 *   out_mv [n_mb][41][2], out_cost [n_mb][41], out_levels [n_mb][7][256] (scan order per 4x4 block)
 *   secs[0] = motion-estimation seconds, secs[1] = transform/quant seconds
 * ---------------------------------------------------------------------------------------------- */
#include <time.h>
static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

static const unsigned char k_pt[41][3] = { /* type, x, y */
  {1,0,0},{2,0,0},{2,0,8},{3,0,0},{3,8,0},{4,0,0},{4,8,0},{4,0,8},{4,8,8},
  {5,0,0},{5,8,0},{5,0,4},{5,8,4},{5,0,8},{5,8,8},{5,0,12},{5,8,12},
  {6,0,0},{6,4,0},{6,8,0},{6,12,0},{6,0,8},{6,4,8},{6,8,8},{6,12,8},
  {7,0,0},{7,4,0},{7,8,0},{7,12,0},{7,0,4},{7,4,4},{7,8,4},{7,12,4},{7,0,8},{7,4,8},{7,8,8},{7,12,8},{7,0,12},{7,4,12},{7,8,12},{7,12,12}};
static const int k_base[8] = {0, 0, 1, 3, 5, 9, 17, 25};

void jmref_run_mbs(void *h, int n_mb, const int16_t *mb_xy, const int16_t *preds, const int *lambda3, int qp,
                   const int *qparams, const uint8_t *scan, const uint8_t *c_cost, int do_tq,
                   int16_t *out_mv, int64_t *out_cost, int16_t *out_levels, double *secs)
{
  jmref_ctx *c = (jmref_ctx *)h; g_ctx = c;
  int lam[3] = { lambda3[0], lambda3[1], lambda3[2] };
  double t_me = 0, t_tq = 0;
  for (int m = 0; m < n_mb; m++) {
    int mbx = mb_xy[2 * m], mby = mb_xy[2 * m + 1];
    double t0 = now_s();
    for (int p = 0; p < 41; p++) {
      MEBlock b; MotionVector pred;
      fill_mv_block(c, &b, k_pt[p][0], mbx + k_pt[p][1], mby + k_pt[p][2], 0);
      pred.mv_x = preds[(m * 41 + p) * 2]; pred.mv_y = preds[(m * 41 + p) * 2 + 1];
      b.mv[0].mv_x = (short)(((pred.mv_x + 2) >> 2) * 4);          /* mv_search.c:931-932 */
      b.mv[0].mv_y = (short)(((pred.mv_y + 2) >> 2) * 4);
      distblk mc = full_search_motion_estimation(c->mb, &pred, &b, DISTBLK_MAX, lam[F_PEL]);
      if (!c->p_Vid->start_me_refinement_hp) mc = DISTBLK_MAX;      /* mv_search.c:971-974 */
      mc = sub_pel_motion_estimation(c->mb, &pred, &b, mc, lam);
      out_mv[(m * 41 + p) * 2] = b.mv[0].mv_x; out_mv[(m * 41 + p) * 2 + 1] = b.mv[0].mv_y;
      out_cost[m * 41 + p] = (int64_t)mc;
      free_mem2Dpel(b.orig_pic);
    }
    double t1 = now_s();
    t_me += t1 - t0;
    if (do_tq) {
      static const int w4[8] = {4, 4, 4, 2, 2, 2, 1, 1}, h4[8] = {4, 4, 2, 4, 2, 1, 2, 1};
      for (int mode = 1; mode <= 7; mode++)
        for (int blk = 0; blk < 16; blk++) {
          int bx4 = blk & 3, by4 = blk >> 2;
          int ux4 = bx4, uy4 = by4;
          if (mode < 5) { ux4 &= ~1; uy4 &= ~1; }                  /* macroblock.c:946-971 */
          int slot = k_base[mode] + (uy4 / h4[mode]) * (4 / w4[mode]) + ux4 / w4[mode];
          int mvx = out_mv[(m * 41 + slot) * 2], mvy = out_mv[(m * 41 + slot) * 2 + 1];
          imgpel *rl = UMVLine4X(c->ref, ((mby + uy4 * 4) << 2) + mvy, ((mbx + ux4 * 4) << 2) + mvx);
          rl += (by4 - uy4) * 4 * c->p_Vid->padded_size_x + (bx4 - ux4) * 4;
          int blk16[16], *rows[4], lev[17], run[17], cost = 0;
          LevelQuantParams qs[16], *qrows[4];
          for (int y = 0; y < 4; y++) {
            rows[y] = blk16 + 4 * y; qrows[y] = qs + 4 * y;
            for (int x = 0; x < 4; x++)
              blk16[4 * y + x] = (int)c->cur[mby + by4 * 4 + y][mbx + bx4 * 4 + x] - (int)rl[y * c->p_Vid->padded_size_x + x];
          }
          for (int i = 0; i < 16; i++) { qs[i].OffsetComp = qparams[3 * i]; qs[i].ScaleComp = qparams[3 * i + 1]; qs[i].InvScaleComp = qparams[3 * i + 2]; }
          forward4x4(rows, rows, 0, 0);
          QuantMethods q; memset(&q, 0, sizeof(q));
          q.qp = qp; q.ACLevel = lev; q.ACRun = run; q.q_params = qrows; q.coeff_cost = &cost;
          q.pos_scan = (const byte (*)[2])scan; q.c_cost = c_cost;
          quant_4x4_normal(c->mb, rows, &q);
          if (out_levels) {
            int16_t *o = out_levels + ((size_t)(m * 7 + mode - 1) * 16 + blk) * 16;
            memset(o, 0, 32);
            for (int i = 0, k = 0; lev[i]; i++) { k += run[i]; o[k++] = (int16_t)lev[i]; }
          }
        }
      t_tq += now_s() - t1;
    }
  }
  secs[0] = t_me; secs[1] = t_tq;
}

/* JM's own median / directional motion-vector predictor (GetMotionVectorPredictorNormal, static in mv_prediction.c, reached
 * through the function pointer init_motion_vector_prediction installs) on three caller-described neighbours A, B, C:
 * nb[k] = {available, ref_idx, mv_x, mv_y}. */
void jmref_mv_predictor(const int *nb, int ref_frame, int mb_x, int mb_y, int bsx, int bsy, int16_t *out)
{
  Macroblock mb;
  PicMotionParams row[3], *rows[1];
  PixelPos block[4];
  MotionVector pmv;
  int k;
  memset(&mb, 0, sizeof(mb)); memset(row, 0, sizeof(row)); memset(block, 0, sizeof(block));
  rows[0] = row;
  for (k = 0; k < 3; k++)
  {
    block[k].available = nb[4 * k];
    block[k].pos_x = (short)k; block[k].pos_y = 0;
    row[k].ref_idx[LIST_0] = (char)nb[4 * k + 1];
    row[k].mv[LIST_0].mv_x = (short)nb[4 * k + 2];
    row[k].mv[LIST_0].mv_y = (short)nb[4 * k + 3];
  }
  init_motion_vector_prediction(&mb, 0);
  mb.GetMVPredictor(&mb, block, &pmv, (short)ref_frame, rows, LIST_0, mb_x, mb_y, bsx, bsy);
  out[0] = pmv.mv_x; out[1] = pmv.mv_y;
}

/* JM's own DeblockFrame (lencod/src/loopFilter.c:63, non-MBAFF functions of loop_filter_normal.c) on a caller-described picture.
 * mb = n records of 176 bytes in the layout of jmb_db_mb (include/jmb200.h); planes are 8-bit samples, filtered in place. */
extern void DeblockFrame(VideoParameters *p_Vid, imgpel **imgY, imgpel ***imgUV);
extern void get_mb_block_pos_normal(BlockPos *PicPos, int mb_addr, short *x, short *y);
void jmref_deblock(uint8_t *luma, uint8_t *cb, uint8_t *cr, int w, int h, int yuv, int slice_type, int direct8x8inf, const uint8_t *mb)
{
  VideoParameters *p_Vid = (VideoParameters *)calloc(1, sizeof(VideoParameters));
  seq_parameter_set_rbsp_t *sps = (seq_parameter_set_rbsp_t *)calloc(1, sizeof(*sps));
  Slice *sl = (Slice *)calloc(1, sizeof(Slice));
  StorablePicture *pic = (StorablePicture *)calloc(1, sizeof(StorablePicture));
  const int mbw = w / 16, mbh = h / 16, n = mbw * mbh, wc = w / 2, hc = yuv == 1 ? h / 2 : h;
  imgpel **imgY = NULL, **uvp[2] = {NULL, NULL}, ***imgUV = NULL;
  int i, k, l, x, y;
  p_Vid->PicSizeInMbs = n; p_Vid->PicWidthInMbs = mbw;
  p_Vid->structure = FRAME; p_Vid->mb_aff_frame_flag = 0; p_Vid->P444_joined = 0;
  p_Vid->yuv_format = yuv; sps->chroma_format_idc = yuv; sps->direct_8x8_inference_flag = direct8x8inf;
  p_Vid->active_sps = sps;
  p_Vid->mb_size[IS_LUMA][0] = p_Vid->mb_size[IS_LUMA][1] = 16;
  p_Vid->mb_size[IS_CHROMA][0] = 8; p_Vid->mb_size[IS_CHROMA][1] = yuv == 1 ? 8 : 16;
  p_Vid->width_padded = w; p_Vid->width_cr = wc; p_Vid->pad_size_uv_x = 0;
  p_Vid->bitdepth_scale[IS_LUMA] = p_Vid->bitdepth_scale[IS_CHROMA] = 1;
  p_Vid->max_pel_value_comp[0] = p_Vid->max_pel_value_comp[1] = p_Vid->max_pel_value_comp[2] = 255;
  p_Vid->get_mb_block_pos = get_mb_block_pos_normal;
  p_Vid->PicPos = (BlockPos *)calloc(n + 1, sizeof(BlockPos));
  for (i = 0; i <= n; i++) { p_Vid->PicPos[i].x = (short)(i % mbw); p_Vid->PicPos[i].y = (short)(i / mbw); }
  p_Vid->mb_data = (Macroblock *)calloc(n, sizeof(Macroblock));
  p_Vid->enc_picture = pic;
  pic->mv_info = (PicMotionParams **)calloc(h / 4, sizeof(PicMotionParams *));
  pic->mv_info[0] = (PicMotionParams *)calloc((size_t)(h / 4) * (w / 4), sizeof(PicMotionParams));
  for (y = 1; y < h / 4; y++) pic->mv_info[y] = pic->mv_info[0] + (size_t)y * (w / 4);
  sl->slice_type = slice_type; sl->p_Vid = p_Vid;
  for (i = 0; i < n; i++)
  {
    const uint8_t *r = mb + (size_t)i * 176;
    Macroblock *m = &p_Vid->mb_data[i];
    const int16_t *mv = (const int16_t *)(r + 16);
    const int8_t *rid = (const int8_t *)(r + 144);
    m->p_Vid = p_Vid; m->p_Slice = sl; m->mbAddrX = i;
    m->mb_type = r[0];
    m->luma_transform_size_8x8_flag = (r[1] & 1) != 0; m->cbp = (r[1] & 2) ? 1 : 0;
    m->mbAvailA = (r[1] & 4) != 0; m->mbAvailB = (r[1] & 8) != 0;
    m->qp = (int8_t)r[2]; m->qpc[0] = (int8_t)r[3]; m->qpc[1] = (int8_t)r[4];
    m->DFDisableIdc = (int8_t)r[5]; m->DFAlphaC0Offset = (int8_t)r[6]; m->DFBetaOffset = (int8_t)r[7];
    m->cbp_blk = *(const uint32_t *)(r + 8);
    for (l = 0; l < 2; l++)
      for (k = 0; k < 16; k++)
      {
        PicMotionParams *p = &pic->mv_info[(i / mbw) * 4 + k / 4][(i % mbw) * 4 + k % 4];
        p->mv[l].mv_x = mv[(l * 16 + k) * 2]; p->mv[l].mv_y = mv[(l * 16 + k) * 2 + 1];
        p->ref_idx[l] = (char)(rid[l * 16 + k] < 0 ? -1 : 0);
        p->ref_pic[l] = rid[l * 16 + k] < 0 ? NULL : (StorablePicture *)(uintptr_t)(4096 + 64 * rid[l * 16 + k]);      /* only ever compared */
      }
  }
  get_mem2Dpel(&imgY, h, w);
  for (y = 0; y < h; y++) for (x = 0; x < w; x++) imgY[y][x] = luma[(size_t)y * w + x];
  if (yuv)
  {
    get_mem2Dpel(&uvp[0], hc, wc); get_mem2Dpel(&uvp[1], hc, wc);
    for (y = 0; y < hc; y++) for (x = 0; x < wc; x++) { uvp[0][y][x] = cb[(size_t)y * wc + x]; uvp[1][y][x] = cr[(size_t)y * wc + x]; }
    imgUV = uvp;
  }
  DeblockFrame(p_Vid, imgY, imgUV);
  for (y = 0; y < h; y++) for (x = 0; x < w; x++) luma[(size_t)y * w + x] = (uint8_t)imgY[y][x];
  if (yuv)
  {
    for (y = 0; y < hc; y++) for (x = 0; x < wc; x++) { cb[(size_t)y * wc + x] = (uint8_t)uvp[0][y][x]; cr[(size_t)y * wc + x] = (uint8_t)uvp[1][y][x]; }
    free_mem2Dpel(uvp[0]); free_mem2Dpel(uvp[1]);
  }
  free_mem2Dpel(imgY);
  free(pic->mv_info[0]); free(pic->mv_info); free(pic); free(p_Vid->mb_data); free(p_Vid->PicPos); free(sl); free(sps); free(p_Vid);
}
