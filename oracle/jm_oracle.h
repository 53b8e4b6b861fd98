/*
 * oracle/jm_oracle.h -- TEST INFRASTRUCTURE ONLY (checker, never shipped, never on the product path).
 *
 * CPU restatement, in plain C, of the JM 19.0 lencod algorithms on the motion-estimation +
 * transform/quantisation hot path.  Every function cites the reference file:line it follows
 * (paths relative to /root/reference).  Parity is PINNED: tests/test_oracle_vs_ref.py checks every
 * function here against the real JM leaf functions (oracle/_ref/libjmref.so, built from the
 * reference's own sources by oracle/Makefile) and against tests/golden/ fixtures that were generated
 * from those same JM functions by tests/golden/make_golden.py.
 */
#ifndef JM_ORACLE_H
#define JM_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JMO_PAD_X 32   /* IMG_PAD_SIZE_X  lencod/inc/defines.h:121 */
#define JMO_PAD_Y 20   /* IMG_PAD_SIZE_Y  lencod/inc/defines.h:122 */

enum { JMO_SAD = 0, JMO_SSE = 1, JMO_SATD = 2 };   /* ERROR_SAD/SSE/SATD */

/* A reference picture: 16 quarter-pel planes [fy][fx], each (h+40) x (w+64) uint16 samples. */
typedef struct jmo_ref {
  int w, h, W, H;          /* picture and padded sizes */
  uint16_t *plane[4][4];   /* plane[fy][fx][ (y+PAD_Y)*W + (x+PAD_X) ] */
} jmo_ref;

jmo_ref *jmo_ref_create(const uint16_t *luma, int w, int h, int stride, int max_value);
void     jmo_ref_destroy(jmo_ref *r);
void     jmo_ref_get_plane(const jmo_ref *r, int fy, int fx, uint16_t *out);

void jmo_spiral(int search_range, int16_t *xy /* 2*(2R+1)^2 */);
int  jmo_mvbits(int v);

/* raw (unscaled, never early-terminated) distortion of a bsx x bsy source block against the
 * reference at absolute quarter-pel position (cand_x, cand_y) */
int jmo_dist(const jmo_ref *r, const uint16_t *src, int bsx, int bsy, int cand_x, int cand_y,
             int metric, int test8x8);

/* the weighted / bi-predictive members of the distortion table (me_distortion.c:434-1520):
 * form 0 plain, 1 weighted, 2 average of two references, 3 weighted average; wp = {weight1, weight2, offset,
 * luma_log_weight_denom, wp_luma_round}; max_value = max_imgpel_value */
int jmo_dist_ex(const jmo_ref *r1, const jmo_ref *r2, const uint16_t *src, int bsx, int bsy, int c1x, int c1y, int c2x, int c2y,
                int metric, int test8x8, int form, const int *wp, int max_value);

int64_t jmo_full_search(const jmo_ref *r, const uint16_t *cur, int cur_stride, int blocktype,
                        int pos_x, int pos_y, int pred_x, int pred_y, int center_x, int center_y,
                        int lambda, int64_t min_mcost, int search_range, int16_t *mv_out);

int64_t jmo_sub_pel(const jmo_ref *r, const uint16_t *cur, int cur_stride, int blocktype,
                    int pos_x, int pos_y, int pred_x, int pred_y, int mv_x, int mv_y,
                    const int *lambda3, int64_t min_mcost, int metric_h, int metric_q,
                    int start_hp, int start_qp, int test8x8, int16_t *mv_out);

/* fast full search: 16 4x4 SAD surfaces with the macroblock-origin clamp, 41 partition surfaces */
void jmo_ffs_center(int pmv_x, int pmv_y, int search_range, const int *max_hmv_q, const int *max_vmv_q,
                    int16_t *center);
void jmo_ffs_setup(const jmo_ref *r, const uint16_t *cur, int cur_stride, int mb_x, int mb_y,
                   int center_x, int center_y, int search_range,
                   uint32_t *block_sad /* [8][16][max_pos], types 1..7 filled */);
int64_t jmo_ffs_search(const uint32_t *block_sad, int search_range, int blocktype, int block_index,
                       int center_x, int center_y, int pred_x, int pred_y, int lambda,
                       int64_t min_mcost, int max_mvd, int16_t *mv_out);

void jmo_forward4x4(int *blk /* 16, in place */);
void jmo_forward8x8(int *blk /* 64, in place */);
int  jmo_hadamard_sad4x4(const int16_t *diff);
int  jmo_hadamard_sad8x8(const int16_t *diff);

/* variants as in oracle/ref_harness.c::jmref_quant */
int jmo_quant_list(int m, int q_bits, int qp_per, int dequant, int clip, int use_cost, int around, int arw,
                   const int *params, const uint8_t *c_cost, int *coef, int *levels, int *runs, int *fadjust, int *coeff_cost);
void jmo_hadamard(int kind, int *vals);
void jmo_inverse4x4(int *blk /* 16, in place */);
void jmo_inverse8x8(int *blk /* 64, in place */);
long long jmo_luma_residual_coding(const uint16_t *src, const uint16_t *pred, int n, int qp, const int *qparams,
                                   const uint8_t *scan, const uint8_t *c_cost, int is_cavlc, int max_value,
                                   short *levels, int *cost8, int *cbp, int *cbp_blk, uint16_t *recon);
int jmo_quant(int variant, int *coef, int qp, const int *qparams, const uint8_t *scan,
              const uint8_t *c_cost, int is_cavlc, int adapt_rnd_weight,
              int *levels, int *runs, int *fadjust, int *coeff_cost);

/* EPZS (me_epzs_int.c:42-426, me_epzs_sub.c:30-213) for a caller-supplied predictor list; the structs have the layout of
 * jmb_epzs_req / jmb_epzs_res of the product ABI so the tests feed both with the same bytes */
#define JMO_EPZS_REF_GT0_FRAME 1
#define JMO_EPZS_ADAPT_PATTERN 2
#define JMO_EPZS_SQUARE_HINT   4
#define JMO_EPZS_DUAL          8
#define JMO_EPZS_SUBPEL       16
#define JMO_EPZS_TEST8X8      32
#define JMO_EPZS_SKIP_INT     64
#define JMO_EPZS_WINDOW_GEN  128   /* segment 2 = JM's window_predictor set around the start mv (EPZSWindowPredictorInit mode 0) */
typedef struct jmo_epzs_req {
  int16_t pos_x, pos_y, pred_x, pred_y, start_x, start_y;
  uint8_t blocktype, ref, flags, pattern;
  uint8_t pattern_dual, jm_ref, reserved_[2];
  uint8_t n_cand[4], gate[4];
  int32_t cand_off;
  int32_t lambda[3];
  int16_t range_x, range_y;
  int64_t stop, medthres, prev_sad, subthres, min_mcost;
} jmo_epzs_req;
typedef struct jmo_epzs_res {
  int16_t mv_x, mv_y, imv_x, imv_y;
  int64_t cost, icost, prev_sad;
  int32_t exit_code, n_evals;
} jmo_epzs_res;
void jmo_epzs(const jmo_ref *r, const uint16_t *cur, int cur_stride, const jmo_epzs_req *q, const int16_t *cands,
              const int *me /* metric_h, metric_q, start_hp, start_qp, search_pos2 */, jmo_epzs_res *o);

void jmo_chroma_pred(const uint8_t *ref_c, int wc, int hc, int stride, int yuv, int mb_cx, int mb_cy, const int16_t *mv16, uint8_t *pred);
int jmo_chroma_rc(const uint8_t *src, const uint8_t *pred, int yuv, int qp_ac, int qp_dc, const int *params_ac, const int *params_dc,
                  const uint8_t *c_cost, int is_cavlc, int16_t *dc_levels, int16_t *ac_levels, unsigned *cbp_bits, uint8_t *recon);
void jmo_epzs_batch(const jmo_ref *r, const uint16_t *cur, int cur_stride, const jmo_epzs_req *reqs, int n, const int16_t *cands,
                    const int *me, jmo_epzs_res *res);
void jmo_mc_tq_modes_mb(const jmo_ref *r, const uint16_t *cur, int cur_stride, int mb_x, int mb_y, const int16_t *mv41, int n, int qp,
                        const int *qparams, const uint8_t *scan, const uint8_t *c_cost, int is_cavlc, unsigned mode_mask, int16_t *levels);

/* deblocking of a whole picture in place (DeblockFrame, loopFilter.c:63); jmo_db_mb has the layout of jmb_db_mb (include/jmb200.h) */
typedef struct jmo_db_mb {
  uint8_t mb_type, flags; int8_t qp, qpc[2], df_disable_idc, df_alpha_c0_offset, df_beta_offset;
  uint32_t cbp_blk, pad_;
  int16_t mv[2][16][2];
  int8_t ref_id[2][16];
} jmo_db_mb;
void jmo_deblock(uint8_t *luma, int pitch, uint8_t *cb, uint8_t *cr, int pitch_c, int w, int h, int yuv, int slice_type, int direct8x8inf,
                 const jmo_db_mb *mbs);

#ifdef __cplusplus
}
#endif
#endif
