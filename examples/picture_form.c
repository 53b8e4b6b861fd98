/* examples/picture_form.c -- the picture form of the libjmb200 C ABI from plain C (what INTEGRATION.md calls "going beyond one
 * call per block"): one reference picture, one current picture, the 41 motion searches of every macroblock and the 4x4
 * transform + quantisation of all seven partition modes, with host buffers.  It is the call sequence bench.py's end-to-end
 * leg times.  Build (after `make -C jm_b200/csrc`):
 *     gcc -O2 -Iinclude examples/picture_form.c -Ljm_b200/lib -ljmb200 -Wl,-rpath,$PWD/jm_b200/lib -lm -o picture_form
 * Needs a B200: without one jmb_create() fails and the program says so (there is no CPU path).
 * The quantiser tables below are the H.264 ones JM keeps in q_matrix.c / q_offsets.c (flat matrices, inter offsets). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "jmb200.h"

#define CHECK(call) do { int rc_ = (call); if (rc_) { fprintf(stderr, "%s -> %d: %s\n", #call, rc_, jmb_last_error(ctx)); return 1; } } while (0)

/* quant_coef / dequant_coef (lencod/src/q_matrix.c:20-36), positions classed as in the standard */
static const int QUANT[6][3] = {{13107, 5243, 8066}, {11916, 4660, 7490}, {10082, 4194, 6554}, {9362, 3647, 5825}, {8192, 3355, 5243}, {7282, 2893, 4559}};
static const int DEQUANT[6][3] = {{10, 16, 13}, {11, 18, 14}, {13, 20, 16}, {14, 23, 18}, {16, 25, 20}, {18, 29, 23}};
static const unsigned char ZIGZAG[16][2] = {{0,0},{1,0},{0,1},{0,2},{1,1},{2,0},{3,0},{2,1},{1,2},{0,3},{1,3},{2,2},{3,1},{3,2},{2,3},{3,3}};
static const unsigned char COEFF_COST[16] = {3, 2, 2, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};

static void quant_desc_4x4(jmb_quant_desc *q, int qp)
{
  int i, j, q_bits = 15 + qp / 6;
  memset(q, 0, sizeof(*q));
  q->n = 4; q->qp = qp; q->is_cavlc = 1;
  for (j = 0; j < 4; j++)
    for (i = 0; i < 4; i++)
    {
      int cls = ((i & 1) && (j & 1)) ? 1 : (!(i & 1) && !(j & 1)) ? 0 : 2;
      q->qparams[j * 4 + i][0] = 342 << (q_bits - 11);            /* inter offset 1/6 in Q11, q_offsets.c:238-249 */
      q->qparams[j * 4 + i][1] = QUANT[qp % 6][cls];
      q->qparams[j * 4 + i][2] = DEQUANT[qp % 6][cls] << 4;
    }
  memcpy(q->scan, ZIGZAG, sizeof(ZIGZAG));
  memcpy(q->c_cost, COEFF_COST, sizeof(COEFF_COST));
}

int main(void)
{
  enum { W = 352, H = 288, MBW = W / 16, NMB = (W / 16) * (H / 16), R = 16, QP = 28 };
  static const unsigned char type_of[41] = {1, 2,2, 3,3, 4,4,4,4, 5,5,5,5,5,5,5,5, 6,6,6,6,6,6,6,6, 7,7,7,7,7,7,7,7,7,7,7,7,7,7,7,7};
  static const unsigned char w4[8] = {0, 4, 4, 2, 2, 2, 1, 1}, h4[8] = {0, 4, 2, 4, 2, 1, 2, 1}, base[8] = {0, 0, 1, 3, 5, 9, 17, 25};
  jmb_ctx *ctx = NULL;
  jmb_me_config cfg = {R, 0, {JMB_SAD, JMB_SATD, JMB_SATD}, 0, 1, 9, 9};
  jmb_quant_desc qd;
  uint16_t *ref = malloc(sizeof(uint16_t) * W * H), *cur = malloc(sizeof(uint16_t) * W * H);
  jmb_me_req *req = calloc((size_t)NMB * 41, sizeof(*req));
  jmb_me_res *res = calloc((size_t)NMB * 41, sizeof(*res));
  int16_t *levels = malloc(sizeof(int16_t) * 7 * NMB * 256);
  int32_t *cost = malloc(sizeof(int32_t) * 7 * NMB * 4);
  uint32_t *cbp_blk = malloc(sizeof(uint32_t) * 7 * NMB);
  int slot = 0, x, y, mb, k, lambda = (int)(32.0 * sqrt(0.85 * pow(2.0, (QP - 12) / 3.0)) + 0.5), bits;
  long long coded = 0;

  if (jmb_create(0, &ctx)) { fprintf(stderr, "jmb_create: %s\n", jmb_last_error(NULL)); return 2; }
  for (bits = 1; (1 << bits) < 4 * (2 * R + 3) + 1; bits++) ;            /* max_mvd, mv_search.c:325-329 */
  cfg.max_mvd = (1 << ((3 + 2 * bits) >> 1)) - 1;
  CHECK(jmb_me_configure(ctx, &cfg));

  for (y = 0; y < H; y++)                                                 /* a texture, and the same texture moved by (3, -2) */
    for (x = 0; x < W; x++)
    {
      ref[y * W + x] = (uint16_t)(128 + 60 * sin(x * 0.21) * cos(y * 0.17) + 30 * sin((x + 2 * y) * 0.05));
      cur[y * W + x] = (uint16_t)(128 + 60 * sin((x - 3) * 0.21) * cos((y + 2) * 0.17) + 30 * sin(((x - 3) + 2 * (y + 2)) * 0.05));
    }
  CHECK(jmb_ref_put(ctx, slot, ref, W, H, W, 8, JMB_HOST));               /* quarter-pel planes: getSubImagesLuma */
  CHECK(jmb_pic_begin(ctx, cur, W, H, W, JMB_HOST, &slot, 1));

  for (mb = 0; mb < NMB; mb++)                                            /* 41 requests per macroblock, canonical order */
    for (k = 0; k < 41; k++)
    {
      jmb_me_req *q = &req[mb * 41 + k];
      int t = type_of[k], i = k - base[t], per_row = 4 / w4[t];
      q->blocktype = (uint8_t)t;
      q->pos_x = (int16_t)((mb % MBW) * 16 + (i % per_row) * w4[t] * 4);
      q->pos_y = (int16_t)((mb / MBW) * 16 + (i / per_row) * h4[t] * 4);
      q->pred_x = q->pred_y = 0;                                          /* the caller's predictor (GetMVPredictor) */
      q->center_x = q->center_y = 0;                                      /* ((pred + 2) >> 2) * 4 */
      q->mode = JMB_SEARCH_FULL; q->flags = JMB_REQ_SUBPEL;
      q->lambda[0] = q->lambda[1] = q->lambda[2] = lambda;
      q->min_mcost = (int64_t)0x7fffffff << 5;                            /* DISTBLK_MAX */
    }
  CHECK(jmb_me_search_frame(ctx, req, NMB, res, JMB_HOST));
  quant_desc_4x4(&qd, QP);
  CHECK(jmb_mc_tq_modes(ctx, NULL, NMB, 0x7F, &qd, levels, cost, cbp_blk, JMB_HOST));   /* from the resident results */

  for (mb = 0; mb < NMB; mb++) coded += cbp_blk[mb] != 0;
  printf("macroblock 0: 16x16 mv (%d,%d) quarter-pels, cost %lld; %lld of %d macroblocks have coded 16x16-mode blocks; %llu kernels launched\n",
         res[0].mv_x, res[0].mv_y, (long long)res[0].cost, coded, (int)NMB, (unsigned long long)jmb_launch_count(ctx));
  jmb_destroy(ctx);
  free(ref); free(cur); free(req); free(res); free(levels); free(cost); free(cbp_blk);
  return 0;
}
