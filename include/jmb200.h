/*
 * jmb200.h -- C ABI of libjmb200.so: B200 (sm_100a) kernels for the JM 19.0 lencod motion-estimation
 * and integer transform/quantisation hot path.
 *
 * Plain C, plain pointers and sizes; no CUDA or torch types.  Every entry point names the JM
 * interface it stands behind (paths relative to the JM 19.0 tree, shihuade/JM).  The reference-side
 * binding (GNU ld --wrap stubs compiled against JM's own headers) is jm_b200/shim/jm_wrap.c and is
 * described in INTEGRATION.md.
 *
 * Conventions (all JM's):
 *   - samples are uint16_t (imgpel, lcommon/inc/typedefs.h:36) holding bit-depth-8 values;
 *   - motion vectors and predictors are in quarter-pel units (MotionVector, 2 x short);
 *   - costs are J = (D << 5) + lambda * bits  (LAMBDA_ACCURACY_BITS 5, lencod/inc/defines.h:130);
 *   - block types 1..7 = 16x16,16x8,8x16,8x8,8x4,4x8,4x4 (lencod/inc/macroblock.h:58-68);
 *   - reference planes are padded by 32 x 20 samples (IMG_PAD_SIZE_X/Y, lencod/inc/defines.h:121-122)
 *     and addressed with JM's UMVLine4X origin clamp (lencod/inc/refbuf.h:22-26).
 *
 * Memory location of every data pointer is given by a `loc` argument: JMB_HOST (pageable or pinned
 * host memory; the call copies in/out and returns when the results are in the caller's buffers),
 * JMB_HOST_ASYNC (pinned host memory, nothing waited for) or JMB_DEVICE (device memory of the
 * context's GPU; the call only enqueues work on the context's stream -- use jmb_sync()).  There is NO CPU implementation behind this ABI: without a CUDA device
 * jmb_create() fails with JMB_ERR_NO_DEVICE and nothing else can be called.
 *
 * Return value: 0 on success, a negative JMB_ERR_* otherwise; jmb_last_error() gives the text.
 * JM's own convention for fatal conditions is error(text, code) -> exit (lencod/src/filehandle.c:37);
 * the shim forwards every non-zero return to it.
 */
#ifndef JMB200_H
#define JMB200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JMB_ABI_VERSION 2

enum { JMB_HOST = 0, JMB_DEVICE = 1,
       JMB_HOST_ASYNC = 2 };   /* PINNED host memory; copies are only enqueued on the context's stream: the buffers must stay
                                  valid, and hold no results, until jmb_sync() -- accepted by the picture-form entry points
                                  (jmb_ref_put*, jmb_pic_begin*, jmb_me_search_frame*, jmb_mc_tq_modes) */
enum { JMB_SAD = 0, JMB_SSE = 1, JMB_SATD = 2 };            /* ERROR_SAD/SSE/SATD, lencod/inc/defines.h */
enum { JMB_SEARCH_FULL = 0,                                 /* full_search_motion_estimation  */
       JMB_SEARCH_FAST_FULL = 1 };                          /* fast_full_search_motion_estimation */
enum {
  JMB_OK = 0,
  JMB_ERR_NO_DEVICE = -1,
  JMB_ERR_CUDA = -2,
  JMB_ERR_ARG = -3,
  JMB_ERR_UNSUPPORTED = -4,
  JMB_ERR_STATE = -5
};

#define JMB_MAX_REFS 16
#define JMB_REQ_SUBPEL   1   /* run the half-/quarter-pel refinement after the integer search */
#define JMB_REQ_TEST8X8  2   /* SATD on 8x8 sub-blocks (MEBlock.test8x8, set when Transform8x8Mode) */
#define JMB_REQ_SKIP_INT 4   /* no integer search: refine (center_x, center_y) only (SubPelME call) */

typedef struct jmb_ctx jmb_ctx;

/* Search-engine configuration = the fields of InputParameters / VideoParameters the searches read
 * (init_motion_search_module, lencod/src/mv_search.c:315-515). */
typedef struct jmb_me_config {
  int32_t search_range;      /* SearchRange, integer pels (positions per search = (2R+1)^2) */
  int32_t max_mvd;           /* p_Vid->max_mvd (mv_search.c:325-329); FAST_FULL guard uses max_mvd-1 */
  int32_t metric[3];         /* MEDistortionFPel/HPel/QPel */
  int32_t start_hp;          /* p_Vid->start_me_refinement_hp (mv_search.c:445) */
  int32_t start_qp;          /* p_Vid->start_me_refinement_qp (mv_search.c:446) */
  int32_t search_pos2;       /* MEBlock.search_pos2 (9) */
  int32_t search_pos4;       /* MEBlock.search_pos4 (9) */
} jmb_me_config;

/* One motion search = one BlockMotionSearch leaf call (lencod/src/mv_search.c:858-1024):
 * currMB->IntPelME (:960) followed, when JMB_REQ_SUBPEL, by currMB->SubPelME (:975). 40 bytes. */
typedef struct jmb_me_req {
  int16_t pos_x, pos_y;        /* MEBlock.pos_x/pos_y: block origin in the picture, pels */
  int16_t pred_x, pred_y;      /* pred_mv, quarter-pel */
  int16_t center_x, center_y;  /* search centre as an mv, quarter-pel, multiple of 4
                                  (FULL: MEBlock.mv[list] on entry, mv_search.c:931-957;
                                   FAST_FULL: p_ffast_me->search_center, me_fullfast.c:309-327) */
  uint8_t blocktype;           /* 1..7 */
  uint8_t ref;                 /* index into the reference list given to jmb_pic_begin */
  uint8_t mode;                /* JMB_SEARCH_* */
  uint8_t flags;               /* JMB_REQ_* */
  int32_t lambda[3];           /* lambda_factor[F_PEL,H_PEL,Q_PEL], each 0..65535 (JM's largest is < 6000) */
  int32_t reserved_;           /* keeps min_mcost 8-byte aligned; set to 0 */
  int64_t min_mcost;           /* incoming minimum cost (DISTBLK_MAX from BlockMotionSearch) */
} jmb_me_req;

/* 24 bytes */
typedef struct jmb_me_res {
  int16_t mv_x, mv_y;          /* final mv (after sub-pel when requested), quarter-pel */
  int16_t imv_x, imv_y;        /* mv after the integer search */
  int64_t cost;                /* value SubPelME (or IntPelME) returns */
  int64_t icost;               /* value IntPelME returns */
} jmb_me_res;

/* Per-macroblock inter prediction description for jmb_mc_tq: mirrors what luma_residual_coding
 * (lencod/src/macroblock.c:1182) reads from currMB->b8x8[] and currSlice->all_mv. 72 bytes. */
typedef struct jmb_mb_pred {
  int16_t mv[16][2];           /* mv of each 4x4 block, raster order inside the MB, quarter-pel */
  uint8_t b8mode[4];           /* partition mode of each 8x8 quadrant, 1..7 (1..3 = whole-MB modes) */
  uint8_t ref[4];              /* reference index of each 8x8 quadrant */
} jmb_mb_pred;

/* Quantiser description = struct quant_methods + LevelQuantParams (lcommon/inc/quant_params.h:17-51). */
typedef struct jmb_quant_desc {
  int32_t n;                   /* 4 or 8 (transform size) */
  int32_t qp;                  /* qp_scaled; qp_per = qp/6 (p_Quant->qp_per_matrix) */
  int32_t is_cavlc;            /* currSlice->symbol_mode == CAVLC: clip levels to 2063; for n=8 use the
                                  4-way interleaved CAVLC level/run lists (quant_8x8cavlc_*) */
  int32_t around;              /* 0: *_normal, 1: *_around (adaptive rounding: also writes fadjust) */
  int32_t adapt_rnd_weight;    /* p_Vid->AdaptRndWeight */
  int32_t qparams[64][3];      /* [j*n+i] = {OffsetComp, ScaleComp, InvScaleComp} */
  uint8_t scan[64][2];         /* pos_scan: {i, j} per scan position */
  uint8_t c_cost[64];          /* COEFF_COST4x4 / COEFF_COST8x8 row in use */
} jmb_quant_desc;

/* List quantiser description: the DC / AC members of JM's quantiser family -- quant_ac4x4_normal/_around
 * (lencod/src/quant4x4_normal.c:117, quant4x4_around.c:132), quant_dc4x4_normal (:200), quant_dc2x2_normal/_around and
 * quant_dc4x2_normal/_around (lencod/src/quantChroma_normal.c, quantChroma_around.c) -- are one loop over m coefficients
 * taken in scan order.  The caller gathers the coefficients in that order and gives per-position parameters. */
enum { JMB_DQ_LEVEL = 0,         /* coefficient := level                               (quant_dc4x4_*)            */
       JMB_DQ_SHIFT = 1,         /* coefficient := (level * InvScale) << qp_per        (quant_dc2x2_*, _dc4x2_*)  */
       JMB_DQ_SHIFT_RND4 = 2 };  /* coefficient := ((level * InvScale) << qp_per + 8) >> 4   (quant_ac4x4_*)      */
typedef struct jmb_qlist_desc {
  int32_t m;                     /* coefficients per list, 1..16 */
  int32_t q_bits;                /* forward shift: Q_BITS + qp_per (+1 for the DC forms) */
  int32_t qp_per;
  int32_t dequant;               /* JMB_DQ_* */
  int32_t clip;                  /* CAVLC: clip levels to 2063 */
  int32_t use_cost;              /* accumulate coeff_cost (c_cost[run] / MAX_VALUE), quant_ac4x4_* only */
  int32_t around;                /* also write fadjust (quant_ac4x4_around) */
  int32_t adapt_rnd_weight;
  int32_t params[16][3];         /* per list position {Offset (already doubled for the DC forms), Scale, InvScale} */
  uint8_t c_cost[16];
} jmb_qlist_desc;

enum { JMB_HAD_4X4 = 0, JMB_IHAD_4X4 = 1, JMB_HAD_4X2 = 2, JMB_IHAD_4X2 = 3, JMB_HAD_2X2 = 4, JMB_IHAD_2X2 = 5 };

/* ---- context --------------------------------------------------------------------------------- */
int         jmb_abi_version(void);
int         jmb_create(int device, jmb_ctx **out);
void        jmb_destroy(jmb_ctx *ctx);
const char *jmb_last_error(const jmb_ctx *ctx);     /* ctx may be NULL for jmb_create failures */
int         jmb_sync(jmb_ctx *ctx);
void       *jmb_stream(jmb_ctx *ctx);               /* the cudaStream_t all work is enqueued on */
uint64_t    jmb_launch_count(const jmb_ctx *ctx);   /* kernels launched by this context so far */
int         jmb_host_alloc(jmb_ctx *ctx, size_t bytes, void **out);   /* pinned host memory */
int         jmb_host_free(jmb_ctx *ctx, void *p);

/* ---- reference pictures: stands behind getSubImagesLuma (lencod/src/img_luma.c:611), called from
 * UnifiedOneForthPix (lencod/src/image.c:2187) when a picture enters the DPB --------------------- */
int jmb_ref_put(jmb_ctx *ctx, int slot, const uint16_t *luma, int width, int height, int stride,
                int bitdepth, int loc);
int jmb_ref_drop(jmb_ctx *ctx, int slot);           /* free_storable_picture, lencod/src/mbuffer.c:729 */
/* copy one quarter-pel plane [fy][fx] back as JM lays it out: (height+40) x (width+64) uint16_t */
int jmb_ref_get_plane(jmb_ctx *ctx, int slot, int fy, int fx, uint16_t *out, int loc);

/* The same picture as 8-bit samples (one byte each): JM's imgpel is uint16_t but every BASELINE configuration codes 8-bit
 * video, so a caller that holds bytes (the YUV file JM reads, lcommon/src/input.c) sends half the bytes over the host link. */
int jmb_ref_put_u8(jmb_ctx *ctx, int slot, const uint8_t *luma, int width, int height, int stride, int loc);

/* Reconstructed-reference exchange between GPUs of one box WITHOUT a host round trip or a collective library: the owner
 * exports an IPC handle of a device buffer holding its reconstructed luma (uint16_t or uint8_t samples); a peer process
 * opens it once (jmb_peer_open, cached per handle) and then builds ITS OWN quarter-pel planes reading the owner's samples
 * straight over NVLink (the sub-pel kernel's loads are the transfer: jmb_ref_put* with loc = JMB_DEVICE on the mapped
 * pointer).  Stands where config 5 of BASELINE.json asks for the reconstructed-reference broadcast.
 *   handle: 64 bytes (cudaIpcMemHandle_t), to be carried to the peers by any means (a file, a socket, torch.distributed). */
#define JMB_IPC_HANDLE_BYTES 64
int jmb_dev_alloc(jmb_ctx *ctx, size_t bytes, void **out);          /* device memory of the context's GPU (exportable) */
int jmb_dev_free(jmb_ctx *ctx, void *p);
int jmb_dev_copy(jmb_ctx *ctx, void *dst, const void *src, size_t bytes, int dst_loc, int src_loc);   /* on the context's stream */
int jmb_peer_export(jmb_ctx *ctx, const void *dev_ptr, unsigned char handle[JMB_IPC_HANDLE_BYTES]);
int jmb_peer_open(jmb_ctx *ctx, const unsigned char handle[JMB_IPC_HANDLE_BYTES], void **mapped);
int jmb_peer_close(jmb_ctx *ctx, void *mapped);

/* ---- current picture: p_Vid->pCurImg, read by get_original_block (lencod/src/mv_search.c:786) and
 * setup_fast_full_search (lencod/src/me_fullfast.c:333-337) ------------------------------------- */
int jmb_pic_begin(jmb_ctx *ctx, const uint16_t *cur, int width, int height, int stride, int loc,
                  const int *ref_slots, int nref);

int jmb_pic_begin_u8(jmb_ctx *ctx, const uint8_t *cur, int width, int height, int stride, int loc,
                     const int *ref_slots, int nref);

/* ---- motion search --------------------------------------------------------------------------- */
int jmb_me_configure(jmb_ctx *ctx, const jmb_me_config *cfg);
/* n searches; requests of one macroblock and reference that sit next to each other in `reqs`
 * share their 4x4 SAD evaluations (what setup_fast_full_search + update_full_search_large_blocks
 * do for one macroblock, lencod/src/me_fullfast.c:196-608). */
int jmb_me_search(jmb_ctx *ctx, const jmb_me_req *reqs, int n, jmb_me_res *res, int loc);
/* Whole-picture form: exactly 41 requests per macroblock, macroblocks in any order, the 41 in
 * canonical partition order (type 1..7, partitions of a type in raster order inside the
 * macroblock: slot = base[type] + (by4/h4)*(4/w4) + bx4/w4 with base = {0,1,3,5,9,17,25}).
 * No grouping pass is needed, so with JMB_DEVICE nothing touches the host. */
int jmb_me_search_frame(jmb_ctx *ctx, const jmb_me_req *reqs, int n_mb, jmb_me_res *res, int loc);

/* Whole-picture form with the requests GENERATED ON THE DEVICE: the caller sends only what BlockMotionSearch cannot know
 * without it -- the 41 motion-vector predictors of every macroblock (GetMVPredictor, lencod/src/mv_prediction.c:192), 4 bytes
 * each -- and the slice-level constants; block geometry, search centre, flags and bounds follow from JM's own rules:
 *   centre, FULL:      ((pred + 2) >> 2) * 4 per partition (mv_search.c:931-932), clipped to the mv range (:957, clip_mv_range
 *                      lencod/src/conformance.c:640);
 *   centre, FAST_FULL: the rounded 16x16 predictor for all 41 partitions, clipped to [min + 4R, max - 4R] (me_fullfast.c:309-327);
 *   min_mcost = DISTBLK_MAX (mv_search.c:871);  JMB_REQ_TEST8X8 only for block types <= 4 (mv_search.c:1630);
 *   final mv clipped to the mv range again (mv_search.c:981).
 * Macroblocks 0 .. n_mb-1 of the current picture in raster order.  Results: 8 bytes per search. */
typedef struct jmb_mb_mvpred { int16_t pred[41][2]; } jmb_mb_mvpred;            /* canonical partition order, quarter-pel; 164 bytes */
typedef struct jmb_frame_params {
  int32_t lambda[3];               /* lambda_factor[F_PEL, H_PEL, Q_PEL] of the slice (mode_decision.c:192) */
  int32_t mode;                    /* JMB_SEARCH_FULL / JMB_SEARCH_FAST_FULL */
  int32_t flags;                   /* JMB_REQ_SUBPEL, JMB_REQ_TEST8X8 */
  int32_t ref;                     /* index in the picture's reference list */
  int32_t mv_min_x, mv_max_x;      /* p_Vid->MaxHmvR[4], [5]: quarter-pel mv range of the level (conformance.c:604) */
  int32_t mv_min_y, mv_max_y;      /* p_Vid->MaxVmvR[4], [5] */
} jmb_frame_params;
typedef struct jmb_me_res8 {
  int16_t mv_x, mv_y;              /* final mv, quarter-pel */
  int32_t cost;                    /* the cost SubPelME (or IntPelME) returned; INT32_MAX when it does not fit (nothing beat DISTBLK_MAX) */
} jmb_me_res8;
/* res may be NULL (results stay on the device for jmb_mc_tq_modes* / jmb_luma_residual_coding_modes with res = NULL) */
int jmb_me_search_frame_pred(jmb_ctx *ctx, const jmb_mb_mvpred *pred, int n_mb, const jmb_frame_params *fp, jmb_me_res8 *res, int loc);

/* ---- EPZS (SearchMode 3) --------------------------------------------------------------------------------------------
 * One request = one EPZS_integer_motion_estimation call (lencod/src/me_epzs_int.c:42-426) followed, with JMB_EPZS_SUBPEL, by
 * what BlockMotionSearch does next (mv_search.c:964-976): EPZS_sub_pel_motion_estimation (lencod/src/me_epzs_sub.c:30-213).
 * The whole state machine runs on the device, one warp per request: centre check, early terminations against the previous
 * distortions, the ordered predictor list, the refinement pattern walk (pattern_data, me_epzs_common.c:48-76) incl. the
 * second-best ("dual") pass, and the sub-pel stage.  The caller supplies what JM's HOST state holds:
 *   - the predictor list in up to four segments, in JM's generator order, each behind the cost gate its generator sits
 *     behind in JM: segment s is checked when gate[s] == 0 or centre cost > gate[s] * stop
 *     (spatial / spatial-memory / hierarchical / co-located: always; the co-located block's neighbours
 *     me_epzs_common.c:1556: gate 1; window predictors me_epzs_int.c:193-203: gate 0 or 3; block-type predictors :211: 0 or 2);
 *   - stop = EPZSDetermineStopCriterion(...) (me_epzs_common.c:1874), medthres / subthres of the block type (:454-457),
 *     prev_sad = *prevSad (p_EPZS->distortion[list][blocktype-1][pos_x>>2]);
 *   - range_x/y = mv_block->searchRange.max_x/max_y in quarter-pel (candidates further from the start mv are skipped).
 * Candidate lists: cands[2 * (cand_off + i)] = mv_x, [+1] = mv_y (motion vectors, quarter-pel).
 * Sub-pel stage metrics / start positions come from jmb_me_configure (metric[1], metric[2], start_hp, start_qp,
 * search_pos2); start_qp must be 1 (JM indexes next_start_pos[][-1] otherwise, me_epzs_sub.c:141,182). */
enum { JMB_EPZS_PAT_SDIAMOND = 0, JMB_EPZS_PAT_SQUARE = 1, JMB_EPZS_PAT_EDIAMOND = 2, JMB_EPZS_PAT_LDIAMOND = 3,
       JMB_EPZS_PAT_SBDIAMOND = 4, JMB_EPZS_PAT_PMVFAST = 5 };
#define JMB_EPZS_REF_GT0_FRAME 1   /* ref > 0 in a frame picture: the prevSad early exits apply (:103, :254, :362) */
#define JMB_EPZS_ADAPT_PATTERN 2   /* EPZSPattern != 0 (:286) */
#define JMB_EPZS_SQUARE_HINT   4   /* ref > 0 && blocktype != 1 (:296) */
#define JMB_EPZS_DUAL          8   /* EPZSDualRefinement > 0 (:384) */
#define JMB_EPZS_SUBPEL       16   /* continue with the sub-pel stage */
#define JMB_EPZS_TEST8X8      32   /* MEBlock.test8x8: SATD on 8x8 sub-blocks */
#define JMB_EPZS_SKIP_INT     64   /* sub-pel stage only: start_x/y = the integer-stage mv, min_mcost = the cost handed to SubPelME */
#define JMB_EPZS_WINDOW_GEN  128   /* segment 2 is JM's window_predictor set (EPZSWindowPredictorInit mode 0, me_epzs_common.c:352-371) and
                                      is generated on the device, not read from the list: n_cand[2] = 8 * rings - 1 points around the start
                                      mv at distances range_x >> (rings - 1), ..., range_x >> 0, eight per ring in JM's order */
typedef struct jmb_epzs_req {      /* 88 bytes */
  int16_t pos_x, pos_y;            /* block origin, pels */
  int16_t pred_x, pred_y;          /* pred_mv, quarter-pel */
  int16_t start_x, start_y;        /* *mv on entry, quarter-pel */
  uint8_t blocktype, ref, flags, pattern;   /* pattern = p_EPZS->searchPattern (JMB_EPZS_PAT_*) */
  uint8_t pattern_dual, jm_ref, reserved_[2];   /* p_EPZS->searchPatternD; jm_ref = mv_block->ref_idx: JM's rules that tell reference 0
                                                   from the others read this, `ref` only selects the picture in the device list */
  uint8_t n_cand[4], gate[4];
  int32_t cand_off;
  int32_t lambda[3];               /* lambda_factor[F_PEL, H_PEL, Q_PEL] */
  int16_t range_x, range_y;
  int64_t stop, medthres, prev_sad, subthres, min_mcost;
} jmb_epzs_req;
typedef struct jmb_epzs_res {      /* 40 bytes */
  int16_t mv_x, mv_y;              /* final mv */
  int16_t imv_x, imv_y;            /* mv the integer stage leaves in *mv (= what goes to p_motion) */
  int64_t cost, icost;             /* returned by SubPelME / by EPZS_integer_motion_estimation */
  int64_t prev_sad;                /* *prevSad as the integer stage leaves it */
  int32_t exit_code;               /* which return of the integer stage: 1 :103, 2 :135, 3 :254, 4 :362, 5 the end */
  int32_t n_evals;                 /* distortions evaluated (diagnostic; parallel evaluation may count a few more than JM) */
} jmb_epzs_res;
int jmb_epzs_search(jmb_ctx *ctx, const jmb_epzs_req *reqs, int n, const int16_t *cands, int n_cands, jmb_epzs_res *res, int loc);

/* Whole-picture form with the requests generated on the device: per macroblock the 41 predictors (= start mvs, EPZSSubPelGrid)
 * and n_shared candidate mvs checked by every partition of the macroblock (what the spatial / co-located generators yield for
 * a macroblock whose neighbours' motion the caller knows), plus the window predictors around the start mv when `window` is set
 * (gate 3); thresholds per block type.  Results as jmb_me_search_frame_pred. */
typedef struct jmb_epzs_frame_params {
  int32_t lambda[3];
  int32_t flags;                   /* JMB_EPZS_ADAPT_PATTERN | JMB_EPZS_DUAL | JMB_EPZS_SUBPEL | JMB_EPZS_TEST8X8 (block types <= 4) */
  int32_t ref;
  int32_t pattern, pattern_dual;
  int32_t n_shared;                /* candidates per macroblock in `shared` (<= 32) */
  int32_t window;                  /* 0: none; else window predictors +-(4 << k), k = 0..window-1, 8 per ring (EPZSWindowPredictorInit) */
  int32_t range;                   /* searchRange.max_x = max_y, quarter-pel */
  int32_t medthres[8], minthres[8], maxthres[8], subthres[8];
  int32_t mv_min_x, mv_max_x, mv_min_y, mv_max_y;
} jmb_epzs_frame_params;
int jmb_epzs_search_frame(jmb_ctx *ctx, const jmb_mb_mvpred *pred, const int16_t *shared, int n_mb, const jmb_epzs_frame_params *fp,
                          jmb_me_res8 *res, int loc);

/* ---- macroblock-resident surfaces + per-partition arg-min: the form JM's sequential call sites can use -----------------------
 * jmb_mb_surfaces: the sixteen 4x4 SAD surfaces of macroblock (mb_x, mb_y) against reference `ref` for every displacement within
 * `radius` integer pels of the centre (quarter-pel, multiple of 4) -- setup_fast_full_search's BlockSAD (me_fullfast.c:492-556),
 * kept as uint16 in HBM, one set per reference, valid until the next jmb_pic_begin / jmb_ref_put.  Only enqueues work.
 * jmb_mb_search: one partition's search over the resident surfaces -- fast_full_search_motion_estimation (me_fullfast.c:618-689,
 * JMB_SEARCH_FAST_FULL: macroblock-origin clamp, max_mvd guard) or full_search_motion_estimation (me_fullsearch.c:39-103,
 * JMB_SEARCH_FULL: the request's own centre, partition-origin clamp) -- followed by the sub-pel refinement with JMB_REQ_SUBPEL
 * (or that alone with JMB_REQ_SKIP_INT).  Synchronous; the answer comes back through a host-mapped mailbox (no copy calls).
 * JMB_ERR_STATE when the request's window is not covered by the resident surfaces (the caller then uses jmb_me_search). */
int jmb_mb_surfaces(jmb_ctx *ctx, int ref, int mb_x, int mb_y, int center_x, int center_y, int radius);
int jmb_mb_search(jmb_ctx *ctx, const jmb_me_req *req, jmb_me_res *res);

/* jmb_mb_chain: up to JMB_CHAIN_MAX searches of ONE macroblock and ONE reference over the resident surfaces in one synchronous
 * call, where later searches take their motion-vector predictor from earlier ones -- what PartitionMotionSearch /
 * SubPartitionMotionSearch do block after block (lencod/src/mv_search.c:1560-1850).  Per search the device does what
 * BlockMotionSearch does (mv_search.c:858-1024): GetMVPredictor from the three neighbours (lcommon/src/mv_prediction.c:192-300,
 * non-MBAFF; the caller resolves get_neighbors, macroblock.c, incl. the C -> D replacement) -> search centre (FULL: the
 * predictor rounded to integer pels, :931-934, clipped to the mv range, :957; FAST_FULL: req.center) -> IntPelME -> SubPelME
 * (JMB_REQ_SUBPEL) -> clip_mv_range (:981); the clipped mv is what later searches of the chain see at that position
 * (set_me_parameters, mv_search.c:100).  A neighbour is either given (mv, ref_idx as they stand in enc_picture->mv_info) or
 * names an earlier search of the same chain (dep); searches with different chain ids are independent and run side by side.
 * req.pred_x/y (and, for FULL, req.center) are outputs here: the result echoes what the device derived, so that a caller who
 * ran ahead of JM's own call sequence can check each answer against the predictor JM computes when it gets there.
 * status: DONE; UNCOVERED = the search window leaves the resident surfaces (not searched; use jmb_me_search);
 * SKIPPED = a search it depends on was not done. */
#define JMB_CHAIN_MAX 16
enum { JMB_CHAIN_DONE = 0, JMB_CHAIN_UNCOVERED = 1, JMB_CHAIN_SKIPPED = 2 };
typedef struct jmb_chain_nb {      /* 8 bytes: one neighbour (A, B or C) of a block */
  int16_t mv_x, mv_y;              /* mv_info[pos_y][pos_x].mv[list] */
  int8_t ref_idx;                  /* mv_info[pos_y][pos_x].ref_idx[list] (for dep >= 0: the reference being searched) */
  int8_t available;                /* PixelPos.available */
  int8_t dep;                      /* -1, or the index (in this call) of the earlier search whose block covers the position */
  int8_t pad_;
} jmb_chain_nb;
typedef struct jmb_chain_req {     /* 72 bytes */
  jmb_me_req req;
  jmb_chain_nb nb[3];              /* A (left), B (up), C (up-right, or D where JM substitutes it) */
  int8_t jm_ref;                   /* ref_frame GetMVPredictor compares the neighbours' ref_idx with */
  int8_t chain;                    /* chain id, 0 .. (number of chains - 1); the searches of a chain appear in order */
  int8_t pad_[6];
} jmb_chain_req;
typedef struct jmb_chain_res {     /* 40 bytes */
  jmb_me_res res;
  int16_t pred_x, pred_y, center_x, center_y;
  int32_t status, pad_;
} jmb_chain_res;
/* mv_limits = {MaxHmvR[4], MaxHmvR[5], MaxVmvR[4], MaxVmvR[5]} (quarter-pel); int_divide = JM_INT_DIVIDE (lencod/inc/defines.h) */
int jmb_mb_chain(jmb_ctx *ctx, const jmb_chain_req *reqs, int n, const int32_t mv_limits[4], int int_divide, jmb_chain_res *res);

/* BlockSAD surfaces of one macroblock exactly as setup_fast_full_search leaves them:
 * out[(blocktype*16 + slot) * max_pos + pos], blocktype 1..7, uint32 (distpel), spiral order.
 * (lencod/src/me_fullfast.c:59-81 allocation, :492-556, :196-260) */
int jmb_ffs_surfaces(jmb_ctx *ctx, int ref, int mb_x, int mb_y, int center_x, int center_y,
                     uint32_t *out, int loc);

/* distortion of n candidates of one block: computeSAD / computeSSE / computeSATD
 * (lencod/src/me_distortion.c:349,1190,745).  cand = absolute quarter-pel positions (x,y pairs);
 * out[i] = full (never early-terminated) distortion, not scaled by 32. */
int jmb_dist(jmb_ctx *ctx, int ref, int metric, int blocktype, int pos_x, int pos_y,
             const int16_t *cand_xy, int n, int test8x8, int32_t *out, int loc);

/* The other nine members of JM's distortion table (p_Vid->computeUniPred[3..5], computeBiPred1[], computeBiPred2[],
 * lencod/src/mv_search.c:486-506): the block is compared with a PREDICTION formed from one or two references.
 *   JMB_PRED_PLAIN            ref1                                                      computeSAD/SSE/SATD
 *   JMB_PRED_WEIGHTED         clip(((w1*ref1 + round) >> denom) + offset)               computeSADWP :434, SATDWP :833, SSEWP :1261
 *   JMB_PRED_AVERAGE          (ref1 + ref2 + 1) >> 1                                    computeBiPredSAD1 :525, SATD1 :943, SSE1 :1353
 *   JMB_PRED_WEIGHTED_AVERAGE clip(((w1*ref1 + w2*ref2 + 2*round) >> (denom+1)) + offset)   computeBiPredSAD2 :624, SATD2 :1038, SSE2 :1438
 * weight1/offset = mv_block->weight_luma/offset_luma (WEIGHTED) or weight1/weight2/offsetBi (WEIGHTED_AVERAGE,
 * PrepareBiPredMEParams mv_search.c:205); log_weight_denom / wp_round = currSlice->luma_log_weight_denom / wp_luma_round.
 * The second reference is read at ONE candidate (cand2, absolute quarter-pel) for the whole call, as JM's bi-predictive
 * searches do (me_fullsearch.c:112-178, :299-390: one list moves, the other stands still); each reference gets its own
 * UMVLine4X clamp (partition origin for SAD/SSE, every sub-block origin for SATD). */
enum { JMB_PRED_PLAIN = 0, JMB_PRED_WEIGHTED = 1, JMB_PRED_AVERAGE = 2, JMB_PRED_WEIGHTED_AVERAGE = 3 };
typedef struct jmb_dist_pred {
  int32_t form;                       /* JMB_PRED_* */
  int32_t ref2;                       /* index in the picture's reference list (AVERAGE forms) */
  int32_t cand2_x, cand2_y;           /* absolute quarter-pel position read in ref2 */
  int32_t weight1, weight2, offset;
  int32_t log_weight_denom, wp_round;
} jmb_dist_pred;
int jmb_dist_ex(jmb_ctx *ctx, int ref, const jmb_dist_pred *pred, int metric, int blocktype, int pos_x, int pos_y,
                const int16_t *cand_xy, int n, int test8x8, int32_t *out, int loc);

/* The mode-decision distortion back-ends p_Vid->distortion4x4 / distortion8x8 (select_distortion, lencod/src/me_distortion.c:
 * 38-170): SAD / SSE / Hadamard SAD of nblk difference blocks of n x n int16 (n = 4 or 8) the caller formed -- the skip, direct
 * and bi-predictive candidates of GetSkipCostMB / BPredPartitionCost / BIDPartitionCost (mv_search.c:589-675, :1159-1325), the
 * 8x8 transform-size decision (macroblock.c:1413) and the intra chroma mode cost (intra_chroma.c:443).  out[i] is NOT scaled by
 * 32.  thres (may be NULL; 8x8 SAD only) = dist_down(min_cost) per block: distortion8x8SADthres stops summing rows once the
 * running sum exceeds it and returns that partial sum. */
int jmb_block_distortion(jmb_ctx *ctx, int metric, int n, const int16_t *diff, int nblk, const int32_t *thres, int32_t *out, int loc);

/* ---- transform + quantisation ---------------------------------------------------------------- */
/* forward4x4 / forward8x8 (lcommon/src/transform.c:20,353) on nblk blocks of n*n int32, in place */
int jmb_forward_transform(jmb_ctx *ctx, int32_t *blocks, int nblk, int n, int loc);

/* forward transform (optional) + quant_{4x4,8x8,8x8cavlc}_{normal,around} on nblk blocks.
 *   coef   [nblk][n*n] int32: residual (do_transform=1) or transformed coefficients (0), row-major;
 *                             on return the dequantised coefficients, as JM leaves them in tblock
 *   levels/runs [nblk][lr_stride] int32 with lr_stride = 17 (n=4), 65 (n=8), 68 (n=8 cavlc: 4 x 17)
 *   fadjust [nblk][n*n] (around only, may be NULL), coeff_cost[nblk] (added to, like *coeff_cost),
 *   nonzero [nblk] */
int jmb_quant_blocks(jmb_ctx *ctx, const jmb_quant_desc *q, int do_transform, int32_t *coef, int nblk,
                     int32_t *levels, int32_t *runs, int32_t *fadjust, int32_t *coeff_cost,
                     int32_t *nonzero, int loc);

/* whole-picture inter residual coding: luma_prediction (lencod/src/mc_prediction.c:144) ->
 * compute_residue -> forward4x4|8x8 -> quant, for every macroblock of the current picture.
 *   pred   [n_mb] macroblocks in raster order
 *   levels [n_mb][256] int16: quantised levels, per 4x4 (or 8x8) block in scan order
 *                              (4x4: block b = by*4+bx at [b*16 + k]; 8x8: block b8 at [b8*64 + k])
 *   coeff_cost [n_mb][4] int32 per 8x8 quadrant, cbp_blk [n_mb] uint32: bit b set if 4x4 block b has a
 *   nonzero level (8x8 transform: bits of the quadrant set together, macroblock.c:1004)
 *   pred == NULL: use the prediction table the last jmb_pred_from_results(..., pred = NULL, ...) call left
 *   on the device (no host round trip of an intermediate JM itself only keeps in all_mv). */
int jmb_mc_tq(jmb_ctx *ctx, const jmb_mb_pred *pred, int n_mb, const jmb_quant_desc *q,
              int16_t *levels, int32_t *coeff_cost, uint32_t *cbp_blk, int loc);

/* Residual coding of EVERY inter partition mode in one launch -- what JM's RDO loop does per inter candidate
 * (luma_residual_coding called from RDCost_for_macroblocks, lencod/src/rdopt.c:1861, once per mode): for each mode m
 * with bit m-1 set in mode_mask, prediction with the mode's motion vectors (taken from the 41 search results of the
 * macroblock, reference 0) -> residual -> forward transform -> quantisation.  Outputs are mode-major and always sized
 * for 7 modes: levels [7][n_mb][256], coeff_cost [7][n_mb][4], cbp_blk [7][n_mb] (layouts as jmb_mc_tq).
 *   res == NULL: the results of the last jmb_me_search_frame call, still resident on the device. */
int jmb_mc_tq_modes(jmb_ctx *ctx, const jmb_me_res *res, int n_mb, unsigned mode_mask, const jmb_quant_desc *q,
                    int16_t *levels, int32_t *coeff_cost, uint32_t *cbp_blk, int loc);

/* jmb_mc_tq_modes with the output in JM's own shape -- (level, run) lists, quant4x4_normal.c:94-112 -- instead of dense
 * levels: per (mode, macroblock) one 16-byte head and, only for nonzero levels, 4-byte tokens.  Tokens of one (mode,
 * macroblock) are contiguous at tokens[head.token_off .. +head.n_tokens), ordered by transform block then scan position
 * = ACLevel/ACRun order; blk = 4x4 block index by*4+bx (n = 4), quadrant b8 (n = 8), b8*4 + list (n = 8 CAVLC: the four
 * interleaved lists of quant_8x8cavlc_normal).  cost8 = coeff_cost per 8x8 quadrant saturated at 255 (JM only compares it
 * with _LUMA_COEFF_COST_ = 4 and _LUMA_MB_COEFF_COST_ = 5, macroblock.c:1238-1255; one level > 1 alone adds 999999).
 * heads [7][n_mb] mode-major; *n_tokens = tokens produced; JMB_ERR_STATE if they exceed token_cap (nothing lost on the device:
 * call again with a larger buffer).  JMB_HOST copies the heads, then exactly the tokens produced. */
typedef struct jmb_tq_head { uint32_t cbp_blk; uint32_t token_off; uint16_t n_tokens; uint8_t cost8[4]; uint16_t reserved_; } jmb_tq_head;
typedef struct jmb_tq_token { int16_t level; uint8_t run; uint8_t blk; } jmb_tq_token;
int jmb_mc_tq_modes_compact(jmb_ctx *ctx, const jmb_me_res *res, int n_mb, unsigned mode_mask, const jmb_quant_desc *q,
                            jmb_tq_head *heads, jmb_tq_token *tokens, uint32_t token_cap, uint32_t *n_tokens, int loc);

/* ---- chroma of inter macroblocks: motion-compensated prediction + residual coding ------------------------
 * jmb_ref_put_chroma / jmb_pic_chroma: the U and V planes of a reference slot / of the current picture (8-bit samples when
 * sample_bytes = 1, uint16_t when 2), width_c x height_c each (4:2:0: w/2 x h/2; 4:2:2: w/2 x h).
 * jmb_chroma_residual_coding: for macroblocks first_mb .. first_mb + n_mb - 1 of a P slice, both components:
 *   prediction      OneComponentChromaPrediction4x4 (lencod/src/mc_prediction.c:292-352): every chroma sample takes the motion
 *                   vector of the luma 4x4 block above it, bilinear at 1/8 sample (1/4 vertically for 4:2:2), picture clamp;
 *   residual coding residual_transform_quant_chroma_4x4 (lencod/src/block.c:954-1202): forward4x4 of the 4 / 8 blocks, DC through
 *                   hadamard2x2 + quant_dc2x2_normal (4:2:0) or hadamard4x2 + quant_dc4x2_normal at qp + 3 (4:2:2), quant_ac4x4_normal
 *                   with one running coeff_cost per component, the _CHROMA_COEFF_COST_ threshold, inverse4x4 + sample_reconstruct.
 * Motion: pred[i] (as jmb_mc_tq) or, with pred == NULL, partition mode `mode` of the resident search results, reference 0.
 * Outputs per macroblock: dc_levels [2][8] and ac_levels [2][8][15] int16, dense in scan order (SCAN_YUV420 / SCAN_YUV422, zig-zag
 * positions 1..15; block b = by * 2 + bx); cbp_blk_chroma = bits 16.. of currMB->cbp_blk shifted down by 16 (cbp_blk_chroma,
 * block.c:158); cr_cbp = what chroma_residual_coding adds to currMB->cbp >> 4 (0, 1, 2); recon [2][16][8] uint8 (may be NULL). */
typedef struct jmb_chroma_desc {
  int32_t yuv_format;              /* 1 = 4:2:0, 2 = 4:2:2 */
  int32_t is_cavlc;
  int32_t qp_ac[2], qp_dc[2];      /* currMB->qpc[uv] + bitdepth_chroma_qp_scale; the DC qp is that + 3 for 4:2:2 */
  int32_t params_ac[2][16][3];     /* q_params_4x4[uv + 1][0][qp_ac][j][i] = {OffsetComp, ScaleComp, InvScaleComp} at [j * 4 + i] */
  int32_t params_dc[2][3];         /* q_params_4x4[uv + 1][0][qp_dc][0][0] */
  uint8_t c_cost[16];              /* COEFF_COST4x4[disthres] */
} jmb_chroma_desc;
int jmb_ref_put_chroma(jmb_ctx *ctx, int slot, const void *u, const void *v, int sample_bytes, int width_c, int height_c, int stride, int loc);
int jmb_pic_chroma(jmb_ctx *ctx, const void *u, const void *v, int sample_bytes, int width_c, int height_c, int stride, int loc);
int jmb_chroma_residual_coding(jmb_ctx *ctx, const jmb_mb_pred *pred, int mode, int first_mb, int n_mb, const jmb_chroma_desc *d,
                               int16_t *dc_levels, int16_t *ac_levels, uint32_t *cbp_blk_chroma, uint32_t *cr_cbp, uint8_t *recon, int loc);

/* nlist lists of q->m coefficients (scan order), in place: on return the dequantised coefficients; levels / runs [nlist][17]
 * (terminated by level 0), fadjust [nlist][m] (around only, may be NULL), coeff_cost [nlist] (added to; may be NULL), nonzero [nlist] */
int jmb_quant_list(jmb_ctx *ctx, const jmb_qlist_desc *q, int32_t *coef, int nlist, int32_t *levels, int32_t *runs, int32_t *fadjust,
                   int32_t *coeff_cost, int32_t *nonzero, int loc);

/* hadamard4x4 / ihadamard4x4 / hadamard4x2 / ihadamard4x2 / hadamard2x2 / ihadamard2x2 (lcommon/src/transform.c:121-330) on nblk
 * blocks, in place.  Flat layouts: 4x4 = 16 row-major; 4x2 = 8 (forward: 2 rows x 4 in and out; inverse: 2 rows x 4 in,
 * 4 rows x 2 out); 2x2 = 4 ({b[0][0], b[0][4], b[4][0], b[4][4]} in for the forward, as JM reads them). */
int jmb_hadamard(jmb_ctx *ctx, int kind, int32_t *vals, int nblk, int loc);

/* inverse4x4 / inverse8x8 (lcommon/src/transform.c:70, :450) on nblk blocks of n*n int32 (dequantised coefficients), in place */
int jmb_inverse_transform(jmb_ctx *ctx, int32_t *blocks, int nblk, int n, int loc);

/* luma_residual_coding (lencod/src/macroblock.c:1182-1257) of a non-skipped inter macroblock of a P slice, for every
 * partition mode with its bit set in mode_mask and every macroblock, in one launch: prediction from the macroblock's 41
 * search results (reference 0) -> residual -> forward transform -> quant_{4x4,8x8,8x8cavlc}_normal -> per block inverse
 * transform + sample_reconstruct (lcommon/src/blk_prediction.c:48) -> JM's coefficient thresholding (quadrant cost <=
 * _LUMA_COEFF_COST_ -> reset_block, macroblock.c:806; macroblock cost <= _LUMA_MB_COEFF_COST_ -> luma cbp cleared,
 * :1248-1255) -> distortion.  Outputs are mode-major, sized for 7 modes:
 *   levels [7][n_mb][256] (as jmb_mc_tq, zeroed where reset_block applies), cost8 [7][n_mb][4] (after the resets),
 *   cbp_blk [7][n_mb] (bits 0..15), cbp [7][n_mb] (bits 0..3), recon [7][n_mb][16][16] uint8 (may be NULL),
 *   sse [7][n_mb] = sum (source - reconstruction)^2, the luma distortion RDCost_for_macroblocks charges (rdopt.c:1886).
 * Adaptive rounding is not available here (its offsets are updated from macroblock to macroblock, q_around.c). */
int jmb_luma_residual_coding_modes(jmb_ctx *ctx, const jmb_me_res *res, int n_mb, unsigned mode_mask, const jmb_quant_desc *q,
                                   int16_t *levels, int32_t *cost8, uint32_t *cbp_blk, uint32_t *cbp, uint8_t *recon, int32_t *sse, int loc);

/* The same for n_mb macroblocks starting at picture address first_mb whose prediction is given explicitly
 * (pred[i] describes macroblock first_mb + i: partition mode, reference and 4x4 mvs per quadrant, as in jmb_mc_tq) --
 * what luma_residual_coding sees in currMB->b8x8[] / currSlice->all_mv.  Outputs are [n_mb]-sized. */
int jmb_luma_residual_coding(jmb_ctx *ctx, const jmb_mb_pred *pred, int first_mb, int n_mb, const jmb_quant_desc *q,
                             int16_t *levels, int32_t *cost8, uint32_t *cbp_blk, uint32_t *cbp, uint8_t *recon, int32_t *sse, int loc);

/* all_mv fill of BlockMotionSearch (lencod/src/mv_search.c:1005-1014): turn the results of a
 * jmb_me_search_frame call (41 per macroblock, canonical order) into the jmb_mb_pred of partition
 * mode `mode` (1..7) for every macroblock, reference 0.
 *   res == NULL : use the results of the last jmb_me_search / jmb_me_search_frame call, still resident on the device;
 *   pred == NULL: keep the table on the device for jmb_mc_tq(pred = NULL). */
int jmb_pred_from_results(jmb_ctx *ctx, const jmb_me_res *res, int n_mb, int mode, jmb_mb_pred *pred, int loc);

/* ---- measurement: CUDA events recorded on the context's stream around every kernel launch ----- */
/* ---- deblocking: DeblockFrame (lencod/src/loopFilter.c:63-299) with the non-MBAFF strength and edge functions of
 * lencod/src/loop_filter_normal.c (GetStrengthVer :53, GetStrengthHor :182, EdgeLoopLumaVer :310, EdgeLoopLumaHor :447,
 * EdgeLoopChromaVer :585, EdgeLoopChromaHor :677), frame pictures, 8-bit, 4:0:0 / 4:2:0 / 4:2:2.  The picture is filtered IN
 * PLACE, macroblock after macroblock in raster order as the standard has it (a macroblock's left edge reads samples its left
 * neighbour's horizontal edges have already changed); the device walks the macroblocks as a wavefront x + 2y with the same
 * result.  Per macroblock the caller gives what DeblockMb and the strength functions read: */
#define JMB_DB_T8X8     1          /* Macroblock.luma_transform_size_8x8_flag */
#define JMB_DB_CBP      2          /* Macroblock.cbp != 0 */
#define JMB_DB_AVAIL_A  4          /* Macroblock.mbAvailA / mbAvailB: only read with DFDisableIdc 2 (no filtering across slice boundaries) */
#define JMB_DB_AVAIL_B  8
typedef struct jmb_db_mb {         /* 176 bytes */
  uint8_t  mb_type;                /* Macroblock.mb_type: 0 PSKIP / BSKIP_DIRECT, 1 P16x16, 2 P16x8, 3 P8x16, 8 P8x8, 9 I4MB, 10 I16MB, 13 I8MB, 14 IPCM */
  uint8_t  flags;                  /* JMB_DB_* */
  int8_t   qp, qpc[2];             /* Macroblock.qp, qpc[0..1] (0 for IPCM, loopFilter.c:40-47) */
  int8_t   df_disable_idc, df_alpha_c0_offset, df_beta_offset;      /* DFDisableIdc, DFAlphaC0Offset, DFBetaOffset */
  uint32_t cbp_blk;                /* low 16 bits of Macroblock.cbp_blk: the coded luma 4x4 blocks, raster order */
  uint32_t pad_;
  int16_t  mv[2][16][2];           /* enc_picture->mv_info[][].mv[list] of the sixteen 4x4 blocks (raster order), quarter-pel */
  int8_t   ref_id[2][16];          /* which picture mv_info[][].ref_pic[list] is: -1 = none (ref_idx -1), else an id equal exactly for equal pictures */
} jmb_db_mb;
/* luma / cb / cr: u8 planes of width x height (chroma: width/2 x height/2 or height); cb = cr = NULL with yuv_format 0.
 * slice_type: JM's (0 P, 1 B, 2 I); SP / SI slices and MBAFF / field pictures are refused.  mbs: width/16 * height/16 entries, raster order. */
int jmb_deblock_picture(jmb_ctx *ctx, uint8_t *luma, int pitch, uint8_t *cb, uint8_t *cr, int pitch_c, int width, int height, int yuv_format,
                        int slice_type, int direct_8x8_inference, const jmb_db_mb *mbs, int loc);

/* kernel names: subpel_planes, pack_cur, int_search, subpel_refine, dist, ffs_surfaces, forward,
 * quant_blocks, mc_tq, pred_from_results, gen_requests (EPZS request generation and result packing), epzs, chroma, deblock, argmin */
int jmb_timing_enable(jmb_ctx *ctx, int on);      /* also clears the accumulated samples */
int jmb_timing_get(jmb_ctx *ctx, const char *kernel, double *total_ms, int *launches);

#ifdef __cplusplus
}
#endif
#endif /* JMB200_H */
