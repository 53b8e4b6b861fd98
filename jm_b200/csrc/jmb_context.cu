// jmb_context.cu -- context, memory and picture management of libjmb200 (host side, C++).
#include <stdarg.h>
#include <stdlib.h>
#include "jmb_internal.h"

static char g_create_err[512] = "";

int jmb_fail(jmb_ctx *ctx, int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(ctx ? ctx->err : g_create_err, 512, fmt, ap);
  va_end(ap);
  return code;
}

int jmb_make_tmap_u8(jmb_ctx *ctx, CUtensorMap *out, const void *base, int width, int height, int pitch, int box_w, int box_h) {
  typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static encode_fn encode = nullptr;
  if (!encode) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    JMB_CUDA(ctx, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) return jmb_fail(ctx, JMB_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    encode = (encode_fn)fn;
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)width, (cuuint64_t)height}, gstride[1] = {(cuuint64_t)pitch};
  const cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)box_h}, estride[2] = {1, 1};
  CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void *)base, gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return jmb_fail(ctx, JMB_ERR_CUDA, "cuTensorMapEncodeTiled(%dx%d pitch %d box %dx%d) -> %d", width, height, pitch, box_w, box_h, (int)r);
  return 0;
}

int jmb_reserve_host(jmb_ctx *ctx, void **p, size_t *cap, size_t bytes) {
  if (bytes <= *cap) return 0;
  if (*p) JMB_CUDA(ctx, cudaFreeHost(*p));
  *p = nullptr; *cap = 0;
  size_t want = bytes + bytes / 4 + 4096;
  JMB_CUDA(ctx, cudaHostAlloc(p, want, cudaHostAllocDefault));
  *cap = want;
  return 0;
}

int jmb_reserve_dev(jmb_ctx *ctx, void **p, size_t *cap, size_t bytes) {
  if (bytes <= *cap) return 0;
  if (*p) { JMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); JMB_CUDA(ctx, cudaFree(*p)); }
  *p = nullptr; *cap = 0;
  size_t want = bytes + bytes / 4 + 4096;
  JMB_CUDA(ctx, cudaMalloc(p, want));
  *cap = want;
  return 0;
}

void jmb_time_begin(jmb_ctx *ctx, int kid) {
  if (!ctx->timing) return;
  if (ctx->ev_n[kid] == ctx->ev_cap[kid]) {
    int cap = ctx->ev_cap[kid] ? ctx->ev_cap[kid] * 2 : 64;
    ctx->ev[kid] = (jmb_ctx::EvPair *)realloc(ctx->ev[kid], cap * sizeof(jmb_ctx::EvPair));
    for (int i = ctx->ev_cap[kid]; i < cap; i++) { cudaEventCreate(&ctx->ev[kid][i].a); cudaEventCreate(&ctx->ev[kid][i].b); }
    ctx->ev_cap[kid] = cap;
  }
  cudaEventRecord(ctx->ev[kid][ctx->ev_n[kid]].a, ctx->stream);
}
void jmb_time_end(jmb_ctx *ctx, int kid) {
  if (!ctx->timing) return;
  cudaEventRecord(ctx->ev[kid][ctx->ev_n[kid]].b, ctx->stream);
  ctx->ev_n[kid]++;
}

// Reads the device-side validation word (and clears it).  Called by every entry point that synchronises.
int jmb_check_device_errors(jmb_ctx *ctx) {
  JMB_CUDA(ctx, cudaMemcpyAsync(ctx->h_err, ctx->d_err, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  JMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->h_err[0]) {
    const int code = ctx->h_err[0], idx = ctx->h_err[1];
    JMB_CUDA(ctx, cudaMemsetAsync(ctx->d_err, 0, 2 * sizeof(int), ctx->stream));
    return jmb_fail(ctx, JMB_ERR_ARG, "motion-search request %d rejected on the device (code 0x%x:%s%s%s%s%s%s%s%s%s); its result was not written", idx, code,
                    code & JMB_REQERR_BLOCKTYPE ? " blocktype" : "", code & JMB_REQERR_REF ? " ref" : "", code & JMB_REQERR_POS ? " position/alignment" : "",
                    code & JMB_REQERR_CENTER ? " centre-not-integer-pel" : "", code & JMB_REQERR_MODE ? " mode" : "", code & JMB_REQERR_LAMBDA ? " lambda" : "",
                    code & JMB_REQERR_MINCOST ? " min_mcost" : "", code & JMB_REQERR_LAYOUT ? " frame-layout" : "",
                    code & JMB_REQERR_FPEL_METRIC ? " full-search-with-MEDistortionFPel-other-than-SAD" : "");
  }
  return JMB_OK;
}

extern "C" {

static const char *k_names[JMB_K_COUNT] = {"subpel_planes", "pack_cur", "int_search", "subpel_refine", "dist", "ffs_surfaces",
                                           "forward", "quant_blocks", "mc_tq", "pred_from_results", "gen_requests", "epzs",
                                           "chroma", "deblock", "argmin"};

int jmb_timing_enable(jmb_ctx *ctx, int on) {
  JMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->timing = on != 0;
  for (int k = 0; k < JMB_K_COUNT; k++) ctx->ev_n[k] = 0;
  return JMB_OK;
}

int jmb_timing_get(jmb_ctx *ctx, const char *kernel, double *total_ms, int *launches) {
  JMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (int k = 0; k < JMB_K_COUNT; k++)
    if (!strcmp(kernel, k_names[k])) {
      double t = 0;
      for (int i = 0; i < ctx->ev_n[k]; i++) { float ms = 0; JMB_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev[k][i].a, ctx->ev[k][i].b)); t += ms; }
      *total_ms = t; *launches = ctx->ev_n[k];
      return JMB_OK;
    }
  return jmb_fail(ctx, JMB_ERR_ARG, "jmb_timing_get: unknown kernel '%s'", kernel);
}

int jmb_abi_version(void) { return JMB_ABI_VERSION; }

const char *jmb_last_error(const jmb_ctx *ctx) { return ctx ? ctx->err : g_create_err; }

int jmb_create(int device, jmb_ctx **out) {
  if (!out) return jmb_fail(nullptr, JMB_ERR_ARG, "jmb_create: out is NULL");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return jmb_fail(nullptr, JMB_ERR_NO_DEVICE,
                    "jmb_create: no CUDA device (%s); libjmb200 has no CPU path",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
  if (device < 0 || device >= count)
    return jmb_fail(nullptr, JMB_ERR_ARG, "jmb_create: device %d out of range (0..%d)", device, count - 1);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major < 10)
    return jmb_fail(nullptr, JMB_ERR_UNSUPPORTED,
                    "jmb_create: device %d is sm_%d%d; libjmb200 is built for sm_100a only", device,
                    prop.major, prop.minor);
  jmb_ctx *ctx = new jmb_ctx();
  ctx->device = device;
  if (cudaSetDevice(device) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
    jmb_fail(nullptr, JMB_ERR_CUDA, "jmb_create: cannot create a stream on device %d: %s", device,
             cudaGetErrorString(cudaGetLastError()));
    delete ctx;
    return JMB_ERR_CUDA;
  }
  // defaults = the bundled encoder cfgs: SearchRange 32, SAD / SATD / SATD
  ctx->me.search_range = 32; ctx->me.max_mvd = 1023;
  ctx->me.metric[0] = JMB_SAD; ctx->me.metric[1] = JMB_SATD; ctx->me.metric[2] = JMB_SATD;
  ctx->me.start_hp = 0; ctx->me.start_qp = 1; ctx->me.search_pos2 = 9; ctx->me.search_pos4 = 9;
  if (cudaMalloc(&ctx->d_err, 2 * sizeof(int)) != cudaSuccess || cudaMemset(ctx->d_err, 0, 2 * sizeof(int)) != cudaSuccess ||
      cudaHostAlloc(&ctx->h_err, 2 * sizeof(int), cudaHostAllocDefault) != cudaSuccess) {
    jmb_fail(nullptr, JMB_ERR_CUDA, "jmb_create: cannot allocate the error word: %s", cudaGetErrorString(cudaGetLastError()));
    jmb_destroy(ctx);
    return JMB_ERR_CUDA;
  }
  ctx->h_err[0] = ctx->h_err[1] = 0;
  *out = ctx;
  return JMB_OK;
}

void jmb_destroy(jmb_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (int i = 0; i < JMB_MAX_REFS; i++) { if (ctx->refs[i].planes) cudaFree(ctx->refs[i].planes); if (ctx->refs[i].chroma) cudaFree(ctx->refs[i].chroma); }
  if (ctx->cur_c) cudaFree(ctx->cur_c);
  if (ctx->cur) cudaFree(ctx->cur);
  if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
  if (ctx->h_groups) cudaFreeHost(ctx->h_groups);
  if (ctx->d_stage) cudaFree(ctx->d_stage);
  if (ctx->d_stage2) cudaFree(ctx->d_stage2);
  if (ctx->d_groups) cudaFree(ctx->d_groups);
  if (ctx->d_reftab) cudaFree(ctx->d_reftab);
  if (ctx->d_qdesc) cudaFree(ctx->d_qdesc);
  if (ctx->d_stage3) cudaFree(ctx->d_stage3);
  if (ctx->d_stage4) cudaFree(ctx->d_stage4);
  if (ctx->d_stage5) cudaFree(ctx->d_stage5);
  if (ctx->d_res_keep) cudaFree(ctx->d_res_keep);
  if (ctx->d_pred_keep) cudaFree(ctx->d_pred_keep);
  for (int i = 0; i < ctx->n_peers; i++) cudaIpcCloseMemHandle(ctx->peers[i].mapped);
  for (int i = 0; i < JMB_MAX_REFS; i++) if (ctx->surf[i].buf) cudaFree(ctx->surf[i].buf);
  if (ctx->mbox) cudaFreeHost(ctx->mbox);
  if (ctx->d_one) cudaFree(ctx->d_one);
  if (ctx->d_db) cudaFree(ctx->d_db);
  if (ctx->d_mvpred) cudaFree(ctx->d_mvpred);
  if (ctx->d_res8) cudaFree(ctx->d_res8);
  if (ctx->d_heads) cudaFree(ctx->d_heads);
  if (ctx->d_tokens) cudaFree(ctx->d_tokens);
  if (ctx->d_tok_count) cudaFree(ctx->d_tok_count);
  if (ctx->h_tok_count) cudaFreeHost(ctx->h_tok_count);
  if (ctx->d_err) cudaFree(ctx->d_err);
  if (ctx->h_err) cudaFreeHost(ctx->h_err);
  for (int k = 0; k < 16; k++) {
    for (int i = 0; i < ctx->ev_cap[k]; i++) { cudaEventDestroy(ctx->ev[k][i].a); cudaEventDestroy(ctx->ev[k][i].b); }
    free(ctx->ev[k]);
  }
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

int jmb_sync(jmb_ctx *ctx) {
  return jmb_check_device_errors(ctx);
}

void *jmb_stream(jmb_ctx *ctx) { return (void *)ctx->stream; }
uint64_t jmb_launch_count(const jmb_ctx *ctx) { return ctx->launches; }

int jmb_host_alloc(jmb_ctx *ctx, size_t bytes, void **out) {
  JMB_CUDA(ctx, cudaHostAlloc(out, bytes, cudaHostAllocDefault));
  return JMB_OK;
}
int jmb_host_free(jmb_ctx *ctx, void *p) {
  JMB_CUDA(ctx, cudaFreeHost(p));
  return JMB_OK;
}

int jmb_me_configure(jmb_ctx *ctx, const jmb_me_config *cfg) {
  if (!cfg) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_me_configure: cfg is NULL");
  if (cfg->search_range < 1 || cfg->search_range > JMB_MAX_SEARCH_RANGE)
    return jmb_fail(ctx, JMB_ERR_UNSUPPORTED, "jmb_me_configure: search_range %d not in 1..%d", cfg->search_range, JMB_MAX_SEARCH_RANGE);
  if (cfg->max_mvd < 2) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_me_configure: max_mvd %d (must be > 1)", cfg->max_mvd);
  // start_me_refinement_hp/qp are 0 or 1 (mv_search.c:445-446); the refinement kernel indexes its nine candidates with them
  if ((cfg->start_hp & ~1) || (cfg->start_qp & ~1))
    return jmb_fail(ctx, JMB_ERR_ARG, "jmb_me_configure: start_hp %d / start_qp %d must be 0 or 1", cfg->start_hp, cfg->start_qp);
  for (int i = 0; i < 3; i++)
    if (cfg->metric[i] < 0 || cfg->metric[i] > 2)
      return jmb_fail(ctx, JMB_ERR_ARG, "jmb_me_configure: metric[%d]=%d", i, cfg->metric[i]);
  // metric[0] != SAD is accepted here (sub-pel-only requests and jmb_dist do not care); a FULL-search request under it
  // is rejected on the device (JMB_REQERR_FPEL_METRIC): the search kernel is a SAD kernel
  if (cfg->search_pos2 < 1 || cfg->search_pos2 > 9 || cfg->search_pos4 < 1 || cfg->search_pos4 > 9)
    return jmb_fail(ctx, JMB_ERR_ARG, "jmb_me_configure: search_pos2/4 must be 1..9");
  ctx->me = *cfg;
  ctx->me_configured = true;
  return JMB_OK;
}

// Bring `bytes` of caller data (host or device) to the device; returns the device pointer to read.
static int to_device(jmb_ctx *ctx, const void *src, size_t bytes, int loc, void **dptr, void **scratch, size_t *cap) {
  if (loc == JMB_DEVICE) { *dptr = (void *)src; return 0; }
  if (!jmb_is_host(loc)) return jmb_fail(ctx, JMB_ERR_ARG, "loc %d is not JMB_HOST / JMB_DEVICE / JMB_HOST_ASYNC", loc);
  int rc = jmb_reserve_dev(ctx, scratch, cap, bytes);
  if (rc) return rc;
  JMB_CUDA(ctx, cudaMemcpyAsync(*scratch, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  *dptr = *scratch;
  return 0;
}

static int ref_put_impl(jmb_ctx *ctx, int slot, const void *luma, int sample_bytes, int width, int height, int stride, int loc) {
  if (slot < 0 || slot >= JMB_MAX_REFS) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_ref_put: slot %d", slot);
  if (width < 16 || height < 16 || (width & 15) || (height & 15) || stride < width)
    return jmb_fail(ctx, JMB_ERR_ARG, "jmb_ref_put: %dx%d stride %d (need multiples of 16)", width, height, stride);
  if (!luma) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_ref_put: NULL picture");
  JMB_CUDA(ctx, cudaSetDevice(ctx->device));
  jmb_ref *r = &ctx->refs[slot];
  int W = width + 2 * JMB_PAD_X, H = height + 2 * JMB_PAD_Y;
  int pitch = (W + 127) & ~127;
  size_t plane_bytes = (size_t)pitch * H;
  if (!r->planes || r->w != width || r->h != height) {
    if (r->planes) { JMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); JMB_CUDA(ctx, cudaFree(r->planes)); r->planes = nullptr; }
    JMB_CUDA(ctx, cudaMalloc(&r->planes, plane_bytes * 16 + 64));   // + slack: unaligned 4-sample reads fetch the next word
    r->w = width; r->h = height; r->W = W; r->H = H; r->pitch = pitch; r->plane_bytes = plane_bytes;
    int rc = jmb_make_tmap_u8(ctx, &r->tmap_int, r->planes, W, H, pitch, JMB_WIN_BOX_W, JMB_WIN_BOX_H);
    if (rc) return rc;
  }
  void *d_src = nullptr;
  int rc = to_device(ctx, luma, (size_t)stride * height * sample_bytes, loc, &d_src, &ctx->d_stage, &ctx->d_stage_cap);
  if (rc) return rc;
  rc = jmb_launch_subpel(ctx, d_src, sample_bytes, stride, r);
  if (rc) return rc;
  r->valid = true;
  ctx->pic_serial++;
  if (loc == JMB_HOST) JMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return JMB_OK;
}

int jmb_ref_put(jmb_ctx *ctx, int slot, const uint16_t *luma, int width, int height, int stride,
                int bitdepth, int loc) {
  if (bitdepth != 8)
    return jmb_fail(ctx, JMB_ERR_UNSUPPORTED, "jmb_ref_put: bit depth %d (this build packs samples to 8 bits)", bitdepth);
  return ref_put_impl(ctx, slot, luma, 2, width, height, stride, loc);
}

int jmb_ref_put_u8(jmb_ctx *ctx, int slot, const uint8_t *luma, int width, int height, int stride, int loc) {
  return ref_put_impl(ctx, slot, luma, 1, width, height, stride, loc);
}

int jmb_ref_drop(jmb_ctx *ctx, int slot) {
  if (slot < 0 || slot >= JMB_MAX_REFS) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_ref_drop: slot %d", slot);
  jmb_ref *r = &ctx->refs[slot];
  JMB_CUDA(ctx, cudaSetDevice(ctx->device));
  if (r->planes) { JMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); JMB_CUDA(ctx, cudaFree(r->planes)); }
  if (r->chroma) JMB_CUDA(ctx, cudaFree(r->chroma));
  *r = jmb_ref();
  return JMB_OK;
}

__global__ void k_plane_to_u16(const uint8_t *__restrict__ plane, int pitch, int W, int H, uint16_t *__restrict__ out) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x < W && y < H) out[(size_t)y * W + x] = plane[(size_t)y * pitch + x];
}

int jmb_ref_get_plane(jmb_ctx *ctx, int slot, int fy, int fx, uint16_t *out, int loc) {
  if (slot < 0 || slot >= JMB_MAX_REFS || !ctx->refs[slot].valid)
    return jmb_fail(ctx, JMB_ERR_STATE, "jmb_ref_get_plane: slot %d holds no picture", slot);
  if ((fy | fx) & ~3) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_ref_get_plane: plane [%d][%d]", fy, fx);
  jmb_ref *r = &ctx->refs[slot];
  size_t bytes = (size_t)r->W * r->H * sizeof(uint16_t);
  uint16_t *d_out = out;
  if (loc == JMB_HOST) {
    int rc = jmb_reserve_dev(ctx, &ctx->d_stage2, &ctx->d_stage2_cap, bytes);
    if (rc) return rc;
    d_out = (uint16_t *)ctx->d_stage2;
  }
  dim3 grid((r->W + 255) / 256, r->H);
  k_plane_to_u16<<<grid, 256, 0, ctx->stream>>>(r->planes + (size_t)(fy * 4 + fx) * r->plane_bytes, r->pitch, r->W, r->H, d_out);
  JMB_LAUNCH_CHECK(ctx);
  if (loc == JMB_HOST) {
    JMB_CUDA(ctx, cudaMemcpyAsync(out, d_out, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    JMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return JMB_OK;
}

// u16 samples -> u8 plane with a 128-byte-aligned pitch (the layout every search kernel reads)
__global__ void k_pack_cur(const uint16_t *__restrict__ src, int stride, int w, int h, uint8_t *__restrict__ dst, int pitch) {
  int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y;
  if (x4 >= w || y >= h) return;
  const uint16_t *s = src + (size_t)y * stride + x4;
  uint32_t v = (uint32_t)(s[0] & 255) | ((uint32_t)(s[1] & 255) << 8) | ((uint32_t)(s[2] & 255) << 16) | ((uint32_t)(s[3] & 255) << 24);
  *(uint32_t *)(dst + (size_t)y * pitch + x4) = v;
}

static int pic_begin_impl(jmb_ctx *ctx, const void *cur, int sample_bytes, int width, int height, int stride, int loc,
                          const int *ref_slots, int nref) {
  if (width < 16 || height < 16 || (width & 15) || (height & 15) || stride < width)
    return jmb_fail(ctx, JMB_ERR_ARG, "jmb_pic_begin: %dx%d stride %d", width, height, stride);
  if (nref < 0 || nref > JMB_MAX_REFS) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_pic_begin: nref %d", nref);
  if (!cur) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_pic_begin: NULL picture");
  for (int i = 0; i < nref; i++) {
    int s = ref_slots[i];
    if (s < 0 || s >= JMB_MAX_REFS || !ctx->refs[s].valid)
      return jmb_fail(ctx, JMB_ERR_STATE, "jmb_pic_begin: reference slot %d holds no picture", s);
    if (ctx->refs[s].w != width || ctx->refs[s].h != height)
      return jmb_fail(ctx, JMB_ERR_ARG, "jmb_pic_begin: reference slot %d is %dx%d, picture is %dx%d", s,
                      ctx->refs[s].w, ctx->refs[s].h, width, height);
    ctx->ref_list[i] = s;
  }
  ctx->nref = nref;
  JMB_CUDA(ctx, cudaSetDevice(ctx->device));
  int pitch = (width + 127) & ~127;
  size_t bytes = (size_t)pitch * height;
  if (bytes > ctx->cur_cap) {
    if (ctx->cur) { JMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); JMB_CUDA(ctx, cudaFree(ctx->cur)); ctx->cur = nullptr; }
    JMB_CUDA(ctx, cudaMalloc(&ctx->cur, bytes));
    ctx->cur_cap = bytes;
    ctx->cur_w = 0;
  }
  if (ctx->cur_w != width || ctx->cur_h != height) {
    int rc = jmb_make_tmap_u8(ctx, &ctx->tmap_cur, ctx->cur, width, height, pitch, 16, 16);
    if (rc) return rc;
  }
  ctx->cur_w = width; ctx->cur_h = height; ctx->cur_pitch = pitch;
  ctx->pic_serial++;
  if (sample_bytes == 1) {      // bytes already: one strided copy straight into the search layout, no kernel
    JMB_CUDA(ctx, cudaMemcpy2DAsync(ctx->cur, pitch, cur, stride, width, height,
                                    loc == JMB_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream));
  } else {
    void *d_src = nullptr;
    int rc = to_device(ctx, cur, (size_t)stride * height * sizeof(uint16_t), loc, &d_src, &ctx->d_stage, &ctx->d_stage_cap);
    if (rc) return rc;
    dim3 grid((width / 4 + 127) / 128, height);
    jmb_time_begin(ctx, JMB_K_PACK);
    k_pack_cur<<<grid, 128, 0, ctx->stream>>>((const uint16_t *)d_src, stride, width, height, ctx->cur, pitch);
    jmb_time_end(ctx, JMB_K_PACK);
    JMB_LAUNCH_CHECK(ctx);
  }
  if (loc == JMB_HOST) JMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return JMB_OK;
}

int jmb_pic_begin(jmb_ctx *ctx, const uint16_t *cur, int width, int height, int stride, int loc,
                  const int *ref_slots, int nref) {
  return pic_begin_impl(ctx, cur, 2, width, height, stride, loc, ref_slots, nref);
}

int jmb_pic_begin_u8(jmb_ctx *ctx, const uint8_t *cur, int width, int height, int stride, int loc,
                     const int *ref_slots, int nref) {
  if (!jmb_is_host(loc) && loc != JMB_DEVICE) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_pic_begin_u8: loc %d", loc);
  return pic_begin_impl(ctx, cur, 1, width, height, stride, loc, ref_slots, nref);
}

// ---- plain device memory + peer (NVLink) mapping -------------------------------------------------------------------
int jmb_dev_alloc(jmb_ctx *ctx, size_t bytes, void **out) {
  if (!out || !bytes) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_dev_alloc: bad argument");
  JMB_CUDA(ctx, cudaSetDevice(ctx->device));
  JMB_CUDA(ctx, cudaMalloc(out, bytes));
  return JMB_OK;
}
int jmb_dev_free(jmb_ctx *ctx, void *p) {
  JMB_CUDA(ctx, cudaSetDevice(ctx->device));
  JMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  JMB_CUDA(ctx, cudaFree(p));
  return JMB_OK;
}
int jmb_dev_copy(jmb_ctx *ctx, void *dst, const void *src, size_t bytes, int dst_loc, int src_loc) {
  if (!dst || !src) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_dev_copy: NULL pointer");
  JMB_CUDA(ctx, cudaSetDevice(ctx->device));
  JMB_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, ctx->stream));
  if (dst_loc == JMB_HOST || src_loc == JMB_HOST) JMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return JMB_OK;
}
int jmb_peer_export(jmb_ctx *ctx, const void *dev_ptr, unsigned char handle[JMB_IPC_HANDLE_BYTES]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == JMB_IPC_HANDLE_BYTES, "IPC handle size");
  if (!dev_ptr || !handle) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_peer_export: NULL pointer");
  JMB_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaIpcMemHandle_t h;
  JMB_CUDA(ctx, cudaIpcGetMemHandle(&h, (void *)dev_ptr));
  memcpy(handle, &h, sizeof(h));
  return JMB_OK;
}
int jmb_peer_open(jmb_ctx *ctx, const unsigned char handle[JMB_IPC_HANDLE_BYTES], void **mapped) {
  if (!handle || !mapped) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_peer_open: NULL pointer");
  for (int i = 0; i < ctx->n_peers; i++)
    if (!memcmp(ctx->peers[i].handle, handle, JMB_IPC_HANDLE_BYTES)) { *mapped = ctx->peers[i].mapped; return JMB_OK; }
  if (ctx->n_peers == 32) return jmb_fail(ctx, JMB_ERR_STATE, "jmb_peer_open: 32 peer buffers are open already");
  JMB_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  void *p = nullptr;
  JMB_CUDA(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  memcpy(ctx->peers[ctx->n_peers].handle, handle, JMB_IPC_HANDLE_BYTES);
  ctx->peers[ctx->n_peers++].mapped = p;
  *mapped = p;
  return JMB_OK;
}
int jmb_peer_close(jmb_ctx *ctx, void *mapped) {
  for (int i = 0; i < ctx->n_peers; i++)
    if (ctx->peers[i].mapped == mapped) {
      JMB_CUDA(ctx, cudaSetDevice(ctx->device));
      JMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      JMB_CUDA(ctx, cudaIpcCloseMemHandle(mapped));
      ctx->peers[i] = ctx->peers[--ctx->n_peers];
      return JMB_OK;
    }
  return jmb_fail(ctx, JMB_ERR_ARG, "jmb_peer_close: not a pointer jmb_peer_open returned");
}

}  // extern "C"
