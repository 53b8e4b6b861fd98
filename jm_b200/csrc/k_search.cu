// k_search.cu -- K1 + K2 + K3: integer-pel full search for every partition of a macroblock.
//
// One CTA per (macroblock, reference) "group".  The CTA stages the reference search window and the
// 16x16 source block in shared memory, evaluates the sixteen 4x4 SADs of every displacement with
// packed-byte VABSDIFF4 (4 samples per instruction), sums them to the 41 partition SADs exactly as
// update_full_search_large_blocks does (lencod/src/me_fullfast.c:196-260) and keeps, per
// partition, the minimum of  J = (SAD << 5) + lambda * (mvbits[dx] + mvbits[dy])  with JM's
// tie-break: the first position in spiral order wins (strict '<' at me_fullsearch.c:89 and
// me_fullfast.c:677).  The 4x4 SADs are shared by all partitions of the macroblock; nothing but
// the 41 (mv, cost) results leaves the SM.
//
// Semantics reproduced:
//  * FULL search  (full_search_motion_estimation, lencod/src/me_fullsearch.c:39-103): each
//    partition has its own centre; computeSAD clamps the PARTITION origin (me_distortion.c:367).
//  * FAST FULL search (setup_fast_full_search + fast_full_search_motion_estimation,
//    lencod/src/me_fullfast.c:269-689): one centre per macroblock; the MACROBLOCK origin is clamped
//    (:498); candidates with GetMaxMVD >= max_mvd-1 are skipped (:671).
//  Both clamps are "displacement clamps": the block(s) are read at D = clamp(d, Dlo, Dhi) per axis,
//  so a thread that owns displacement D serves every candidate d that clamps onto it.
//  JM's early terminations only replace a losing cost by min_mcost, so complete sums decide alike.
#include "jmb_dist_dev.cuh"

namespace {

constexpr int NPART = 41;
#ifndef JMB_IS_NT
#define JMB_IS_NT 128
#endif
constexpr int NT = JMB_IS_NT;      // threads per CTA (multiple of 32, >= 64)
constexpr int CW = 92;             // chunk of displacements handled per staging pass (15 + CW + 15 + 4 <= the 128-byte TMA box)
constexpr int CH = 72;             // multiple of 4 (a thread owns 4 vertically adjacent displacements)
constexpr int WIN_PITCH = JMB_WIN_BOX_W;   // bytes per staged window row = the TMA box width (>= CW + 15 + the fifth word)
constexpr int WIN_ROWS = CH + 15;
static_assert(3 * NPART <= NT, "the stage-1 scan wants three threads per partition");
static_assert(WIN_ROWS == JMB_WIN_BOX_H && 15 + CW + 15 + 4 <= WIN_PITCH, "TMA box and window geometry disagree");
#ifndef JMB_S1_COLS
#define JMB_S1_COLS 8
#define JMB_S1_RGS 2
#endif
constexpr int S1_COLS = JMB_S1_COLS, S1_RGS = JMB_S1_RGS, S1_ITEMS = S1_COLS * S1_RGS, S1_PITCH = 43;   // stage-1 neighbourhood
constexpr int ADJ_PITCH = 44;      // words per column (>= NPART; 176 B keeps the 128-bit row reads conflict-free)
#ifndef JMB_IS_PACKED
#define JMB_IS_PACKED 1            // the sweep's gate on two partitions per register (see "packed gate" in k_int_search)
#endif
constexpr int PK_N = 19, PK_PITCH = 20;      // packed pairs of the 38 partitions of at most 8x16 samples (SAD < 2^15); words per table row
constexpr unsigned PK_ADJ_MAX = 8191;        // mv-cost terms are capped (a smaller term only loosens the gate) so that no 16-bit field overflows
constexpr int WQ_CAP = 192;          // per-warp queue of gate hits awaiting their exact evaluation
constexpr int HS_PITCH = NPART + 1;   // u16 per thread: SADs of the partitions of a displacement that met the gate
constexpr int S1_BYTES = S1_ITEMS * 4 * S1_PITCH * 2, SWEEP_BYTES = (NT / 32) * WQ_CAP * 8 + NT * HS_PITCH * 2;   // stage-1 sums and the sweep's queue / hit SADs share memory
constexpr unsigned S1_NONE = 0x3fffffffu, S1_BAD = 0x40000000u;   // real costs stay far below (lambda <= 65535)
constexpr int S1T_ROWS = S1_COLS + 4 * S1_RGS;      // exact mv-cost terms of the stage-1 columns and rows, per partition
constexpr int PK_BYTES = JMB_IS_PACKED ? (CW + CH / 4) * PK_PITCH * 4 : 0;      // packed copies of the column / row-group gate tables
constexpr int INT_SEARCH_DYN_SMEM = (CW + CH / 4 + S1T_ROWS) * ADJ_PITCH * 4 + (S1_BYTES > SWEEP_BYTES ? S1_BYTES : SWEEP_BYTES) + PK_BYTES;
#ifndef JMB_IS_SMEM_PAD
#define JMB_IS_SMEM_PAD 0      // tuning builds only: extra dynamic shared memory lowers the CTAs per SM (occupancy probe at 128 registers: 4 CTAs 0.72 ms, 3 CTAs 0.77, 2 CTAs 0.92)
#endif
constexpr int IDX_BITS = 15;       // spiral index field of the packed (cost << IDX_BITS) | index keys
constexpr int MAX_SEARCH_RANGE = 64;   // JM's largest SearchRange: (2*64+1)^2 = 16641 positions
static_assert((2 * MAX_SEARCH_RANGE + 1) * (2 * MAX_SEARCH_RANGE + 1) <= (1 << IDX_BITS), "spiral indices must fit IDX_BITS");
static_assert(48 + IDX_BITS <= 63, "min_mcost <= 2^48 (jmb_req_check) shifted by IDX_BITS must stay below 2^63");
static_assert(MAX_SEARCH_RANGE == JMB_MAX_SEARCH_RANGE, "jmb_me_configure and the search kernel must agree on the limit");

struct PartGeom { unsigned char type, bx, by, w4, h4; };
// canonical partition order: by type, then raster order of the partitions inside the macroblock
__constant__ PartGeom c_part[NPART] = {
  {1,0,0,4,4},
  {2,0,0,4,2},{2,0,2,4,2},
  {3,0,0,2,4},{3,2,0,2,4},
  {4,0,0,2,2},{4,2,0,2,2},{4,0,2,2,2},{4,2,2,2,2},
  {5,0,0,2,1},{5,2,0,2,1},{5,0,1,2,1},{5,2,1,2,1},{5,0,2,2,1},{5,2,2,2,1},{5,0,3,2,1},{5,2,3,2,1},
  {6,0,0,1,2},{6,1,0,1,2},{6,2,0,1,2},{6,3,0,1,2},{6,0,2,1,2},{6,1,2,1,2},{6,2,2,1,2},{6,3,2,1,2},
  {7,0,0,1,1},{7,1,0,1,1},{7,2,0,1,1},{7,3,0,1,1},{7,0,1,1,1},{7,1,1,1,1},{7,2,1,1,1},{7,3,1,1,1},
  {7,0,2,1,1},{7,1,2,1,1},{7,2,2,1,1},{7,3,2,1,1},{7,0,3,1,1},{7,1,3,1,1},{7,2,3,1,1},{7,3,3,1,1}};

struct ReqS {
  int active;
  int cx, cy;              // search centre, integer-pel displacement
  int px, py;              // predictor, quarter-pel
  int dlo_x, dhi_x, dlo_y, dhi_y;   // displacement clamp range of this request
  int lam;
  int ffs;                 // FAST_FULL semantics (max_mvd guard)
  int req;                 // index of the request in the caller's array
  unsigned long long init; // min_mcost << IDX_BITS
};

// everything the slow path needs, in shared memory
struct Grp {
  ReqS rq[NPART];
  unsigned long long best[NPART];      // (cost << IDX_BITS) | spiral index of the current winner
  __align__(16) unsigned thr[NPART + 3];   // gate: a SAD can only win if sad < thr (see bound_of)
  __align__(16) int4 inner[NPART];     // displacements that stand for exactly one candidate: x0, x1, y0, y1 (inclusive)
  __align__(16) unsigned thr2[PK_PITCH];   // packed gate: min(thr, 32767) of two partitions per word (c_pk_half tells which half)
  int nbig;                            // active partitions among those whose thr is still above 32767 (the packed gate stands back until 0)
  int R, max_mvd_m1;
};

// Packed gate: the partitions of up to 8x16 samples in pairs (low half, high half), in the order the packed sums come out of
// the sixteen 4x4 SADs: E_j = (a[4j], a[4j+2]), O_j = (a[4j+1], a[4j+3]), 8x4 rows E_j + O_j, 4x8 columns E_0 + E_1 ..., 8x8
// pairs, the 8x16 pair.  c_pk_half[p] = 2 * word + half for partition p (the three larger partitions are compared as scalars).
__constant__ signed char c_pk_half[NPART] = {
  -1, -1, -1,                  // 16x16, 16x8 x 2: scalars
  36, 37,                      // 8x16: word 18
  32, 33, 34, 35,              // 8x8: words 16, 17
  16, 17, 18, 19, 20, 21, 22, 23,      // 8x4 (9..16): words 8..11 = rows 0..3, (left, right)
  24, 26, 25, 27, 28, 30, 29, 31,      // 4x8 (17..24): words 12 (s6[0], s6[2]), 13 (s6[1], s6[3]), 14 (s6[4], s6[6]), 15 (s6[5], s6[7])
  0, 2, 1, 3, 4, 6, 5, 7, 8, 10, 9, 11, 12, 14, 13, 15};      // 4x4 (25..40): words 2j = E_j (a[4j], a[4j+2]), 2j+1 = O_j (a[4j+1], a[4j+3])

// lowers partition p's half of its packed word to min(thr, 32767) (never raises it: thr only falls)
__device__ __forceinline__ void pk_set_thr(Grp *g, int p, unsigned thr) {
  const int h = c_pk_half[p];
  if (h < 0) return;
  const unsigned sh = (h & 1) * 16, val = min(thr, 32767u);
  unsigned *w = &g->thr2[h >> 1], old = *(volatile unsigned *)w;
  for (;;) {
    if (((old >> sh) & 0xffffu) <= val) return;
    const unsigned seen = atomicCAS(w, old, (old & ~(0xffffu << sh)) | (val << sh));
    if (seen == old) return;
    old = seen;
  }
}

// Largest SAD that can still win against the key `k`: a winner needs (sad << 5) + lambda * bits <= cost(k),
// and every candidate pays at least 2 bits (mvbits >= 1 per component).  The gate is sad < bound.
__device__ __forceinline__ unsigned bound_of(unsigned long long k, int lam) {
  const unsigned long long cost = k >> IDX_BITS, floor_ = 2ull * (unsigned)lam;
  if (cost < floor_) return 0u;
  return (unsigned)min(((cost - floor_) >> 5) + 1, 0x7fffffffull);   // compared as signed int after the column adjustment
}

// ---- TMA (cp.async.bulk.tensor) + mbarrier plumbing ---------------------------------------------------------------
struct TMaps { CUtensorMap cur; CUtensorMap ref[JMB_MAX_REFS]; };   // kernel parameter (__grid_constant__)
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *b, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *b, unsigned parity) {
  unsigned ok;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
  } while (!ok);
}
// one tile of a 2-D u8 tensor -> shared memory; coordinates may lie outside the tensor (zero fill)
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int x, int y, unsigned long long *bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(smem_u32(dst)), "l"((unsigned long long)map), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ unsigned sad4(unsigned a, unsigned b, unsigned c) {
  unsigned d;
  asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

__device__ __forceinline__ void publish(Grp *g, int p, unsigned long long k) {
  if (k < atomicMin(&g->best[p], k)) {
    const unsigned b = bound_of(k, g->rq[p].lam), old = atomicMin(&g->thr[p], b);
#if JMB_IS_PACKED
    if (b < old) {      // the packed copy follows; the first to bring thr under 2^15 counts it off
      pk_set_thr(g, p, *(volatile unsigned *)&g->thr[p]);
      if (old > 32767u && b <= 32767u && c_pk_half[p] >= 0) atomicSub(&g->nbig, 1);
    }
#endif
  }
}

// Exact cost evaluation of the candidates that read the block(s) at displacement (Dx,Dy).  Reached only
// when the SAD passed the gate (rare), so it is kept out of line.  Interior displacements stand for
// exactly one candidate; a displacement ON the clamp boundary stands for every candidate beyond it.
__device__ __noinline__ void eval_slow(Grp *g, int p, unsigned sad, int Dx, int Dy) {
  const ReqS *q = &g->rq[p];
  const int R = g->R, max_mvd_m1 = g->max_mvd_m1;
  if (Dx < q->dlo_x || Dx > q->dhi_x || Dy < q->dlo_y || Dy > q->dhi_y) return;
  int x0 = Dx, x1 = Dx, y0 = Dy, y1 = Dy;
  if (Dx == q->dlo_x) x0 = q->cx - R;     // every d <= Dlo clamps to Dlo
  if (Dx == q->dhi_x) x1 = q->cx + R;
  if (Dy == q->dlo_y) y0 = q->cy - R;
  if (Dy == q->dhi_y) y1 = q->cy + R;
  x0 = max(x0, q->cx - R); x1 = min(x1, q->cx + R);
  y0 = max(y0, q->cy - R); y1 = min(y1, q->cy + R);
  unsigned long long k = ~0ull;
  for (int dy = y0; dy <= y1; dy++)
    for (int dx = x0; dx <= x1; dx++) {
      int mx = 4 * dx - q->px, my = 4 * dy - q->py;
      if (q->ffs && max(abs(mx), abs(my)) >= max_mvd_m1) continue;
      unsigned long long cost = ((unsigned long long)sad << 5) + (unsigned long long)((long long)q->lam * (jmb_mvbits(mx) + jmb_mvbits(my)));
      k = min(k, (cost << IDX_BITS) | (unsigned)jmb_spiral_index(dx - q->cx, dy - q->cy));
    }
  if (k < *(volatile unsigned long long *)&g->best[p]) publish(g, p, k);
}

// A SAD passed the gate: exact cost of the one candidate an interior displacement stands for; everything on a
// clamp boundary or outside the partition's own window goes to eval_slow.
__device__ __noinline__ void level2(Grp *g, int p, unsigned sad, int Dx, int Dy) {
  const int4 in = g->inner[p];
  if (Dx < in.x || Dx > in.y || Dy < in.z || Dy > in.w) { eval_slow(g, p, sad, Dx, Dy); return; }
  const ReqS *q = &g->rq[p];
  const int mx = 4 * Dx - q->px, my = 4 * Dy - q->py;
  if (q->ffs && max(abs(mx), abs(my)) >= g->max_mvd_m1) return;
  const unsigned long long cost = ((unsigned long long)sad << 5) + (unsigned long long)((long long)q->lam * (jmb_mvbits(mx) + jmb_mvbits(my)));
  const unsigned long long cur = *(volatile unsigned long long *)&g->best[p];
  if (cost > (cur >> IDX_BITS)) return;
  const unsigned long long k = (cost << IDX_BITS) | (unsigned)jmb_spiral_index(Dx - q->cx, Dy - q->cy);
  if (k < cur) publish(g, p, k);
}

// inverse of jmb_spiral_index
__device__ void spiral_xy(int idx, int *dx, int *dy) {
  if (idx == 0) { *dx = 0; *dy = 0; return; }
  int s = (int)sqrtf((float)idx);
  while (s * s > idx) s--;
  while ((s + 1) * (s + 1) <= idx) s++;
  int l = (s + 1) / 2, base = (2 * l - 1) * (2 * l - 1), off = idx - base;
  if (off < 2 * (2 * l - 1)) { *dx = (off >> 1) - l + 1; *dy = (off & 1) ? l : -l; }
  else { off -= 2 * (2 * l - 1); *dy = (off >> 1) - l; *dx = (off & 1) ? l : -l; }
}

// compile-time copy of the partition geometry (bx, by, w4, h4 in 4x4 units) for part_sum<P>
__host__ __device__ constexpr int pg_first(int t) { return t == 1 ? 0 : t == 2 ? 1 : t == 3 ? 3 : t == 4 ? 5 : t == 5 ? 9 : t == 6 ? 17 : 25; }
__host__ __device__ constexpr int pg_type(int p) { return p < 1 ? 1 : p < 3 ? 2 : p < 5 ? 3 : p < 9 ? 4 : p < 17 ? 5 : p < 25 ? 6 : 7; }
__host__ __device__ constexpr int pg_w4(int p) { return pg_type(p) <= 2 ? 4 : pg_type(p) <= 5 ? 2 : 1; }
__host__ __device__ constexpr int pg_h4(int p) { return (pg_type(p) == 1 || pg_type(p) == 3) ? 4 : (pg_type(p) == 2 || pg_type(p) == 4 || pg_type(p) == 6) ? 2 : 1; }
__host__ __device__ constexpr int pg_bx(int p) { return ((p - pg_first(pg_type(p))) % (4 / pg_w4(p))) * pg_w4(p); }
__host__ __device__ constexpr int pg_by(int p) { return ((p - pg_first(pg_type(p))) / (4 / pg_w4(p))) * pg_h4(p); }
template <int P>
__device__ __forceinline__ unsigned part_sum(const unsigned *a) {
  unsigned v = 0;
#pragma unroll
  for (int y = 0; y < pg_h4(P); y++)
#pragma unroll
    for (int x = 0; x < pg_w4(P); x++) v += a[(pg_by(P) + y) * 4 + pg_bx(P) + x];
  return v;
}

// the 41 partition SADs of one displacement from its sixteen 4x4 SADs (update_full_search_large_blocks,
// lencod/src/me_fullfast.c:207-259), handed to F(partition, sad) in canonical order
template <typename F>
__device__ __forceinline__ void for_each_partition(const unsigned *a, F &&f) {
  unsigned s6[8], s5[8], s4[4];
#pragma unroll
  for (int i = 0; i < 4; i++) { s6[i] = a[i] + a[4 + i]; s6[4 + i] = a[8 + i] + a[12 + i]; }
#pragma unroll
  for (int i = 0; i < 8; i++) s5[i] = a[2 * i] + a[2 * i + 1];
  s4[0] = s6[0] + s6[1]; s4[1] = s6[2] + s6[3]; s4[2] = s6[4] + s6[5]; s4[3] = s6[6] + s6[7];
  const unsigned s3a = s4[0] + s4[2], s3b = s4[1] + s4[3], s2a = s4[0] + s4[1], s2b = s4[2] + s4[3];
  f(0, s2a + s2b);
  f(1, s2a); f(2, s2b);
  f(3, s3a); f(4, s3b);
#pragma unroll
  for (int i = 0; i < 4; i++) f(5 + i, s4[i]);
#pragma unroll
  for (int i = 0; i < 8; i++) f(9 + i, s5[i]);      // 8x4: (bx 0|2, by 0..3) raster = pairs (2i, 2i+1)
#pragma unroll
  for (int i = 0; i < 8; i++) f(17 + i, s6[i]);     // 4x8: (bx 0..3, by 0|2)
#pragma unroll
  for (int i = 0; i < 16; i++) f(25 + i, a[i]);
}

// Sixty-four 4x4 SADs of one thread item: window column xo (bytes from the staged row start), displacement
// rows row0 .. row0+3.  Window row row0+wr is byte-aligned once (4 PRMT) and meets source row wr-dy for each dy.
__device__ __forceinline__ void sad_item(const uint8_t *win, const unsigned *ssrc, int row0, int xo, unsigned (&acc)[4][16]) {
  const unsigned sel = 0x3210u + 0x1111u * (xo & 3);
  const unsigned *wrow = (const unsigned *)(win + row0 * WIN_PITCH + (xo & ~3));
#pragma unroll
  for (int s = 0; s < 4; s++)
#pragma unroll
    for (int b = 0; b < 16; b++) acc[s][b] = 0;
  unsigned srow[4][4];
#pragma unroll
  for (int wr = 0; wr < 19; wr++) {
    if (wr < 16) {
      const uint4 v = *(const uint4 *)&ssrc[wr * 4];
      srow[wr & 3][0] = v.x; srow[wr & 3][1] = v.y; srow[wr & 3][2] = v.z; srow[wr & 3][3] = v.w;
    }
    const unsigned *wp = wrow + wr * (WIN_PITCH / 4);
    const unsigned w0 = wp[0], w1 = wp[1], w2 = wp[2], w3 = wp[3], w4 = wp[4];
    unsigned al[4];
    al[0] = __byte_perm(w0, w1, sel); al[1] = __byte_perm(w1, w2, sel);
    al[2] = __byte_perm(w2, w3, sel); al[3] = __byte_perm(w3, w4, sel);
#pragma unroll
    for (int dy = 0; dy < 4; dy++) {
      const int r = wr - dy;                 // source row matched with this window row at displacement row0+dy
      if (r >= 0 && r < 16) {
#pragma unroll
        for (int c = 0; c < 4; c++) acc[dy][(r >> 2) * 4 + c] = sad4(al[c], srow[r & 3][c], acc[dy][(r >> 2) * 4 + c]);
      }
    }
  }
}

// groups == nullptr: frame layout, group g = requests [41g, 41g+41) in canonical partition order.
//
// Work decomposition: a thread owns ONE column of displacements and FOUR vertically adjacent rows, so the
// byte alignment (PRMT) of a window row is done once and feeds four displacements; its sixty-four 4x4 SADs
// live in registers.  Winners are tracked per partition in shared memory; a thread only leaves the
// arithmetic loop when one of its SADs beats the current bound (sad < thr), which after the seeding step
// below is rare.  ALU-pipe work per displacement: 64 VABSDIFF4 + 19 PRMT + 41 ISETP.
#ifndef JMB_IS_MINB
#define JMB_IS_MINB 3      // 3 CTAs of 128 threads: 167 registers, nothing spilled (4 CTAs cap the kernel at 128 registers: 0.72 -> 0.70 ms)
#endif
__global__ void __launch_bounds__(NT, JMB_IS_MINB)
k_int_search(const jmb_me_req *__restrict__ reqs, const int *__restrict__ groups, jmb_me_res *__restrict__ res,
             const __grid_constant__ TMaps tm, int w, int h, int R, int max_mvd_m1, int nref, int fpel_metric, int *__restrict__ err,
             const jmb_frame_gen gen) {
  __shared__ Grp G;
  __shared__ __align__(128) uint8_t win[WIN_ROWS * WIN_PITCH];   // TMA destination: the search window tile
  __shared__ __align__(128) unsigned ssrc[16 * 4];               // TMA destination: the 16x16 source macroblock
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ int sbox[12];
  __shared__ int wq_n[NT / 32];
  // gate tables (dynamic shared memory, INT_SEARCH_DYN_SMEM bytes): per column / per row and partition,
  // floor(lambda * (bits_x - 1) / 32) and floor(lambda * (bits_y - 1) / 32)
  extern __shared__ __align__(16) unsigned dyn_smem[];
  unsigned *const adjx = dyn_smem, *const adjy4 = dyn_smem + CW * ADJ_PITCH;   // adjy4: min over the 4 rows of an item
  unsigned *const s1x = adjy4 + (CH / 4) * ADJ_PITCH, *const s1y = s1x + S1_COLS * ADJ_PITCH;   // lambda * bits of stage 1's columns / rows
  unsigned long long (*const wq)[WQ_CAP] = (unsigned long long (*)[WQ_CAP])(s1y + 4 * S1_RGS * ADJ_PITCH);
  unsigned short *const S1 = (unsigned short *)wq;      // stage 1 is over (barrier) before the sweep touches wq / hitsad
  unsigned short (*const hitsad)[HS_PITCH] = (unsigned short (*)[HS_PITCH])(wq + NT / 32);
  unsigned *const adjx2 = (unsigned *)((char *)wq + (S1_BYTES > SWEEP_BYTES ? S1_BYTES : SWEEP_BYTES)), *const adjy2 = adjx2 + CW * PK_PITCH;

  const int tid = threadIdx.x, g = blockIdx.x;
  if (tid < NT / 32) wq_n[tid] = 0;
  if (tid == 0) {
    sbox[0] = sbox[2] = 1 << 30; sbox[1] = sbox[3] = -(1 << 30); sbox[7] = 0; sbox[10] = NPART;
    mbar_init(&mbar, 1);
    G.R = R; G.max_mvd_m1 = max_mvd_m1; G.thr[NPART] = G.thr[NPART + 1] = G.thr[NPART + 2] = 0;
    G.nbig = 0;
    for (int i = 0; i < PK_PITCH; i++) G.thr2[i] = 0xffffffffu;      // (every partition lowers its half below; the spare word is never read)
  }
  __syncthreads();
  if (tid < NPART) {
    int ri = groups ? groups[g * NPART + tid] : g * NPART + tid;
    ReqS q; q.active = 0; q.req = ri;
    if (ri >= 0) {
      jmb_me_req r = gen.pred ? jmb_frame_request(gen, g, tid) : reqs[ri];      // picture form: formed here, never in memory
      int bad = jmb_req_check(r, w, h, nref);
      if (!bad && !groups && !gen.pred) {      // frame layout: request k of a group must be partition k of one macroblock and reference
        const PartGeom pg = c_part[tid];
        const jmb_me_req r0 = reqs[g * NPART];
        if (r.blocktype != pg.type || (r.pos_x & 15) != pg.bx * 4 || (r.pos_y & 15) != pg.by * 4 || ((r.pos_x ^ r0.pos_x) & ~15) ||
            ((r.pos_y ^ r0.pos_y) & ~15) || r.ref != r0.ref) bad = JMB_REQERR_LAYOUT;
      }
      // this kernel is a SAD kernel: full_search_motion_estimation takes computePredFPel = the MEDistortionFPel metric, and
      // setup_fast_full_search builds squared-error surfaces for any metric other than SAD (dist_method, me_fullfast.c:274)
      if (!bad && !(r.flags & JMB_REQ_SKIP_INT) && fpel_metric != JMB_SAD) bad = JMB_REQERR_FPEL_METRIC;
      jmb_req_report(err, bad, ri);
      if (!bad && !(r.flags & JMB_REQ_SKIP_INT)) {
        q.active = 1;
        q.ffs = (r.mode == JMB_SEARCH_FAST_FULL);
        int ox = q.ffs ? (r.pos_x & ~15) : r.pos_x, oy = q.ffs ? (r.pos_y & ~15) : r.pos_y;   // clamped origin
        q.dlo_x = -JMB_PAD_X - ox; q.dhi_x = (w + JMB_PAD_X - 1 - 16) - ox;      // UMVLine4X, refbuf.h:25
        q.dlo_y = -JMB_PAD_Y - oy; q.dhi_y = (h + JMB_PAD_Y - 1 - 16) - oy;
        q.cx = r.center_x >> 2; q.cy = r.center_y >> 2;
        q.px = r.pred_x; q.py = r.pred_y; q.lam = r.lambda[0];
        q.init = (unsigned long long)r.min_mcost << IDX_BITS;
        // union of the (clamped) displacement ranges of the group's requests
        atomicMin(&sbox[0], jmb_clip(q.dlo_x, q.dhi_x, q.cx - R)); atomicMax(&sbox[1], jmb_clip(q.dlo_x, q.dhi_x, q.cx + R));
        atomicMin(&sbox[2], jmb_clip(q.dlo_y, q.dhi_y, q.cy - R)); atomicMax(&sbox[3], jmb_clip(q.dlo_y, q.dhi_y, q.cy + R));
        atomicMin(&sbox[10], tid);
        sbox[4] = r.pos_x & ~15; sbox[5] = r.pos_y & ~15; sbox[6] = r.ref; sbox[7] = 1;   // same for every request of a group
      }
    }
    G.rq[tid] = q;
    G.best[tid] = q.active ? q.init : 0ull;
    G.thr[tid] = q.active ? bound_of(q.init, q.lam) : 0u;
#if JMB_IS_PACKED
    pk_set_thr(&G, tid, G.thr[tid]);
    if (q.active && c_pk_half[tid] >= 0 && G.thr[tid] > 32767u) atomicAdd(&G.nbig, 1);
#endif
    int4 in = make_int4(1, 0, 1, 0);
    if (q.active) in = make_int4(max(q.dlo_x + 1, q.cx - R), min(q.dhi_x - 1, q.cx + R), max(q.dlo_y + 1, q.cy - R), min(q.dhi_y - 1, q.cy + R));
    G.inner[tid] = in;
  }
  __syncthreads();
  if (!sbox[7]) return;   // nothing but sub-pel-only requests in this group
  const int bx0 = sbox[0], bx1 = sbox[1], by0 = sbox[2], by1 = sbox[3], mbx = sbox[4], mby = sbox[5];
  const CUtensorMap *ref_map = &tm.ref[sbox[6]];

  bool seeded = false;
  unsigned phase = 0;
  for (int cy0 = by0; cy0 <= by1; cy0 += CH) {
    const int ch = min(CH, by1 - cy0 + 1);
    for (int cx0 = bx0; cx0 <= bx1; cx0 += CW) {
      const int cw = min(CW, bx1 - cx0 + 1);
      __syncthreads();
      // Stage the search window with ONE 2-D TMA tile load.  TMA wants the innermost start coordinate on a 16-byte
      // boundary, so the tile starts xoff0 (0..15) samples left of the window: staged byte (r, c) = plane sample
      // (mby + cy0 + r + PAD_Y, mbx + cx0 + PAD_X - xoff0 + c); samples outside the padded plane arrive as zeros (no
      // valid candidate reads them).  The first chunk also fetches the 16x16 source macroblock.  The copy runs while the
      // threads build the gate tables below; everybody meets at the mbarrier before the first SAD.
      const int ax = mbx + cx0 + JMB_PAD_X, xoff0 = ax & 15;
      if (tid == 0) {
        mbar_expect_tx(&mbar, WIN_ROWS * WIN_PITCH + (seeded ? 0 : 256));
        tma_load_2d(win, ref_map, ax - xoff0, mby + cy0 + JMB_PAD_Y, &mbar);
        if (!seeded) tma_load_2d(ssrc, &tm.cur, mbx, mby, &mbar);
      }
      const int nrg = (ch + 3) >> 2, mid = nrg >> 1;
      const bool s1 = !seeded;
      seeded = true;
      // Stage 1 (first chunk only): exact, gate-free evaluation of a small neighbourhood of the search centre
      // (S1_COLS columns x 4*S1_RGS rows) so that the bounds are tight before the sweep starts.  Phase a: warp 0,
      // one thread per item, computes the SADs and leaves the 41 partition sums of its 4 displacements in shared
      // memory -- while the other warps build the column table of the gate.  Phase b: three threads per partition
      // scan the sums with the exact cost and publish the winner.
      const int ncol = min(S1_COLS, cw), nrgs = min(S1_RGS, nrg);
      const int col0 = jmb_clip(0, cw - ncol, G.rq[sbox[10]].cx - cx0 - ncol / 2);
      const int rg0 = jmb_clip(0, nrg - nrgs, ((G.rq[sbox[10]].cy - cy0) >> 2) - nrgs / 2);
      if (s1 && tid < 32) {
        mbar_wait(&mbar, phase);
        if (tid < ncol * nrgs) {
          const int ic = col0 + tid % ncol, rg = rg0 + tid / ncol;
          unsigned acc[4][16];
          sad_item(win, ssrc, rg * 4, ic + xoff0, acc);
#pragma unroll
          for (int s = 0; s < 4; s++) {
            unsigned short *o = S1 + (tid * 4 + s) * S1_PITCH;
            for_each_partition(acc[s], [&](int p, unsigned v) { o[p] = (unsigned short)v; });
          }
        }
      } else {
        // Gate tables.  Column part: every candidate in column Dx pays at least lambda * (bits_x(Dx) + 1), i.e.
        // lambda * (bits_x - 1) more than the constant folded into thr; row part likewise, minimised over the 4 rows
        // of an item.  A column / row ON the clamp boundary stands for every candidate beyond it: no term there.
        // A thread keeps one partition and strides over the columns / row groups.
        const int t0 = s1 ? tid - 32 : tid, nthr = s1 ? NT - 32 : NT, lanes_per_p = nthr / NPART;   // nthr >= 41
        const int p = t0 % NPART, sub = t0 / NPART;
        if (sub < lanes_per_p) {
          const ReqS &q = G.rq[p];
          const int4 in = G.inner[p];
          const unsigned lam = (unsigned)q.lam;
          // (an inactive request has an empty `inner`, so its terms are all zero)
          for (int ic = sub; ic < cw; ic += lanes_per_p) {
            const int Dx = cx0 + ic;
            unsigned a = min(65535u, (lam * (unsigned)(jmb_mvbits(4 * Dx - q.px) - 1)) >> 5);
            if (Dx < in.x || Dx > in.y) a = 0;
            adjx[ic * ADJ_PITCH + p] = a;
#if JMB_IS_PACKED
            if (c_pk_half[p] >= 0) ((unsigned short *)adjx2)[ic * PK_PITCH * 2 + c_pk_half[p]] = (unsigned short)min(a, PK_ADJ_MAX);
#endif
          }
          // the row term is monotone in |4 Dy - py|: of the (up to) 4 rows of an item the one nearest py / 4 has the minimum
          for (int rg = sub; rg < ((ch + 3) >> 2); rg += lanes_per_p) {
            const int D0 = cy0 + 4 * rg, D1 = cy0 + min(4 * rg + 3, ch - 1), Dn = q.py >> 2;
            const int am = min(abs(4 * jmb_clip(D0, D1, Dn) - q.py), abs(4 * jmb_clip(D0, D1, Dn + 1) - q.py));
            unsigned a = min(65535u, (lam * (unsigned)(jmb_mvbits(am) - 1)) >> 5);
            if (D0 < in.z || D1 > in.w) a = 0;
            adjy4[rg * ADJ_PITCH + p] = a;
#if JMB_IS_PACKED
            if (c_pk_half[p] >= 0) ((unsigned short *)adjy2)[rg * PK_PITCH * 2 + c_pk_half[p]] = (unsigned short)min(a, PK_ADJ_MAX);
#endif
          }
          if (s1) {      // stage 1's exact terms; S1_BAD marks a column / row that is not a plain candidate of this partition
            for (int c = sub; c < ncol; c += lanes_per_p) {
              const int Dx = cx0 + col0 + c, mx = 4 * Dx - q.px;
              const bool ok = q.active && Dx >= in.x && Dx <= in.y && !(q.ffs && abs(mx) >= max_mvd_m1);
              s1x[c * ADJ_PITCH + p] = ok ? lam * (unsigned)jmb_mvbits(mx) : S1_BAD;
            }
            for (int r = sub; r < 4 * nrgs; r += lanes_per_p) {
              const int Dy = cy0 + rg0 * 4 + r, my = 4 * Dy - q.py;
              const bool ok = q.active && Dy >= in.z && Dy <= in.w && Dy < cy0 + ch && !(q.ffs && abs(my) >= max_mvd_m1);
              s1y[r * ADJ_PITCH + p] = ok ? lam * (unsigned)jmb_mvbits(my) : S1_BAD;
            }
          }
        }
      }
      if (tid == 0) sbox[11] = 0;      // next warp item of the sweep
      mbar_wait(&mbar, phase);         // the window (and source) tiles have landed
      phase ^= 1;
      __syncthreads();
      if (s1) {
        // three threads per partition (41 x 3 <= NT): thread j scans displacement rows j, j+3, j+6 of the neighbourhood;
        // the spiral index is only worked out on a tie; each thread publishes its own winner (shared-memory atomics)
        const int p = tid / 3, j = tid - 3 * p;
        if (p < NPART && G.rq[p].active) {
          const ReqS &q = G.rq[p];
          unsigned bcost = S1_NONE; int bc = 0, bDy = 0, bidx = -1;
          for (int r = j; r < 4 * nrgs; r += 3) {
            const unsigned ycost = s1y[r * ADJ_PITCH + p];
            if (ycost == S1_BAD) continue;
            const int Dy = cy0 + rg0 * 4 + r;
            const unsigned short *sp = S1 + (((r >> 2) * ncol) * 4 + (r & 3)) * S1_PITCH + p;
            for (int c = 0; c < ncol; c++) {
              const unsigned cost = ((unsigned)sp[c * 4 * S1_PITCH] << 5) + ycost + s1x[c * ADJ_PITCH + p];
              if (cost > bcost) continue;
              if (cost == bcost) {
                if (bidx < 0) bidx = jmb_spiral_index(cx0 + col0 + bc - q.cx, bDy - q.cy);
                const int idx = jmb_spiral_index(cx0 + col0 + c - q.cx, Dy - q.cy);
                if (idx >= bidx) continue;
                bidx = idx;
              } else bidx = -1;
              bcost = cost; bc = c; bDy = Dy;
            }
          }
          if (bcost != S1_NONE) {
            if (bidx < 0) bidx = jmb_spiral_index(cx0 + col0 + bc - q.cx, bDy - q.cy);
            const unsigned long long k = ((unsigned long long)bcost << IDX_BITS) | (unsigned)bidx;
            if (k < *(volatile unsigned long long *)&G.best[p]) publish(&G, p, k);
          }
        }
        __syncthreads();
      }

      // the sweep: warps draw batches of 32 items from a shared counter (a batch that meets the gate runs much
      // longer than one that does not, so a static split would leave warps waiting at the final barrier)
      const int nitems = nrg * cw, lane = tid & 31;
      const unsigned cw_rcp = cw > 1 ? 0xffffffffu / (unsigned)cw + 1 : 0u;      // it / cw == umulhi(it, cw_rcp) for it, cw < 2^16 (cw > 1)
      for (;;) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&sbox[11], 32);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= nitems) break;
        const int it = base + lane, warp = tid >> 5;
        if (it < nitems) {
        const int k = cw > 1 ? (int)__umulhi((unsigned)it, cw_rcp) : it, ic = it - k * cw;
        const int rg = (k & 1) ? mid - ((k + 1) >> 1) : mid + (k >> 1);   // centre rows first: tight bounds early
        const int row0 = rg * 4, xo = ic + xoff0;
        unsigned acc[4][16];
        sad_item(win, ssrc, row0, xo, acc);
#if JMB_IS_PACKED
        // Packed gate.  A partition meets the gate when sad + (column term + row-group term) - thr < 0.  For the 38 partitions of at
        // most 8x16 samples that is done two per register: N = adjx2 + adjy2 - thr2 once per item, then per displacement the
        // packed sums (8 multiply-adds to pair up the 4x4 SADs, 11 adds for the larger shapes instead of 22), one add of N each
        // OR-ed into one word (LOP3 takes two at a time); a set sign bit in either half is a hit.  The words are
        // plain integers lo + 65536 hi: a negative low field borrows 1 from the high one, which can only turn a high field
        // negative when the low one already is a hit -- never the other way round.  Terms are capped at PK_ADJ_MAX and thr at
        // 32767 so that no field wraps (a cap only loosens the gate); while some thr is still above 32767 (nbig) every
        // displacement goes to the exact re-visit.  The three large partitions are compared as before.
        int t0, t1, t2;
        {
          unsigned th[4];
          asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(th[0]), "=r"(th[1]), "=r"(th[2]), "=r"(th[3]) : "r"((unsigned)__cvta_generic_to_shared(&G.thr[0])));
          const uint4 v = *(const uint4 *)&adjx[ic * ADJ_PITCH], u = *(const uint4 *)&adjy4[rg * ADJ_PITCH];
          t0 = (int)(th[0] - v.x - u.x); t1 = (int)(th[1] - v.y - u.y); t2 = (int)(th[2] - v.z - u.z);
        }
        unsigned N[PK_PITCH];
#pragma unroll
        for (int i = 0; i < PK_PITCH / 4; i++) {
          unsigned th[4];
          asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(th[0]), "=r"(th[1]), "=r"(th[2]), "=r"(th[3]) : "r"((unsigned)__cvta_generic_to_shared(&G.thr2[4 * i])));
          const uint4 v = *(const uint4 *)&adjx2[ic * PK_PITCH + 4 * i], u = *(const uint4 *)&adjy2[rg * PK_PITCH + 4 * i];
          N[4 * i] = v.x + u.x - th[0]; N[4 * i + 1] = v.y + u.y - th[1]; N[4 * i + 2] = v.z + u.z - th[2]; N[4 * i + 3] = v.w + u.w - th[3];
        }
        const bool big = *(volatile int *)&G.nbig != 0;
        bool hit[4];
#pragma unroll
        for (int s = 0; s < 4; s++) {
          hit[s] = false;
          if (row0 + s < ch) {
            const unsigned *a = acc[s];
            unsigned w[PK_N];
#pragma unroll
            for (int j = 0; j < 4; j++) { w[2 * j] = a[4 * j] + a[4 * j + 2] * 65536u; w[2 * j + 1] = a[4 * j + 1] + a[4 * j + 3] * 65536u; }
#pragma unroll
            for (int j = 0; j < 4; j++) w[8 + j] = w[2 * j] + w[2 * j + 1];                  // 8x4 rows
            w[12] = w[0] + w[2]; w[13] = w[1] + w[3]; w[14] = w[4] + w[6]; w[15] = w[5] + w[7];   // 4x8 columns
            w[16] = w[12] + w[13]; w[17] = w[14] + w[15];                                        // 8x8
            w[18] = w[16] + w[17];                                                               // 8x16
            const unsigned s2a = __dp2a_lo(w[16], 0x0101u, 0u), s2b = __dp2a_lo(w[17], 0x0101u, 0u);   // 16x8: the two halves of an 8x8 pair added up
            unsigned m = w[0] + N[0];      // the sign bits of all halves OR-ed together (one LOP3 per two words)
#pragma unroll
            for (int i = 1; i < PK_N; i += 2) m |= (w[i] + N[i]) | (w[i + 1] + N[i + 1]);
            hit[s] = big | ((m & 0x80008000u) != 0u) | ((int)(s2a + s2b) < t0) | ((int)s2a < t1) | ((int)s2b < t2);
          }
        }
#else
        // gate: does any of the 4 x 41 partition SADs beat its current bound?
        unsigned t[NPART + 3];
#pragma unroll
        for (int i = 0; i < (NPART + 3) / 4; i++) {
          asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                       : "=r"(t[4 * i]), "=r"(t[4 * i + 1]), "=r"(t[4 * i + 2]), "=r"(t[4 * i + 3])
                       : "r"((unsigned)__cvta_generic_to_shared(&G.thr[4 * i])));
        }
        {
          const uint4 *ap = (const uint4 *)&adjx[ic * ADJ_PITCH], *bp = (const uint4 *)&adjy4[rg * ADJ_PITCH];
#pragma unroll
          for (int i = 0; i < (NPART + 3) / 4; i++) {
            const uint4 v = ap[i], u = bp[i];
            t[4 * i] -= v.x + u.x; t[4 * i + 1] -= v.y + u.y; t[4 * i + 2] -= v.z + u.z; t[4 * i + 3] -= v.w + u.w;
          }
        }
        // one flag per displacement keeps the common case to 41 compares
        bool hit[4];
#pragma unroll
        for (int s = 0; s < 4; s++) {
          hit[s] = false;
          if (row0 + s < ch) for_each_partition(acc[s], [&](int p, unsigned v) { hit[s] |= (int)v < (int)t[p]; });
        }
#endif
        // Re-visit of a displacement that met the gate: collect the partitions that did in a bit mask (straight-line,
        // no calls), then walk the set bits: one jump per hit to the few adds that re-sum that partition's SAD, one more
        // check with the row term of the mv cost, and only then the exact evaluation.  Cost is proportional to the number
        // of hits, and only the code of partitions that hit is ever fetched (the flat-SAD regime lives here).
        const int Dx = cx0 + ic;
#pragma unroll
        for (int s = 0; s < 4; s++) {
          if (!hit[s]) continue;
          unsigned mlo = 0, mhi = 0;
#if JMB_IS_PACKED
          unsigned t[NPART + 3];      // the exact gate, partition by partition (the packed one above only says "look here")
#pragma unroll
          for (int i = 0; i < (NPART + 3) / 4; i++) {
            asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(t[4 * i]), "=r"(t[4 * i + 1]), "=r"(t[4 * i + 2]), "=r"(t[4 * i + 3])
                         : "r"((unsigned)__cvta_generic_to_shared(&G.thr[4 * i])));
            const uint4 v = *(const uint4 *)&adjx[ic * ADJ_PITCH + 4 * i], u = *(const uint4 *)&adjy4[rg * ADJ_PITCH + 4 * i];
            t[4 * i] -= v.x + u.x; t[4 * i + 1] -= v.y + u.y; t[4 * i + 2] -= v.z + u.z; t[4 * i + 3] -= v.w + u.w;
          }
          for_each_partition(acc[s], [&](int p, unsigned v) {
            if ((int)v < (int)t[p]) { if (p < 32) mlo |= 1u << (p & 31); else mhi |= 1u << (p & 31); hitsad[tid][p] = (unsigned short)v; }
          });
#else
          for_each_partition(acc[s], [&](int p, unsigned v) {
            if ((int)v < (int)t[p]) { if (p < 32) mlo |= 1u << (p & 31); else mhi |= 1u << (p & 31); hitsad[tid][p] = (unsigned short)v; }
          });
#endif
          const int Dy = cy0 + row0 + s;
          const unsigned *ax = adjx + ic * ADJ_PITCH;
          while (mlo | mhi) {
            int p;
            if (mlo) { p = __ffs(mlo) - 1; mlo &= mlo - 1; } else { p = 32 + __ffs(mhi) - 1; mhi &= mhi - 1; }
            const unsigned v = hitsad[tid][p];
            // second look with the exact row term of the mv cost (none on a clamp-boundary row)
            const ReqS &q = G.rq[p];
            unsigned ay = 0;
            if (Dy >= G.inner[p].z && Dy <= G.inner[p].w) ay = min(65535u, ((unsigned)q.lam * (unsigned)(jmb_mvbits(4 * Dy - q.py) - 1)) >> 5);
            const int bound = (int)(*(volatile unsigned *)&G.thr[p] - ax[p] - ay);
            if ((int)v < bound) {
              // hits are parked in the warp's queue and evaluated below by all 32 lanes together (a lone lane walking
              // its own hits would hold the other 31 idle); a full queue falls back to evaluating in place
              const int slot = atomicAdd(&wq_n[warp], 1);
              if (slot < WQ_CAP) wq[warp][slot] = ((unsigned long long)(unsigned)((p << 16) | ((Dy - cy0) << 8) | (Dx - cx0)) << 32) | v;
              else level2(&G, p, v, Dx, Dy);
            }
          }
        }
        }
        __syncwarp();
        const int nq = min(*(volatile int *)&wq_n[warp], WQ_CAP);
        for (int e = lane; e < nq; e += 32) {
          const unsigned long long ent = wq[warp][e];
          const unsigned hi = (unsigned)(ent >> 32);
          level2(&G, (int)(hi >> 16), (unsigned)ent, cx0 + (int)(hi & 255), cy0 + (int)((hi >> 8) & 255));
        }
        __syncwarp();
        if (lane == 0) wq_n[warp] = 0;
        __syncwarp();
      }
    }
  }
  __syncthreads();
  if (tid < NPART && G.rq[tid].active) {
    const ReqS &q = G.rq[tid];
    unsigned long long k = G.best[tid];
    int dx = 0, dy = 0;
    long long cost = (long long)(k >> IDX_BITS);
    if (k != q.init) spiral_xy((int)(k & ((1u << IDX_BITS) - 1)), &dx, &dy);
    jmb_me_res o;
    o.imv_x = o.mv_x = (int16_t)(4 * (q.cx + dx));
    o.imv_y = o.mv_y = (int16_t)(4 * (q.cy + dy));
    o.icost = o.cost = cost;
    res[q.req] = o;
  }
}

// BlockSAD surfaces of one macroblock in JM's layout and spiral order (jmb_ffs_surfaces).
__global__ void k_ffs_surfaces(const uint8_t *__restrict__ cur, int cur_pitch, const uint8_t *__restrict__ ref, int ref_pitch,
                               int w, int h, int mbx, int mby, int cx, int cy, int R, uint32_t *__restrict__ out) {
  const int max_pos = (2 * R + 1) * (2 * R + 1);
  const int pos = blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= max_pos) return;
  int l = 0, dx = 0, dy = 0;
  if (pos) {
    int s = (int)sqrtf((float)pos);
    while (s * s > pos) s--;
    while ((s + 1) * (s + 1) <= pos) s++;
    l = (s + 1) / 2;
    int base = (2 * l - 1) * (2 * l - 1), off = pos - base;
    if (off < 2 * (2 * l - 1)) { dx = (off >> 1) - l + 1; dy = (off & 1) ? l : -l; }
    else { off -= 2 * (2 * l - 1); dy = (off >> 1) - l; dx = (off & 1) ? l : -l; }
  }
  // macroblock origin clamped once per position, me_fullfast.c:498
  int X = jmb_clip(-JMB_PAD_X, w + JMB_PAD_X - 1 - 16, mbx + cx + dx);
  int Y = jmb_clip(-JMB_PAD_Y, h + JMB_PAD_Y - 1 - 16, mby + cy + dy);
  unsigned a[16];
  for (int b = 0; b < 16; b++) {
    int bx = (b & 3) * 4, by = (b >> 2) * 4;
    unsigned s = 0;
    for (int y = 0; y < 4; y++) {
      unsigned sv = *(const unsigned *)(cur + (size_t)(mby + by + y) * cur_pitch + mbx + bx);
      const uint8_t *rp = ref + (size_t)(Y + by + y + JMB_PAD_Y) * ref_pitch + (X + bx + JMB_PAD_X);
      unsigned rv = rp[0] | (rp[1] << 8) | (rp[2] << 16) | ((unsigned)rp[3] << 24);
      s = __vsadu4(sv, rv) + s;
    }
    a[b] = s;
  }
#define O(t, i) out[((size_t)(t) * 16 + (i)) * max_pos + pos]
  for (int i = 0; i < 16; i++) O(7, i) = a[i];
  unsigned s6[16], s4[16];
  for (int i = 0; i < 4; i++) { s6[i] = a[i] + a[i + 4]; s6[8 + i] = a[8 + i] + a[12 + i]; O(6, i) = s6[i]; O(6, 8 + i) = s6[8 + i]; }
  for (int i = 0; i < 16; i += 2) O(5, i) = a[i] + a[i + 1];
  s4[0] = s6[0] + s6[1]; s4[2] = s6[2] + s6[3]; s4[8] = s6[8] + s6[9]; s4[10] = s6[10] + s6[11];
  O(4, 0) = s4[0]; O(4, 2) = s4[2]; O(4, 8) = s4[8]; O(4, 10) = s4[10];
  O(3, 0) = s4[0] + s4[8]; O(3, 2) = s4[2] + s4[10];
  O(2, 0) = s4[0] + s4[2]; O(2, 8) = s4[8] + s4[10];
  O(1, 0) = s4[0] + s4[2] + s4[8] + s4[10];
#undef O
}


// ---- macroblock-resident SAD surfaces + per-partition arg-min: the drop-in path of JM's strictly sequential call sites --------
// JM decides one partition at a time (each search's predictor needs the previous decisions), but the expensive half
// of a search -- the sixteen 4x4 SADs of every displacement -- does not depend on the predictor.  jmb_mb_surfaces computes them
// ONCE per macroblock and reference over a window a little larger than the search window (setup_fast_full_search's BlockSAD,
// me_fullfast.c:492-556, kept as u16 and never leaving HBM); jmb_mb_search then is what fast_full_search_motion_estimation
// (:618-689) / full_search_motion_estimation (me_fullsearch.c:39-103) do per partition: partition sums
// (update_full_search_large_blocks :196-260, here on the fly), mv cost with the call's own predictor, arg-min with JM's
// tie-break -- one small launch, answer through a host-mapped mailbox.
constexpr int SF_NT = 96, SF_BH = 4, SF_PITCH = 96;      // surface kernel: threads, displacement rows per CTA, u16 per surface row
static_assert(SF_PITCH >= CW, "a surface row holds one staging chunk of displacements");

__global__ void __launch_bounds__(SF_NT)
k_mb_surfaces(const __grid_constant__ TMaps tm, int ref, int mbx, int mby, int x0, int y0, int ncol, int nrow, unsigned short *__restrict__ out) {
  __shared__ __align__(128) uint8_t win[WIN_ROWS * WIN_PITCH];
  __shared__ __align__(128) unsigned ssrc[16 * 4];
  __shared__ __align__(8) unsigned long long mbar;
  const int tid = threadIdx.x, band = blockIdx.x;
  const int by0 = y0 + band * SF_BH, bh = min(SF_BH, nrow - band * SF_BH);      // displacement rows of this CTA
  const int ax = mbx + x0 + JMB_PAD_X, xoff0 = ax & 15;
  if (tid == 0) {
    mbar_init(&mbar, 1);
    mbar_expect_tx(&mbar, WIN_ROWS * WIN_PITCH + 256);
    tma_load_2d(win, &tm.ref[ref], ax - xoff0, mby + by0 + JMB_PAD_Y, &mbar);
    tma_load_2d(ssrc, &tm.cur, mbx, mby, &mbar);
  }
  __syncthreads();
  mbar_wait(&mbar, 0);
  const int nrg = (bh + 3) >> 2;
  for (int it = tid; it < nrg * ncol; it += SF_NT) {
    const int rg = it / ncol, ic = it - rg * ncol;
    unsigned acc[4][16];
    sad_item(win, ssrc, rg * 4, ic + xoff0, acc);
#pragma unroll
    for (int s = 0; s < 4; s++) {
      const int r = band * SF_BH + rg * 4 + s;
      if (rg * 4 + s < bh)      // the 41 partition SADs of this displacement (update_full_search_large_blocks), one surface per partition
        for_each_partition(acc[s], [&](int p, unsigned v) { out[((size_t)p * nrow + r) * SF_PITCH + ic] = (unsigned short)v; });
    }
  }
}

struct SurfView { const unsigned short *p; int x0, y0, ncol, nrow; };      // surfaces of displacements [x0, x0+ncol) x [y0, y0+nrow)

constexpr int AM_NT = 512;
__constant__ signed char c_sp9[9][2] = {{0,0},{0,-1},{0,1},{-1,-1},{1,-1},{-1,0},{1,0},{-1,1},{1,1}};      // spiral_search[0..8], mv_search.c:410-442

// One partition's search: arg-min over its resident SAD surface, then -- JMB_REQ_SUBPEL -- the half- / quarter-pel
// refinement of sub_pel_motion_estimation (me_fullsearch.c:186-289, as k_subpel_refine does for whole pictures).
// Called by all AM_NT threads of a CTA; the answer is returned in thread 0.
struct MbSearchS {
  unsigned long long wbest[AM_NT / 32];
  int sums[9];
  int mvx, mvy;
  long long mn, icost;
};

__device__ __forceinline__ void mb_search_body(MbSearchS &S, const jmb_me_req &r, int slot, const SurfView &sv, const uint8_t *__restrict__ cur, int cur_pitch,
                                               const RefView &rv, int w, int h, int R, int max_mvd_m1, const jmb_me_config &me, jmb_me_res &o) {
  const int tid = threadIdx.x;
  const long long DISTBLK_MAX = (long long)0x7fffffff << 5;
  __syncthreads();      // (S may still be read by a previous call)
  if (!(r.flags & JMB_REQ_SKIP_INT)) {
    const bool ffs = r.mode == JMB_SEARCH_FAST_FULL;
    const int ox = ffs ? (r.pos_x & ~15) : r.pos_x, oy = ffs ? (r.pos_y & ~15) : r.pos_y;      // origin the clamp applies to
    const int dlo_x = -JMB_PAD_X - ox, dhi_x = (w + JMB_PAD_X - 1 - 16) - ox, dlo_y = -JMB_PAD_Y - oy, dhi_y = (h + JMB_PAD_Y - 1 - 16) - oy;
    const int cx = r.center_x >> 2, cy = r.center_y >> 2, px = r.pred_x, py = r.pred_y, lam = r.lambda[0];
    const int side = 2 * R + 1;
    const unsigned short *surf = sv.p + (size_t)slot * sv.nrow * SF_PITCH;
    // costs fit 32 bits (lambda <= 65535, mvbits <= 33 each, SAD < 2^16) unless nothing has been found yet; the spiral index
    // (the tie-break) is only worked out for a candidate that reaches the thread's current best cost
    unsigned long long best = (unsigned long long)r.min_mcost << IDX_BITS;
    unsigned bcost = (unsigned)min((unsigned long long)r.min_mcost, 0xffffffffull);
    const unsigned side_rcp = 0xffffffffu / (unsigned)side + 1u;      // i / side == umulhi(i, side_rcp) for i, side < 2^16
    for (int i = tid; i < side * side; i += AM_NT) {
      const int dyi = (int)__umulhi((unsigned)i, side_rcp), dxi = i - dyi * side;
      const int dx = cx - R + dxi, dy = cy - R + dyi;                  // the candidate (integer pels)
      const int mx = 4 * dx - px, my = 4 * dy - py;
      if (ffs && max(abs(mx), abs(my)) >= max_mvd_m1) continue;       // me_fullfast.c:671
      const int Dx = jmb_clip(dlo_x, dhi_x, dx) - sv.x0, Dy = jmb_clip(dlo_y, dhi_y, dy) - sv.y0;   // where the block is read (UMVLine4X)
      const unsigned sad = surf[(size_t)Dy * SF_PITCH + Dx];
      const unsigned cost = (sad << 5) + (unsigned)lam * (unsigned)(jmb_mvbits(mx) + jmb_mvbits(my));
      if (cost > bcost) continue;
      const unsigned long long k = ((unsigned long long)cost << IDX_BITS) | (unsigned)jmb_spiral_index(dx - cx, dy - cy);
      if (k < best) { best = k; bcost = cost; }
    }
#pragma unroll
    for (int sh = 16; sh; sh >>= 1) {
      const unsigned long long o2 = ((unsigned long long)__shfl_xor_sync(0xffffffffu, (unsigned)(best >> 32), sh) << 32) | __shfl_xor_sync(0xffffffffu, (unsigned)best, sh);
      best = min(best, o2);
    }
    if ((tid & 31) == 0) S.wbest[tid >> 5] = best;
    __syncthreads();
    if (tid == 0) {
      for (int i = 1; i < AM_NT / 32; i++) best = min(best, S.wbest[i]);
      int dx = 0, dy = 0;
      if (best != ((unsigned long long)r.min_mcost << IDX_BITS)) spiral_xy((int)(best & ((1u << IDX_BITS) - 1)), &dx, &dy);
      S.mvx = 4 * (cx + dx); S.mvy = 4 * (cy + dy);
      S.icost = (long long)(best >> IDX_BITS);
      S.mn = me.start_hp ? S.icost : DISTBLK_MAX;                    // BlockMotionSearch, mv_search.c:971-974
    }
  } else if (tid == 0) { S.mvx = r.center_x; S.mvy = r.center_y; S.icost = S.mn = r.min_mcost; }
  __syncthreads();
  const int imx = S.mvx, imy = S.mvy;
  if (r.flags & JMB_REQ_SUBPEL) {
#pragma unroll 1
    for (int stage = 0; stage < 2; stage++) {
      const int metric = me.metric[1 + stage], step = stage ? 1 : 2;
      const int pos0 = stage ? me.start_qp : me.start_hp;
      const int pos1 = stage ? me.search_pos4 : (!me.start_hp ? max(1, me.search_pos2) : me.search_pos2);
      const int nn = (metric == JMB_SATD && (r.flags & JMB_REQ_TEST8X8)) ? 8 : 4;
      const int nsx = c_bsx[r.blocktype] / nn, nsub = nsx * (c_bsy[r.blocktype] / nn);
      if (tid < 9) S.sums[tid] = 0;
      __syncthreads();
      const int mvx = S.mvx, mvy = S.mvy, ncand = pos1 - pos0;
      for (int it = tid; it < ncand * nsub; it += AM_NT) {
        const int c = it / nsub, sb = it - c * nsub, sbx = sb % nsx, sby = sb / nsx, pos = pos0 + c;
        SrcBlk src;
        load_src(src, cur, cur_pitch, r.pos_x + sbx * nn, r.pos_y + sby * nn, nn);
        atomicAdd(&S.sums[pos], subblock_dist(rv, src, (r.pos_x << 2) + mvx + step * c_sp9[pos][0], (r.pos_y << 2) + mvy + step * c_sp9[pos][1], sbx, sby, nn, metric));
      }
      __syncthreads();
      if (tid < 32) {      // JM's sequential strict-'<' selection (me_fullsearch.c:221-289): the first of the cheapest candidates, if it beats the bound
        long long mn = S.mn;
        if (stage == 1 && !me.start_qp) mn = DISTBLK_MAX;
        unsigned key = 0xffffffffu;      // (cost << 4) | position: costs stay below 2^23
        if (tid >= pos0 && tid < pos1) {
          const int cxq = mvx + step * c_sp9[tid][0], cyq = mvy + step * c_sp9[tid][1];
          key = (((unsigned)r.lambda[1 + stage] * (unsigned)(jmb_mvbits(cxq - r.pred_x) + jmb_mvbits(cyq - r.pred_y)) + ((unsigned)S.sums[tid] << 5)) << 4) | (unsigned)tid;
        }
        key = __reduce_min_sync(0xffffffffu, key);
        if (tid == 0) {
          int best = 0;
          if (key != 0xffffffffu && (long long)(key >> 4) < mn) { mn = key >> 4; best = (int)(key & 15); }
          S.mn = mn; S.mvx = mvx + step * c_sp9[best][0]; S.mvy = mvy + step * c_sp9[best][1];
        }
      }
      __syncthreads();
    }
  }
  if (tid == 0) {
    o.imv_x = (int16_t)imx; o.imv_y = (int16_t)imy; o.icost = S.icost;
    if (r.flags & JMB_REQ_SUBPEL) { o.mv_x = (int16_t)S.mvx; o.mv_y = (int16_t)S.mvy; o.cost = S.mn; }
    else { o.mv_x = o.imv_x; o.mv_y = o.imv_y; o.cost = S.icost; }
  }
}

// one search in one launch, answer and completion flag written straight into host-mapped memory
__global__ void __launch_bounds__(AM_NT)
k_mb_search(jmb_me_req r, int slot, SurfView sv, const uint8_t *__restrict__ cur, int cur_pitch, RefView rv, int w, int h, int R, int max_mvd_m1,
            jmb_me_config me, jmb_me_res *__restrict__ mailbox, volatile int *flag, int seq) {
  __shared__ MbSearchS S;
  jmb_me_res o;
  mb_search_body(S, r, slot, sv, cur, cur_pitch, rv, w, h, R, max_mvd_m1, me, o);
  if (threadIdx.x == 0) {
    *mailbox = o;
    __threadfence_system();
    *flag = seq;
  }
}

// ---- chains: what PartitionMotionSearch / SubPartitionMotionSearch do block after block (mv_search.c:1560-1850) -----------
// A search of a chain takes its predictor from neighbours some of which are EARLIER searches of the same chain (the upper 16x8
// block for the lower one, the 4x4 blocks of a quadrant for each other): GetMVPredictor (lcommon/src/mv_prediction.c:192-300,
// non-MBAFF) -> search centre (mv_search.c:925-957) -> IntPelME -> SubPelME -> clip_mv_range (:981) -> set_me_parameters (the
// final mv is what later blocks see).  One CTA per chain runs its searches in order; chains of one call run side by side.
struct ChainArgs { jmb_chain_req q[JMB_CHAIN_MAX]; int n, nchains, mv_lim[4], int_divide; };

__device__ __forceinline__ int median3(int a, int b, int c) { return max(min(a, b), min(max(a, b), c)); }

__global__ void __launch_bounds__(AM_NT)
k_mb_chain(const __grid_constant__ ChainArgs A, SurfView sv, const uint8_t *__restrict__ cur, int cur_pitch, RefView rv, int w, int h, int R, int max_mvd_m1,
           jmb_me_config me, jmb_chain_res *__restrict__ mailbox, volatile int *flag, int seq, unsigned *__restrict__ done_count) {
  __shared__ MbSearchS S;
  __shared__ jmb_me_req cur_req;
  __shared__ int s_status;
  __shared__ short fin[JMB_CHAIN_MAX][2];      // final (clipped) mv of the searches done so far
  __shared__ signed char fin_ok[JMB_CHAIN_MAX];
  const int tid = threadIdx.x, chain = blockIdx.x;
  for (int i = 0; i < A.n; i++) {
    const jmb_chain_req &cq = A.q[i];
    if (cq.chain != chain) continue;
    if (tid == 0) {
      int st = 0, ax[3], ay[3], rf[3], av[3];
      for (int k = 0; k < 3; k++) {
        const jmb_chain_nb &nb = cq.nb[k];
        av[k] = nb.available; rf[k] = nb.available ? nb.ref_idx : -1; ax[k] = nb.mv_x; ay[k] = nb.mv_y;
        if (nb.available && nb.dep >= 0) {
          if (nb.dep >= i || A.q[nb.dep].chain != chain || !fin_ok[nb.dep]) st = JMB_CHAIN_SKIPPED;
          else { ax[k] = fin[nb.dep][0]; ay[k] = fin[nb.dep][1]; }
        }
        if (!nb.available) ax[k] = ay[k] = 0;
      }
      // GetMotionVectorPredictorNormal (mv_prediction.c:192-300)
      const int ref = cq.jm_ref, bsx = c_bsx[cq.req.blocktype], bsy = c_bsy[cq.req.blocktype], mbx = cq.req.pos_x & 15, mby = cq.req.pos_y & 15;
      int type = 0;      // 0 median, 1 L, 2 U, 3 UR
      if (rf[0] == ref && rf[1] != ref && rf[2] != ref) type = 1;
      else if (rf[0] != ref && rf[1] == ref && rf[2] != ref) type = 2;
      else if (rf[0] != ref && rf[1] != ref && rf[2] == ref) type = 3;
      if (bsx == 8 && bsy == 16) { if (mbx == 0) { if (rf[0] == ref) type = 1; } else if (rf[2] == ref) type = 3; }
      else if (bsx == 16 && bsy == 8) { if (mby == 0) { if (rf[1] == ref) type = 2; } else if (rf[0] == ref) type = 1; }
      int px, py;
      if (type == 0) {
        if (!(av[1] || av[2])) { px = ax[0]; py = ay[0]; }
        else { px = median3(ax[0], ax[1], ax[2]); py = median3(ay[0], ay[1], ay[2]); }
      } else { px = ax[type - 1]; py = ay[type - 1]; }
      jmb_me_req r = cq.req;
      r.pred_x = (int16_t)px; r.pred_y = (int16_t)py;
      if (r.mode == JMB_SEARCH_FULL) {      // mv_search.c:931-932 (JM_INT_DIVIDE), :957
        r.center_x = (int16_t)jmb_clip(A.mv_lim[0], A.mv_lim[1], A.int_divide ? ((px + 2) >> 2) * 4 : (px / 4) * 4);
        r.center_y = (int16_t)jmb_clip(A.mv_lim[2], A.mv_lim[3], A.int_divide ? ((py + 2) >> 2) * 4 : (py / 4) * 4);
      }
      if (!st) {      // every position the search reads (after UMVLine4X's clamp) must lie inside the resident surfaces
        const bool ffs = r.mode == JMB_SEARCH_FAST_FULL;
        const int ox = ffs ? (r.pos_x & ~15) : r.pos_x, oy = ffs ? (r.pos_y & ~15) : r.pos_y;
        const int dlo_x = -JMB_PAD_X - ox, dhi_x = (w + JMB_PAD_X - 1 - 16) - ox, dlo_y = -JMB_PAD_Y - oy, dhi_y = (h + JMB_PAD_Y - 1 - 16) - oy;
        const int cx = r.center_x >> 2, cy = r.center_y >> 2;
        const int lx = jmb_clip(dlo_x, dhi_x, cx - R), hx = jmb_clip(dlo_x, dhi_x, cx + R), ly = jmb_clip(dlo_y, dhi_y, cy - R), hy = jmb_clip(dlo_y, dhi_y, cy + R);
        if (lx < sv.x0 || hx >= sv.x0 + sv.ncol || ly < sv.y0 || hy >= sv.y0 + sv.nrow || ((r.center_x | r.center_y) & 3)) st = JMB_CHAIN_UNCOVERED;
      }
      cur_req = r; s_status = st;
      fin_ok[i] = 0;
    }
    __syncthreads();
    const int st = s_status;
    jmb_me_res o;
    if (!st) {
      const jmb_me_req r = cur_req;
      const int t = r.blocktype, bx = (r.pos_x & 15) >> 2, by = (r.pos_y & 15) >> 2;      // canonical slot of the partition (order of c_part)
      const int base = t == 1 ? 0 : t == 2 ? 1 : t == 3 ? 3 : t == 4 ? 5 : t == 5 ? 9 : t == 6 ? 17 : 25;
      const int w4 = c_bsx[t] >> 2, h4 = c_bsy[t] >> 2;
      mb_search_body(S, r, base + (by / h4) * (4 / w4) + bx / w4, sv, cur, cur_pitch, rv, w, h, R, max_mvd_m1, me, o);
    }
    if (tid == 0) {
      jmb_chain_res cr;
      memset(&cr, 0, sizeof(cr));
      cr.status = st; cr.pred_x = cur_req.pred_x; cr.pred_y = cur_req.pred_y; cr.center_x = cur_req.center_x; cr.center_y = cur_req.center_y;
      if (!st) {
        cr.res = o;
        fin[i][0] = (short)jmb_clip(A.mv_lim[0], A.mv_lim[1], o.mv_x); fin[i][1] = (short)jmb_clip(A.mv_lim[2], A.mv_lim[3], o.mv_y);      // mv_search.c:981
        fin_ok[i] = 1;
      }
      mailbox[i] = cr;
    }
    __syncthreads();
  }
  if (tid == 0) {
    __threadfence_system();
    if (atomicAdd(done_count, 1u) == (unsigned)A.nchains - 1u) { *done_count = 0; __threadfence_system(); *flag = seq; }
  }
}

}  // namespace

int jmb_launch_refine(jmb_ctx *ctx, const jmb_me_req *d_reqs, jmb_me_res *d_res, int n, const uint8_t *const *d_ref_planes);

// canonical slot of a request inside its macroblock group (order of c_part)
static int part_slot(const jmb_me_req &r) {
  static const int base[8] = {0, 0, 1, 3, 5, 9, 17, 25};
  static const int w4[8] = {4, 4, 4, 2, 2, 2, 1, 1}, h4[8] = {4, 4, 2, 4, 2, 1, 2, 1};
  int t = r.blocktype, bx = (r.pos_x & 15) >> 2, by = (r.pos_y & 15) >> 2;
  return base[t] + (by / h4[t]) * (4 / w4[t]) + bx / w4[t];
}

static int validate_req(jmb_ctx *ctx, const jmb_me_req &r, int i) {
  static const int bsx[8] = {0, 16, 16, 8, 8, 8, 4, 4}, bsy[8] = {0, 16, 8, 16, 8, 4, 8, 4};
  if (r.blocktype < 1 || r.blocktype > 7) return jmb_fail(ctx, JMB_ERR_ARG, "request %d: blocktype %d", i, r.blocktype);
  if (r.ref >= ctx->nref) return jmb_fail(ctx, JMB_ERR_ARG, "request %d: ref %d of %d", i, r.ref, ctx->nref);
  if (r.pos_x < 0 || r.pos_y < 0 || r.pos_x + bsx[r.blocktype] > ctx->cur_w || r.pos_y + bsy[r.blocktype] > ctx->cur_h ||
      (r.pos_x % bsx[r.blocktype]) || (r.pos_y % bsy[r.blocktype]))
    return jmb_fail(ctx, JMB_ERR_ARG, "request %d: block (%d,%d) type %d outside / misaligned in %dx%d", i, r.pos_x, r.pos_y,
                    r.blocktype, ctx->cur_w, ctx->cur_h);
  if (!(r.flags & JMB_REQ_SKIP_INT) && ((r.center_x | r.center_y) & 3))
    return jmb_fail(ctx, JMB_ERR_ARG, "request %d: search centre (%d,%d) is not integer-pel", i, r.center_x, r.center_y);
  if (r.mode > JMB_SEARCH_FAST_FULL) return jmb_fail(ctx, JMB_ERR_ARG, "request %d: mode %d", i, r.mode);
  for (int k = 0; k < 3; k++)
    if (r.lambda[k] < 0 || r.lambda[k] > 65535) return jmb_fail(ctx, JMB_ERR_ARG, "request %d: lambda[%d]=%d not in 0..65535", i, k, r.lambda[k]);
  if (r.min_mcost < 0 || r.min_mcost > ((int64_t)1 << 48)) return jmb_fail(ctx, JMB_ERR_ARG, "request %d: min_mcost out of range", i);
  return 0;
}

static int upload_ref_table(jmb_ctx *ctx, const uint8_t *const **d_tab, int plane) {
  // table of plane[0][0] (or all-16 base) pointers of the picture's reference list, in d_groups' tail
  (void)plane;
  const uint8_t *tab[JMB_MAX_REFS];
  for (int i = 0; i < JMB_MAX_REFS; i++) tab[i] = i < ctx->nref ? ctx->refs[ctx->ref_list[i]].planes : nullptr;
  static_assert(sizeof(tab) == JMB_MAX_REFS * sizeof(void *), "");
  int rc = jmb_reserve_dev(ctx, &ctx->d_reftab, &ctx->d_reftab_cap, sizeof(tab));
  if (rc) return rc;
  JMB_CUDA(ctx, cudaMemcpyAsync(ctx->d_reftab, tab, sizeof(tab), cudaMemcpyHostToDevice, ctx->stream));
  *d_tab = (const uint8_t *const *)ctx->d_reftab;
  return 0;
}

// the launches of one search call: integer search of n_groups (macroblock, reference) groups, then the refinement
static int me_search_launch(jmb_ctx *ctx, const jmb_me_req *d_reqs, const int *d_groups, int n_groups, int n, jmb_me_res *d_res,
                            bool any_subpel, const uint8_t *const *d_tab) {
  jmb_time_begin(ctx, JMB_K_INT_SEARCH);
  if (!ctx->smem_opt_in) {      // per context: the attribute belongs to the device the context runs on
    JMB_CUDA(ctx, cudaFuncSetAttribute(k_int_search, cudaFuncAttributeMaxDynamicSharedMemorySize, INT_SEARCH_DYN_SMEM + JMB_IS_SMEM_PAD));
    ctx->smem_opt_in = true;
  }
  TMaps tm;
  memset(&tm, 0, sizeof(tm));
  tm.cur = ctx->tmap_cur;
  for (int i = 0; i < ctx->nref; i++) tm.ref[i] = ctx->refs[ctx->ref_list[i]].tmap_int;
  jmb_frame_gen gen; memset(&gen, 0, sizeof(gen));
  if (ctx->gen_pred) { gen.pred = ctx->gen_pred; gen.fp = ctx->gen_fp; gen.R = ctx->gen_R; gen.mb_w = ctx->gen_mb_w; }
  k_int_search<<<n_groups, NT, INT_SEARCH_DYN_SMEM + JMB_IS_SMEM_PAD, ctx->stream>>>(d_reqs, d_groups, d_res, tm, ctx->cur_w, ctx->cur_h, ctx->me.search_range,
                                                                  ctx->me.max_mvd - 1, ctx->nref, ctx->me.metric[0], ctx->d_err, gen);
  jmb_time_end(ctx, JMB_K_INT_SEARCH);
  JMB_LAUNCH_CHECK(ctx);
  if (any_subpel) { int rc = jmb_launch_refine(ctx, d_reqs, d_res, n, d_tab); if (rc) return rc; }
  ctx->last_res = d_res; ctx->last_res_n = n;
  return JMB_OK;
}

static int me_search_impl(jmb_ctx *ctx, const jmb_me_req *reqs, int n, jmb_me_res *res, int loc, bool frame_layout) {
  if (n <= 0) return JMB_OK;
  if (!ctx->cur || ctx->nref == 0) return jmb_fail(ctx, JMB_ERR_STATE, "jmb_me_search: call jmb_pic_begin with >= 1 reference first");
  if (ctx->me.search_range > MAX_SEARCH_RANGE) return jmb_fail(ctx, JMB_ERR_UNSUPPORTED, "search_range %d > %d", ctx->me.search_range, MAX_SEARCH_RANGE);
  if (loc == JMB_HOST_ASYNC && !frame_layout) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_me_search: JMB_HOST_ASYNC needs the frame layout (the grouping pass reads the requests on the host)");
  if (!reqs || !res) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_me_search: NULL buffer");
  JMB_CUDA(ctx, cudaSetDevice(ctx->device));
  const uint8_t *const *d_tab = nullptr;
  int rc = upload_ref_table(ctx, &d_tab, 0);
  if (rc) return rc;

  const jmb_me_req *d_reqs = reqs; jmb_me_res *d_res = res;
  const int *d_groups = nullptr; int n_groups = 0;
  bool any_subpel = false;
  const jmb_me_req *h_reqs = nullptr;
  const bool host = jmb_is_host(loc);

  if (host) h_reqs = reqs;
  else if (!frame_layout) {
    // grouping needs the request headers on the host
    rc = jmb_reserve_host(ctx, &ctx->h_stage, &ctx->h_stage_cap, (size_t)n * sizeof(jmb_me_req)); if (rc) return rc;
    JMB_CUDA(ctx, cudaMemcpyAsync(ctx->h_stage, reqs, (size_t)n * sizeof(jmb_me_req), cudaMemcpyDeviceToHost, ctx->stream));
    JMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    h_reqs = (const jmb_me_req *)ctx->h_stage;
  }
  // the frame layout is validated on the device (jmb_req_check); ungrouped requests are walked here anyway
  if (h_reqs && !frame_layout) {
    for (int i = 0; i < n; i++) { rc = validate_req(ctx, h_reqs[i], i); if (rc) return rc; any_subpel |= (h_reqs[i].flags & JMB_REQ_SUBPEL) != 0; }
  } else any_subpel = true;

  if (frame_layout) {
    if (n % NPART) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_me_search_frame: n=%d is not a multiple of 41", n);
    n_groups = n / NPART;
  } else {
    // consecutive requests of one (macroblock, ref) form a group; a repeated partition starts a new one
    rc = jmb_reserve_host(ctx, &ctx->h_groups, &ctx->h_groups_cap, (size_t)n * NPART * sizeof(int)); if (rc) return rc;
    int *hg = (int *)ctx->h_groups;
    int kx = -1, ky = -1, kr = -1;
    for (int i = 0; i < n; i++) {
      const jmb_me_req &q = h_reqs[i];
      int slot = part_slot(q);
      bool fresh = n_groups == 0 || (q.pos_x & ~15) != kx || (q.pos_y & ~15) != ky || q.ref != kr || hg[(n_groups - 1) * NPART + slot] >= 0;
      if (fresh) {
        for (int p = 0; p < NPART; p++) hg[n_groups * NPART + p] = -1;
        n_groups++; kx = q.pos_x & ~15; ky = q.pos_y & ~15; kr = q.ref;
      }
      hg[(n_groups - 1) * NPART + slot] = i;
    }
    rc = jmb_reserve_dev(ctx, &ctx->d_groups, &ctx->d_groups_cap, (size_t)n_groups * NPART * sizeof(int)); if (rc) return rc;
    JMB_CUDA(ctx, cudaMemcpyAsync(ctx->d_groups, hg, (size_t)n_groups * NPART * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    d_groups = (const int *)ctx->d_groups;
  }
  if (host) {
    rc = jmb_reserve_dev(ctx, &ctx->d_stage, &ctx->d_stage_cap, (size_t)n * sizeof(jmb_me_req)); if (rc) return rc;
    rc = jmb_reserve_dev(ctx, &ctx->d_res_keep, &ctx->d_res_keep_cap, (size_t)n * sizeof(jmb_me_res)); if (rc) return rc;
    JMB_CUDA(ctx, cudaMemcpyAsync(ctx->d_stage, reqs, (size_t)n * sizeof(jmb_me_req), cudaMemcpyHostToDevice, ctx->stream));
    d_reqs = (const jmb_me_req *)ctx->d_stage; d_res = (jmb_me_res *)ctx->d_res_keep;
  }
  rc = me_search_launch(ctx, d_reqs, d_groups, n_groups, n, d_res, any_subpel, d_tab);
  if (rc) return rc;
  if (host) {
    JMB_CUDA(ctx, cudaMemcpyAsync(res, d_res, (size_t)n * sizeof(jmb_me_res), cudaMemcpyDeviceToHost, ctx->stream));
    if (loc == JMB_HOST) return jmb_check_device_errors(ctx);
  }
  return JMB_OK;
}

namespace {
// Requests of a whole picture from the 41 predictors of every macroblock (jmb_me_search_frame_pred): block geometry from the
// partition index, search centre and flags by the rules of BlockMotionSearch / setup_fast_full_search.
// (40-byte records: a block of 256 stages them in shared memory and writes 16-byte words, 640 per block, fully coalesced)
// final clip of the mv (mv_search.c:981) applied to the resident results, and their 8-byte form (24-byte records read through
// shared memory as 16-byte words)
__global__ void __launch_bounds__(256)
k_pack_results(jmb_me_res *__restrict__ res, int n, jmb_frame_params fp, jmb_me_res8 *__restrict__ out) {
  __shared__ __align__(16) jmb_me_res sr[256];
  static_assert(sizeof(jmb_me_res) == 24, "staging copies 256 x 24 bytes as 384 x 16");
  const int t0 = blockIdx.x * 256, t = t0 + threadIdx.x, cnt = min(256, n - t0);
  const uint4 *src = (const uint4 *)(res + t0);
  for (int i = threadIdx.x; i < cnt * 24 / 16; i += 256) ((uint4 *)sr)[i] = src[i];
  for (int i = (cnt * 24 / 16) * 16 + threadIdx.x * 8; i < cnt * 24; i += 256 * 8) *(uint2 *)((char *)sr + i) = *(const uint2 *)((const char *)src + i);
  __syncthreads();
  if (t >= n) return;
  const jmb_me_res r = sr[threadIdx.x];
  const int mx = jmb_clip(fp.mv_min_x, fp.mv_max_x, r.mv_x), my = jmb_clip(fp.mv_min_y, fp.mv_max_y, r.mv_y);
  if (mx != r.mv_x || my != r.mv_y) { res[t].mv_x = (int16_t)mx; res[t].mv_y = (int16_t)my; }
  if (out) {
    jmb_me_res8 o;
    o.mv_x = (int16_t)mx; o.mv_y = (int16_t)my;
    o.cost = r.cost > 0x7fffffffLL ? 0x7fffffff : (int32_t)r.cost;
    out[t] = o;
  }
}
}  // namespace

extern "C" {

int jmb_me_search(jmb_ctx *ctx, const jmb_me_req *reqs, int n, jmb_me_res *res, int loc) {
  return me_search_impl(ctx, reqs, n, res, loc, false);
}

int jmb_me_search_frame(jmb_ctx *ctx, const jmb_me_req *reqs, int n_mb, jmb_me_res *res, int loc) {
  return me_search_impl(ctx, reqs, n_mb * NPART, res, loc, true);
}

int jmb_me_search_frame_pred(jmb_ctx *ctx, const jmb_mb_mvpred *pred, int n_mb, const jmb_frame_params *fp, jmb_me_res8 *res, int loc) {
  if (n_mb <= 0) return JMB_OK;
  if (!pred || !fp) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_me_search_frame_pred: NULL argument");
  if (!ctx->cur || ctx->nref == 0) return jmb_fail(ctx, JMB_ERR_STATE, "jmb_me_search_frame_pred: call jmb_pic_begin with >= 1 reference first");
  const int mb_w = ctx->cur_w / 16, mb_total = mb_w * (ctx->cur_h / 16), R = ctx->me.search_range;
  if (n_mb > mb_total) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_me_search_frame_pred: n_mb %d (picture has %d)", n_mb, mb_total);
  if (fp->mode < JMB_SEARCH_FULL || fp->mode > JMB_SEARCH_FAST_FULL || fp->ref < 0 || fp->ref >= ctx->nref)
    return jmb_fail(ctx, JMB_ERR_ARG, "jmb_me_search_frame_pred: mode %d ref %d", fp->mode, fp->ref);
  if (fp->mv_min_x > fp->mv_max_x || fp->mv_min_y > fp->mv_max_y || fp->mv_min_x < -32768 || fp->mv_max_x > 32767 || fp->mv_min_y < -32768 ||
      fp->mv_max_y > 32767 || ((fp->mv_min_x | fp->mv_min_y) & 3) ||
      (fp->mode == JMB_SEARCH_FAST_FULL && (fp->mv_min_x + 4 * R > fp->mv_max_x - 4 * R || fp->mv_min_y + 4 * R > fp->mv_max_y - 4 * R)))
    return jmb_fail(ctx, JMB_ERR_ARG, "jmb_me_search_frame_pred: mv range x %d..%d y %d..%d", fp->mv_min_x, fp->mv_max_x, fp->mv_min_y, fp->mv_max_y);
  for (int k = 0; k < 3; k++)
    if (fp->lambda[k] < 0 || fp->lambda[k] > 65535) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_me_search_frame_pred: lambda[%d]=%d not in 0..65535", k, fp->lambda[k]);
  if (R > MAX_SEARCH_RANGE) return jmb_fail(ctx, JMB_ERR_UNSUPPORTED, "search_range %d > %d", R, MAX_SEARCH_RANGE);
  JMB_CUDA(ctx, cudaSetDevice(ctx->device));
  const bool host = jmb_is_host(loc);
  const int n = n_mb * NPART;
  const jmb_mb_mvpred *d_pred = pred;
  int rc;
  if (host) {
    rc = jmb_reserve_dev(ctx, &ctx->d_mvpred, &ctx->d_mvpred_cap, (size_t)n_mb * sizeof(jmb_mb_mvpred)); if (rc) return rc;
    JMB_CUDA(ctx, cudaMemcpyAsync(ctx->d_mvpred, pred, (size_t)n_mb * sizeof(jmb_mb_mvpred), cudaMemcpyHostToDevice, ctx->stream));
    d_pred = (const jmb_mb_mvpred *)ctx->d_mvpred;
  }
  rc = jmb_reserve_dev(ctx, &ctx->d_res_keep, &ctx->d_res_keep_cap, (size_t)n * sizeof(jmb_me_res)); if (rc) return rc;
  const uint8_t *const *d_tab = nullptr;
  rc = upload_ref_table(ctx, &d_tab, 0); if (rc) return rc;
  jmb_me_res *d_res = (jmb_me_res *)ctx->d_res_keep;
  jmb_me_res8 *d_out = res;
  if (host && res) {
    rc = jmb_reserve_dev(ctx, &ctx->d_res8, &ctx->d_res8_cap, (size_t)n * sizeof(jmb_me_res8)); if (rc) return rc;
    d_out = (jmb_me_res8 *)ctx->d_res8;
  }
  // The 41 requests of a macroblock are formed inside the kernels from its predictor record (jmb_frame_request): no request
  // array, no generator launch.  With the sub-pel stage on, its kernel also clips the final mv and writes the 8-byte results.
  const bool subpel = (fp->flags & JMB_REQ_SUBPEL) != 0;
  ctx->gen_pred = d_pred; ctx->gen_fp = *fp; ctx->gen_R = R; ctx->gen_mb_w = mb_w;
  ctx->pack_on = subpel; ctx->pack_out = d_out;
  rc = me_search_launch(ctx, nullptr, nullptr, n_mb, n, d_res, subpel, d_tab);
  ctx->gen_pred = nullptr; ctx->pack_on = false; ctx->pack_out = nullptr;
  if (rc) return rc;
  if (!subpel) {
    jmb_time_begin(ctx, JMB_K_GEN);
    k_pack_results<<<(n + 255) / 256, 256, 0, ctx->stream>>>(d_res, n, *fp, d_out);
    jmb_time_end(ctx, JMB_K_GEN);
    JMB_LAUNCH_CHECK(ctx);
  }
  if (host) {
    if (res) JMB_CUDA(ctx, cudaMemcpyAsync(res, d_out, (size_t)n * sizeof(jmb_me_res8), cudaMemcpyDeviceToHost, ctx->stream));
    if (loc == JMB_HOST) return jmb_check_device_errors(ctx);
  }
  return JMB_OK;
}

int jmb_mb_surfaces(jmb_ctx *ctx, int ref, int mb_x, int mb_y, int center_x, int center_y, int radius) {
  if (!ctx->cur || ref < 0 || ref >= ctx->nref) return jmb_fail(ctx, JMB_ERR_STATE, "jmb_mb_surfaces: no picture / bad ref %d", ref);
  if ((mb_x & 15) || (mb_y & 15) || mb_x < 0 || mb_y < 0 || mb_x + 16 > ctx->cur_w || mb_y + 16 > ctx->cur_h || ((center_x | center_y) & 3))
    return jmb_fail(ctx, JMB_ERR_ARG, "jmb_mb_surfaces: macroblock (%d,%d) centre (%d,%d)", mb_x, mb_y, center_x, center_y);
  if (radius < 1 || 2 * radius + 1 > CW) return jmb_fail(ctx, JMB_ERR_UNSUPPORTED, "jmb_mb_surfaces: radius %d (window wider than %d displacements)", radius, CW);
  JMB_CUDA(ctx, cudaSetDevice(ctx->device));
  jmb_ctx::Surf &sf = ctx->surf[ref];
  const int n = 2 * radius + 1;
  const size_t bytes = (size_t)NPART * n * SF_PITCH * sizeof(unsigned short);
  int rc = jmb_reserve_dev(ctx, &sf.buf, &sf.cap, bytes); if (rc) return rc;
  sf.valid = true; sf.mb_x = mb_x; sf.mb_y = mb_y; sf.x0 = (center_x >> 2) - radius; sf.y0 = (center_y >> 2) - radius; sf.n = n;
  sf.pic_serial = ctx->pic_serial;
  TMaps tm;
  memset(&tm, 0, sizeof(tm));
  tm.cur = ctx->tmap_cur;
  for (int i = 0; i < ctx->nref; i++) tm.ref[i] = ctx->refs[ctx->ref_list[i]].tmap_int;
  jmb_time_begin(ctx, JMB_K_FFS_SURF);
  k_mb_surfaces<<<(n + SF_BH - 1) / SF_BH, SF_NT, 0, ctx->stream>>>(tm, ref, mb_x, mb_y, sf.x0, sf.y0, n, n, (unsigned short *)sf.buf);
  jmb_time_end(ctx, JMB_K_FFS_SURF);
  JMB_LAUNCH_CHECK(ctx);
  return JMB_OK;
}

static int mailbox_init(jmb_ctx *ctx) {
  if (ctx->mbox) return 0;
  JMB_CUDA(ctx, cudaHostAlloc(&ctx->mbox, 1024, cudaHostAllocMapped));      // one answer at 0, the flag at 128, chain answers from 256
  memset(ctx->mbox, 0, 1024);
  JMB_CUDA(ctx, cudaMalloc(&ctx->d_one, 16));
  JMB_CUDA(ctx, cudaMemset(ctx->d_one, 0, 16));
  JMB_CUDA(ctx, cudaHostGetDevicePointer(&ctx->d_mbox, ctx->mbox, 0));
  return 0;
}

// wait for the mailbox flag (the kernels write the result into host-mapped memory, then the flag); falls back to a stream
// synchronisation when the flag does not show up in time (e.g. a launch failure, reported by the synchronisation)
static int mailbox_wait(jmb_ctx *ctx, int seq) {
  volatile int *flag = (volatile int *)((char *)ctx->mbox + 128);
  for (long spin = 0; *flag != seq; spin++) {
    if (spin > 2000000L) { JMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); if (*flag != seq) return jmb_fail(ctx, JMB_ERR_CUDA, "mailbox: no answer from the device"); }
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
  }
  return 0;
}

int jmb_mb_search(jmb_ctx *ctx, const jmb_me_req *req, jmb_me_res *res) {
  if (!req || !res) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_mb_search: NULL argument");
  if (!ctx->cur || ctx->nref == 0) return jmb_fail(ctx, JMB_ERR_STATE, "jmb_mb_search: call jmb_pic_begin first");
  int rc = validate_req(ctx, *req, 0); if (rc) return rc;
  const bool skip_int = (req->flags & JMB_REQ_SKIP_INT) != 0;
  if (!skip_int && ctx->me.metric[0] != JMB_SAD) return jmb_fail(ctx, JMB_ERR_UNSUPPORTED, "jmb_mb_search: the surfaces are SAD surfaces (MEDistortionFPel = SAD)");
  if ((req->flags & JMB_REQ_TEST8X8) && req->blocktype > 4) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_mb_search: JMB_REQ_TEST8X8 needs blocktype <= 4");
  JMB_CUDA(ctx, cudaSetDevice(ctx->device));
  rc = mailbox_init(ctx); if (rc) return rc;
  jmb_me_res *mb_res = (jmb_me_res *)ctx->d_mbox;
  volatile int *d_flag = (volatile int *)((char *)ctx->d_mbox + 128);
  const int seq = ++ctx->mbox_seq;
  const int R = ctx->me.search_range, w = ctx->cur_w, h = ctx->cur_h;
  SurfView sv{nullptr, 0, 0, 0, 0};
  if (!skip_int) {
    const jmb_ctx::Surf &sf = ctx->surf[req->ref];
    const bool ffs = req->mode == JMB_SEARCH_FAST_FULL;
    if (!sf.valid || sf.pic_serial != ctx->pic_serial || sf.mb_x != (req->pos_x & ~15) || sf.mb_y != (req->pos_y & ~15))
      return jmb_fail(ctx, JMB_ERR_STATE, "jmb_mb_search: no surfaces resident for macroblock (%d,%d) reference %d", req->pos_x & ~15, req->pos_y & ~15, req->ref);
    // every position the search reads (after UMVLine4X's clamp) must lie inside the resident surfaces
    const int ox = ffs ? (req->pos_x & ~15) : req->pos_x, oy = ffs ? (req->pos_y & ~15) : req->pos_y;
    const int dlo_x = -JMB_PAD_X - ox, dhi_x = (w + JMB_PAD_X - 1 - 16) - ox, dlo_y = -JMB_PAD_Y - oy, dhi_y = (h + JMB_PAD_Y - 1 - 16) - oy;
    const int cx = req->center_x >> 2, cy = req->center_y >> 2;
    auto clip = [](int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); };
    const int lx = clip(dlo_x, dhi_x, cx - R), hx = clip(dlo_x, dhi_x, cx + R), ly = clip(dlo_y, dhi_y, cy - R), hy = clip(dlo_y, dhi_y, cy + R);
    if (lx < sf.x0 || hx >= sf.x0 + sf.n || ly < sf.y0 || hy >= sf.y0 + sf.n)
      return jmb_fail(ctx, JMB_ERR_STATE, "jmb_mb_search: window of the request (centre %d,%d) is not covered by the resident surfaces", req->center_x, req->center_y);
    sv = SurfView{(const unsigned short *)sf.buf, sf.x0, sf.y0, sf.n, sf.n};
  }
  const jmb_ref &rr = ctx->refs[ctx->ref_list[req->ref]];
  RefView rv{rr.planes, rr.plane_bytes, rr.pitch, rr.w, rr.h};
  jmb_time_begin(ctx, JMB_K_ARGMIN);
  k_mb_search<<<1, AM_NT, 0, ctx->stream>>>(*req, part_slot(*req), sv, ctx->cur, ctx->cur_pitch, rv, w, h, R, ctx->me.max_mvd - 1, ctx->me, mb_res, d_flag, seq);
  jmb_time_end(ctx, JMB_K_ARGMIN);
  JMB_LAUNCH_CHECK(ctx);
  rc = mailbox_wait(ctx, seq); if (rc) return rc;
  *res = *(const jmb_me_res *)ctx->mbox;
  return JMB_OK;
}

int jmb_mb_chain(jmb_ctx *ctx, const jmb_chain_req *reqs, int n, const int32_t mv_limits[4], int int_divide, jmb_chain_res *res) {
  static_assert(sizeof(jmb_chain_req) == 72 && sizeof(jmb_chain_res) == 40 && 256 + JMB_CHAIN_MAX * sizeof(jmb_chain_res) <= 1024, "chain mailbox layout");
  if (!reqs || !res || !mv_limits || n < 1 || n > JMB_CHAIN_MAX) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_mb_chain: %d requests (1..%d)", n, JMB_CHAIN_MAX);
  if (!ctx->cur || ctx->nref == 0) return jmb_fail(ctx, JMB_ERR_STATE, "jmb_mb_chain: call jmb_pic_begin first");
  if (ctx->me.metric[0] != JMB_SAD) return jmb_fail(ctx, JMB_ERR_UNSUPPORTED, "jmb_mb_chain: the surfaces are SAD surfaces (MEDistortionFPel = SAD)");
  if (mv_limits[0] > mv_limits[1] || mv_limits[2] > mv_limits[3] || mv_limits[0] < -32768 || mv_limits[1] > 32767 || mv_limits[2] < -32768 || mv_limits[3] > 32767)
    return jmb_fail(ctx, JMB_ERR_ARG, "jmb_mb_chain: mv range x %d..%d y %d..%d", mv_limits[0], mv_limits[1], mv_limits[2], mv_limits[3]);
  ChainArgs A;
  memset(&A, 0, sizeof(A));
  int nchains = 0;
  for (int i = 0; i < n; i++) {
    const jmb_chain_req &q = reqs[i];
    int rc = validate_req(ctx, q.req, i); if (rc) return rc;
    if (q.req.flags & JMB_REQ_SKIP_INT) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_mb_chain: request %d: JMB_REQ_SKIP_INT has no place in a chain", i);
    if ((q.req.flags & JMB_REQ_TEST8X8) && q.req.blocktype > 4) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_mb_chain: request %d: JMB_REQ_TEST8X8 needs blocktype <= 4", i);
    if (q.req.ref != reqs[0].req.ref || (q.req.pos_x & ~15) != (reqs[0].req.pos_x & ~15) || (q.req.pos_y & ~15) != (reqs[0].req.pos_y & ~15))
      return jmb_fail(ctx, JMB_ERR_ARG, "jmb_mb_chain: request %d belongs to another macroblock / reference than request 0", i);
    if (q.chain < 0 || q.chain >= JMB_CHAIN_MAX) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_mb_chain: request %d: chain %d", i, q.chain);
    for (int k = 0; k < 3; k++)
      if (q.nb[k].available && q.nb[k].dep >= 0 && (q.nb[k].dep >= i || reqs[q.nb[k].dep].chain != q.chain))
        return jmb_fail(ctx, JMB_ERR_ARG, "jmb_mb_chain: request %d depends on %d, which is not an earlier search of its chain", i, q.nb[k].dep);
    nchains = nchains > q.chain + 1 ? nchains : q.chain + 1;
    A.q[i] = q;
  }
  A.n = n; A.nchains = nchains; A.int_divide = int_divide;
  for (int k = 0; k < 4; k++) A.mv_lim[k] = mv_limits[k];
  JMB_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = mailbox_init(ctx); if (rc) return rc;
  const jmb_me_req &r0 = reqs[0].req;
  const jmb_ctx::Surf &sf = ctx->surf[r0.ref];
  if (!sf.valid || sf.pic_serial != ctx->pic_serial || sf.mb_x != (r0.pos_x & ~15) || sf.mb_y != (r0.pos_y & ~15))
    return jmb_fail(ctx, JMB_ERR_STATE, "jmb_mb_chain: no surfaces resident for macroblock (%d,%d) reference %d", r0.pos_x & ~15, r0.pos_y & ~15, r0.ref);
  SurfView sv{(const unsigned short *)sf.buf, sf.x0, sf.y0, sf.n, sf.n};
  const jmb_ref &rr = ctx->refs[ctx->ref_list[r0.ref]];
  RefView rv{rr.planes, rr.plane_bytes, rr.pitch, rr.w, rr.h};
  const int seq = ++ctx->mbox_seq;
  jmb_time_begin(ctx, JMB_K_ARGMIN);
  k_mb_chain<<<nchains, AM_NT, 0, ctx->stream>>>(A, sv, ctx->cur, ctx->cur_pitch, rv, ctx->cur_w, ctx->cur_h, ctx->me.search_range, ctx->me.max_mvd - 1, ctx->me,
                                                 (jmb_chain_res *)((char *)ctx->d_mbox + 256), (volatile int *)((char *)ctx->d_mbox + 128), seq, (unsigned *)ctx->d_one);
  jmb_time_end(ctx, JMB_K_ARGMIN);
  JMB_LAUNCH_CHECK(ctx);
  rc = mailbox_wait(ctx, seq); if (rc) return rc;
  memcpy(res, (const char *)ctx->mbox + 256, (size_t)n * sizeof(jmb_chain_res));
  return JMB_OK;
}

int jmb_ffs_surfaces(jmb_ctx *ctx, int ref, int mb_x, int mb_y, int center_x, int center_y, uint32_t *out, int loc) {
  if (!ctx->cur || ref < 0 || ref >= ctx->nref) return jmb_fail(ctx, JMB_ERR_STATE, "jmb_ffs_surfaces: no picture / bad ref %d", ref);
  if ((mb_x & 15) || (mb_y & 15) || mb_x < 0 || mb_y < 0 || mb_x + 16 > ctx->cur_w || mb_y + 16 > ctx->cur_h || ((center_x | center_y) & 3))
    return jmb_fail(ctx, JMB_ERR_ARG, "jmb_ffs_surfaces: macroblock (%d,%d) centre (%d,%d)", mb_x, mb_y, center_x, center_y);
  JMB_CUDA(ctx, cudaSetDevice(ctx->device));
  const jmb_ref &r = ctx->refs[ctx->ref_list[ref]];
  const int R = ctx->me.search_range, max_pos = (2 * R + 1) * (2 * R + 1);
  size_t bytes = (size_t)8 * 16 * max_pos * sizeof(uint32_t);
  uint32_t *d_out = out;
  if (loc == JMB_HOST) {
    int rc = jmb_reserve_dev(ctx, &ctx->d_stage2, &ctx->d_stage2_cap, bytes); if (rc) return rc;
    d_out = (uint32_t *)ctx->d_stage2;
  }
  jmb_time_begin(ctx, JMB_K_FFS_SURF);
  k_ffs_surfaces<<<(max_pos + 127) / 128, 128, 0, ctx->stream>>>(ctx->cur, ctx->cur_pitch, r.planes, r.pitch, r.w, r.h, mb_x, mb_y,
                                                                 center_x >> 2, center_y >> 2, R, d_out);
  jmb_time_end(ctx, JMB_K_FFS_SURF);
  JMB_LAUNCH_CHECK(ctx);
  if (loc == JMB_HOST) {
    JMB_CUDA(ctx, cudaMemcpyAsync(out, d_out, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    JMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return JMB_OK;
}

}  // extern "C"
