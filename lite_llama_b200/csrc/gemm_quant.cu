// Weight-only quantised GEMMs on tcgen05: W4A16 (int4 groups + fp32 scale/zero) and W8A16
// (fp8-e4m3 or int8 bytes + fp32 block scales).  Replaces
//   lite_llama/kernels/quantization/w4a16.py:28-207   (scalar outer-product Triton kernel)
//   lite_llama/kernels/quantization/w8a16.py:48-216
// and consumes exactly their weight layouts:
//   int4 : qweight int32 [N, K/8], nibble j of word i = K index 8i+j; scales, zeros fp32 [N, K/g]
//   8-bit: qweight uint8(e4m3)/int8 [N, K]; scales fp32 [ceil(N/gn), ceil(K/gk)]
//
// Pipeline per CTA (tile orientation: gemm_tc.cuh -- 128 weight rows on the TMEM lanes, tokens on UMMA N):
//   TMA producer     raw quantised weight tile (128 rows x 128 BYTES, 128B-swizzled) + the matching
//                    activation sub-tiles -> deep shared-memory ring (HBM latency hiding)
//   8 dequant warps  two groups of 4 warps alternate over the 64-element k-blocks; thread = weight row:
//                    raw bytes -> exact fp16 values in registers -> tcgen05.st into a TMEM operand stage
//   2 MMA issuers    tcgen05.mma with the A operand read from TENSOR MEMORY and B (activations) from
//                    shared memory, accumulating into one of four TMEM accumulator buffers; a buffer is
//                    closed at every scale-group boundary ("segment"); the two warps take alternate segments
//   4 epilogue warps drain each closed buffer:  acc += scale * (P - zero * xsum)   in fp32; for int4 they also
//                    compute xsum[m] = sum_k x[m,k] of the segment from the activation tiles in shared memory
//
// Why TMEM for the dequantised operand: staging it in shared memory (v2 of this kernel, see
// profiles/) costs 16 KiB of st.shared + 16-32 KiB of tensor-core reads per 4 KiB of int4 fetched from
// HBM -- an 8x traffic amplification that caps the kernel below the HBM roofline at the 128 B/clk/SM
// shared-memory port.  In TMEM the operand never touches shared memory.
//
// Numerics: the tensor core multiplies the *unscaled* integer (or exact e4m3) weight values, so products
// and the per-group partial sums are exact in fp32; scale and zero point are applied in fp32 afterwards,
// exactly where the reference applies them (w4a16.py:104, w8a16.py:116).
// Measured history of this kernel (device timelines + ncu under profiles/): shared-memory operand staging,
// per-segment global scale loads, single-lane role loops and serial barrier polling each capped it at
// 5-20 % of HBM peak before the current structure.
#include "gemm_tc.cuh"
#include <type_traits>
#include <stdlib.h>

namespace gemm {

constexpr int kQuantThreads = 16 * 32;      // 2 x 4 dequant warps, 4 epilogue, 2 producers, 2 MMA issuers
// Warp roles.  The SM sub-partition arbiter favours the HIGHEST warp id among eligible warps, so the
// latency-critical single-instruction-stream roles (MMA issue, TMA issue, accumulator drain) get the high ids
// and the ALU-heavy dequant warps the low ones; with the opposite order the issuers waited ~500 cycles for a
// slot behind the dequant warps of their sub-partition.  warp % 4 = TMEM lane quadrant for dequant / epilogue.
constexpr int kWarpEpi0 = 8;                // epilogue warps 8..11
constexpr int kWarpXProd = 12, kWarpWProd = 13, kWarpMma1 = 14, kWarpMma0 = 15;
constexpr int kMaxOpStages = 6;             // TMEM operand stages (two 64-element sub-blocks = 64 columns each)
constexpr int kMaxRawStages = 8;
// operand ring depth: 6 stages (three raw blocks of int4) when they fit next to the four accumulator buffers
__host__ __device__ constexpr int op_stages_for(int bt) { return 4 * bt + 6 * 64 <= 512 ? 6 : 4; }
constexpr int kTabPitch = 129;            // floats per scale-table column (128 weight rows + 1: conflict-free both ways)

enum QuantKind { Q_INT4 = 0, Q_FP8 = 1, Q_INT8 = 2 };

struct QuantParams {
  void* y; int64_t ldy;
  const void* bias; int bias_is_f32;
  const float* scales; int64_t s_stride_n, s_stride_k;
  const float* zeros;                      // int4 only (same strides as scales)
  const void* x; int64_t ldx;              // activations (global): only read by the non-integer zero-point fallback
  int force_xsum;                          // scale group is not 64 * 2^n: zero points cannot be folded per sub-block
  int seg_shift;                           // log2(group_k / 64) when force_xsum == 0
  int group_n, group_k;                    // int4: group_n = 1, group_k = group_size
  uint32_t* ws; int* counters;
  int M, N, K;
  int splits, kb_per_split, num_kb;
  // Stream-K schedule (sk_ctas > 0; grid.x = sk_ctas, grid.z = 1): the (tile, raw block) space is linearised
  // tile-major and CTA c owns the contiguous range [c*total/sk_ctas, (c+1)*total/sk_ctas) -- every SM streams the
  // same number of weight bytes (+-1 raw block) whatever the tile count, instead of whole tiles in uneven waves.
  // A range covers the tail of one tile, whole tiles, and the head of another ("pieces"); a tile shared by several
  // CTAs is completed through the split-K workspace + ticket, slots in CTA order (deterministic).
  int sk_ctas, sk_slots;                   // sk_slots: workspace slots per tile (max CTAs sharing one tile)
  int n_tiles, nrb_tile;                   // weight-row tiles, raw blocks per tile (K / raw-block K)
  int w_stages, x_stages;                  // ring depths: raw weight blocks / activation blocks
  int max_groups;                          // scale groups one CTA's k-range can touch (sizes the smem scale table)
  int aligned;                             // no segment spans two raw blocks (two MMA warps may alternate blocks)
  int zeros_foldable;                      // packed layout: the packer verified that every zero point is an integer in [0, 1023]
  int swiglu;                              // epilogue: rows are (gate_i, up_i) pairs; y[m, i] = silu(gate) * up (N/2 columns)
  long long* timeline;                     // debug: [16][256] clock64 stamps of CTA (0,0,0), or null
  int debug;                               // debug experiments (timing only, results are wrong): see DBG_*
};
// DBG_RAW_OPERAND: skip the int4 -> fp16 arithmetic (operand = raw bits): how much of the time is the ALU pipe?
// DBG_PAIR_K4: extract nibble pairs (k, k+4) with one LOP3 + HFMA2 each (5 ALU-pipe instructions per word instead
// of 8) WITHOUT permuting the activations -- prices a pre-shuffled activation layout.  DBG_NO_STREAMK: host side.
// DBG_SKIP_DRAIN: only the last segment of a piece is read back from TMEM (prices a scale-folded operand).
enum { DBG_FORCE_XSUM = 8, DBG_RAW_OPERAND = 16, DBG_PAIR_K4 = 32, DBG_NO_STREAMK = 64, DBG_SKIP_DRAIN = 128,
       DBG_FEW_MMA = 256 /* issue 1 of every 4 MMAs */, DBG_NO_STTM = 512 /* skip the TMEM operand store */,
       DBG_WAIT_SPIN = 1024 /* mbarrier.test_wait polling */, DBG_WAIT_SHORT = 2048 /* try_wait without the long suspend hint */,
       DBG_NO_XLOAD = 4096 /* activation tiles are never fetched */, DBG_HALF_RING = 8192 /* half the weight ring depth */,
       DBG_LEGACY_OPRING = 16384 /* round-1 operand-ring indexing (phase-aliasing hang repro; results may be wrong) */,
       DBG_FORCE = 32768 /* no experiment: just run the DBG instantiation (prices its code size) */ };
// Wait flavour of the hand-off barriers (0: try_wait with the long suspend hint, 1: test_wait polling, 2: try_wait
// with the default suspend window).  B2L_QWAIT_MODE covers every wait, B2L_QWAIT_LAT_MODE the latency-critical
// ones (operand stage full / free, accumulator full): A/B builds, see tools/build_variants.py.
#ifndef B2L_QWAIT_MODE
#define B2L_QWAIT_MODE 0
#endif
#ifndef B2L_QWAIT_LAT_MODE
#define B2L_QWAIT_LAT_MODE B2L_QWAIT_MODE
#endif
// Dequant warps: 0 = wait for the operand stage before converting a sub-block, 1 = convert first, wait before the store
#ifndef B2L_DQ_ORDER
#define B2L_DQ_ORDER 0
#endif
template <int MODE>
__device__ __forceinline__ void qwait_at(uint64_t* bar, uint32_t parity, int debug, int line) {
  constexpr unsigned int fh = b2l_file_hash(__FILE__);
  if ((debug & DBG_WAIT_SPIN) || MODE == 1) mbar_wait_spin_at(bar, parity, fh, line);
  else if ((debug & DBG_WAIT_SHORT) || MODE == 2) mbar_wait_short_at(bar, parity, fh, line);
  else mbar_wait_at(bar, parity, fh, line);
}
#define qwait(bar, parity, debug) qwait_at<B2L_QWAIT_MODE>((bar), (parity), (debug), __LINE__)
#define qwait_lat(bar, parity, debug) qwait_at<B2L_QWAIT_LAT_MODE>((bar), (parity), (debug), __LINE__)
constexpr int kTlCta = 4096;    // timeline: [1024][3] per-CTA {globaltimer at entry, at exit, SM id}
constexpr int kTlSub = 3072 + 32;   // timeline: [96][5] per-sub-block stamps of dequant warp 0
constexpr int kTlMisc = 3072;   // timeline: start-up / shutdown stamps of CTA (0,0,0), see tools/timeline.py

// B2L_TLITE (A/B builds only, tools/build_variants.py): a handful of clock stamps per CTA in the PRODUCT instantiation
// (the DBG instantiation's own bookkeeping makes it ~50 % slower, so its timeline misplaces the critical path)
#ifdef B2L_TLITE
#define LT(i) do { if (p.timeline) p.timeline[(size_t)blockIdx.x * 32 + (i)] = clock64(); } while (0)
#else
#define LT(i) do {} while (0)
#endif

// role rows of the debug timeline
enum { TL_PROD = 0, TL_MMA_READY = 1, TL_MMA_ISSUED = 2, TL_DQ_START = 3, TL_DQ_DONE = 4, TL_EPI_START = 5, TL_EPI_DONE = 6 };
#define TL_STAMP(role, idx)                                                                   \
  do {                                                                                        \
    if (tl && (idx) < 256 && lane == 0) tl[(role) * 256 + (idx)] = clock64();                 \
  } while (0)

static long long* g_timeline = nullptr;   // set through b2l_debug_set_timeline (tools/timeline.py)
static int g_debug = 0;                   // b2l_debug_set_flags

struct __align__(8) QControl {
  uint64_t w_full[kMaxRawStages];
  uint64_t w_empty[kMaxRawStages];
  uint64_t x_full[kMaxRawStages];
  uint64_t x_empty[kMaxRawStages];
  uint64_t op_full[kMaxOpStages];
  uint64_t op_empty[kMaxOpStages];
  uint64_t acc_full[4];
  uint64_t acc_empty[4];
  uint64_t tab_full;                       // packed layout: the scale / zero-point tables have landed (bulk copies)
  uint32_t tmem_base;
  int is_last;
  // tiles this CTA finished last (stream-K / split-K): reduced by ALL threads after the role loops
  int nfix;
  int fix_tile[4], fix_ntile[4], fix_sharers[4];
  // shared tiles whose partial sums are written but whose tickets have not been taken yet (see flush_tickets)
  int pend_tile[4], pend_ntile[4], pend_sharers[4];
};
static_assert(sizeof(QControl) <= 1024, "control block too large");

// One CTA's share of the work: pieces (tile, raw-block range) in processing order.  Legacy grid: exactly one piece
// (tile = blockIdx.x, k-range of split blockIdx.z).  Everything that indexes rings, TMEM stages, accumulator buffers
// and the scale table counts raw blocks / segments LINEARLY over the CTA's pieces, so only the code that touches
// global memory (TMA coordinates, table staging, output) looks at the tile.
struct Piece {
  int n_tile;
  int rb_lo, rb_hi;      // raw blocks [rb_lo, rb_hi) of the tile's K range
};
struct WorkRange {
  int npieces;
  int total_rb;          // raw blocks of this CTA over all pieces
  int t0, g0, g1;        // stream-K: first tile, linear raw-block range
  int nrb_tile;
  int legacy_tile, legacy_lo, legacy_hi;
  __device__ __forceinline__ Piece piece(int i) const {
    Piece pc;
    if (nrb_tile == 0) {
      pc.n_tile = legacy_tile; pc.rb_lo = legacy_lo; pc.rb_hi = legacy_hi;
    } else {
      pc.n_tile = t0 + i;
      const int base = pc.n_tile * nrb_tile;
      pc.rb_lo = g0 > base ? g0 - base : 0;
      pc.rb_hi = g1 - base < nrb_tile ? g1 - base : nrb_tile;
    }
    return pc;
  }
};
// first / last stream-K CTA that owns a raw block of tile t
// (32-bit arithmetic: the host only plans stream-K when total * (ctas + 1) fits; a 64-bit division is ~120
// instructions and every thread of the CTA used to do two of them before the first weight request could leave)
__device__ __forceinline__ int sk_owner(int lin_rb, int ctas, int total) {
  return (int)((((unsigned)lin_rb + 1u) * (unsigned)ctas - 1u) / (unsigned)total);
}

// Four TMEM accumulator buffers: the epilogue's drain latency (wake-up, tcgen05.ld, FMAs, arrive) is several
// hundred cycles per segment; with only two buffers the MMA warps stalled on acc_empty (ncu: profiles/).

// A raw stage holds 128 weight rows x 128 BYTES of quantised K (TMA needs full 128-byte rows to stream HBM
// efficiently) = RW operand k-blocks of 64 elements (int4: 256 elements -> 4, 8-bit: 128 -> 2), followed by
// the RW activation sub-tiles [BT x 64] (128B-swizzled, K-major: the UMMA B operand).
template <int QK> __host__ __device__ constexpr int raw_sub() { return QK == Q_INT4 ? 4 : 2; }
template <int QK, int BT> __host__ __device__ constexpr int raw_stage_bytes() {
  return kABytes + raw_sub<QK>() * BT * kStageKBytes;
}

// ---- dequantisers: packed weights -> fp16 pairs in K order (one 32-bit TMEM column per pair) ----
__device__ __forceinline__ uint32_t lop3_and_or(uint32_t a, uint32_t mask, uint32_t orv) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(r) : "r"(a), "r"(mask), "r"(orv));  // (a & mask) | orv
  return r;
}
// one int32 word (8 nibbles) -> 4 half2 {n0,n1} {n2,n3} {n4,n5} {n6,n7}
__device__ __forceinline__ void dequant_int4_word(uint32_t w, __half2 add, uint32_t* out) {
  // byte j holds nibbles (2j, 2j+1).  Replicate the byte into both halves, keep nibble 2j in the low
  // half's mantissa and nibble 2j+1 (still shifted left by 4) in the high half's, OR in the fp16
  // exponent of 1024: low = 1024 + n0, high = 1024 + 16*n1; one HFMA2 maps both back to exact integers.
  const __half2 mul = __halves2half2(__float2half(1.0f), __float2half(0.0625f));
  // add = {-(1024 + z), -(64 + z)}: an integer zero point z is subtracted in the same (exact) HFMA2
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t rep = __byte_perm(w, w, 0x1111 * j);
    const uint32_t h = lop3_and_or(rep, 0x00F0000Fu, 0x64006400u);
    const __half2 v = __hfma2(*reinterpret_cast<const __half2*>(&h), mul, add);
    out[j] = *reinterpret_cast<const uint32_t*>(&v);
  }
}
__device__ __forceinline__ uint32_t cvt_e4m3x2_f16x2(uint16_t v) {
  uint32_t r;
  asm("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(r) : "h"(v));
  return r;
}
// one 32-bit word (4 bytes) -> 2 half2
__device__ __forceinline__ void dequant_w8_word(uint32_t w, bool fp8, uint32_t* out) {
  if (fp8) {
    out[0] = cvt_e4m3x2_f16x2((uint16_t)(w & 0xFFFFu));
    out[1] = cvt_e4m3x2_f16x2((uint16_t)(w >> 16));
  } else {
    // b -> u = b + 128 (xor 0x80), fp16(0x6400 | u) = 1024 + u, minus 1152 = b   (all exact)
    const __half2 off = __float2half2_rn(-1152.0f);
    const uint32_t u = w ^ 0x80808080u;
    const uint32_t h0 = __byte_perm(u, 0x64646464u, 0x4140u);   // [0x64, byte1, 0x64, byte0]
    const uint32_t h1 = __byte_perm(u, 0x64646464u, 0x4342u);
    const __half2 v0 = __hadd2(*reinterpret_cast<const __half2*>(&h0), off);
    const __half2 v1 = __hadd2(*reinterpret_cast<const __half2*>(&h1), off);
    out[0] = *reinterpret_cast<const uint32_t*>(&v0);
    out[1] = *reinterpret_cast<const uint32_t*>(&v1);
  }
}

// registers -> TMEM: lane of the warp writes its own TMEM lane, 32 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem desc]   (A: 128 lanes x 8 columns = 128 x 16 fp16, K-major)
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// same approximate ops as elementwise.cu's swiglu (Triton's lowering of tl.sigmoid on sm_100a)
__device__ __forceinline__ float ex2_approx_q(float x) {
  float r;
  asm("ex2.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float div_full_q(float a, float b) {
  float r;
  asm("div.full.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ float lds_f32(const float* p) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(smem_u32(p)));
  return v;
}
__device__ __forceinline__ uint4 lds_128(const void* p) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_u32(p)));
  return v;
}
__device__ __forceinline__ int ld_acquire_smem(const int* p) {
  int v;
  asm volatile("ld.acquire.cta.shared.b32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_smem(int* p, int v) {
  asm volatile("st.release.cta.shared.b32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}

// Hand-off granularity.  Weights arrive per RAW BLOCK (one 128-byte-wide TMA tile: 256 int4 / 128 8-bit K
// values); the dequantised operand is handed to the MMA warps per OPERAND STAGE of two 64-element sub-blocks
// (64 TMEM columns, 8 MMA steps) through a ring of 6 stages (4 when the token tile is 64).  History (device
// timelines under profiles/): whole-block stages -- only 3 fitted -- stayed occupied for dequant + MMA latency;
// single sub-block stages paid one barrier wait, two tcgen05 fences, one tcgen05.wait::st and one arrive per
// KiB of weights and warp, ~400 of the ~800 cycles a sub-block took.

// MMA issue of one full raw block whose K range holds whole segments of SPS sub-blocks each (scale group divides the
// raw block): one operand stage (two sub-blocks, 8 MMA steps) per elected-lane block, every offset a compile-time
// constant.  The issuing warp shares its scheduler with two dequant warps and gets an issue slot every ~4-6
// cycles, so its instruction count -- not the tensor pipe (16 cycles per MMA at 32 tokens: tools/mma_rate.cu) --
// sets the pace of the accumulator hand-offs: ~25 instructions per sub-block here against ~90 in a rolled loop
// with run-time stage arithmetic (700 cycles per segment on the device timeline, profiles/).
// Operand-ring slot of operand stage `sp` (0 .. SUBST-1) of raw block `rb`.  The ring is split into two halves of
// NST/2 stages: even raw blocks (dequant warps 0-3, MMA warp 0) cycle through the first half, odd raw blocks
// (dequant warps 4-7, MMA warp 1) through the second.  Every op_full / op_empty barrier is then produced and
// waited on by ONE dequant group and ONE MMA warp which both see every one of its phases in order.  (Round 1
// indexed the ring as (2*rb + sp) % 6: a stage alternated between the two groups, so a waiter skipped every
// other phase of "its" barrier and a parity wait could be satisfied by the completion of the phase BEFORE the
// one it needed whenever one group ran a raw block ahead of the other -- stale operands, extra arrivals and,
// eventually, a dead-locked pipeline: the intermittent decode hangs at 16- and 32-token tiles.)
template <int NST, int SUBST>
__device__ __forceinline__ void op_slot(int rb, int sp, int& stage, uint32_t& parity, int debug) {
  constexpr int H = NST / 2;
  static_assert(NST % 2 == 0 && H >= 1, "operand ring must split into two halves");
  if (debug & DBG_LEGACY_OPRING) {            // round-1 indexing, kept in the DBG instantiation to reproduce the hang
    const int gst = rb * SUBST + sp;
    stage = gst % NST;
    parity = (uint32_t)(gst / NST) & 1u;
    return;
  }
  const int u = (rb >> 1) * SUBST + sp;       // running operand-stage count of this raw block's group
  const int q = u / H;
  stage = (rb & 1) * H + (u - q * H);
  parity = (uint32_t)q & 1u;
}

template <int QK, int BT, int NST, int SPS>
__device__ __forceinline__ void issue_block(QControl* ctl, uint32_t tmem_base, uint32_t a_base, uint64_t bdesc0,
                                            int seg0, int rb, uint64_t* x_empty, int debug) {
  constexpr int RW = raw_sub<QK>();
  constexpr int NB = 4;
  constexpr int kXBytes = BT * kStageKBytes;
  constexpr uint32_t idesc = umma_idesc(1u, 0u, 0u, 0u, 0u, kBlockW, BT);
#pragma unroll
  for (int sp = 0; sp < RW / 2; ++sp) {
    int ostage; uint32_t oparity;
    op_slot<NST, RW / 2>(rb, sp, ostage, oparity, debug);
    // segments touched by this stage: sub-blocks 2sp and 2sp+1
    const int segA = seg0 + (2 * sp) / SPS, segB = seg0 + (2 * sp + 1) / SPS;
    if ((2 * sp) % SPS == 0) qwait(&ctl->acc_empty[segA & (NB - 1)], ((segA / NB) & 1) ^ 1, debug);
    if (SPS == 1) qwait(&ctl->acc_empty[segB & (NB - 1)], ((segB / NB) & 1) ^ 1, debug);
    qwait_lat(&ctl->op_full[ostage], oparity, debug);
    tc_fence_after();
    if (elect_one()) {
      const uint32_t a_addr = a_base + ostage * 64;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int sub = 2 * sp + h;
        const int sseg = h ? segB : segA;
        const uint32_t d = tmem_base + (sseg & (NB - 1)) * BT;
        const uint64_t bd = bdesc0 + (uint64_t)((sub * kXBytes) >> 4);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (!(debug & DBG_FEW_MMA) || j == 3)
            umma_f16_ts(d, a_addr + h * 32 + j * 8, bd + 2 * j, idesc, (sub % SPS != 0 || j > 0) ? 1u : 0u);
        if (sub % SPS == SPS - 1) tc_commit(&ctl->acc_full[sseg & (NB - 1)]);
      }
      tc_commit(&ctl->op_empty[ostage]);
      if (sp == RW / 2 - 1) tc_commit(x_empty);
    }
    __syncwarp();
  }
}

// LAYOUT (int4 only): 0 = the reference's tensors as they are; 1 = the load-time repack of b2l_w4a16_pack_* (nibbles
// of a word interleaved so that one LOP3 extracts a K-consecutive pair, scale / zero tables tiled [tile][group][128
// rows] so that a CTA's share is one contiguous bulk copy per piece).
// EPI: 0 = store the projection; 1 = SiLU(gate) * up over row-interleaved gate|up weights (its own instantiation: the
// activation code unrolled over the token tile grew the kernel by ~10 % and that alone slowed every GEMM, profiles/).
template <int QK, int BT, bool DBG, int LAYOUT, int EPI = 0>
__global__ void __launch_bounds__(kQuantThreads, 1)
gemm_quant_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x,
                  const QuantParams p) {
  using T = __half;
  constexpr int OPS = op_stages_for(BT);
  constexpr int kOpStages = OPS;
  constexpr int TP = LAYOUT == 1 ? kBlockW : kTabPitch;   // floats per table column
  // the debug experiments and the timeline stamps only exist in the DBG instantiation (selected by the host when a
  // debug flag or a timeline buffer is set): the product kernel carries none of their code
  const int debug = DBG ? p.debug : 0;
  long long* const timeline = DBG ? p.timeline : nullptr;
  // first thing: the TMA unit starts fetching the weight descriptor while the CTA works out its range (the first
  // weight request otherwise waits for the descriptor AND for the data: two dependent trips to memory)
  if (threadIdx.x == kWarpWProd * 32) tma_prefetch_desc(&tmap_w);
#ifdef B2L_TLITE
  if (threadIdx.x == 0) LT(23);
#endif
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  QControl* ctl = reinterpret_cast<QControl*>(smem);
  // Two rings with different lifetimes.  A raw weight block (128 rows x 128 quantised bytes, from HBM) is free as
  // soon as the dequant warps hold it in registers; the activation block (RW sub-tiles of BT tokens x 64 K, from
  // L2) must stay until its MMAs retire.  One combined stage kept the weight bytes parked for the whole
  // dequant -> MMA latency and starved the HBM stream (5 stages x ~5000 cycles of lifetime = 1000 cycles/block).
  uint8_t* w_ring = smem + 1024;                       // w_stages x 16 KiB
  constexpr int kXBytes = BT * kStageKBytes;           // one activation sub-tile (64 K elements)
  constexpr int RW = raw_sub<QK>();                    // 64-element k-blocks per raw block
  constexpr int kRawK = RW * 64;
  constexpr int kXStage = RW * kXBytes;
  uint8_t* x_ring = w_ring + (size_t)p.w_stages * kABytes;
  // scale (and zero-point) table of this CTA's k-range, TRANSPOSED: [scale group][kTabPitch = 129] fp32 with the weight
  // row as the fast index, so that the row-per-thread readers (dequant, epilogue) are bank-conflict free.  (The first
  // layout, row-major with one cp.async stream per row, made every access a 32-way bank conflict and took 5000
  // cycles just to issue: ncu + device timeline under profiles/.)
  float* scale_tab = reinterpret_cast<float*>(x_ring + (size_t)p.x_stages * kXStage);
  const int tab_floats = (p.max_groups * kTabPitch + 3) & ~3;
  float* zero_tab = scale_tab + tab_floats;
  float* xs_smem = zero_tab + (QK == Q_INT4 ? tab_floats : 0);   // [2][BT] per-segment xsum exchange
  // TMEM map (columns): [NB accumulator buffers of BT] [kOpStages operand stages of 64]
  constexpr int NB = 4;
  constexpr int kAColBase = NB * BT;
  constexpr int kColsUsed = kAColBase + kOpStages * 64;
  constexpr int kTmemCols = kColsUsed <= 32 ? 32 : kColsUsed <= 64 ? 64 : kColsUsed <= 128 ? 128
                          : kColsUsed <= 256 ? 256 : 512;
  static_assert(kColsUsed <= 512, "TMEM budget exceeded");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tile = blockIdx.y;
  long long* tl = (timeline && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) ? timeline : nullptr;
  const int seg_len = p.group_k;                  // K elements per scale group (multiple of 16)
  // ---- this CTA's work: pieces (tile, raw-block range); see struct Piece ----
  WorkRange wr;
  int k_begin, steps_total;                       // legacy grid only: the single piece's first K and its MMA steps
  if (p.sk_ctas > 0) {
    const unsigned total = (unsigned)(p.n_tiles * p.nrb_tile);
    wr.nrb_tile = p.nrb_tile;
    wr.g0 = (int)((blockIdx.x * total) / (unsigned)p.sk_ctas);
    wr.g1 = (int)(((blockIdx.x + 1u) * total) / (unsigned)p.sk_ctas);
    wr.t0 = (int)((unsigned)wr.g0 / (unsigned)p.nrb_tile);
    wr.npieces = (int)((unsigned)(wr.g1 - 1) / (unsigned)p.nrb_tile) - wr.t0 + 1;
    wr.total_rb = wr.g1 - wr.g0;
    wr.legacy_tile = wr.legacy_lo = wr.legacy_hi = 0;
    k_begin = 0;
    steps_total = wr.total_rb * RW * 4;            // stream-K needs K to be a whole number of raw blocks
  } else {
    const int kb0 = blockIdx.z * p.kb_per_split;  // kb_per_split is a multiple of RW
    const int kb1 = kb0 + p.kb_per_split < p.num_kb ? kb0 + p.kb_per_split : p.num_kb;
    wr.nrb_tile = 0; wr.g0 = wr.g1 = wr.t0 = 0;
    wr.npieces = 1;
    wr.legacy_tile = blockIdx.x;
    wr.legacy_lo = kb0 / RW;
    wr.legacy_hi = (kb1 + RW - 1) / RW;
    wr.total_rb = wr.legacy_hi - wr.legacy_lo;
    k_begin = kb0 * 64;
    int k_end = kb1 * 64; if (k_end > p.K) k_end = p.K;
    // segments: maximal runs of 16-element MMA steps inside one scale group and inside [k_begin, k_end)
    steps_total = (k_end - k_begin + 15) / 16;
  }
  const int nrb = wr.total_rb;                    // raw blocks of this CTA
  // two MMA warps (alternate raw blocks) when no segment spans two raw blocks, else warp 1 alone
  const int mma_warps = p.aligned ? 2 : 1;

  if (tl && threadIdx.x == 0) tl[kTlMisc + 0] = clock64();
#ifdef B2L_TLITE
  if (threadIdx.x == 0) {
    LT(0);
    if (p.timeline) {
      unsigned long long gt; unsigned smid;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      p.timeline[(size_t)blockIdx.x * 32 + 16] = (long long)gt;
      p.timeline[(size_t)blockIdx.x * 32 + 18] = (long long)smid;
    }
  }
#endif
  // every CTA: wall-clock (globaltimer, ns) at entry and exit + its SM id -> who is the straggler, and what is the
  // real SM clock (CTA 0 also stamps clock64 at both ends)
  if (timeline && threadIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && blockIdx.x < 1024) {
    unsigned long long gt; unsigned smid;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    timeline[kTlCta + 3 * blockIdx.x] = (long long)gt;
    timeline[kTlCta + 3 * blockIdx.x + 2] = (long long)smid;
  }

  // weight producer state (warp kWarpWProd): the first ring-full of TMA loads is issued BEFORE the set-up barrier
  PipeState ws(p.w_stages);
  int w_pi = 0, w_lin = 0;
  Piece w_pc = wr.piece(0);
  int w_rb = w_pc.rb_lo;
  auto w_issue = [&](bool wait) {
    if (wait) qwait(&ctl->w_empty[ws.stage], ws.phase ^ 1, debug);
    if (elect_one()) {
      mbar_arrive_expect_tx(&ctl->w_full[ws.stage], kABytes);     // OOB parts are zero-filled but still counted
      const int kb = w_rb * RW;                                      // first 64-element k-block of this raw block
      tma_load_2d_hint(w_ring + (size_t)ws.stage * kABytes, &tmap_w, QK == Q_INT4 ? kb * 8 : kb * 64,
                       w_pc.n_tile * kBlockW, &ctl->w_full[ws.stage], kEvictFirst);
    }
    __syncwarp();
    TL_STAMP(TL_PROD, w_lin);
    ws.advance();
    ++w_lin;
    if (++w_rb == w_pc.rb_hi && w_lin < nrb) { w_pc = wr.piece(++w_pi); w_rb = w_pc.rb_lo; }
  };

  // int4: this thread saw a zero point that cannot be folded (or the group size rules folding out altogether)
  int bad_zero = (QK == Q_INT4 && (p.force_xsum || (debug & DBG_FORCE_XSUM) || (LAYOUT == 1 && !p.zeros_foldable))) ? 1 : 0;
  constexpr int kFillWarps = kWarpEpi0 + 4;
  if (LAYOUT == 0 && warp < kFillWarps) {
    // ---- warps 0..11 (idle until the first weights land): stage the scale / zero-point table ----
    // A lane owns table columns `lane` and `lane + 32` (a CTA keeps at most 64 scale groups: host), warp w copies
    // rows w, w+12, ...: global reads are coalesced along a weight row's groups and the transposed shared-memory
    // writes hit 32 different banks.  EVERY load is issued before the first store and before the weight producer
    // is released (named barrier 5): table requests that queue behind a ring-full of weight traffic per SM took
    // ~10000 cycles to come back (device timeline, profiles/).  The tables are weights (static), so all of this
    // runs before griddepcontrol.wait.
    constexpr int kRowsPerWarp = (kBlockW + kFillWarps - 1) / kFillWarps;
    int col_tile[2] = {0, 0}, col_g[2] = {0, 0};
    bool col_ok[2] = {false, false};
    {
      int tcol = 0;                                    // table column = scale group, linear over the pieces
      for (int pi = 0; pi < wr.npieces; ++pi) {
        const Piece pc = wr.piece(pi);
        int g_first, g_count;
        if (p.sk_ctas > 0) {
          g_first = pc.rb_lo * kRawK / seg_len;
          g_count = (pc.rb_hi - pc.rb_lo) * kRawK / seg_len;
        } else {
          g_first = k_begin / seg_len;
          g_count = (k_begin + steps_total * 16 - 1) / seg_len - g_first + 1;
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int tc = lane + 32 * h;
          if (tc >= tcol && tc < tcol + g_count) { col_ok[h] = true; col_tile[h] = pc.n_tile; col_g[h] = g_first + tc - tcol; }
        }
        tcol += g_count;
      }
    }
    // branch-free issue: out-of-range elements read element 0 and are zeroed afterwards (rows past N: benign
    // values); the scale row of weight row n is n / group_n -- a shift for the power-of-two groups in use
    float sv[kRowsPerWarp][2], zv[kRowsPerWarp][2];
    const int gn_shift = (p.group_n & (p.group_n - 1)) == 0 ? 31 - __clz(p.group_n) : -1;
    const float* sbase[2];
    const float* zbase[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      sbase[h] = p.scales + (int64_t)col_g[h] * p.s_stride_k;
      zbase[h] = QK == Q_INT4 ? p.zeros + (int64_t)col_g[h] * p.s_stride_k : nullptr;
    }
#pragma unroll
    for (int i = 0; i < kRowsPerWarp; ++i) {
      const int r = warp + i * kFillWarps;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int n = col_tile[h] * kBlockW + r;
        const bool ok = col_ok[h] && r < kBlockW && n < p.N;
        const int sn = QK == Q_INT4 ? n : (gn_shift >= 0 ? (n >> gn_shift) : n / p.group_n);   // int4: group_n = 1
        sv[i][h] = __ldg(sbase[h] + (int64_t)(ok ? sn : 0) * p.s_stride_n);
        if (QK == Q_INT4) zv[i][h] = __ldg(zbase[h] + (int64_t)(ok ? n : 0) * p.s_stride_n);
      }
    }
#pragma unroll
    for (int i = 0; i < kRowsPerWarp; ++i) {
      const int r = warp + i * kFillWarps;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const bool ok = col_ok[h] && r < kBlockW && col_tile[h] * kBlockW + r < p.N;
        if (!ok) { sv[i][h] = 0.f; if (QK == Q_INT4) zv[i][h] = 0.f; }
      }
    }
    asm volatile("bar.arrive 5, %0;" ::"n"((kFillWarps + 1) * 32) : "memory");   // loads are out: release the weight producer
#pragma unroll
    for (int i = 0; i < kRowsPerWarp; ++i) {
      const int r = warp + i * kFillWarps;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int tc = lane + 32 * h;
        if (col_ok[h] && r < kBlockW) {
          scale_tab[tc * kTabPitch + r] = sv[i][h];
          if (QK == Q_INT4) {
            const float z = zv[i][h];
            zero_tab[tc * kTabPitch + r] = z;
            bad_zero |= !(z >= 0.f && z <= 1023.f && z == rintf(z));
          }
        }
      }
    }
    if (tl && warp == kWarpEpi0 && lane == 0) tl[kTlMisc + 7] = clock64();
  }

  if (LAYOUT == 1 && warp == kWarpXProd && lane == 0) {
    // packed tables: a piece's groups x 128 rows are contiguous in global memory -> one bulk copy per table and
    // piece.  Requested by the activation producer's lane (idle until griddepcontrol.wait; the tables are weights)
    // while the weight producer initialises its ring and sends its first requests: issued one after the other by
    // a single lane, the table copies alone kept the first weight request back by ~0.7 us (per-CTA stamps, profiles/).
    mbar_init(&ctl->tab_full, 1);
    mbar_fence_init();
    fence_proxy_async_smem();
    const int G = p.K / seg_len;
    const int gpr = RW >> p.seg_shift;               // scale groups per raw block (stream-K: the raw block holds whole groups)
    int total_cols = 0;
    for (int pi = 0; pi < wr.npieces; ++pi) {
      const Piece pc = wr.piece(pi);
      if (p.sk_ctas > 0) total_cols += (pc.rb_hi - pc.rb_lo) * gpr;
      else total_cols += (k_begin + steps_total * 16 - 1) / seg_len - k_begin / seg_len + 1;
    }
    mbar_arrive_expect_tx(&ctl->tab_full, (uint32_t)total_cols * kBlockW * 4 * 2);
    int tcol = 0;
    for (int pi = 0; pi < wr.npieces; ++pi) {
      const Piece pc = wr.piece(pi);
      int g_first, g_count;
      if (p.sk_ctas > 0) {
        g_first = pc.rb_lo * gpr;
        g_count = (pc.rb_hi - pc.rb_lo) * gpr;
      } else {
        g_first = k_begin / seg_len;
        g_count = (k_begin + steps_total * 16 - 1) / seg_len - g_first + 1;
      }
      const size_t src = ((size_t)pc.n_tile * G + g_first) * kBlockW;
      bulk_g2s(zero_tab + tcol * kBlockW, p.zeros + src, (uint32_t)g_count * kBlockW * 4, &ctl->tab_full);
      bulk_g2s(scale_tab + tcol * kBlockW, p.scales + src, (uint32_t)g_count * kBlockW * 4, &ctl->tab_full);
      tcol += g_count;
    }
    LT(19);
  }
  if (warp == kWarpWProd) {
    if (lane == 0) {
      // Order matters for the launch latency of the short GEMMs (a CTA of the qkv / o projections converts 3-5 raw
      // blocks in all): only what the first requests need is initialised before they are issued -- the weight ring's
      // barriers, then the first ring-full of weights -- and the other ~30 barriers afterwards (the device timeline
      // showed the first weight request leaving 2400 cycles after kernel entry, most of it this lane's set-up).
      LT(24);
      for (int s = 0; s < p.w_stages; ++s) {
        mbar_init(&ctl->w_full[s], 1);
        mbar_init(&ctl->w_empty[s], 4);        // the 4 dequant warps that convert the block
      }
      ctl->nfix = 0;
      LT(25);
      mbar_fence_init();
      fence_proxy_async_smem();
      LT(26);
      if (tl) tl[kTlMisc + 8] = clock64();
    }
    __syncwarp();
    // Weights are static: under programmatic dependent launch this stream starts while the previous kernel is
    // still running (no griddepcontrol.wait on this path) -- and before the rest of the CTA has finished setting up.
    if (LAYOUT == 0) asm volatile("bar.sync 5, %0;" ::"n"((kFillWarps + 1) * 32) : "memory");
    const int pre = nrb < p.w_stages ? nrb : p.w_stages;
    for (int i = 0; i < pre; ++i) w_issue(false);
    if (lane == 0) LT(1);
    if (lane == 0) {
      // the rest of the set-up, behind the first requests; published to the other warps by the set-up barrier below
      tma_prefetch_desc(&tmap_x);
      for (int s = 0; s < p.x_stages; ++s) {
        mbar_init(&ctl->x_full[s], 1);
        mbar_init(&ctl->x_empty[s], 1);        // tcgen05.commit of the MMA warp that issued the block
      }
      for (int s = 0; s < kOpStages; ++s) {
        mbar_init(&ctl->op_full[s], 4);
        mbar_init(&ctl->op_empty[s], 1);
      }
      for (int s = 0; s < NB; ++s) {
        mbar_init(&ctl->acc_full[s], 1);
        mbar_init(&ctl->acc_empty[s], 4);
      }
      mbar_fence_init();
    }
    __syncwarp();
  }
  if (warp == kWarpMma0) {
    tmem_alloc(&ctl->tmem_base, kTmemCols);
    tmem_relinquish();
    // The next kernel may set itself up while this one streams weights (common.cuh).  The CTA's trigger is issued by
    // the warp that owns the tensor memory, AFTER the allocation: a dependent CTA that became co-resident before it
    // could take the SM's TMEM columns and park in griddepcontrol.wait for this grid, while this CTA blocks in
    // tcgen05.alloc (dead-lock).  No other warp triggers, so the CTA counts as triggered only from here on.
    pdl_launch_dependents();
    if (tl && lane == 0) tl[kTlMisc + 9] = clock64();
  }
  if (tl && threadIdx.x == 0) tl[kTlMisc + 10] = clock64();
  tc_fence_before();
  // the set-up barrier also publishes the table and ORs the "zero point cannot be folded" flags of the fill threads
  const bool cta_need_xsum = __syncthreads_or(bad_zero) != 0;
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;
  if (tl && threadIdx.x == 0) tl[kTlMisc + 1] = clock64();
  if (threadIdx.x == 0) LT(2);

  if (warp == kWarpWProd) {
    // ===================== weight producer (whole warp loops, one elected lane issues) =====================
    while (w_lin < nrb) w_issue(true);
  } else if (warp == kWarpXProd) {
    // ===================== activation producer =====================
    pdl_wait_prior();                        // the activations are the previous kernel's output
    if (lane == 0) LT(3);
    PipeState xs(p.x_stages);
    for (int pi = 0; pi < wr.npieces; ++pi) {
      const Piece pc = wr.piece(pi);
      for (int rb = pc.rb_lo; rb < pc.rb_hi; ++rb) {
        qwait(&ctl->x_empty[xs.stage], xs.phase ^ 1, debug);
        if (debug & DBG_NO_XLOAD) {
          if (elect_one()) mbar_arrive(&ctl->x_full[xs.stage]);
        } else if (elect_one()) {
          mbar_arrive_expect_tx(&ctl->x_full[xs.stage], kXStage);
          const int kb = rb * RW;
#pragma unroll
          for (int sub = 0; sub < RW; ++sub)
            tma_load_2d_hint(x_ring + (size_t)xs.stage * kXStage + sub * kXBytes, &tmap_x, (kb + sub) * 64, m_tile * BT,
                             &ctl->x_full[xs.stage], kEvictLast);
        }
        __syncwarp();
        xs.advance();
      }
    }
  } else if (warp == kWarpMma0 || warp == kWarpMma1) {
    // ===================== MMA issuers =====================
    const int mma_idx = kWarpMma0 - warp;             // 0 or 1
    if (mma_idx < mma_warps) {
      constexpr uint32_t idesc = umma_idesc(1u, 0u, 0u, 0u, 0u, kBlockW, BT);
      const uint64_t x_desc0 = umma_desc_sw128_kmajor(smem_u32(x_ring));
      const uint32_t a_base = tmem_base + kAColBase;
      const int seg_steps = seg_len / 16;
      // running segment state (only the elected lane's copy is ever used; election always picks the same lane)
      int seg = 0;
      int seg_left = ((k_begin / seg_len + 1) * seg_len - k_begin) / 16;   // steps until the open segment closes
      if (seg_left > steps_total) seg_left = steps_total;
      bool seg_open = false;
      int rstage = mma_idx % p.x_stages;
      uint32_t rphase = (mma_idx / p.x_stages) & 1;
      for (int rb = mma_idx; rb < nrb; rb += mma_warps) {
        const int gsub0 = rb * RW;                                       // first operand sub-block of this raw block
        TL_STAMP(7, rb);
        qwait(&ctl->x_full[rstage], rphase, debug);                         // activation sub-tiles (TMA)
        tc_fence_after();
        TL_STAMP(TL_MMA_READY, rb);
        if (rb == 0 && lane == 0) LT(4);
        int steps_here = steps_total - rb * (RW * 4);
        if (steps_here > RW * 4) steps_here = RW * 4;
        if (mma_warps == 2) {
          // raw blocks hold whole segments in this mode: state restarts at the block boundary
          seg = rb * (kRawK / seg_len);
          seg_left = seg_steps < steps_here ? seg_steps : steps_here;
          seg_open = false;
        }
        const uint64_t bdesc0 = x_desc0 + (uint64_t)((rstage * kXStage) >> 4);
        if (mma_warps == 2 && steps_here == RW * 4) {
          // aligned mode: the raw block holds 1 or 2 whole segments
          if (seg_steps == RW * 4)
            issue_block<QK, BT, OPS, RW>(ctl, tmem_base, a_base, bdesc0, seg, rb, &ctl->x_empty[rstage], debug);
          else
            issue_block<QK, BT, OPS, RW / 2>(ctl, tmem_base, a_base, bdesc0, seg, rb, &ctl->x_empty[rstage], debug);
        } else if (elect_one()) {
#pragma unroll
          for (int st = 0; st < RW * 4; ++st) {
            if (st < steps_here) {
              const int buf = seg % NB;
              int ostage; uint32_t oparity;                      // operand stage = two sub-blocks
              op_slot<kOpStages, RW / 2>(rb, st >> 3, ostage, oparity, debug);
              if ((st & 7) == 0) {
                qwait_lat(&ctl->op_full[ostage], oparity, debug);   // dequantised stage
                tc_fence_after();
              }
              if (!seg_open) {
                // accumulator buffer `buf` must have been drained (segment seg-NB) by all four epilogue warps
                qwait(&ctl->acc_empty[buf], ((seg / NB) & 1) ^ 1, debug);
                tc_fence_after();
              }
              const uint64_t bdesc = bdesc0 + (uint64_t)(((st >> 2) * kXBytes) >> 4) + 2 * (st & 3);
              umma_f16_ts(tmem_base + buf * BT, a_base + ostage * 64 + (st & 7) * 8, bdesc, idesc, seg_open ? 1u : 0u);
              if ((st & 7) == 7 || st == steps_here - 1) tc_commit(&ctl->op_empty[ostage]);
              --seg_left;
              if (seg_left == 0) {
                tc_commit(&ctl->acc_full[buf]);
                ++seg;
                seg_open = false;
                // the next segment may extend beyond this raw block (single-warp mode); it is cut by the k-range end only
                const int total_left = steps_total - rb * (RW * 4) - (st + 1);
                seg_left = seg_steps < total_left ? seg_steps : total_left;
              } else {
                seg_open = true;
              }
            }
          }
          tc_commit(&ctl->x_empty[rstage]);
        }
        __syncwarp();
        TL_STAMP(TL_MMA_ISSUED, rb);
        if (rb == 0 && lane == 0) LT(5);
        if (rb + mma_warps >= nrb && lane == 0) LT(6 + mma_idx);
        rstage += mma_warps;
        if (rstage >= p.x_stages) { rstage -= p.x_stages; rphase ^= 1; }
      }
    }
  } else if (warp >= kWarpEpi0) {
    // ===================== epilogue (warps 8..11) =====================
    const int quad = warp & 3;
    const int epi_tid = quad * 32 + lane;
    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
    float acc[BT];
#pragma unroll
    for (int j = 0; j < BT; ++j) acc[j] = 0.f;

    // int4: an integer zero point is subtracted exactly by the dequant warps (same HFMA2 that removes the fp16
    // magic offset), which leaves a scale-only epilogue.  A zero point that is not an integer in [0, 1023]
    // cannot be folded exactly: then the whole CTA falls back to  acc += s * (P - z * xsum)  with xsum = the
    // segment's activation row sums (computed here from global memory; rare, correctness path only).  The
    // decision was taken by the table-fill threads and ORed across the CTA by the set-up barrier.
    const bool need_xsum = QK == Q_INT4 && cta_need_xsum;
    if (p.sk_ctas > 0 && warp == kWarpEpi0) {
      // The reducer of a shared tile (see the end of the kernel) reads the other sharers' workspace slots exactly once,
      // at the very end, with every other warp waiting for it: the first access of this SM to those pages took ~1.3 us
      // whatever the amount of data (per-CTA stamps, profiles/) -- an address translation miss.  One throw-away
      // gpu-scope load per slot now, while this warp has nothing to do, takes the miss off the critical path (no L1
      // allocation: the slots are polled with ordinary loads later).  The value is not used, so it does not matter
      // that the previous kernel may still be clearing the words.
      const Piece lp = wr.piece(wr.npieces - 1);
      const int total = p.n_tiles * p.nrb_tile;
      const int c_lo = sk_owner(lp.n_tile * p.nrb_tile, p.sk_ctas, total);
      const int c_hi = sk_owner((lp.n_tile + 1) * p.nrb_tile - 1, p.sk_ctas, total);
      if (c_lo == (int)blockIdx.x && lane > 0 && lane <= c_hi - c_lo) {
        const uint32_t* a = p.ws + ((size_t)lp.n_tile * p.sk_slots + lane) * (BT * kBlockW);
        uint32_t d;
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(d) : "l"(a) : "memory");
      }
    }
    if (LAYOUT == 1) mbar_wait(&ctl->tab_full, 0);
    if (need_xsum) pdl_wait_prior();   // the fallback reads the activations (the previous kernel's output) from global
    if (tl && warp == kWarpEpi0 && lane == 0) tl[kTlMisc + 3] = clock64();
    constexpr int PARTS = 128 / BT;
    const int xs_m = epi_tid / PARTS, xs_part = epi_tid % PARTS;
    T* y = reinterpret_cast<T*>(p.y);

    int seg = 0;                                           // segments and table columns run linearly over the pieces
    int tcol = 0;
    int npend = 0;                                         // shared tiles written, tickets outstanding (uniform)
    auto flush_tickets = [&]() {
      __threadfence();                                     // every partial tile written so far is visible device-wide ...
      asm volatile("bar.sync 2, 128;" ::: "memory");       // ... for all 128 writers, before any ticket is taken
      if (epi_tid < npend) {
        // Last finisher of a tile: its partial sums are added up by ALL 512 threads once the role loops are over (end
        // of the kernel): 128 epilogue threads walking the slots made the tail of the short GEMMs ~5 us long.
        const int tile = ctl->pend_tile[epi_tid], sh = ctl->pend_sharers[epi_tid];
        const int old = atomicAdd(&p.counters[tile], 1);
        if (old == sh - 1) {
          p.counters[tile] = 0;                            // self-reset: the next launch (or graph replay) starts clean
          const int f = atomicAdd(&ctl->nfix, 1);
          if (f >= 4) __trap();                            // cannot happen: at most two shared pieces per CTA
          ctl->fix_tile[f] = tile; ctl->fix_ntile[f] = ctl->pend_ntile[epi_tid]; ctl->fix_sharers[f] = sh;
        }
      }
      asm volatile("bar.sync 2, 128;" ::: "memory");
      npend = 0;
    };
    for (int pi = 0; pi < wr.npieces; ++pi) {
      const Piece pc = wr.piece(pi);
      const int n = pc.n_tile * kBlockW + epi_tid;
      const bool n_ok = n < p.N;
      // enumerate this piece's segments exactly like the MMA warps do
      int k_pos, k_lim;
      if (p.sk_ctas > 0) { k_pos = pc.rb_lo * kRawK; k_lim = pc.rb_hi * kRawK; }
      else { k_pos = k_begin; k_lim = k_begin + steps_total * 16; }
      int k_group_end = (k_pos / seg_len + 1) * seg_len;   // end of the open segment's scale group
      while (k_pos < k_lim) {
        const int k_next = k_group_end < k_lim ? k_group_end : k_lim;
        if (QK == Q_INT4 && need_xsum) {
          // thread -> token m = tid / PARTS, an interleaved share of the segment's 8-element chunks
          float xsum_m = 0.f;
          const int m = m_tile * BT + xs_m;
          if (m < p.M) {
            const __half* xrow = reinterpret_cast<const __half*>(p.x) + (int64_t)m * p.ldx;
            for (int k = k_pos + xs_part * 8; k < k_next; k += PARTS * 8) {
              if (k + 8 <= p.K) {
                const uint4 v = ldg_128(xrow + k);
                const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const float2 f = __half22float2(h[q]);
                  xsum_m += f.x + f.y;
                }
              } else {
                for (int kk = k; kk < p.K; ++kk) xsum_m += __half2float(xrow[kk]);
              }
            }
          }
#pragma unroll
          for (int o = 1; o < PARTS && o < 32; o <<= 1) xsum_m += __shfl_xor_sync(0xffffffffu, xsum_m, o);
          // exchange: every thread needs all BT sums
          float* xs = xs_smem + (seg & 1) * BT;
          if (xs_part == 0) xs[xs_m] = xsum_m;
          asm volatile("bar.sync 3, 128;" ::: "memory");
        }
        const int buf = seg % NB;
        const float sc = lds_f32(scale_tab + tcol * TP + epi_tid);        // rows past N hold zeros
        float zp = 0.f;
        if (QK == Q_INT4 && need_xsum) zp = lds_f32(zero_tab + tcol * TP + epi_tid);
        qwait_lat(&ctl->acc_full[buf], (seg / NB) & 1, debug);
        tc_fence_after();
        if (warp == kWarpEpi0) TL_STAMP(TL_EPI_START, seg);
        if (warp == kWarpEpi0 && lane == 0 && seg == 0) LT(12);
        if ((debug & DBG_SKIP_DRAIN) && k_next < k_lim) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&ctl->acc_empty[buf]);
          k_pos = k_next; k_group_end += seg_len; ++tcol; ++seg;
          continue;
        }
        const uint32_t ta = taddr + buf * BT;
        constexpr int CH = BT < 32 ? BT : 32;                  // columns drained per tcgen05.wait::ld
        const float4* xs4 = reinterpret_cast<const float4*>(xs_smem + (seg & 1) * BT);
#pragma unroll
        for (int c0 = 0; c0 < BT; c0 += CH) {
          uint32_t pv[CH];
#pragma unroll
          for (int c = 0; c < CH; c += 16) tmem_ld_32x32b_x16(ta + c0 + c, *reinterpret_cast<uint32_t(*)[16]>(pv + c));
          tc_wait_ld();
          if (c0 + CH == BT) {
            // the accumulator is in registers: hand the buffer back before doing the arithmetic
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&ctl->acc_empty[buf]);
          }
          if (QK == Q_INT4 && need_xsum) {
#pragma unroll
            for (int j4 = 0; j4 < CH / 4; ++j4) {
              const float4 x4 = xs4[c0 / 4 + j4];
              float* a4 = acc + c0 + 4 * j4;
              a4[0] = fmaf(sc, fmaf(-zp, x4.x, __uint_as_float(pv[4 * j4 + 0])), a4[0]);
              a4[1] = fmaf(sc, fmaf(-zp, x4.y, __uint_as_float(pv[4 * j4 + 1])), a4[1]);
              a4[2] = fmaf(sc, fmaf(-zp, x4.z, __uint_as_float(pv[4 * j4 + 2])), a4[2]);
              a4[3] = fmaf(sc, fmaf(-zp, x4.w, __uint_as_float(pv[4 * j4 + 3])), a4[3]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < CH; ++j) acc[c0 + j] = fmaf(sc, __uint_as_float(pv[j]), acc[c0 + j]);
          }
        }
        if (warp == kWarpEpi0) TL_STAMP(TL_EPI_DONE, seg);
        k_pos = k_next;
        k_group_end += seg_len;
        ++tcol;
        ++seg;
      }

      // ---- the piece is complete: write the tile, or hand the partial sum to the tile's last finisher ----
      if (tl && warp == kWarpEpi0 && lane == 0 && pi == wr.npieces - 1) tl[kTlMisc + 5] = clock64();
      if (warp == kWarpEpi0 && lane == 0 && pi == wr.npieces - 1) LT(13);
      pdl_wait_prior();              // output / workspace buffers may still be in use by the previous kernel
      auto store_final = [&]() {
        float bv = 0.f;
        if (n_ok && p.bias) bv = p.bias_is_f32 ? reinterpret_cast<const float*>(p.bias)[n]
                                               : __half2float(reinterpret_cast<const __half*>(p.bias)[n]);
        if (EPI == 1) {
          // SiLU(gate) * up in the epilogue (reference: two GEMMs + swiglu_forward, models/base.py:259-264,
          // kernels/swiglu.py:45-65).  The weight rows were interleaved at load time (row 2i = gate_i, 2i+1 = up_i), so
          // the pair sits in neighbouring lanes.  Both projections are rounded to fp16 first -- exactly the values the
          // unfused path writes to memory -- and the activation uses the swiglu kernel's arithmetic: bit-identical.
#pragma unroll
          for (int j = 0; j < BT; ++j) {
            const float mine = __half2float(__float2half_rn(acc[j] + bv));
            const float other = __shfl_xor_sync(0xffffffffu, mine, 1);
            const int m = m_tile * BT + j;
            if (n_ok && (epi_tid & 1) == 0 && m < p.M) {
              const float e = ex2_approx_q(-mine * 1.4426950408889634f);
              const float sl = mine * div_full_q(1.0f, 1.0f + e);
              y[(int64_t)m * p.ldy + (n >> 1)] = __float2half_rn(sl * other);
            }
          }
          return;
        }
        if (!n_ok) return;
        // one running pointer and one row-count test for the whole tile: this runs on the CTA's critical path (the
        // per-row index arithmetic and bounds tests made the tile store ~600 instructions per thread, 1.3 us)
        T* yp = y + (int64_t)(m_tile * BT) * p.ldy + n;
        const int rows = p.M - m_tile * BT;
        if (rows >= BT) {
#pragma unroll
          for (int j = 0; j < BT; ++j) { *yp = __float2half_rn(acc[j] + bv); yp += p.ldy; }
        } else {
#pragma unroll
          for (int j = 0; j < BT; ++j) {
            if (j < rows) *yp = __float2half_rn(acc[j] + bv);
            yp += p.ldy;
          }
        }
      };
      // who shares this tile: legacy grid = the p.splits CTAs along grid.z; stream-K = the CTAs whose ranges
      // overlap the tile's raw blocks, slot = position among them (CTA order => deterministic sum order)
      int sharers, slot, slots_per_tile;
      if (p.sk_ctas > 0) {
        const int total = p.n_tiles * p.nrb_tile;
        const int c_lo = sk_owner(pc.n_tile * p.nrb_tile, p.sk_ctas, total);
        const int c_hi = sk_owner((pc.n_tile + 1) * p.nrb_tile - 1, p.sk_ctas, total);
        sharers = c_hi - c_lo + 1;
        slot = (int)blockIdx.x - c_lo;
        slots_per_tile = p.sk_slots;
      } else {
        sharers = p.splits; slot = blockIdx.z; slots_per_tile = p.splits;
      }
      if (sharers == 1) {
        store_final();
      } else if (p.sk_ctas > 0) {
        // ---- stream-K: flag-in-data hand-over to the tile's first sharer ----
        // A partial value travels as ONE 32-bit word whose lowest mantissa bit is forced to 1 ("present"; a cleared
        // word is 0): a reader that sees the bit has the value, so the writer needs neither a fence nor a ticket --
        // it stores and moves on.  (The bit costs <= 2^-24 relative on a partial sum, 2^13 times below the fp16
        // output's resolution, and is the same in every run: results stay deterministic.)  The reducer is the CTA of
        // slot 0: the tile is the LAST piece of its range (its range starts at or before the tile and ends inside
        // it), so it gets there when everybody else is about done, keeps its own partial sum on chip and reads the
        // other slots once all 512 threads are free (end of the kernel), clearing the words behind it: the workspace
        // is all-zero between launches (contract of the C ABI, like the ticket counters).  Ticket protocol before:
        // fence + atomic (two L2 round trips) on every sharer, then up to two serial tile reductions on whoever came
        // last -- 3 to 5 us on the grid's critical path (per-CTA stamps, profiles/r2n_tlite*.txt).  All
        // sk_ctas <= #SM CTAs are co-resident, so waiting for a sharer cannot dead-lock.
        uint32_t* base32 = p.ws + (size_t)pc.n_tile * slots_per_tile * (BT * kBlockW);
        if (slot != 0) {
          uint32_t* mine = base32 + (size_t)slot * (BT * kBlockW) + epi_tid;
#pragma unroll
          for (int j = 0; j < BT; ++j) mine[j * kBlockW] = __float_as_uint(acc[j]) | 1u;
        } else {
          float* own = reinterpret_cast<float*>(w_ring);     // every raw block has been consumed: the ring is idle
#pragma unroll
          for (int j = 0; j < BT; ++j) own[j * kBlockW + epi_tid] = acc[j];
          if (epi_tid == 0) {
            ctl->fix_tile[0] = pc.n_tile; ctl->fix_ntile[0] = pc.n_tile; ctl->fix_sharers[0] = sharers;
            ctl->nfix = 1;
          }
        }
      } else {
        const int tile_id = m_tile * p.n_tiles + pc.n_tile;
        uint32_t* base = p.ws + (size_t)tile_id * slots_per_tile * (BT * kBlockW);
        uint32_t* mine = base + (size_t)slot * (BT * kBlockW);
#pragma unroll
        for (int j = 0; j < BT; ++j) mine[j * kBlockW + epi_tid] = __float_as_uint(acc[j]);
        // The ticket is NOT taken here: fence + atomic are two L2 round trips (~1.4 us, per-CTA stamps under profiles/)
        // during which no accumulator was drained -- a CTA whose range crosses a tile boundary stalled its MMA warps
        // in the middle of the kernel and became the straggler of the grid.  The tile is queued and all tickets are
        // taken together after the last piece (one fence, overlapping atomics).  At most two pieces of a range are
        // shared (its first and last tile), the queue holds four.
        if (epi_tid == 0) { ctl->pend_tile[npend] = tile_id; ctl->pend_ntile[npend] = pc.n_tile; ctl->pend_sharers[npend] = sharers; }
        if (++npend == 4) flush_tickets();
      }
#pragma unroll
      for (int j = 0; j < BT; ++j) acc[j] = 0.f;
    }
    if (npend > 0) flush_tickets();
    if (warp == kWarpEpi0 && lane == 0) LT(14);
  } else {
    // ===================== dequant warps 0..7: thread = one weight row of every second raw block =====================
    // A warp may only touch the TMEM lane quadrant (warp id % 4).  Warps 0..3 convert the even raw blocks,
    // warps 4..7 the odd ones: the two dequant warps that share a scheduler then sit in different phases
    // (one waits for its shared-memory loads or its tcgen05.st while the other keeps the ALU pipe busy).
    const int half = warp >> 2;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int rx = row & 7;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    bool fold_zero = false;
    const int nsub_total = (steps_total + 3) / 4;          // operand sub-blocks (64 K elements) with at least one MMA step
    // table column of a 64-element sub-block = its scale group counted from the CTA's first one; under stream-K every
    // piece starts on a raw-block (= group) boundary, so the count simply runs on across pieces
    const int sub_first = p.sk_ctas > 0 ? 0 : k_begin / 64;
    const int g_first = p.sk_ctas > 0 ? 0 : k_begin / seg_len;
    const int g_last = (k_begin + steps_total * 16 - 1) / seg_len - k_begin / seg_len;
    if (QK == Q_INT4) fold_zero = !cta_need_xsum;
    if (LAYOUT == 1) mbar_wait(&ctl->tab_full, 0);
    if (tl && warp == 0 && lane == 0) tl[kTlMisc + 4] = clock64();
    // ring position of this group's next raw block, advanced by 2 per iteration (ring depths are even, so no
    // integer division -- it cost ~30 ALU-pipe instructions per raw block on the pipe that bounds this kernel)
    int rstage = half;
    uint32_t rphase = 0;
    for (int rb = half; rb < nrb; rb += 2) {
      // zero points of this block's 64-element sub-blocks (each lies in one scale group), fetched before the barrier
      // waits so that the table latency is off the conversion's critical path.  Reference layout: one add
      // {-(1024+z), -(64+z)} per (low nibble, high nibble) pair; packed layout: add_lo = {-(1024+z)} x2 for the
      // low-nibble pairs, add_hi = {-(64+z)} x2 for the high-nibble pairs.
      __half2 add_lo[RW], add_hi[RW];
      if (QK == Q_INT4) {
#pragma unroll
        for (int sub = 0; sub < RW; ++sub) {
          float z = 0.f;
          if (fold_zero) {
            int g = ((sub_first + rb * RW + sub) >> p.seg_shift) - g_first;
            g = g < g_last ? g : g_last;                     // blocks past K are multiplied by zero activations
            z = lds_f32(zero_tab + g * TP + row);
          }
          if (LAYOUT == 1) {
            add_lo[sub] = __float2half2_rn(-1024.f - z);
            add_hi[sub] = __float2half2_rn(-64.f - z);
          } else {
            add_lo[sub] = __floats2half2_rn(-1024.f - z, -64.f - z);
            add_hi[sub] = add_lo[sub];
          }
        }
      }
      qwait(&ctl->w_full[rstage], rphase, debug);
      const uint8_t* rawrow = w_ring + (size_t)rstage * kABytes + row * 128;
      uint4 raw[8];                                          // the row's 128 raw bytes (8 swizzled 16-byte chunks)
#pragma unroll
      for (int c = 0; c < 8; ++c) raw[c] = lds_128(rawrow + ((c ^ rx) << 4));
      // The weight stage is handed back to the TMA producer at the END of this block, not here: at this point the
      // loads have been issued, not returned, and nothing that follows depends on them until the conversion --
      // a refill that found its weights in L2 could land in the stage first (rows are written in order, so the
      // lowest rows of a tile were the ones hit).  Found in round 2 by tests/test_kernels_gpu.py::
      // test_w4a16_streamk_back_to_back_launches_are_isolated at 33-64 tokens: ~10 % of the launches over a long-K
      // shape had one or a few output columns off by one scale group's contribution (tools/nondet_probe.py).
      if (quad == 0) TL_STAMP(TL_DQ_START, rb);
      if (warp == 0 && lane == 0 && rb == 0) LT(8);
      // one operand sub-stage per 64-element sub-block; the "sub-block ready" arrive of sub-block s is issued
      // after the conversion of s+1, so tcgen05.wait::st never waits for a store that was just issued
      int pending = -1;
#pragma unroll
      for (int sub = 0; sub < RW; ++sub) {
        const int gsub = rb * RW + sub;
        constexpr int CPS = 8 / RW;                            // 16-byte chunks of the row per sub-block
        // DBG: five clock stamps per sub-block of warp 0 (loop top, operand stage free, converted, previous store
        // retired + announced, store issued): tools/timeline.py prints the phase lengths
        long long* const tls = (DBG && tl && warp == 0 && lane == 0 && gsub < 96) ? tl + kTlSub + gsub * 5 : nullptr;
        if (tls) tls[0] = clock64();
        if (gsub < nsub_total) {                               // sub-blocks past the k-range have no MMA step
          int ostage; uint32_t oparity;                        // operand stage = sub-blocks (2i, 2i+1); rb * RW is even
          op_slot<kOpStages, RW / 2>(rb, sub >> 1, ostage, oparity, debug);
          if (B2L_DQ_ORDER == 0 && (sub & 1) == 0) {
            qwait_lat(&ctl->op_empty[ostage], oparity ^ 1, debug);
            tc_fence_after();
          }
          if (tls) tls[1] = clock64();
          uint32_t vals[32];
          if (QK == Q_INT4) {
            const uint4 w0 = raw[sub * CPS], w1 = raw[sub * CPS + CPS - 1];
            const uint32_t ws[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
            if (debug & DBG_RAW_OPERAND) {
#pragma unroll
              for (int j = 0; j < 8; ++j) { vals[4 * j] = ws[j]; vals[4 * j + 1] = ws[j] ^ 1u; vals[4 * j + 2] = ws[j] ^ 2u; vals[4 * j + 3] = ws[j] ^ 3u; }
            } else if (LAYOUT == 1) {
              // packed word: nibble positions hold K indices [0,2,4,6,1,3,5,7] -> each LOP3 yields a K-consecutive
              // pair; low-nibble pairs are 1024 + n, high-nibble pairs 1024 + 16 n (5 ALU-pipe instructions per
              // word instead of 8)
              const __half2 one = __float2half2_rn(1.0f), sixteenth = __float2half2_rn(0.0625f);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const uint32_t w8 = ws[j] >> 8;
                const uint32_t h0 = lop3_and_or(ws[j], 0x000F000Fu, 0x64006400u);
                const uint32_t h1 = lop3_and_or(ws[j], 0x00F000F0u, 0x64006400u);
                const uint32_t h2 = lop3_and_or(w8, 0x000F000Fu, 0x64006400u);
                const uint32_t h3 = lop3_and_or(w8, 0x00F000F0u, 0x64006400u);
                const __half2 v0 = __hfma2(*reinterpret_cast<const __half2*>(&h0), one, add_lo[sub]);
                const __half2 v1 = __hfma2(*reinterpret_cast<const __half2*>(&h1), sixteenth, add_hi[sub]);
                const __half2 v2 = __hfma2(*reinterpret_cast<const __half2*>(&h2), one, add_lo[sub]);
                const __half2 v3 = __hfma2(*reinterpret_cast<const __half2*>(&h3), sixteenth, add_hi[sub]);
                vals[4 * j] = *reinterpret_cast<const uint32_t*>(&v0);
                vals[4 * j + 1] = *reinterpret_cast<const uint32_t*>(&v1);
                vals[4 * j + 2] = *reinterpret_cast<const uint32_t*>(&v2);
                vals[4 * j + 3] = *reinterpret_cast<const uint32_t*>(&v3);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) dequant_int4_word(ws[j], add_lo[sub], vals + 4 * j);
            }
          } else {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const uint4 w = raw[(sub * CPS + c) % 8];
              const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) dequant_w8_word(ws[j], QK == Q_FP8, vals + 8 * c + 2 * j);
            }
          }
          if (tls) {
            long long t;
            asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) : "r"(vals[0]), "r"(vals[15]), "r"(vals[31]) : "memory");
            tls[2] = t;
          }
          // the previous stage is announced after the conversion of this stage's first half: by then its
          // tcgen05.st have retired and tcgen05.wait::st does not wait
          if ((sub & 1) == 0 && pending >= 0) {
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&ctl->op_full[pending]);
            pending = -1;
          }
          // the operand stage is only needed for the store: waiting for it here, after the conversion and after the
          // previous stage has been announced, keeps a late stage from delaying either
          if (B2L_DQ_ORDER == 1 && (sub & 1) == 0) {
            qwait_lat(&ctl->op_empty[ostage], oparity ^ 1, debug);
            tc_fence_after();
          }
          if (tls) tls[3] = clock64();
          if (!(debug & DBG_NO_STTM)) tmem_st_32x32b_x32(lane_base + kAColBase + ostage * 64 + (sub & 1) * 32, vals);
          else asm volatile("" ::"r"(vals[0] ^ vals[7] ^ vals[13] ^ vals[31]));
          if (tls) tls[4] = clock64();
          if ((sub & 1) == 1 || gsub == nsub_total - 1) pending = ostage;
        }
      }
      if (pending >= 0) {
        tc_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ctl->op_full[pending]);
      }
      // every raw word has been converted: the stage can be refilled
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl->w_empty[rstage]);
      if (quad == 0) TL_STAMP(TL_DQ_DONE, rb);
      if (warp == 0 && lane == 0 && rb == 0) LT(9);
      if (quad == 0 && lane == 0 && rb + 2 >= nrb) LT(10 + half);
      rstage += 2;
      if (rstage >= p.w_stages) { rstage -= p.w_stages; rphase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
#ifdef B2L_TLITE
  if (threadIdx.x == 0 && p.timeline) {     // (the clock read is tied to a value loaded after the barrier: ptxas hoists it otherwise)
    long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) : "r"(ctl->nfix) : "memory");
    p.timeline[(size_t)blockIdx.x * 32 + 15] = t;
  }
#endif
  // ---- split tiles this CTA finished last: out = sum of the partial tiles in slot order (deterministic) + bias ----
  {
    const int nfix = ctl->nfix;
    const int slots_per_tile = p.sk_ctas > 0 ? p.sk_slots : p.splits;
    constexpr int JT = kQuantThreads / kBlockW;            // token rows covered per pass (4)
    constexpr int NJ = BT / JT;
    const int n_local = threadIdx.x % kBlockW, j_first = threadIdx.x / kBlockW;
    if (p.sk_ctas == 0 && nfix > 0) __threadfence();
    for (int f = 0; f < nfix; ++f) {
      const int sharers = ctl->fix_sharers[f];
      float sum[NJ];
      if (p.sk_ctas > 0) {
        // ---- stream-K: this CTA is slot 0 of the tile; its own partial sum is in shared memory, the other slots'
        // arrive as tagged words (see the epilogue).  Up to SB slots are requested together: the reduction is bound by
        // the bytes one SM can pull out of L2 (~1.2 us for 128 KiB when the words were 8 bytes wide), not by latency.
        // A word whose tag bit is not there yet is simply read again.  Sum in slot order, then clear what was read.
        uint32_t* base = p.ws + (size_t)ctl->fix_tile[f] * slots_per_tile * (BT * kBlockW) + n_local;
        const float* own = reinterpret_cast<const float*>(w_ring);
#pragma unroll
        for (int i = 0; i < NJ; ++i) sum[i] = own[(j_first + i * JT) * kBlockW + n_local];
#ifdef B2L_TLITE
        if (threadIdx.x == 0 && p.timeline) {
          long long t;
          asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) : "f"(sum[0]) : "memory");
          p.timeline[(size_t)blockIdx.x * 32 + 28] = t;
          p.timeline[(size_t)blockIdx.x * 32 + 29] = 0;
        }
#endif
        constexpr int SB = NJ <= 8 ? 6 : 3;
        for (int s0 = 1; s0 < sharers; s0 += SB) {
          uint32_t v[SB][NJ];
          unsigned long long t0 = 0ull;
          for (;;) {
            // All requests first, the tag bits afterwards.  gpu-scope loads (served by L2, never by L1): the same
            // workspace words are read by whichever CTA reduces this tile in the NEXT launch, possibly on this very SM,
            // and L1 is not invalidated between programmatically dependent launches -- an ordinary load could return
            // the previous launch's tagged words.  (Ordinary loads were not faster: the ~1.3 us is latency.)  A retry
            // was never needed in the measured runs: the reducer's tile is its last piece.
#pragma unroll
            for (int u = 0; u < SB; ++u) {
              const bool on = s0 + u < sharers;
#pragma unroll
              for (int i = 0; i < NJ; ++i) {
                uint32_t w = 1u;                                   // a slot past the last sharer counts as "there"
                const uint32_t* src = base + (size_t)(s0 + u) * (BT * kBlockW) + (j_first + i * JT) * kBlockW;
                if (on) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(w) : "l"(src) : "memory");
                v[u][i] = w;
              }
            }
            uint32_t all = 1u;
#pragma unroll
            for (int u = 0; u < SB; ++u)
#pragma unroll
              for (int i = 0; i < NJ; ++i) all &= v[u][i];
#ifdef B2L_TLITE
            if (threadIdx.x == 0 && p.timeline && s0 == 1) {
              long long t;
              asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) : "r"(all) : "memory");
              p.timeline[(size_t)blockIdx.x * 32 + 30] = t;
            }
#endif
            if (all & 1u) break;
#ifdef B2L_TLITE
            if (threadIdx.x == 0 && p.timeline) p.timeline[(size_t)blockIdx.x * 32 + 29] += 1;
#endif
            // a sharer is still working (or, if this never ends, something is broken: trap with a record)
            const unsigned long long now = b2l_globaltimer();
            if (t0 == 0ull) t0 = now;
            const unsigned long long lim = b2l_hang_ctl.limit_ns;
            if (lim != 0ull && now - t0 > lim)
              b2l_hang_trap((uint32_t)ctl->fix_tile[f], (uint32_t)s0, b2l_file_hash(__FILE__), __LINE__, now - t0);
          }
#pragma unroll
          for (int u = 0; u < SB; ++u) {
            if (s0 + u < sharers) {
#pragma unroll
              for (int i = 0; i < NJ; ++i) {
                sum[i] += __uint_as_float(v[u][i]);
                base[(size_t)(s0 + u) * (BT * kBlockW) + (j_first + i * JT) * kBlockW] = 0u;   // leave the workspace clean
              }
            }
          }
        }
      } else {
        const uint32_t* base = p.ws + (size_t)ctl->fix_tile[f] * slots_per_tile * (BT * kBlockW) + n_local;
#pragma unroll
        for (int i = 0; i < NJ; ++i) sum[i] = 0.f;
        // every partial value of up to SB slots is requested before the first one is used: the loads are L2 round trips
        // (~1 us each when issued one slot at a time: 5 us for the o projection's 5 sharers, per-CTA stamps under
        // profiles/); the additions still run in slot order (+0.0f for the slots past the last sharer changes nothing)
        constexpr int SB = NJ <= 8 ? 8 : 4;
        for (int s0 = 0; s0 < sharers; s0 += SB) {
          float v[SB][NJ];
#pragma unroll
          for (int u = 0; u < SB; ++u) {
            const bool on = s0 + u < sharers;
#pragma unroll
            for (int i = 0; i < NJ; ++i)
              v[u][i] = on ? __uint_as_float(__ldcg(base + (size_t)(s0 + u) * (BT * kBlockW) + (j_first + i * JT) * kBlockW)) : 0.f;
          }
#pragma unroll
          for (int u = 0; u < SB; ++u)
#pragma unroll
            for (int i = 0; i < NJ; ++i) sum[i] += v[u][i];
        }
        // the workspace is shared with the stream-K launches, whose protocol needs it all-zero between launches
        for (int s0 = 0; s0 < sharers; ++s0) {
#pragma unroll
          for (int i = 0; i < NJ; ++i)
            const_cast<uint32_t*>(base)[(size_t)s0 * (BT * kBlockW) + (j_first + i * JT) * kBlockW] = 0u;
        }
      }
#ifdef B2L_TLITE
      if (threadIdx.x == 0 && p.timeline && f == 0) {
        long long t;
        asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) : "f"(sum[0]) : "memory");
        p.timeline[(size_t)blockIdx.x * 32 + 27] = t;
      }
#endif
      const int n = ctl->fix_ntile[f] * kBlockW + n_local;
      const bool n_ok = n < p.N;
      float bv = 0.f;
      if (n_ok && p.bias) bv = p.bias_is_f32 ? reinterpret_cast<const float*>(p.bias)[n]
                                             : __half2float(reinterpret_cast<const __half*>(p.bias)[n]);
      __half* yy = reinterpret_cast<__half*>(p.y);
#pragma unroll
      for (int i = 0; i < NJ; ++i) {
        const int m = blockIdx.y * BT + j_first + i * JT;
        if (EPI == 1) {                                      // same arithmetic as store_final (neighbouring lanes = a pair)
          const float mine = __half2float(__float2half_rn(sum[i] + bv));
          const float other = __shfl_xor_sync(0xffffffffu, mine, 1);
          if (n_ok && (n_local & 1) == 0 && m < p.M) {
            const float e = ex2_approx_q(-mine * 1.4426950408889634f);
            yy[(int64_t)m * p.ldy + (n >> 1)] = __float2half_rn(mine * div_full_q(1.0f, 1.0f + e) * other);
          }
        } else if (n_ok && m < p.M) {
          yy[(int64_t)m * p.ldy + n] = __float2half_rn(sum[i] + bv);
        }
      }
    }
  }
  if (tl && threadIdx.x == 0) tl[kTlMisc + 6] = clock64();
  if (timeline && threadIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && blockIdx.x < 1024) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    timeline[kTlCta + 3 * blockIdx.x + 1] = (long long)gt;
  }
#ifdef B2L_TLITE
  if (threadIdx.x == 0 && p.timeline) {
    LT(20);
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    p.timeline[(size_t)blockIdx.x * 32 + 17] = (long long)gt;
    p.timeline[(size_t)blockIdx.x * 32 + 21] = ctl->nfix;
    p.timeline[(size_t)blockIdx.x * 32 + 22] = nrb;
  }
#endif
  if (warp == kWarpMma0) tmem_dealloc(tmem_base, kTmemCols);
}


constexpr int kMaxGroupsCta = 40;   // scale-table columns per CTA (x 128 rows x 4 B, twice for int4): tile x split grid
constexpr int kMaxGroupsSk = 64;    //   ... under stream-K (a CTA's range is sized by bytes, not by groups)
static int g_streamk = -1;          // env B2L_STREAMK=0 keeps the tile x split grid (A/B measurements)
static inline bool streamk_enabled() {
  if (g_streamk < 0) {
    const char* e = getenv("B2L_STREAMK");
    g_streamk = (e && e[0] == '0') ? 0 : 1;
  }
  return g_streamk != 0;
}

static inline int groups_touched(int kb_per_split, int seg_len) {
  // worst case over split offsets: a k-range of kb_per_split*64 elements starting anywhere on a 64 boundary
  const int span = kb_per_split * 64;
  return ((span + seg_len - 1) / seg_len + 1 + 3) / 4 * 4;   // table pitch: whole 16-byte chunks
}

// ring depths + shared-memory footprint for a CTA that processes `blocks` raw blocks and keeps `max_groups` table columns
static inline void quant_rings(int qk, SplitPlan& pl, int blocks, int max_groups) {
  const int rw = qk == Q_INT4 ? 4 : 2;
  // activation blocks (L2-resident, held until their MMAs retire) get the smaller share, the raw weight ring
  // (HBM latency x bandwidth) everything that is left
  const int x_stage = rw * pl.bt * kStageKBytes;
  const int fixed = 1024 + 1024 + 1024 /*xsum exchange*/ + ((max_groups * kTabPitch + 3) & ~3) * 4 * (qk == Q_INT4 ? 2 : 1);
  const int budget = b2l_max_smem_optin() - fixed;
  int xs = 4, wsn;
  while (xs > 2 && budget - xs * x_stage < 3 * kABytes) --xs;
  if (xs > blocks + 1) xs = blocks + 1;
  if (xs < 2) xs = 2;
  wsn = (budget - xs * x_stage) / kABytes;
  if (wsn > kMaxRawStages) wsn = kMaxRawStages;
  if (wsn > blocks + 1) wsn = blocks + 1;
  if (wsn < 2) wsn = 2;
  // even ring depths: the two dequant groups / MMA warps alternate raw blocks, so with an even depth a stage
  // always belongs to the same group and every waiter follows every phase of "its" barriers (a parity wait
  // that skipped a phase could be satisfied by a completion two phases old)
  if (wsn > 2) wsn &= ~1;
  if (xs > 2) xs &= ~1;
  // spend what is left on activation stages
  while (xs + 2 <= kMaxRawStages && xs + 2 <= blocks + 2 && fixed + wsn * kABytes + (xs + 2) * x_stage <= b2l_max_smem_optin()) xs += 2;
  pl.stages = wsn;
  pl.x_stages = xs;
  pl.smem_bytes = fixed + wsn * kABytes + xs * x_stage;
  pl.max_groups = max_groups;
}

static inline SplitPlan quant_plan(int qk, int64_t M, int64_t N, int64_t K, int seg_len) {
  // token tile capped at 64: the epilogue keeps the running fp32 accumulator in registers
  SplitPlan pl;
#ifndef B2L_QUANT_BT_CAP
#define B2L_QUANT_BT_CAP 64
#endif
  const int64_t mt = M > B2L_QUANT_BT_CAP ? B2L_QUANT_BT_CAP : M;
  pl.bt = pick_bt(mt);
  pl.n_tiles = (int)cdiv64(N, kBlockW);
  pl.m_tiles = (int)cdiv64(M, pl.bt);
  pl.num_kb = (int)cdiv64(K, 64);
  const int tiles = pl.n_tiles * pl.m_tiles;
  const int sms = b2l_num_sms();
  // ---- stream-K: one CTA per SM, equal contiguous shares of the (tile, raw block) space ----
  {
    const int raw_k = qk == Q_INT4 ? 256 : 128;
    const bool aligned = raw_k % seg_len == 0 && raw_k / seg_len <= 2;     // same rule as QuantParams::aligned
    if (streamk_enabled() && !(g_debug & DBG_NO_STREAMK) && aligned && K % raw_k == 0 && pl.m_tiles == 1 &&
        (int64_t)pl.n_tiles * (K / raw_k) < (1 << 22)) {        // the kernel's range arithmetic is 32-bit: total * (ctas + 1)
      const int nrb_tile = (int)(K / raw_k);
      const int64_t total = (int64_t)pl.n_tiles * nrb_tile;
      int ctas = (int)(total / 2 < sms ? total / 2 : sms);                 // at least two raw blocks per CTA
      if (ctas < 1) ctas = 1;
      const int per_cta = (int)((total + ctas - 1) / ctas);
      const int groups = (per_cta * (raw_k / seg_len) + 3) / 4 * 4;
      if (groups <= kMaxGroupsSk) {
        pl.sk_ctas = ctas;
        pl.nrb_tile = nrb_tile;
        // most CTAs that can share one tile: a tile's nrb_tile blocks cut by ranges of >= floor(total/ctas) blocks
        int slots = 1;
        for (int t = 0; t < pl.n_tiles; ++t) {
          const int64_t lo = (((int64_t)t * nrb_tile + 1) * ctas - 1) / total;
          const int64_t hi = (((int64_t)(t + 1) * nrb_tile) * ctas - 1) / total;
          if ((int)(hi - lo + 1) > slots) slots = (int)(hi - lo + 1);
        }
        pl.sk_slots = slots;
        pl.splits = slots > 1 ? slots : 1;       // "needs workspace" for the callers that only look at splits
        pl.kb_per_split = pl.num_kb;
        quant_rings(qk, pl, per_cta, groups);
        return pl;
      }
    }
  }
  int splits = 1;
  if (tiles < sms) {
    splits = sms / tiles;
    int max_splits = pl.num_kb / 8;
    if (max_splits < 1) max_splits = 1;
    if (splits > max_splits) splits = max_splits;
    if (splits > 32) splits = 32;
  }
  const int rw = qk == Q_INT4 ? 4 : 2;     // k-blocks per raw stage: splits must start on a raw-block boundary
  pl.kb_per_split = ((pl.num_kb + splits - 1) / splits + rw - 1) / rw * rw;
  // keep the per-CTA scale table bounded: more splits when groups are fine-grained -- but keep them BALANCED
  // (equal k-ranges), an uneven last split would make one CTA of every tile the straggler
  while (pl.kb_per_split > rw && groups_touched(pl.kb_per_split, seg_len) > kMaxGroupsCta) {
    ++splits;
    pl.kb_per_split = ((pl.num_kb + splits - 1) / splits + rw - 1) / rw * rw;
  }
  pl.splits = (pl.num_kb + pl.kb_per_split - 1) / pl.kb_per_split;
  quant_rings(qk, pl, pl.kb_per_split / rw, groups_touched(pl.kb_per_split, seg_len));
  return pl;
}

static inline int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e && e[0] ? atoi(e) : dflt;
}

template <int QK, int BT, bool DBG, int LAYOUT, int EPI = 0>
static int launch_q_inst(const CUtensorMap& tw, const CUtensorMap& tx, const QuantParams& p, const SplitPlan& pl,
                         cudaStream_t st) {
  auto kern = gemm_quant_kernel<QK, BT, DBG, LAYOUT, EPI>;
  static int configured_smem[64] = {0};
  B2L_CUDA_CHECK(b2l_ensure_dynamic_smem(kern, pl.smem_bytes, configured_smem));
  dim3 grid((unsigned)pl.n_tiles, (unsigned)pl.m_tiles, (unsigned)pl.splits);
  if (pl.sk_ctas > 0) grid = dim3((unsigned)pl.sk_ctas, (unsigned)pl.m_tiles, 1);
  b2l_launch(kern, dim3(grid), dim3(kQuantThreads), (size_t)(pl.smem_bytes), st, tw, tx, p);
  B2L_LAUNCH_CHECK();
  return B2L_OK;
}

template <int QK, int BT>
static int launch_q_bt(const CUtensorMap& tw, const CUtensorMap& tx, const QuantParams& p, const SplitPlan& pl,
                       cudaStream_t st, int layout) {
#ifdef B2L_TLITE
  const bool dbg = p.debug != 0;
#else
  const bool dbg = p.debug != 0 || p.timeline != nullptr;
#endif
  if constexpr (QK == Q_INT4) {
    if (layout == 1) {
      if (p.swiglu) {
        B2L_REQUIRE(!dbg, "w4a16_matmul_packed: the swiglu epilogue has no debug instantiation");
        return launch_q_inst<QK, BT, false, 1, 1>(tw, tx, p, pl, st);
      }
      return dbg ? launch_q_inst<QK, BT, true, 1>(tw, tx, p, pl, st) : launch_q_inst<QK, BT, false, 1>(tw, tx, p, pl, st);
    }
  }
  return dbg ? launch_q_inst<QK, BT, true, 0>(tw, tx, p, pl, st) : launch_q_inst<QK, BT, false, 0>(tw, tx, p, pl, st);
}

template <int QK>
static int launch_q(const CUtensorMap& tw, const CUtensorMap& tx, const QuantParams& p, const SplitPlan& pl,
                    cudaStream_t st, int layout = 0) {
  switch (pl.bt) {
    case 16: return launch_q_bt<QK, 16>(tw, tx, p, pl, st, layout);
    case 32: return launch_q_bt<QK, 32>(tw, tx, p, pl, st, layout);
    default: return launch_q_bt<QK, 64>(tw, tx, p, pl, st, layout);
  }
}

static int quant_common(int qk, QuantParams& p, SplitPlan& pl, void* y, const void* x, const void* qweight,
                        int64_t m, int64_t n, int64_t k, int64_t x_row_stride, int64_t w_row_stride_elems,
                        int64_t y_row_stride, void* workspace, int64_t workspace_bytes, void* counters,
                        int64_t num_counters, CUtensorMap* tw, CUtensorMap* tx, const char* what, int seg_len) {
  pl = quant_plan(qk, m, n, k, seg_len);
  p.max_groups = pl.max_groups;
  p.sk_ctas = pl.sk_ctas; p.sk_slots = pl.sk_slots; p.nrb_tile = pl.nrb_tile; p.n_tiles = pl.n_tiles;
  p.y = y; p.ldy = y_row_stride; p.M = (int)m; p.N = (int)n; p.K = (int)k;
  p.splits = pl.splits; p.kb_per_split = pl.kb_per_split; p.num_kb = pl.num_kb; p.w_stages = pl.stages; p.x_stages = pl.x_stages;
  if ((g_debug & DBG_HALF_RING) && p.w_stages >= 4) p.w_stages = (p.w_stages / 2) & ~1;
  p.timeline = g_timeline;
  p.debug = g_debug;
  p.swiglu = 0;
  p.zeros_foldable = 0;
  p.x = x; p.ldx = x_row_stride;
  p.force_xsum = 0; p.seg_shift = 0;
  if (qk == Q_INT4) {
    const int q64 = seg_len / 64;
    if (seg_len % 64 != 0 || (q64 & (q64 - 1)) != 0) p.force_xsum = 1;
    else while ((1 << p.seg_shift) < q64) ++p.seg_shift;
  }
  // Two MMA warps on alternate raw blocks need (a) whole segments inside a raw block and (b) every accumulator
  // buffer (segment index mod 4) always owned by the same warp, i.e. 1 or 2 segments per raw block: a warp that
  // waits on an acc_empty phase it did not itself follow could be two phases ahead and alias the parity.
  {
    const int raw_k = qk == Q_INT4 ? 256 : 128;
    p.aligned = (raw_k % seg_len == 0 && raw_k / seg_len <= 2) ? 1 : 0;
  }
  p.ws = (uint32_t*)workspace; p.counters = (int*)counters;
  if (pl.splits > 1) {
    const int64_t need = (int64_t)pl.n_tiles * pl.m_tiles * pl.splits * pl.bt * kBlockW * 4;
    B2L_REQUIRE(workspace && workspace_bytes >= need, "%s: workspace too small (%lld < %lld)", what,
                (long long)workspace_bytes, (long long)need);
    B2L_REQUIRE(counters && num_counters >= (int64_t)pl.n_tiles * pl.m_tiles, "%s: split-K counters too small", what);
  }
  int rc;
  if (qk == Q_INT4)
    rc = b2l_make_tmap_2d(tw, qweight, TMA_I32, (uint64_t)n, (uint64_t)(k / 8), (uint64_t)w_row_stride_elems * 4,
                          kBlockW, 32, true);
  else
    rc = b2l_make_tmap_2d(tw, qweight, TMA_U8, (uint64_t)n, (uint64_t)k, (uint64_t)w_row_stride_elems, kBlockW, 128,
                          true);
  if (rc != B2L_OK) return rc;
  return b2l_make_tmap_2d(tx, x, TMA_F16, (uint64_t)m, (uint64_t)k, (uint64_t)x_row_stride * 2, (uint32_t)pl.bt, 64,
                          true);
}

}  // namespace gemm

using namespace gemm;

static inline int seg_len_for(int kind, int64_t k, int64_t group_k) {
  if (kind == Q_INT4) return (int)group_k;
  int64_t gk = group_k < k ? group_k : k;
  if (gk % 16 != 0) gk = k;
  return (int)((gk + 15) / 16 * 16);
}

// Debug hook: device buffer of 8*256 int64 receiving clock64 stamps from CTA (0,0,0) of the next quantised
// GEMM launches (null disables).  Not part of the product path.
extern "C" int b2l_debug_set_timeline(void* device_buffer) {
  gemm::g_timeline = (long long*)device_buffer;
  return B2L_OK;
}
// DBG_* experiments: any non-zero flag runs the DBG instantiation of the kernel
extern "C" int b2l_debug_set_flags(int flags) {
  gemm::g_debug = flags & 0xFFFF;
  return B2L_OK;
}

extern "C" int64_t b2l_quant_gemm_workspace_bytes(int kind, int64_t m, int64_t n, int64_t k, int64_t group_k) {
  if (m <= 0 || n <= 0 || k <= 0 || group_k <= 0) return 0;
  SplitPlan pl = quant_plan(kind, m, n, k, seg_len_for(kind, k, group_k));
  if (pl.splits <= 1) return 0;
  return (int64_t)pl.n_tiles * pl.m_tiles * pl.splits * pl.bt * kBlockW * 4;
}
// Launch plan of a quantised GEMM (pure function of shapes and the SM count; no GPU work):
// out[0..9] = {stream-K CTAs (0 = tile x split grid), workspace slots per tile, raw blocks per tile, weight-row tiles,
//              splits, 64-element k-blocks per split, weight ring depth, activation ring depth, smem bytes, table columns}
extern "C" int b2l_quant_gemm_plan(int kind, int64_t m, int64_t n, int64_t k, int64_t group_k, int* out, void* /*reserved*/) {
  B2L_REQUIRE(out && m > 0 && n > 0 && k > 0 && group_k > 0, "quant_gemm_plan: bad arguments");
  const SplitPlan pl = quant_plan(kind, m, n, k, seg_len_for(kind, k, group_k));
  out[0] = pl.sk_ctas; out[1] = pl.sk_slots; out[2] = pl.nrb_tile; out[3] = pl.n_tiles; out[4] = pl.splits;
  out[5] = pl.kb_per_split; out[6] = pl.stages; out[7] = pl.x_stages; out[8] = pl.smem_bytes; out[9] = pl.max_groups;
  return B2L_OK;
}

extern "C" int64_t b2l_quant_gemm_num_counters(int64_t m, int64_t n) {
  if (m <= 0 || n <= 0) return 0;
  return cdiv64(n, kBlockW) * cdiv64(m, pick_bt(m > 64 ? 64 : m));
}

extern "C" int b2l_w4a16_matmul(void* y, const void* x, const void* qweight, const void* scales, const void* zeros,
                                const void* bias, int bias_is_f32, int64_t m, int64_t n, int64_t k, int group_size,
                                int64_t x_row_stride, int64_t qw_row_stride, int64_t s_row_stride,
                                int64_t s_col_stride, int64_t y_row_stride, void* workspace, int64_t workspace_bytes,
                                void* counters, int64_t num_counters, void* stream) {
  B2L_REQUIRE(y && x && qweight && scales && zeros, "w4a16_matmul: null pointer");
  B2L_REQUIRE(m >= 0 && n > 0 && k > 0 && k % 8 == 0, "w4a16_matmul: bad shape m=%lld n=%lld k=%lld", (long long)m,
              (long long)n, (long long)k);
  B2L_REQUIRE(group_size > 0 && k % group_size == 0, "w4a16_matmul: K (%lld) must be a multiple of group_size (%d)",
              (long long)k, group_size);
  B2L_REQUIRE(group_size % 16 == 0, "w4a16_matmul: group_size (%d) must be a multiple of 16", group_size);
  B2L_REQUIRE(x_row_stride % 8 == 0 && qw_row_stride % 4 == 0,
              "w4a16_matmul: row strides must be 16-byte multiples (K a multiple of 32)");
  if (m == 0) return B2L_OK;
  QuantParams p; SplitPlan pl; CUtensorMap tw, tx;
  int rc = quant_common(Q_INT4, p, pl, y, x, qweight, m, n, k, x_row_stride, qw_row_stride, y_row_stride, workspace,
                        workspace_bytes, counters, num_counters, &tw, &tx, "w4a16_matmul", group_size);
  if (rc != B2L_OK) return rc;
  p.bias = bias; p.bias_is_f32 = bias_is_f32;
  p.scales = (const float*)scales; p.zeros = (const float*)zeros; p.s_stride_n = s_row_stride; p.s_stride_k = s_col_stride;
  p.group_n = 1; p.group_k = group_size;
  return launch_q<Q_INT4>(tw, tx, p, pl, (cudaStream_t)stream);
}

// Same product on the load-time repack of the three tensors (b2l_w4a16_pack_qweight / b2l_w4a16_pack_table):
// qweight_packed int32 [N, K/8] with the nibbles of each word in K order [0,2,4,6,1,3,5,7]; scales_packed /
// zeros_packed fp32 [ceil(N/128)][K/group][128] (rows past N zero).  zeros_foldable = 1 when the caller has checked
// that every zero point is an integer in [0, 1023] (else the row-sum correction path runs).
extern "C" int b2l_w4a16_matmul_packed(void* y, const void* x, const void* qweight_packed, const void* scales_packed,
                                       const void* zeros_packed, const void* bias, int bias_is_f32, int64_t m,
                                       int64_t n, int64_t k, int group_size, int zeros_foldable, int64_t x_row_stride,
                                       int64_t qw_row_stride, int64_t y_row_stride, void* workspace,
                                       int64_t workspace_bytes, void* counters, int64_t num_counters, int epilogue,
                                       void* stream) {
  B2L_REQUIRE(y && x && qweight_packed && scales_packed && zeros_packed, "w4a16_matmul_packed: null pointer");
  B2L_REQUIRE(m >= 0 && n > 0 && k > 0 && k % 8 == 0, "w4a16_matmul_packed: bad shape m=%lld n=%lld k=%lld",
              (long long)m, (long long)n, (long long)k);
  B2L_REQUIRE(group_size > 0 && k % group_size == 0,
              "w4a16_matmul_packed: K (%lld) must be a multiple of group_size (%d)", (long long)k, group_size);
  B2L_REQUIRE(group_size % 16 == 0, "w4a16_matmul_packed: group_size (%d) must be a multiple of 16", group_size);
  B2L_REQUIRE(x_row_stride % 8 == 0 && qw_row_stride % 4 == 0,
              "w4a16_matmul_packed: row strides must be 16-byte multiples (K a multiple of 32)");
  B2L_REQUIRE((uintptr_t)scales_packed % 16 == 0 && (uintptr_t)zeros_packed % 16 == 0,
              "w4a16_matmul_packed: tables must be 16-byte aligned");
  if (m == 0) return B2L_OK;
  QuantParams p; SplitPlan pl; CUtensorMap tw, tx;
  int rc = quant_common(Q_INT4, p, pl, y, x, qweight_packed, m, n, k, x_row_stride, qw_row_stride, y_row_stride,
                        workspace, workspace_bytes, counters, num_counters, &tw, &tx, "w4a16_matmul_packed", group_size);
  if (rc != B2L_OK) return rc;
  p.bias = bias; p.bias_is_f32 = bias_is_f32;
  p.scales = (const float*)scales_packed; p.zeros = (const float*)zeros_packed; p.s_stride_n = 0; p.s_stride_k = 0;
  p.group_n = 1; p.group_k = group_size;
  p.zeros_foldable = zeros_foldable ? 1 : 0;
  B2L_REQUIRE(epilogue == 0 || epilogue == 1, "w4a16_matmul_packed: epilogue %d (0 = none, 1 = swiglu over interleaved rows)", epilogue);
  B2L_REQUIRE(epilogue == 0 || n % 2 == 0, "w4a16_matmul_packed: the swiglu epilogue needs an even number of rows");
  p.swiglu = epilogue;
  return launch_q<Q_INT4>(tw, tx, p, pl, (cudaStream_t)stream, 1);
}

// ---- load-time repack kernels of the packed W4A16 layout (one pass each, bit-exact permutations) ----
namespace {
__global__ void w4_pack_qweight_kernel(uint32_t* __restrict__ out, const uint32_t* __restrict__ in, int64_t n, int64_t k8,
                                       int64_t in_stride, int64_t out_stride) {
  pdl_prologue();
  const int64_t total = n * k8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / k8, c = i % k8;
    const uint32_t w = in[r * in_stride + c];
    // position p of the packed word holds K index order[p], order = [0,2,4,6,1,3,5,7]
    uint32_t o = 0;
#pragma unroll
    for (int pidx = 0; pidx < 8; ++pidx) {
      const int kk = pidx < 4 ? 2 * pidx : 2 * (pidx - 4) + 1;
      o |= ((w >> (4 * kk)) & 0xFu) << (4 * pidx);
    }
    out[r * out_stride + c] = o;
  }
}
__global__ void w4_pack_table_kernel(float* __restrict__ out, const float* __restrict__ in, int64_t n, int64_t g,
                                     int64_t stride_n, int64_t stride_k, int64_t tiles) {
  pdl_prologue();
  const int64_t total = tiles * g * 128;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i % 128, gi = (i / 128) % g, t = i / (128 * g);
    const int64_t row = t * 128 + r;
    out[i] = row < n ? in[row * stride_n + gi * stride_k] : 0.f;
  }
}
}  // namespace

extern "C" int b2l_w4a16_pack_qweight(void* out, const void* in, int64_t n, int64_t k8, int64_t in_row_stride,
                                      int64_t out_row_stride, void* stream) {
  B2L_REQUIRE(out && in && n > 0 && k8 > 0, "w4a16_pack_qweight: bad arguments");
  const int64_t total = n * k8;
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  b2l_launch(w4_pack_qweight_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, (uint32_t*)out,
             (const uint32_t*)in, n, k8, in_row_stride, out_row_stride);
  B2L_LAUNCH_CHECK();
  return B2L_OK;
}
extern "C" int b2l_w4a16_pack_table(void* out, const void* in, int64_t n, int64_t groups, int64_t row_stride,
                                    int64_t col_stride, void* stream) {
  B2L_REQUIRE(out && in && n > 0 && groups > 0, "w4a16_pack_table: bad arguments");
  const int64_t tiles = (n + 127) / 128;
  const int64_t total = tiles * groups * 128;
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  b2l_launch(w4_pack_table_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, (float*)out, (const float*)in, n,
             groups, row_stride, col_stride, tiles);
  B2L_LAUNCH_CHECK();
  return B2L_OK;
}

extern "C" int b2l_w8a16_matmul(void* y, const void* x, const void* qweight, int weight_is_fp8, const void* scales,
                                const void* bias, int bias_is_f32, int64_t m, int64_t n, int64_t k, int64_t group_n,
                                int64_t group_k, int64_t x_row_stride, int64_t qw_row_stride, int64_t s_row_stride,
                                int64_t s_col_stride, int64_t y_row_stride, void* workspace, int64_t workspace_bytes,
                                void* counters, int64_t num_counters, void* stream) {
  B2L_REQUIRE(y && x && qweight && scales, "w8a16_matmul: null pointer");
  B2L_REQUIRE(m >= 0 && n > 0 && k > 0 && k % 16 == 0, "w8a16_matmul: K must be a multiple of 16, got %lld", (long long)k);
  B2L_REQUIRE(group_n >= 1 && group_k >= 1, "w8a16_matmul: bad scale block %lldx%lld", (long long)group_n, (long long)group_k);
  B2L_REQUIRE(group_k >= k || group_k % 16 == 0, "w8a16_matmul: group_k (%lld) must be a multiple of 16 unless it covers K",
              (long long)group_k);
  B2L_REQUIRE(x_row_stride % 8 == 0 && qw_row_stride % 16 == 0, "w8a16_matmul: row strides must be 16-byte multiples");
  if (m == 0) return B2L_OK;
  QuantParams p; SplitPlan pl; CUtensorMap tw, tx;
  const int qk = weight_is_fp8 ? Q_FP8 : Q_INT8;
  const int seg_len = seg_len_for(qk, k, group_k);
  int rc = quant_common(qk, p, pl, y, x, qweight, m, n, k, x_row_stride, qw_row_stride, y_row_stride, workspace,
                        workspace_bytes, counters, num_counters, &tw, &tx, "w8a16_matmul", seg_len);
  if (rc != B2L_OK) return rc;
  p.bias = bias; p.bias_is_f32 = bias_is_f32;
  p.scales = (const float*)scales; p.zeros = nullptr; p.s_stride_n = s_row_stride; p.s_stride_k = s_col_stride;
  p.group_n = (int)(group_n > n ? n : group_n);
  p.group_k = seg_len;
  if (weight_is_fp8) return launch_q<Q_FP8>(tw, tx, p, pl, (cudaStream_t)stream);
  return launch_q<Q_INT8>(tw, tx, p, pl, (cudaStream_t)stream);
}
