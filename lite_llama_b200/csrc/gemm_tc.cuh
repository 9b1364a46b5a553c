// Shared skeleton of the tcgen05 "weight-stationary, swap-AB" GEMM family.
//
//   Y[m, n] = sum_k X[m, k] * W[n, k]          X: [M, K] activations, W: [N, K] weights (K contiguous)
//
// Decode and continuous-batching steps have M = 1..256 tokens while N, K are in the thousands,
// so the weight matrix is what streams from HBM. The tile is therefore laid out "swapped":
// the 128 lanes of the tensor-core accumulator (UMMA M) index weight rows n, the UMMA N dimension
// (16..256) indexes tokens m. One tcgen05.mma computes D^T[128 n, BT m] += W_tile[128, 16] * X_tile[BT, 16]^T.
// Both operands are K-major in shared memory in the canonical 128-byte-swizzled layout that
// TMA produces, so no data shuffling is needed on the way in.
//
// Warp roles (warp-specialised, mbarrier pipelines):
//   warp 0      TMA producer: one elected lane streams W and X tiles into a STAGES-deep ring
//   warp 1      MMA issuer: one elected lane issues tcgen05.mma; owns the TMEM allocation
//   warps 2-5   epilogue: tcgen05.ld accumulator -> registers -> (split-K fix-up) -> global
//   warps 6+    (quantised variants only) dequantise raw weight tiles into the UMMA operand ring
//
// Split-K: CTAs of one output tile write fp32 partials to a workspace; the last CTA to arrive
// (atomic ticket, self-resetting) reduces them in split order => deterministic, graph-safe.
#pragma once
#include "common.cuh"
#include "tma_desc.cuh"

namespace gemm {

constexpr int kBlockW = 128;        // weight rows per CTA (UMMA M)
constexpr int kStageKBytes = 128;   // bytes along K per pipeline stage (= one swizzle atom row)
constexpr int kMaxStages = 12;
constexpr int kABytes = kBlockW * kStageKBytes;  // 16 KiB per stage for the weight operand

struct SplitPlan {
  int bt;             // token tile (UMMA N)
  int n_tiles, m_tiles;
  int num_kb;         // k-blocks in total
  int splits;
  int kb_per_split;
  int stages;
  int x_stages;       // quantised GEMMs: depth of the separate activation ring
  int smem_bytes;
  int sk_ctas = 0;    // quantised GEMMs: stream-K CTA count (0 = tile x split grid)
  int sk_slots = 0;   //   workspace slots per tile under stream-K
  int nrb_tile = 0;   //   raw blocks per tile
  int max_groups = 0; //   scale-table columns per CTA
};

static inline int pick_bt(int64_t m) {
  if (m <= 16) return 16;
  if (m <= 32) return 32;
  if (m <= 64) return 64;
  if (m <= 128) return 128;
  return 256;
}

// Pure function of shapes + SM count (graph-safe).  extra_stage_bytes: additional per-stage smem
// (raw quantised tile), fixed_bytes: smem outside the ring.
static inline SplitPlan make_plan(int64_t M, int64_t N, int64_t K, int k_elems_per_stage, int extra_stage_bytes,
                                  int fixed_bytes, int min_kb_per_split) {
  SplitPlan pl;
  pl.bt = pick_bt(M);
  pl.n_tiles = (int)cdiv64(N, kBlockW);
  const int sms = b2l_num_sms();
  // Medium token counts with few weight-row tiles (e.g. 256 x 4608 x 3584: 36 tiles of 128 x 256): fill the SMs
  // with SMALLER token tiles before resorting to split-K -- a split of a 256-token tile moves 128 KiB of fp32
  // partials per CTA through the workspace, which cost more than the GEMM (measured 57 us vs 34 us for the
  // reference's Triton kernel on this shape, profiles/r2c_triton_vs_b200.md).
  while (pl.bt > 64 && (int64_t)pl.n_tiles * cdiv64(M, pl.bt) < sms) pl.bt /= 2;
  pl.m_tiles = (int)cdiv64(M, pl.bt);
  pl.num_kb = (int)cdiv64(K, k_elems_per_stage);
  const int tiles = pl.n_tiles * pl.m_tiles;
  int splits = 1;
  if (tiles < sms) {
    splits = sms / tiles;
    int max_splits = pl.num_kb / min_kb_per_split;
    if (max_splits < 1) max_splits = 1;
    if (splits > max_splits) splits = max_splits;
    if (splits > 32) splits = 32;
    if (splits < 1) splits = 1;
  }
  pl.kb_per_split = (pl.num_kb + splits - 1) / splits;
  pl.splits = (pl.num_kb + pl.kb_per_split - 1) / pl.kb_per_split;
  const int stage_bytes = kABytes + pl.bt * kStageKBytes + extra_stage_bytes;
  int budget = b2l_max_smem_optin() - fixed_bytes - 1024 /*alignment slack*/;
  int stages = budget / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages > pl.kb_per_split + 1) stages = pl.kb_per_split + 1;
  if (stages < 2) stages = 2;
  pl.stages = stages;
  pl.smem_bytes = stages * stage_bytes + fixed_bytes + 1024;
  return pl;
}

#ifdef __CUDACC__

struct PipeState {
  int stage; uint32_t phase; int nstages;
  __device__ __forceinline__ PipeState(int n) : stage(0), phase(0), nstages(n) {}
  __device__ __forceinline__ void advance() {
    if (++stage == nstages) { stage = 0; phase ^= 1; }
  }
};

// Control block at the start of dynamic shared memory (after 1024-byte alignment).
struct __align__(8) Control {
  uint64_t full[kMaxStages];      // operands of a stage are ready for the MMA warp
  uint64_t empty[kMaxStages];     // MMA has consumed a stage
  uint64_t raw_full[kMaxStages];  // (quantised) raw tile landed
  uint64_t raw_empty[kMaxStages]; // (quantised) raw tile consumed by the dequant warps
  uint64_t acc_full[2];           // accumulator buffer ready for the epilogue
  uint64_t acc_empty[2];          // accumulator buffer drained
  uint32_t tmem_base;
  int is_last;
};
constexpr int kControlBytes = 512;
static_assert(sizeof(Control) <= kControlBytes, "control block too large");

// Deterministic split-K: every CTA of an output tile stores its fp32 (or int32) partial in the
// workspace, then takes a ticket; the CTA drawing the last ticket reduces all partials in split
// order.  Called by the 128 epilogue threads after their partial stores.
__device__ __forceinline__ bool splitk_ticket(int* __restrict__ counters, int tile_id, int splits, int epi_tid,
                                             int* is_last_smem) {
  __threadfence();
  asm volatile("bar.sync 2, 128;" ::: "memory");
  if (epi_tid == 0) {
    const int old = atomicAdd(&counters[tile_id], 1);
    const int last = (old == splits - 1);
    if (last) counters[tile_id] = 0;  // self-reset: the next launch (or graph replay) starts clean
    *is_last_smem = last;
  }
  asm volatile("bar.sync 2, 128;" ::: "memory");
  const bool last = *is_last_smem != 0;
  if (last) __threadfence();
  return last;
}

#endif  // __CUDACC__

}  // namespace gemm
