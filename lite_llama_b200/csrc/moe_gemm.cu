// Grouped (per-expert) GEMM of the fused-MoE block on tcgen05.
//
// Replaces lite_llama/kernels/fused_moe.py:105-293 (`_fused_moe_kernel` + launcher): for every row block
// of the sorted-slot protocol (sorted_token_ids / expert_ids / num_tokens_post_padded, fused_moe.py:45-99)
//   C[slot, :] = A[slot // top_k, :] @ W[expert(block)].T        (optionally * topk_weight[slot])
// Expert weights are fp16 [E, N, K] or 8-bit (uint8 e4m3 / int8) with one fp32 scale per
// (group_n x group_k) block, the same layout w8a16 uses (fused_moe.py:188-197).
//
// Same swap-AB tile as the dense/quantised GEMMs (gemm_tc.cuh): 128 weight rows of one expert on the
// TMEM lanes, the block's BT token slots on the UMMA N dimension.  Differences:
//   * the activation rows of a block are scattered (gathered through sorted_token_ids), so the eight
//     "transform" warps copy them into the 128B-swizzled operand ring (and, for 8-bit experts,
//     dequantise the raw weight tile next to them);
//   * blocks at or past num_tokens_post_padded exit immediately (the grid is sized for the static
//     worst case, so no host sync is needed -- fused_moe.py:156-158);
//   * the epilogue scatters rows back by slot id and can fold the routing weight in fp32 before the
//     cast (fused_moe.py:201-205).
#include "gemm_tc.cuh"

namespace gemm {

constexpr int kMoeXformWarps = 8;
constexpr int kMoeThreads = (2 + 4 + kMoeXformWarps) * 32;
constexpr int kMoeOpStages = 4;
constexpr int kMoeMaxRaw = 12;

enum MoeKind { M_F16 = 0, M_FP8 = 1, M_INT8 = 2 };

struct MoeParams {
  const __half* a; int64_t lda;            // activations [rows, K]
  __half* c; int64_t ldc;                  // output [num_slots, N]
  const int32_t* sorted_ids;               // [EM]
  const int32_t* expert_ids;               // [EM / BT]
  const int32_t* num_post;                 // [1]
  const __half* topk_weights;              // [num_slots] (activation dtype, fused_moe.py:397) or null
  const float* scales; int64_t s_stride_e, s_stride_n, s_stride_k;
  int group_n, group_k;
  int N, K, num_kb;
  int num_valid_slots, top_k;
  int raw_stages, max_groups;
  int op_stages;                           // depth of the operand ring (<= kMoeOpStages)
};

struct __align__(8) MoeControl {
  uint64_t raw_full[kMoeMaxRaw];
  uint64_t raw_empty[kMoeMaxRaw];
  uint64_t op_full[kMoeOpStages];
  uint64_t op_empty[kMoeOpStages];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint32_t tmem_base;
};
static_assert(sizeof(MoeControl) <= 1024, "control block too large");

__device__ __forceinline__ uint32_t moe_cvt_e4m3x2(uint16_t v) {
  uint32_t r;
  asm("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(r) : "h"(v));
  return r;
}
__device__ __forceinline__ uint4 moe_dequant8(uint32_t lo, uint32_t hi, bool fp8) {
  if (fp8)
    return make_uint4(moe_cvt_e4m3x2((uint16_t)(lo & 0xFFFFu)), moe_cvt_e4m3x2((uint16_t)(lo >> 16)),
                      moe_cvt_e4m3x2((uint16_t)(hi & 0xFFFFu)), moe_cvt_e4m3x2((uint16_t)(hi >> 16)));
  const __half2 off = __float2half2_rn(-1152.0f);
  uint32_t out[4];
  const uint32_t src[2] = {lo ^ 0x80808080u, hi ^ 0x80808080u};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t h = __byte_perm(src[j >> 1], 0x64646464u, (j & 1) ? 0x4342u : 0x4140u);
    const __half2 v = __hadd2(*reinterpret_cast<const __half2*>(&h), off);
    out[j] = *reinterpret_cast<const uint32_t*>(&v);
  }
  return make_uint4(out[0], out[1], out[2], out[3]);
}

template <int MK, int BT>
__global__ void __launch_bounds__(kMoeThreads, BT == 16 ? 2 : 1)
moe_gemm_kernel(const __grid_constant__ CUtensorMap tmap_w, const MoeParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  MoeControl* ctl = reinterpret_cast<MoeControl*>(smem);
  constexpr int kXBytes = BT * kStageKBytes;
  // raw stage = 128 rows x 128 bytes (full-line TMA rows): fp16 -> one ready-to-use 64-element operand tile,
  // 8-bit -> 128 elements = RW (2) operand k-blocks
  constexpr int kRawW = kABytes;
  constexpr int RW = MK == M_F16 ? 1 : 2;
  // op ring stage: [dequantised W tile (8-bit only)] [gathered X tile]
  constexpr int kOpStage = (MK == M_F16 ? 0 : kABytes) + kXBytes;
  uint8_t* op_ring = smem + 1024;
  uint8_t* raw_ring = op_ring + p.op_stages * kOpStage;
  float* scale_tab = reinterpret_cast<float*>(raw_ring + (size_t)p.raw_stages * kRawW);
  constexpr int kTmemCols = 2 * BT < 32 ? 32 : 2 * BT;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tile = blockIdx.x, blk = blockIdx.y;
  pdl_wait_prior();                                 // num_post / expert_ids / sorted ids are the align kernel's output
  const int num_post = *p.num_post;
  if (blk * BT >= num_post) return;                 // whole CTA: nothing routed to this block
  const int expert = p.expert_ids[blk];
  const int nkb = p.num_kb;
  const int seg_len = MK == M_F16 ? (nkb * 64) : p.group_k;
  const int steps_total = (p.K + 15) / 16;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_w);
    for (int s = 0; s < p.raw_stages; ++s) {
      mbar_init(&ctl->raw_full[s], 1);
      mbar_init(&ctl->raw_empty[s], MK == M_F16 ? 1 : kMoeXformWarps);
    }
    for (int s = 0; s < p.op_stages; ++s) {
      mbar_init(&ctl->op_full[s], kMoeXformWarps);
      mbar_init(&ctl->op_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&ctl->acc_full[s], 1);
      mbar_init(&ctl->acc_empty[s], 4);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(&ctl->tmem_base, kTmemCols);
    tmem_relinquish();
    // dependents are released by the TMEM-owning warp once the allocation is done (a CTA that returned above counts
    // as triggered): see gemm_dense.cu for the co-residency dead-lock this ordering rules out
    pdl_launch_dependents();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;

  if (warp == 0) {
    // ===================== TMA producer: this expert's weight tile =====================
    PipeState rs(p.raw_stages);
    const int row0 = expert * p.N + n_tile * kBlockW;
    const int nrb = (nkb + RW - 1) / RW;
    for (int rb = 0; rb < nrb; ++rb) {
      mbar_wait(&ctl->raw_empty[rs.stage], rs.phase ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&ctl->raw_full[rs.stage], kRawW);
        tma_load_2d_hint(raw_ring + (size_t)rs.stage * kRawW, &tmap_w, rb * RW * 64, row0, &ctl->raw_full[rs.stage],
                         kEvictFirst);
      }
      __syncwarp();
      rs.advance();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp loops, one elected lane issues) =====================
    constexpr uint32_t idesc = umma_idesc(1u, 0u, 0u, 0u, 0u, kBlockW, BT);
    const uint64_t op_desc0 = umma_desc_sw128_kmajor(smem_u32(op_ring));
    const uint64_t raw_desc0 = umma_desc_sw128_kmajor(smem_u32(raw_ring));
    PipeState rs(p.raw_stages), os(p.op_stages);
    const int seg_steps = seg_len / 16;
    int steps_left = steps_total;
    int seg_left = seg_steps < steps_left ? seg_steps : steps_left;
    int seg = 0;
    bool seg_open = false;
    for (int i = 0; i < nkb; ++i) {
      if (MK == M_F16) mbar_wait(&ctl->raw_full[rs.stage], rs.phase);
      mbar_wait(&ctl->op_full[os.stage], os.phase);
      tc_fence_after();
      const uint64_t opd = op_desc0 + (uint64_t)((os.stage * kOpStage) >> 4);
      const uint64_t adesc = MK == M_F16 ? raw_desc0 + (uint64_t)((rs.stage * kRawW) >> 4) : opd;
      const uint64_t bdesc = opd + (uint64_t)((MK == M_F16 ? 0 : kABytes) >> 4);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (steps_left > 0) {
          const int buf = seg & 1;
          if (!seg_open) {
            mbar_wait(&ctl->acc_empty[buf], ((seg >> 1) & 1) ^ 1);
            tc_fence_after();
          }
          --steps_left;
          --seg_left;
          const bool close = seg_left == 0;
          if (elect_one()) {
            umma_f16(tmem_base + buf * BT, adesc + 2 * k, bdesc + 2 * k, idesc, seg_open ? 1u : 0u);
            if (close) tc_commit(&ctl->acc_full[buf]);
          }
          seg_open = !close;
          if (close) {
            ++seg;
            seg_left = seg_steps < steps_left ? seg_steps : steps_left;
          }
        }
      }
      if (elect_one()) {
        tc_commit(&ctl->op_empty[os.stage]);
        if (MK == M_F16) tc_commit(&ctl->raw_empty[rs.stage]);
      }
      __syncwarp();
      rs.advance();
      os.advance();
    }
  } else if (warp < 6) {
    // ===================== epilogue =====================
    const int quad = warp & 3;
    const int epi_tid = quad * 32 + lane;
    const int n = n_tile * kBlockW + epi_tid;
    const bool n_ok = n < p.N;
    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
    float acc[BT];
#pragma unroll
    for (int j = 0; j < BT; ++j) acc[j] = 0.f;
    if (MK != M_F16) {
      const int ngroups = (steps_total * 16 + seg_len - 1) / seg_len;
      if (n_ok) {
        const float* srow = p.scales + (int64_t)expert * p.s_stride_e + (int64_t)(n / p.group_n) * p.s_stride_n;
        for (int g = 0; g < ngroups; ++g)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(scale_tab + g * kBlockW + epi_tid)),
                       "l"(srow + (int64_t)g * p.s_stride_k) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    int k_pos = 0, seg = 0;
    const int k_lim = steps_total * 16;
    while (k_pos < k_lim) {
      int k_next = (k_pos / seg_len + 1) * seg_len;
      if (k_next > k_lim) k_next = k_lim;
      float sc = 1.0f;
      if (MK != M_F16) sc = n_ok ? scale_tab[(k_pos / seg_len) * kBlockW + epi_tid] : 0.f;
      const int buf = seg & 1;
      mbar_wait(&ctl->acc_full[buf], (seg >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < BT; c += 16) {
        uint32_t pv[16];
        tmem_ld_32x32b_x16(taddr + buf * BT + c, pv);
        tc_wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[c + j] = fmaf(sc, __uint_as_float(pv[j]), acc[c + j]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl->acc_empty[buf]);
      k_pos = k_next;
      ++seg;
    }
    if (n_ok) {
#pragma unroll
      for (int j = 0; j < BT; ++j) {
        const int slot = p.sorted_ids[blk * BT + j];
        if (slot < p.num_valid_slots) {
          float v = acc[j];
          if (p.topk_weights) v *= __half2float(p.topk_weights[slot]);   // fp32 multiply before the cast
          p.c[(int64_t)slot * p.ldc + n] = __float2half_rn(v);
        }
      }
    }
  } else {
    // ===================== transform warps: gather X rows (+ dequantise W) =====================
    const int dt = threadIdx.x - 6 * 32;          // 0..255
    // source row of every tile row this thread copies (fixed for the whole K loop)
    constexpr int kChunks = BT * 8;               // 16-byte chunks of one X stage
    constexpr int kPerThread = (kChunks + 255) / 256;
    const __half* xrow[kPerThread];
    uint32_t xdst[kPerThread];
#pragma unroll
    for (int i = 0; i < kPerThread; ++i) {
      const int c = dt + i * 256;
      xrow[i] = nullptr;
      xdst[i] = 0;
      if (c < kChunks) {
        const int r = c >> 3, j = c & 7;
        const int slot = p.sorted_ids[blk * BT + r];
        if (slot < p.num_valid_slots) xrow[i] = p.a + (int64_t)(slot / p.top_k) * p.lda + j * 8;
        xdst[i] = r * 128 + ((j ^ (r & 7)) << 4);
      }
    }
    const int row = dt & 127, hsel = dt >> 7, rx = row & 7;   // conflict-free swizzled 16-byte accesses
    PipeState rs(p.raw_stages), os(p.op_stages);
    for (int i = 0; i < nkb; ++i) {
      const int sub = i % RW;
      mbar_wait(&ctl->op_empty[os.stage], os.phase ^ 1);
      uint8_t* opst = op_ring + (size_t)os.stage * kOpStage;
      uint8_t* xst = opst + (MK == M_F16 ? 0 : kABytes);
      // activations of this k-block (zero rows for padding slots; K tail zero-filled)
#pragma unroll
      for (int q = 0; q < kPerThread; ++q) {
        if (dt + q * 256 < kChunks) {
          uint4 v = make_uint4(0u, 0u, 0u, 0u);
          const int kcol = i * 64 + ((dt + q * 256) & 7) * 8;
          if (xrow[q] != nullptr && kcol < p.K) v = ldg_128(xrow[q] + i * 64);
          *reinterpret_cast<uint4*>(xst + xdst[q]) = v;
        }
      }
      if (MK != M_F16) {
        if (sub == 0) mbar_wait(&ctl->raw_full[rs.stage], rs.phase);
        const uint8_t* rawrow = raw_ring + (size_t)rs.stage * kRawW + row * 128;
        const int c0 = sub * 4 + hsel * 2;        // this k-block's 64 bytes = logical chunks 4*sub .. 4*sub+3
        const uint4 lo4 = *reinterpret_cast<const uint4*>(rawrow + ((c0 ^ rx) << 4));
        const uint4 hi4 = *reinterpret_cast<const uint4*>(rawrow + (((c0 + 1) ^ rx) << 4));
        const uint32_t ws[8] = {lo4.x, lo4.y, lo4.z, lo4.w, hi4.x, hi4.y, hi4.z, hi4.w};
        uint8_t* arow = opst + row * 128;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int chunk = hsel * 4 + j;
          *reinterpret_cast<uint4*>(arow + ((chunk ^ (row & 7)) << 4)) = moe_dequant8(ws[2 * j], ws[2 * j + 1], MK == M_FP8);
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      const bool raw_done = (sub == RW - 1) || (i + 1 == nkb);
      if (lane == 0) {
        mbar_arrive(&ctl->op_full[os.stage]);
        if (MK != M_F16 && raw_done) mbar_arrive(&ctl->raw_empty[rs.stage]);
      }
      os.advance();
      if (raw_done) rs.advance();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

struct MoePlan { int raw_stages, smem, max_groups, op_stages; };

static MoePlan moe_plan(int mk, int bt, int64_t K, int seg_len) {
  MoePlan pl;
  const int rawW = kABytes;
  const int rw = mk == M_F16 ? 1 : 2;
  const int op_stage = (mk == M_F16 ? 0 : kABytes) + bt * kStageKBytes;
  pl.max_groups = mk == M_F16 ? 0 : (int)((K + seg_len - 1) / seg_len) + 1;
  const int nkb = (int)cdiv64(K, 64);
  // Decode-sized blocks (16 slots): a CTA streams one 128-row weight tile of ONE expert and is gone -- thousands of
  // short CTAs per launch (Qwen3-30B-A3B at 32 tokens: 115 touched experts x 12-16 tiles), each paying set-up, the
  // first-byte latency and the epilogue alone when one CTA fills the SM.  A half-depth operand ring + 4 raw stages
  // fit TWO CTAs per SM (<= 113 KiB, 64 registers per thread): one streams while the other starts or drains.
  pl.op_stages = kMoeOpStages;
  int cap = b2l_max_smem_optin();
  int max_raw = kMoeMaxRaw;
  if (bt == 16) {
    pl.op_stages = 2;
    cap = (b2l_max_smem_optin() + 1024) / 2 - 1024;     // two CTAs (+1 KiB reserved each) per SM
    max_raw = 4;
  }
  const int fixed = 1024 + pl.op_stages * op_stage + pl.max_groups * kBlockW * 4 + 1024;
  int stages = (cap - fixed) / rawW;
  if (stages > max_raw) stages = max_raw;
  if (stages > nkb / rw + 1) stages = nkb / rw + 1;
  if (stages < 2) {   // does not fit twice: one CTA per SM with the full rings
    pl.op_stages = kMoeOpStages;
    const int fixed1 = 1024 + pl.op_stages * op_stage + pl.max_groups * kBlockW * 4 + 1024;
    stages = (b2l_max_smem_optin() - fixed1) / rawW;
    if (stages > kMoeMaxRaw) stages = kMoeMaxRaw;
    if (stages > nkb / rw + 1) stages = nkb / rw + 1;
    if (stages < 2) stages = 2;
    pl.raw_stages = stages;
    pl.smem = fixed1 + stages * rawW;
    return pl;
  }
  pl.raw_stages = stages;
  pl.smem = fixed + stages * rawW;
  return pl;
}

template <int MK, int BT>
static int launch_moe(const CUtensorMap& tw, const MoeParams& p, const MoePlan& pl, int n_tiles, int max_blocks,
                      cudaStream_t st) {
  auto kern = moe_gemm_kernel<MK, BT>;
  static int configured[64] = {0};
  B2L_CUDA_CHECK(b2l_ensure_dynamic_smem(kern, pl.smem, configured));
  dim3 grid((unsigned)n_tiles, (unsigned)max_blocks);
  b2l_launch(kern, dim3(grid), dim3(kMoeThreads), (size_t)(pl.smem), st, tw, p);
  B2L_LAUNCH_CHECK();
  return B2L_OK;
}

}  // namespace gemm

using namespace gemm;

// block_size must be 16, 32 or 64 and equal to the block size moe_align_block_size was run with.
extern "C" int b2l_moe_gemm(void* c, const void* a, const void* w, int weight_kind, const void* scales,
                            const void* topk_weights, const void* sorted_token_ids, const void* expert_ids,
                            const void* num_tokens_post_padded, int64_t em, int64_t num_valid_slots, int top_k,
                            int block_size, int num_experts, int64_t n, int64_t k, int64_t a_row_stride,
                            int64_t w_row_stride, int64_t c_row_stride, int64_t group_n, int64_t group_k,
                            int64_t s_stride_e, int64_t s_stride_n, int64_t s_stride_k, void* stream) {
  B2L_REQUIRE(c && a && w && sorted_token_ids && expert_ids && num_tokens_post_padded, "moe_gemm: null pointer");
  B2L_REQUIRE(weight_kind == M_F16 || weight_kind == M_FP8 || weight_kind == M_INT8, "moe_gemm: weight kind %d", weight_kind);
  B2L_REQUIRE(block_size == 16 || block_size == 32 || block_size == 64, "moe_gemm: block_size %d (16, 32 or 64)", block_size);
  B2L_REQUIRE(top_k >= 1 && num_experts >= 1 && n > 0 && k > 0, "moe_gemm: bad shape");
  B2L_REQUIRE(a_row_stride % 8 == 0 && (uintptr_t)a % 16 == 0 && k % 8 == 0, "moe_gemm: activations must be 16-byte aligned rows");
  B2L_REQUIRE(weight_kind == M_F16 || scales, "moe_gemm: quantised experts need scales");
  if (em == 0) return B2L_OK;
  int seg_len = (int)k;
  if (weight_kind != M_F16) {
    int64_t gk = group_k < k ? group_k : k;
    B2L_REQUIRE(group_n >= 1 && group_k >= 1, "moe_gemm: bad scale block");
    B2L_REQUIRE(gk >= k || gk % 16 == 0, "moe_gemm: group_k (%lld) must be a multiple of 16 unless it covers K", (long long)group_k);
    seg_len = (int)((gk + 15) / 16 * 16);
    B2L_REQUIRE(k % 16 == 0 && w_row_stride % 16 == 0, "moe_gemm: 8-bit experts need K and row stride multiples of 16");
  }
  MoePlan pl = moe_plan(weight_kind, block_size, k, seg_len);
  MoeParams p;
  p.a = (const __half*)a; p.lda = a_row_stride; p.c = (__half*)c; p.ldc = c_row_stride;
  p.sorted_ids = (const int32_t*)sorted_token_ids; p.expert_ids = (const int32_t*)expert_ids;
  p.num_post = (const int32_t*)num_tokens_post_padded; p.topk_weights = (const __half*)topk_weights;
  p.scales = (const float*)scales; p.s_stride_e = s_stride_e; p.s_stride_n = s_stride_n; p.s_stride_k = s_stride_k;
  p.group_n = (int)(group_n > 0 ? (group_n > n ? n : group_n) : 1); p.group_k = seg_len;
  p.N = (int)n; p.K = (int)k; p.num_kb = (int)cdiv64(k, 64);
  p.num_valid_slots = (int)num_valid_slots; p.top_k = top_k;
  p.raw_stages = pl.raw_stages; p.max_groups = pl.max_groups; p.op_stages = pl.op_stages;
  CUtensorMap tw;
  int rc;
  if (weight_kind == M_F16)
    rc = b2l_make_tmap_2d(&tw, w, TMA_F16, (uint64_t)num_experts * n, (uint64_t)k, (uint64_t)w_row_stride * 2, kBlockW, 64, true);
  else
    rc = b2l_make_tmap_2d(&tw, w, TMA_U8, (uint64_t)num_experts * n, (uint64_t)k, (uint64_t)w_row_stride, kBlockW, 128, true);
  if (rc != B2L_OK) return rc;
  const int n_tiles = (int)cdiv64(n, kBlockW);
  const int max_blocks = (int)((em + block_size - 1) / block_size);   // a trailing partial block is never < num_post
  cudaStream_t st = (cudaStream_t)stream;
#define MOE_DISPATCH(MKV)                                                                      \
  switch (block_size) {                                                                        \
    case 16: return launch_moe<MKV, 16>(tw, p, pl, n_tiles, max_blocks, st);                    \
    case 32: return launch_moe<MKV, 32>(tw, p, pl, n_tiles, max_blocks, st);                    \
    default: return launch_moe<MKV, 64>(tw, p, pl, n_tiles, max_blocks, st);                    \
  }
  if (weight_kind == M_F16) { MOE_DISPATCH(M_F16) }
  if (weight_kind == M_FP8) { MOE_DISPATCH(M_FP8) }
  MOE_DISPATCH(M_INT8)
#undef MOE_DISPATCH
}
