// flash_attention2_no_pad on the 5th-generation tensor cores (tcgen05 + TMEM), head_dim 128.
//
// Replaces lite_llama/kernels/flashattention2_nopad.py:45-231 for the shapes the decoder uses (D = 128, heads packed
// inside a token row); same contract as flash_attention_nopad.cu, which stays the path for other head sizes.
// The reference's Triton build already lowers this op to tcgen05.mma on sm_100a, so the legacy
// mma.sync kernel lost to it 2x (profiles/r2c_triton_vs_b200.md); this kernel is the B200-native one.
//
// One CTA = 128 query rows of one (sequence, q head), 192 threads:
//   warp 0      loader: Q once, then 128-key K and V tiles through separate 3-deep rings (2-D TMA, 128B swizzle;
//               K is free as soon as S = QK^T has retired, V only after PV)
//   warp 1      tensor-core issuer + TMEM owner:  S_j = Q K_j^T  (A, B from shared memory, K-major) and
//               O_j = P_j V_j  (A = P_j read from TENSOR MEMORY, B = V_j from shared memory, MN-major descriptor:
//               V stays [keys][d] exactly as TMA wrote it) -- 16 tcgen05.mma of 128x128x16 per tile
//   warps 2-5   softmax: thread = query row = TMEM lane, so row max / row sum need no shuffles.  Per tile: ONE read of
//               the whole score row S_j from TMEM into registers, softmax in the log2 domain (fp32), P_j (rounded to
//               the activation dtype, flashattention2_nopad.py:156) written back over the columns S_j occupied.
//               The output accumulates IN TENSOR MEMORY across tiles (O += P_j V_j); the running maximum used for
//               scaling is only raised when a tile exceeds it by more than 2^8 ("lazy rescale"), and only then is the
//               O row read, multiplied and written back -- exact, because numerator and denominator share one scale.
//               (v1 read S twice and folded every tile's O through registers: 16 exposed TMEM round trips per tile with
//               one warp per scheduler; ~7700 cycles per tile against 1024 of tensor-core time.)
// TMEM map (512 columns): S0 | S1 | O, each 128 fp32 columns; P_j aliases the first 64 columns of S_j.
// S is double buffered, so the tensor core computes S_{j+1} while the softmax warps work on S_j.
// Causality: only tiles up to the diagonal are loaded; only the diagonal tile is masked.
#include "common.cuh"
#include "tma_desc.cuh"
#include <type_traits>

namespace fatc {

constexpr int kBM = 128, kBN = 128, kD = 128;
constexpr int kStages = 3;
constexpr int kTileBytes = kBN * kD * 2;          // 32 KiB: two 128B-swizzled halves of 64 d each
constexpr int kHalfBytes = kTileBytes / 2;
constexpr int kThreads = 192;
constexpr int kSmemBytes = 1024 /*align*/ + 1024 /*control*/ + kTileBytes * (1 + 2 * kStages);

struct __align__(8) Ctl {
  uint64_t q_full;
  uint64_t k_full[kStages], k_empty[kStages];
  uint64_t v_full[kStages], v_empty[kStages];
  uint64_t s_full[2], p_full[2], o_full;
  uint32_t tmem_base;
};
static_assert(sizeof(Ctl) <= 1024, "control block too large");

struct Params {
  void* out; int64_t o_ts, o_hs;
  const void* start_loc; int start_w;
  const void* seq_len; int len_w;
  int G;                       // q heads per kv head
  float scale_log2;
};

__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
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
__device__ __forceinline__ void tc_wait_st_() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}

template <typename T>
__global__ void __launch_bounds__(kThreads, 1)
fa_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
             const __grid_constant__ CUtensorMap tmap_v, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  Ctl* ctl = reinterpret_cast<Ctl*>(smem);
  uint8_t* q_s = smem + 1024;
  uint8_t* k_ring = q_s + kTileBytes;
  uint8_t* v_ring = k_ring + kStages * kTileBytes;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // heavy (late) query blocks first: under causality block m needs m + 1 key tiles
  const int mblk = (int)gridDim.x - 1 - (int)blockIdx.x;
  const int head = blockIdx.y, b = blockIdx.z;
  const int kvh = head / p.G;
  pdl_wait_prior();                              // lengths, q, k, v are upstream data
  const int64_t seq_len = load_index(p.seq_len, b, p.len_w);
  const int64_t q0 = (int64_t)mblk * kBM;
  if (q0 >= seq_len) return;                     // whole CTA, before any barrier or TMEM state exists
  const int64_t seq_start = load_index(p.start_loc, b, p.start_w);
  const int64_t q_hi = q0 + kBM < seq_len ? q0 + kBM : seq_len;
  const int ntiles = (int)((q_hi + kBN - 1) / kBN);          // keys 0 .. q_hi - 1

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    mbar_init(&ctl->q_full, 1);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&ctl->k_full[s], 1); mbar_init(&ctl->k_empty[s], 1);
      mbar_init(&ctl->v_full[s], 1); mbar_init(&ctl->v_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&ctl->s_full[s], 1); mbar_init(&ctl->p_full[s], 4);
    }
    mbar_init(&ctl->o_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(&ctl->tmem_base, 512);
    tmem_relinquish();
    pdl_launch_dependents();                     // only once this CTA owns its tensor memory (gemm_dense.cu)
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;

  if (warp == 0) {
    // ===================== loader =====================
    if (elect_one()) {
      const int tok0 = (int)(seq_start + q0);
      mbar_arrive_expect_tx(&ctl->q_full, kTileBytes);
      tma_load_2d(q_s, &tmap_q, head * kD, tok0, &ctl->q_full);
      tma_load_2d(q_s + kHalfBytes, &tmap_q, head * kD + 64, tok0, &ctl->q_full);
    }
    __syncwarp();
    for (int j = 0; j < ntiles; ++j) {
      const int st = j % kStages;
      const uint32_t ph = (uint32_t)(j / kStages) & 1u;
      const int tok = (int)(seq_start + (int64_t)j * kBN);
      mbar_wait(&ctl->k_empty[st], ph ^ 1u);
      if (elect_one()) {
        mbar_arrive_expect_tx(&ctl->k_full[st], kTileBytes);
        tma_load_2d(k_ring + st * kTileBytes, &tmap_k, kvh * kD, tok, &ctl->k_full[st]);
        tma_load_2d(k_ring + st * kTileBytes + kHalfBytes, &tmap_k, kvh * kD + 64, tok, &ctl->k_full[st]);
      }
      __syncwarp();
      mbar_wait(&ctl->v_empty[st], ph ^ 1u);
      if (elect_one()) {
        mbar_arrive_expect_tx(&ctl->v_full[st], kTileBytes);
        tma_load_2d(v_ring + st * kTileBytes, &tmap_v, kvh * kD, tok, &ctl->v_full[st]);
        tma_load_2d(v_ring + st * kTileBytes + kHalfBytes, &tmap_v, kvh * kD + 64, tok, &ctl->v_full[st]);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== tensor-core issuer =====================
    constexpr bool bf = std::is_same<T, __nv_bfloat16>::value;
    constexpr uint32_t idesc_s = umma_idesc(1u, bf ? 1u : 0u, bf ? 1u : 0u, 0u, 0u, kBM, kBN);   // Q K^T: both K-major
    constexpr uint32_t idesc_o = umma_idesc(1u, bf ? 1u : 0u, bf ? 1u : 0u, 0u, 1u, kBM, kD);    // P V: V is MN-major
    const uint64_t q_desc = umma_desc_sw128_kmajor(smem_u32(q_s));
    auto issue_s = [&](int j) {
      const int st = j % kStages;
      mbar_wait(&ctl->k_full[st], (uint32_t)(j / kStages) & 1u);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t k_desc = umma_desc_sw128_kmajor(smem_u32(k_ring + st * kTileBytes));
        const uint32_t d = tmem_base + (uint32_t)(j & 1) * 128u;
#pragma unroll
        for (int kk = 0; kk < kD / 16; ++kk) {
          // 16 d-values per step: 32 bytes inside a 64-wide swizzle atom, the second atom column is 16 KiB away
          const uint64_t off = (uint64_t)(((kk >> 2) * kHalfBytes + (kk & 3) * 32) >> 4);
          umma_f16(d, q_desc + off, k_desc + off, idesc_s, kk ? 1u : 0u);
        }
        tc_commit(&ctl->k_empty[st]);
        tc_commit(&ctl->s_full[j & 1]);
      }
      __syncwarp();
    };
    auto issue_pv = [&](int j) {
      const int st = j % kStages;
      mbar_wait(&ctl->v_full[st], (uint32_t)(j / kStages) & 1u);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t v_addr = smem_u32(v_ring + st * kTileBytes);
        const uint32_t d = tmem_base + 256u;                           // the output accumulates across tiles
        const uint32_t a = tmem_base + (uint32_t)(j & 1) * 128u;       // P_j: 64 columns of packed pairs
#pragma unroll
        for (int kk = 0; kk < kBN / 16; ++kk) {
          // 16 keys per step = two 8-row groups of 1024 B; the d 64..127 half is one LBO (16 KiB) away
          const uint64_t v_desc = umma_desc_sw128_mnmajor(v_addr + kk * 2048, kHalfBytes, 1024);
          umma_ts(d, a + kk * 8, v_desc, idesc_o, (j | kk) ? 1u : 0u);
        }
        tc_commit(&ctl->v_empty[st]);
        tc_commit(&ctl->o_full);
      }
      __syncwarp();
    };
    mbar_wait(&ctl->q_full, 0);
    tc_fence_after();
    issue_s(0);
    for (int j = 0; j < ntiles; ++j) {
      if (j + 1 < ntiles) issue_s(j + 1);        // the tensor core works on S_{j+1} while the softmax warps are on S_j
      // P_j is in place and -- the softmax warps arrive only after any rescale of O -- the output may be added to
      mbar_wait(&ctl->p_full[j & 1], (uint32_t)(j >> 1) & 1u);
      tc_fence_after();
      issue_pv(j);
    }
  } else {
    // ===================== softmax warps: thread = query row = TMEM lane =====================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int64_t qpos = q0 + row;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    float m_used = -INFINITY, l_run = 0.f;      // m_used: the maximum the accumulated P / O / l are scaled by
    constexpr float kLazy = 8.0f;               // log2 units: rescale only when a tile exceeds m_used by 2^8

    auto tile = [&](const int j, auto masked) {
      constexpr bool kMasked = decltype(masked)::value;
      mbar_wait(&ctl->s_full[j & 1], (uint32_t)(j >> 1) & 1u);
      tc_fence_after();
      const uint32_t sb = lane_base + (uint32_t)(j & 1) * 128u;
      const int64_t key0 = (int64_t)j * kBN;
      // the whole score row in one round trip
      uint32_t sv[kBN];
#pragma unroll
      for (int c = 0; c < kBN; c += 32) tmem_ld_32x32b_x32(sb + c, *reinterpret_cast<uint32_t(*)[32]>(sv + c));
      tc_wait_ld();
      // causal mask of the diagonal tile as a per-thread column limit; branch-free selects (the first version's
      // 64-bit predicates compiled into ~600 integer compares and a branch around every exp2 per tile)
      const int limit = (int)(qpos - key0);                      // masked tile: key column i is visible iff i <= limit
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < kBN; ++i) {
        float v = __uint_as_float(sv[i]);
        if (kMasked) v = i <= limit ? v : -INFINITY;
        mx = fmaxf(mx, v);
      }
      const float m_tile = mx * p.scale_log2;
      const bool grow = m_tile > m_used + kLazy;                 // first tile: m_used = -inf -> true
      const float m_new = grow ? m_tile : m_used;
      const float base = m_new == -INFINITY ? 0.f : m_new;       // fully masked (padding) rows: avoid inf - inf
      const float alpha = grow ? fast_exp2(m_used - base) : 1.0f;
      float lsum = 0.f;
      {
#pragma unroll
        for (int c = 0; c < kBN; c += 32) {
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float p0 = fast_exp2(fmaf(__uint_as_float(sv[c + i]), p.scale_log2, -base));
            float p1 = fast_exp2(fmaf(__uint_as_float(sv[c + i + 1]), p.scale_log2, -base));
            if (kMasked) {                                         // selects, not branches around the exp2
              p0 = c + i <= limit ? p0 : 0.f;
              p1 = c + i + 1 <= limit ? p1 : 0.f;
            }
            lsum += p0 + p1;
            pk[i >> 1] = pack2<T>(p0, p1);
          }
          tmem_st_32x32b_x16(sb + (c >> 1), pk);
        }
      }
      l_run = l_run * alpha + lsum;
      m_used = m_new;
      if (j > 0) {
        // every softmax warp follows EVERY phase of o_full (no skipped phases, no parity aliasing); by now PV_{j-1}
        // has long retired, so this wait is free in the common case
        mbar_wait(&ctl->o_full, (uint32_t)(j - 1) & 1u);
        if (__any_sync(0xffffffffu, grow)) {
          tc_fence_after();
          const uint32_t ob = lane_base + 256u;
#pragma unroll
          for (int c = 0; c < kD; c += 32) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(ob + c, v);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
            tmem_st_32x32b_x32(ob + c, v);
          }
        }
      }
      tc_wait_st_();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl->p_full[j & 1]);
    };
    // only the LAST tile of a query block touches the diagonal: every other tile runs the copy without mask code
    for (int j = 0; j + 1 < ntiles; ++j) tile(j, std::false_type{});
    tile(ntiles - 1, std::true_type{});
    // the last PV, then the normalised row goes out
    mbar_wait(&ctl->o_full, (uint32_t)(ntiles - 1) & 1u);
    tc_fence_after();
    const bool row_ok = qpos < seq_len;
    {
      const float inv = 1.0f / l_run;
      T* orow = reinterpret_cast<T*>(p.out) + (seq_start + (row_ok ? qpos : 0)) * p.o_ts + (int64_t)head * p.o_hs;
#pragma unroll
      for (int c = 0; c < kD; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(lane_base + 256u + c, v);          // warp-collective: outside the row predicate
        tc_wait_ld();
        if (row_ok) {
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            uint4 w;
            w.x = pack2<T>(__uint_as_float(v[i]) * inv, __uint_as_float(v[i + 1]) * inv);
            w.y = pack2<T>(__uint_as_float(v[i + 2]) * inv, __uint_as_float(v[i + 3]) * inv);
            w.z = pack2<T>(__uint_as_float(v[i + 4]) * inv, __uint_as_float(v[i + 5]) * inv);
            w.w = pack2<T>(__uint_as_float(v[i + 6]) * inv, __uint_as_float(v[i + 7]) * inv);
            *reinterpret_cast<uint4*>(orow + c + i) = w;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------------------
// Prompts of at most 128 tokens (the BASELINE prefill shape: 32 x 128): every (sequence, head) is ONE 128 x 128 tile,
// so the multi-tile pipeline above is all set-up.  This variant keeps the whole score row in registers (one TMEM read,
// no running output), needs 96 KiB of shared memory and 256 TMEM columns, and is compiled for two resident CTAs per
// SM: the loads of one CTA overlap the tensor-core and softmax work of the other.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kSmemOneTile = 1024 + 1024 + 3 * kTileBytes;

struct __align__(8) Ctl1 {
  uint64_t qk_full, v_full, s_full, p_full, o_full;
  uint32_t tmem_base;
};

template <typename T>
__global__ void __launch_bounds__(kThreads, 2)
fa_tc_one_tile_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                      const __grid_constant__ CUtensorMap tmap_v, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  Ctl1* ctl = reinterpret_cast<Ctl1*>(smem);
  uint8_t* q_s = smem + 1024;
  uint8_t* k_s = q_s + kTileBytes;
  uint8_t* v_s = k_s + kTileBytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.y, b = blockIdx.z;
  const int kvh = head / p.G;
  pdl_wait_prior();
  const int64_t seq_len = load_index(p.seq_len, b, p.len_w);
  if (seq_len <= 0) return;
  const int64_t seq_start = load_index(p.start_loc, b, p.start_w);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    mbar_init(&ctl->qk_full, 1); mbar_init(&ctl->v_full, 1); mbar_init(&ctl->s_full, 1);
    mbar_init(&ctl->p_full, 4); mbar_init(&ctl->o_full, 1);
    mbar_fence_init();
    // the loads do not need the tensor memory: issue them before the set-up barrier
    const int tok = (int)seq_start;
    mbar_arrive_expect_tx(&ctl->qk_full, 2 * kTileBytes);
    tma_load_2d(q_s, &tmap_q, head * kD, tok, &ctl->qk_full);
    tma_load_2d(q_s + kHalfBytes, &tmap_q, head * kD + 64, tok, &ctl->qk_full);
    tma_load_2d(k_s, &tmap_k, kvh * kD, tok, &ctl->qk_full);
    tma_load_2d(k_s + kHalfBytes, &tmap_k, kvh * kD + 64, tok, &ctl->qk_full);
    mbar_arrive_expect_tx(&ctl->v_full, kTileBytes);
    tma_load_2d(v_s, &tmap_v, kvh * kD, tok, &ctl->v_full);
    tma_load_2d(v_s + kHalfBytes, &tmap_v, kvh * kD + 64, tok, &ctl->v_full);
  }
  if (warp == 1) {
    tmem_alloc(&ctl->tmem_base, 256);
    tmem_relinquish();
    pdl_launch_dependents();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;

  if (warp == 1) {
    constexpr bool bf = std::is_same<T, __nv_bfloat16>::value;
    constexpr uint32_t idesc_s = umma_idesc(1u, bf ? 1u : 0u, bf ? 1u : 0u, 0u, 0u, kBM, kBN);
    constexpr uint32_t idesc_o = umma_idesc(1u, bf ? 1u : 0u, bf ? 1u : 0u, 0u, 1u, kBM, kD);
    mbar_wait(&ctl->qk_full, 0);
    tc_fence_after();
    if (elect_one()) {
      const uint64_t q_desc = umma_desc_sw128_kmajor(smem_u32(q_s));
      const uint64_t k_desc = umma_desc_sw128_kmajor(smem_u32(k_s));
#pragma unroll
      for (int kk = 0; kk < kD / 16; ++kk) {
        const uint64_t off = (uint64_t)(((kk >> 2) * kHalfBytes + (kk & 3) * 32) >> 4);
        umma_f16(tmem_base, q_desc + off, k_desc + off, idesc_s, kk ? 1u : 0u);
      }
      tc_commit(&ctl->s_full);
    }
    __syncwarp();
    mbar_wait(&ctl->p_full, 0);
    mbar_wait(&ctl->v_full, 0);
    tc_fence_after();
    if (elect_one()) {
      const uint32_t v_addr = smem_u32(v_s);
#pragma unroll
      for (int kk = 0; kk < kBN / 16; ++kk) {
        const uint64_t v_desc = umma_desc_sw128_mnmajor(v_addr + kk * 2048, kHalfBytes, 1024);
        umma_ts(tmem_base + 128u, tmem_base + kk * 8, v_desc, idesc_o, kk ? 1u : 0u);
      }
      tc_commit(&ctl->o_full);
    }
    __syncwarp();
  } else if (warp >= 2) {
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    mbar_wait(&ctl->s_full, 0);
    tc_fence_after();
    float s[kBN];
#pragma unroll
    for (int c = 0; c < kBN; c += 32) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(lane_base + c, v);
      tc_wait_ld();
#pragma unroll
      for (int i = 0; i < 32; ++i) s[c + i] = __uint_as_float(v[i]);
    }
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < kBN; ++i) mx = fmaxf(mx, i <= row ? s[i] : -INFINITY);   // causal: key i visible iff i <= row
    const float base = mx * p.scale_log2;
    float lsum = 0.f;
#pragma unroll
    for (int c = 0; c < kBN; c += 32) {
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        float p0 = fast_exp2(fmaf(s[c + i], p.scale_log2, -base));          // selects, not branches around the exp2
        float p1 = fast_exp2(fmaf(s[c + i + 1], p.scale_log2, -base));
        p0 = (c + i <= row) ? p0 : 0.f;
        p1 = (c + i + 1 <= row) ? p1 : 0.f;
        lsum += p0 + p1;
        pk[i >> 1] = pack2<T>(p0, p1);
      }
      tmem_st_32x32b_x16(lane_base + (c >> 1), pk);
    }
    tc_wait_st_();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&ctl->p_full);
    const float inv = 1.0f / lsum;
    mbar_wait(&ctl->o_full, 0);
    tc_fence_after();
    // tcgen05.ld is warp-collective: every lane loads its row, only rows inside the sequence are stored
    const bool row_ok = row < seq_len;
    T* orow = reinterpret_cast<T*>(p.out) + (seq_start + (row_ok ? row : 0)) * p.o_ts + (int64_t)head * p.o_hs;
#pragma unroll
    for (int c = 0; c < kD; c += 32) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(lane_base + 128u + c, v);
      tc_wait_ld();
      if (row_ok) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 w;
          w.x = pack2<T>(__uint_as_float(v[i]) * inv, __uint_as_float(v[i + 1]) * inv);
          w.y = pack2<T>(__uint_as_float(v[i + 2]) * inv, __uint_as_float(v[i + 3]) * inv);
          w.z = pack2<T>(__uint_as_float(v[i + 4]) * inv, __uint_as_float(v[i + 5]) * inv);
          w.w = pack2<T>(__uint_as_float(v[i + 6]) * inv, __uint_as_float(v[i + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + c + i) = w;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 256);
}

template <typename T>
static int launch(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const Params& p, int64_t batch,
                  int num_q_heads, int64_t max_len, cudaStream_t st) {
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    B2L_CUDA_CHECK(cudaFuncSetAttribute(fa_tc_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_set[dev] = true;
  }
  if (max_len <= kBM) {
    static bool attr1_set[64] = {false};
    if (dev >= 0 && dev < 64 && !attr1_set[dev]) {
      B2L_CUDA_CHECK(cudaFuncSetAttribute(fa_tc_one_tile_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemOneTile));
      attr1_set[dev] = true;
    }
    dim3 grid1(1, (unsigned)num_q_heads, (unsigned)batch);
    b2l_launch(fa_tc_one_tile_kernel<T>, dim3(grid1), dim3(kThreads), (size_t)kSmemOneTile, st, tq, tk, tv, p);
    B2L_LAUNCH_CHECK();
    return B2L_OK;
  }
  dim3 grid((unsigned)cdiv64(max_len, kBM), (unsigned)num_q_heads, (unsigned)batch);
  b2l_launch(fa_tc_kernel<T>, dim3(grid), dim3(kThreads), (size_t)kSmemBytes, st, tq, tk, tv, p);
  B2L_LAUNCH_CHECK();
  return B2L_OK;
}

}  // namespace fatc

// Returns B2L_OK when the tcgen05 kernel took the call, 1 when the shape is not its to take (the caller falls back
// to the mma.sync kernel), a negative code on error.
int b2l_fa_tc_try(void* out, const void* q, const void* k, const void* v, const void* b_start_loc, int start_w,
                  const void* b_seq_len, int len_w, int64_t batch, int num_q_heads, int num_kv_heads, int head_dim,
                  int64_t max_seq_len, int64_t total_tokens, float sm_scale_log2, int64_t q_ts, int64_t q_hs,
                  int64_t k_ts, int64_t k_hs, int64_t v_ts, int64_t v_hs, int64_t o_ts, int64_t o_hs, int dtype,
                  cudaStream_t st) {
  using namespace fatc;
  if (head_dim != kD || total_tokens <= 0) return 1;
  if (q_hs != kD || k_hs != kD || v_hs != kD) return 1;                       // heads packed inside a token row
  if (q_ts % 8 || k_ts % 8 || v_ts % 8 || o_ts % 8 || o_hs % 8) return 1;
  if (((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)out) % 16) return 1;
  if (dtype != B2L_F16 && dtype != B2L_BF16) return 1;
  if (total_tokens >= (1ll << 31)) return 1;
  const TmaElem el = dtype == B2L_F16 ? TMA_F16 : TMA_BF16;
  CUtensorMap tq, tk, tv;
  int rc = b2l_make_tmap_2d(&tq, q, el, (uint64_t)total_tokens, (uint64_t)num_q_heads * kD, (uint64_t)q_ts * 2, kBM, 64, true);
  if (rc != B2L_OK) return rc;
  rc = b2l_make_tmap_2d(&tk, k, el, (uint64_t)total_tokens, (uint64_t)num_kv_heads * kD, (uint64_t)k_ts * 2, kBN, 64, true);
  if (rc != B2L_OK) return rc;
  rc = b2l_make_tmap_2d(&tv, v, el, (uint64_t)total_tokens, (uint64_t)num_kv_heads * kD, (uint64_t)v_ts * 2, kBN, 64, true);
  if (rc != B2L_OK) return rc;
  Params p;
  p.out = out; p.o_ts = o_ts; p.o_hs = o_hs;
  p.start_loc = b_start_loc; p.start_w = start_w; p.seq_len = b_seq_len; p.len_w = len_w;
  p.G = num_q_heads / num_kv_heads; p.scale_log2 = sm_scale_log2;
  if (dtype == B2L_F16) return launch<__half>(tq, tk, tv, p, batch, num_q_heads, max_seq_len, st);
  return launch<__nv_bfloat16>(tq, tk, tv, p, batch, num_q_heads, max_seq_len, st);
}
