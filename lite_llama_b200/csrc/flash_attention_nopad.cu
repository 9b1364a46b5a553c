// flash_attention2_no_pad: causal variable-length prefill attention (FlashAttention-2 schedule).
//
// Replaces lite_llama/kernels/flashattention2_nopad.py:45-231.  Contract kept:
//   * q [T, Hq, D], k/v [T, Hkv, D] (strided views of the fused KV projection output),
//     sequence i occupies tokens [b_start_loc[i], b_start_loc[i] + b_seq_len[i]);
//   * sm_scale already contains log2(e): probabilities are exp2(qk*sm_scale - max)   (:136-141);
//   * P is rounded to the activation dtype before the PV product                   (:156);
//   * rows past b_seq_len are left untouched (the reference leaves them uninitialised, :169,192).
//
// One CTA = 64 query rows of one (sequence, q-head); a producer warp streams 64-token K/V tiles
// with 1-D bulk async copies (TMA engine) into a 3-stage mbarrier ring, four consumer warps own
// 16 query rows each and run QK^T / PV on tensor cores with fp32 online softmax.  Tiles above the
// causal diagonal are never loaded.
#include "common.cuh"
#include <stdlib.h>

namespace {

constexpr int kBM = 64;
constexpr int kBN = 64;
constexpr int kStages = 3;
constexpr int kConsumerWarps = 4;
constexpr int kThreads = (kConsumerWarps + 1) * 32;

struct PrefillParams {
  const void* q; int64_t q_ts, q_hs;
  const void* k; int64_t k_ts, k_hs;
  const void* v; int64_t v_ts, v_hs;
  void* out; int64_t o_ts, o_hs;
  const void* start_loc; int start_w;
  const void* seq_len; int len_w;
  int Hq, Hkv, G;
  float scale_log2;
};

template <int D> __host__ __device__ constexpr int row_pitch() { return D * 2 + 16; }
template <int D> __host__ __device__ constexpr int tile_bytes() { return kBN * row_pitch<D>(); }
template <int D> __host__ __device__ constexpr int stage_bytes() { return 2 * tile_bytes<D>(); }
template <int D> __host__ __device__ constexpr int smem_bytes() { return kStages * stage_bytes<D>() + 128; }

template <typename T, int D>
__global__ void __launch_bounds__(kThreads) flash_prefill_kernel(const PrefillParams p) {
  pdl_prologue();   // everything below reads what earlier kernels of the step wrote (q, KV rows, lengths, table)
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty_bar = full_bar + kStages;
  uint8_t* stage_base = smem + 128;

  const int mblk = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int kvh = head / p.G;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t seq_len = load_index(p.seq_len, b, p.len_w);
  const int64_t q0 = (int64_t)mblk * kBM;
  if (q0 >= seq_len) return;
  const int64_t seq_start = load_index(p.start_loc, b, p.start_w);
  const int64_t q_end = q0 + kBM < seq_len ? q0 + kBM : seq_len;   // exclusive
  const int ntiles = (int)((q_end + kBN - 1) / kBN);               // keys 0 .. q_end-1 (causal)

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kConsumerWarps);
    }
    mbar_fence_init();
  }
  __syncthreads();

  constexpr int PITCH = row_pitch<D>();
  constexpr uint32_t ROW_BYTES = D * 2;

  if (warp == kConsumerWarps) {
    // ===================== producer warp =====================
    const char* kbase = reinterpret_cast<const char*>(p.k) + (int64_t)kvh * p.k_hs * 2;
    const char* vbase = reinterpret_cast<const char*>(p.v) + (int64_t)kvh * p.v_hs * 2;
    for (int t = 0; t < ntiles; ++t) {
      const int s = t % kStages;
      mbar_wait(&empty_bar[s], ((t / kStages) & 1) ^ 1);
      if (lane == 0) mbar_arrive_expect_tx(&full_bar[s], 2u * kBN * ROW_BYTES);
      __syncwarp();
      uint8_t* ks = stage_base + s * stage_bytes<D>();
      uint8_t* vs = ks + tile_bytes<D>();
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int r = lane + 32 * i;
        int64_t pos = (int64_t)t * kBN + r;
        pos = pos < seq_len ? pos : seq_len - 1;   // padded rows duplicate a valid row; they are masked
        const int64_t tok = seq_start + pos;
        bulk_g2s(ks + r * PITCH, kbase + tok * p.k_ts * 2, ROW_BYTES, &full_bar[s]);
        bulk_g2s(vs + r * PITCH, vbase + tok * p.v_ts * 2, ROW_BYTES, &full_bar[s]);
      }
    }
  } else {
    // ===================== consumer warps: 16 query rows each =====================
    const int g = lane >> 2, tq = lane & 3;
    const int64_t r0 = q0 + warp * 16 + g, r1 = r0 + 8;      // sequence-relative query positions
    const bool ok0 = r0 < seq_len, ok1 = r1 < seq_len;
    uint32_t qf[D / 16][4];
    {
      const T* qb = reinterpret_cast<const T*>(p.q) + (int64_t)head * p.q_hs;
      const T* qr0 = qb + (seq_start + (ok0 ? r0 : 0)) * p.q_ts;
      const T* qr1 = qb + (seq_start + (ok1 ? r1 : 0)) * p.q_ts;
#pragma unroll
      for (int kk = 0; kk < D / 16; ++kk) {
        const int c = kk * 16 + 2 * tq;
        qf[kk][0] = ok0 ? *reinterpret_cast<const uint32_t*>(qr0 + c) : 0u;
        qf[kk][1] = ok1 ? *reinterpret_cast<const uint32_t*>(qr1 + c) : 0u;
        qf[kk][2] = ok0 ? *reinterpret_cast<const uint32_t*>(qr0 + c + 8) : 0u;
        qf[kk][3] = ok1 ? *reinterpret_cast<const uint32_t*>(qr1 + c + 8) : 0u;
      }
    }
    float o[D / 8][4];
#pragma unroll
    for (int i = 0; i < D / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m_run[2] = {-INFINITY, -INFINITY};
    float l_run[2] = {0.f, 0.f};
    const int64_t warp_last_row = q0 + warp * 16 + 15;   // keys beyond this are masked for the whole warp

    for (int t = 0; t < ntiles; ++t) {
      const int s = t % kStages;
      mbar_wait(&full_bar[s], (t / kStages) & 1);
      const int64_t key0 = (int64_t)t * kBN;
      if (key0 <= warp_last_row) {   // warp-uniform: skip tiles fully above this warp's diagonal
        const uint32_t ks = smem_u32(stage_base + s * stage_bytes<D>());
        const uint32_t vs = ks + tile_bytes<D>();
        float sc[kBN / 8][4];
#pragma unroll
        for (int j = 0; j < kBN / 8; ++j) sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f;
        {
          const int mj = lane >> 3, r = lane & 7;
          const uint32_t a0 = ks + ((mj >> 1) * 8 + r) * PITCH + (mj & 1) * 16;
#pragma unroll
          for (int kk = 0; kk < D / 16; ++kk) {
#pragma unroll
            for (int jp = 0; jp < kBN / 16; ++jp) {
              uint32_t bf[4];
              ldmatrix_x4(bf, a0 + jp * 16 * PITCH + kk * 32);
              mma_m16n8k16<T>(sc[2 * jp], qf[kk], bf[0], bf[1]);
              mma_m16n8k16<T>(sc[2 * jp + 1], qf[kk], bf[2], bf[3]);
            }
          }
        }
        // scale + causal mask (reference masks with -1e8 after scaling; -inf is equivalent under exp2)
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int j = 0; j < kBN / 8; ++j) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int64_t key = key0 + j * 8 + 2 * tq + (e & 1);
            const int64_t row = (e >> 1) ? r1 : r0;
            float v = sc[j][e] * p.scale_log2;
            v = key <= row ? v : -INFINITY;
            sc[j][e] = v;
            mx[e >> 1] = fmaxf(mx[e >> 1], v);
          }
        }
        float alpha[2], base[2];
#pragma unroll
        for (int r2 = 0; r2 < 2; ++r2) {
          mx[r2] = fmaxf(mx[r2], __shfl_xor_sync(0xffffffffu, mx[r2], 1));
          mx[r2] = fmaxf(mx[r2], __shfl_xor_sync(0xffffffffu, mx[r2], 2));
          const float mnew = fmaxf(m_run[r2], mx[r2]);
          base[r2] = mnew == -INFINITY ? 0.f : mnew;
          alpha[r2] = fast_exp2(m_run[r2] - base[r2]);
          m_run[r2] = mnew;
        }
        float ps[2] = {0.f, 0.f};
#pragma unroll
        for (int j = 0; j < kBN / 8; ++j) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float pv = fast_exp2(sc[j][e] - base[e >> 1]);
            sc[j][e] = pv;
            ps[e >> 1] += pv;
          }
        }
        l_run[0] = l_run[0] * alpha[0] + ps[0];
        l_run[1] = l_run[1] * alpha[1] + ps[1];
#pragma unroll
        for (int i = 0; i < D / 8; ++i) {
          o[i][0] *= alpha[0]; o[i][1] *= alpha[0];
          o[i][2] *= alpha[1]; o[i][3] *= alpha[1];
        }
        {
          const int mj = lane >> 3, r = lane & 7;
          const uint32_t a0 = vs + ((mj & 1) * 8 + r) * PITCH + (mj >> 1) * 16;
#pragma unroll
          for (int kt = 0; kt < kBN / 16; ++kt) {
            uint32_t pf[4];
            pf[0] = pack2<T>(sc[2 * kt][0], sc[2 * kt][1]);
            pf[1] = pack2<T>(sc[2 * kt][2], sc[2 * kt][3]);
            pf[2] = pack2<T>(sc[2 * kt + 1][0], sc[2 * kt + 1][1]);
            pf[3] = pack2<T>(sc[2 * kt + 1][2], sc[2 * kt + 1][3]);
#pragma unroll
            for (int dn = 0; dn < D / 16; ++dn) {
              uint32_t bf[4];
              ldmatrix_x4_trans(bf, a0 + kt * 16 * PITCH + dn * 32);
              mma_m16n8k16<T>(o[2 * dn], pf, bf[0], bf[1]);
              mma_m16n8k16<T>(o[2 * dn + 1], pf, bf[2], bf[3]);
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);
    }

#pragma unroll
    for (int r2 = 0; r2 < 2; ++r2) {
      l_run[r2] += __shfl_xor_sync(0xffffffffu, l_run[r2], 1);
      l_run[r2] += __shfl_xor_sync(0xffffffffu, l_run[r2], 2);
    }
    const float inv0 = 1.0f / l_run[0], inv1 = 1.0f / l_run[1];
    T* ob = reinterpret_cast<T*>(p.out) + (int64_t)head * p.o_hs;
    if (ok0) {
      T* orow = ob + (seq_start + r0) * p.o_ts;
#pragma unroll
      for (int i = 0; i < D / 8; ++i)
        *reinterpret_cast<uint32_t*>(orow + i * 8 + 2 * tq) = pack2<T>(o[i][0] * inv0, o[i][1] * inv0);
    }
    if (ok1) {
      T* orow = ob + (seq_start + r1) * p.o_ts;
#pragma unroll
      for (int i = 0; i < D / 8; ++i)
        *reinterpret_cast<uint32_t*>(orow + i * 8 + 2 * tq) = pack2<T>(o[i][2] * inv1, o[i][3] * inv1);
    }
  }
}

template <typename T, int D>
int launch_prefill(const PrefillParams& p, int64_t batch, int64_t max_len, cudaStream_t st) {
  constexpr int smem = smem_bytes<D>();
  static int configured[64] = {0};
  B2L_CUDA_CHECK(b2l_ensure_dynamic_smem(flash_prefill_kernel<T, D>, smem, configured));
  dim3 grid((unsigned)cdiv64(max_len, kBM), (unsigned)p.Hq, (unsigned)batch);
  b2l_launch(flash_prefill_kernel<T, D>, dim3(grid), dim3(kThreads), (size_t)(smem), st, p);
  B2L_LAUNCH_CHECK();
  return B2L_OK;
}

}  // namespace

// flash_attention_tc.cu: the tcgen05 kernel (head_dim 128, packed heads); returns 1 when the shape is not its to take
int b2l_fa_tc_try(void* out, const void* q, const void* k, const void* v, const void* b_start_loc, int start_w,
                  const void* b_seq_len, int len_w, int64_t batch, int num_q_heads, int num_kv_heads, int head_dim,
                  int64_t max_seq_len, int64_t total_tokens, float sm_scale_log2, int64_t q_ts, int64_t q_hs,
                  int64_t k_ts, int64_t k_hs, int64_t v_ts, int64_t v_hs, int64_t o_ts, int64_t o_hs, int dtype,
                  cudaStream_t st);

static int g_fa_tc = -1;   // env B2L_FA_TC=0 keeps the mma.sync kernel (A/B measurements)

extern "C" int b2l_flash_attention2_no_pad(void* out, const void* q, const void* k, const void* v,
                                           const void* b_start_loc, int start_loc_width, const void* b_seq_len,
                                           int seq_len_width, int64_t batch, int num_q_heads, int num_kv_heads,
                                           int head_dim, int64_t max_seq_len, int64_t total_tokens,
                                           float sm_scale_log2, int64_t q_tok_stride,
                                           int64_t q_head_stride, int64_t k_tok_stride, int64_t k_head_stride,
                                           int64_t v_tok_stride, int64_t v_head_stride, int64_t o_tok_stride,
                                           int64_t o_head_stride, int dtype, void* stream) {
  B2L_REQUIRE(out && q && k && v && b_start_loc && b_seq_len, "flash_attention2_no_pad: null pointer");
  B2L_REQUIRE(num_q_heads > 0 && num_kv_heads > 0 && num_q_heads % num_kv_heads == 0,
              "flash_attention2_no_pad: num_q_heads (%d) must be a multiple of num_kv_heads (%d)", num_q_heads,
              num_kv_heads);
  B2L_REQUIRE((start_loc_width == 4 || start_loc_width == 8) && (seq_len_width == 4 || seq_len_width == 8),
              "flash_attention2_no_pad: index widths must be 4 or 8");
  B2L_REQUIRE(k_tok_stride % 8 == 0 && v_tok_stride % 8 == 0 && k_head_stride % 8 == 0 && v_head_stride % 8 == 0 &&
                  (uintptr_t)k % 16 == 0 && (uintptr_t)v % 16 == 0,
              "flash_attention2_no_pad: K/V rows must be 16-byte aligned");
  B2L_REQUIRE(q_tok_stride % 2 == 0 && q_head_stride % 2 == 0 && o_tok_stride % 2 == 0 && o_head_stride % 2 == 0 &&
                  (uintptr_t)q % 4 == 0 && (uintptr_t)out % 4 == 0,
              "flash_attention2_no_pad: q/out rows must be 4-byte aligned");
  if (batch == 0 || max_seq_len <= 0) return B2L_OK;
  if (g_fa_tc < 0) {
    const char* e = getenv("B2L_FA_TC");
    g_fa_tc = (e && e[0] == '0') ? 0 : 1;
  }
  if (g_fa_tc) {
    const int rc = b2l_fa_tc_try(out, q, k, v, b_start_loc, start_loc_width, b_seq_len, seq_len_width, batch, num_q_heads,
                                 num_kv_heads, head_dim, max_seq_len, total_tokens, sm_scale_log2, q_tok_stride,
                                 q_head_stride, k_tok_stride, k_head_stride, v_tok_stride, v_head_stride, o_tok_stride,
                                 o_head_stride, dtype, (cudaStream_t)stream);
    if (rc <= 0) return rc;      // taken (0) or failed (< 0); 1 = not this kernel's shape
  }
  PrefillParams p;
  p.q = q; p.q_ts = q_tok_stride; p.q_hs = q_head_stride;
  p.k = k; p.k_ts = k_tok_stride; p.k_hs = k_head_stride;
  p.v = v; p.v_ts = v_tok_stride; p.v_hs = v_head_stride;
  p.out = out; p.o_ts = o_tok_stride; p.o_hs = o_head_stride;
  p.start_loc = b_start_loc; p.start_w = start_loc_width;
  p.seq_len = b_seq_len; p.len_w = seq_len_width;
  p.Hq = num_q_heads; p.Hkv = num_kv_heads; p.G = num_q_heads / num_kv_heads;
  p.scale_log2 = sm_scale_log2;
  cudaStream_t st = (cudaStream_t)stream;
#define DISPATCH_D(TT)                                                            \
  switch (head_dim) {                                                             \
    case 32: return launch_prefill<TT, 32>(p, batch, max_seq_len, st);            \
    case 64: return launch_prefill<TT, 64>(p, batch, max_seq_len, st);            \
    case 128: return launch_prefill<TT, 128>(p, batch, max_seq_len, st);          \
    default: B2L_UNSUPPORTED("flash_attention2_no_pad: head_dim %d (supported: 32, 64, 128)", head_dim); \
  }
  if (dtype == B2L_F16) { DISPATCH_D(__half) }
  if (dtype == B2L_BF16) { DISPATCH_D(__nv_bfloat16) }
#undef DISPATCH_D
  B2L_UNSUPPORTED("flash_attention2_no_pad: dtype %d", dtype);
}
