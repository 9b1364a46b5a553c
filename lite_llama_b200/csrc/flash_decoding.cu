// flash_decoding: single-token decode attention over the token-granular paged KV cache.
//
// Replaces lite_llama/kernels/flashdecoding.py:23-380 (stage-1 + stage-2 Triton kernels).
// Same contract: history of batch row b is the cache rows
//   table[b_req_idx[b], 0 .. b_seq_len[b]-1]            (flashdecoding.py:66-69)
// K/V are strided views of one [max_tokens, 2*Hkv, D] pool (models/base.py:119-123).
//
// B200 design:
//  * one CTA per (split, kv-head, batch row): the whole GQA group (G = Hq/Hkv query
//    heads, padded to an m16 MMA tile) shares every K/V byte -> each cached element is
//    read from HBM exactly once (the reference re-reads it G times, one program per q head);
//  * rows are gathered with 1-D bulk async copies (cp.async.bulk -> SASS UBLKCP, the TMA
//    engine) issued by a dedicated producer warp, landing in a 3-stage shared-memory ring
//    guarded by mbarriers; consumers never touch the LSU for K/V;
//  * QK^T and PV run on tensor cores with fp32 accumulation (fp16/bf16 products are exact
//    in fp32, which keeps the reference's "fp32 product" precision, flashdecoding.py:115-120);
//  * split-KV sized from host-known quantities only (graph-safe), fp32 partials + log-sum-exp,
//    merged by a second tiny kernel.
#include "common.cuh"

namespace {

constexpr int kTile = 64;        // tokens per pipeline stage
constexpr int kStages = 3;
constexpr int kConsumerWarps = 4;  // each owns 16 token rows of every tile
constexpr int kThreads = (kConsumerWarps + 1) * 32;

struct DecodeParams {
  const void* q; int64_t q_bs, q_hs;
  const void* k; int64_t k_ts, k_hs;
  const void* v; int64_t v_ts, v_hs;
  const int32_t* table; int64_t tab_bs, tab_ss;
  const void* req_idx; int req_w;
  const void* seq_len; int len_w;
  void* out; int64_t o_bs, o_hs;
  float* ws_o;    // [B, Hq, splits, D]
  float* ws_lse;  // [B, Hq, splits]   (log2 domain)
  int Hq, Hkv, G, qblocks;
  int chunk;      // tokens per split (multiple of kTile)
  int splits;
  float scale_log2;  // qk_scale * log2(e)
};

template <int D> __host__ __device__ constexpr int row_pitch() { return D * 2 + 16; }  // bytes; +16 keeps ldmatrix conflict-free
template <int D> __host__ __device__ constexpr int tile_bytes() { return kTile * row_pitch<D>(); }
template <int D> __host__ __device__ constexpr int stage_bytes() { return 2 * tile_bytes<D>(); }
template <int D> __host__ __device__ constexpr int merge_bytes() { return kConsumerWarps * (16 * (D + 8) + 32) * 4; }
template <int D> __host__ __device__ constexpr int smem_bytes() {
  return (kStages * stage_bytes<D>() > merge_bytes<D>() ? kStages * stage_bytes<D>() : merge_bytes<D>()) + 128;
}

template <typename T, int D>
__global__ void __launch_bounds__(kThreads) flash_decode_split_kernel(const DecodeParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);            // [kStages]
  uint64_t* empty_bar = full_bar + kStages;                          // [kStages]
  uint8_t* stage_base = smem + 128;

  const int split = blockIdx.x;
  const int kvh = blockIdx.y / p.qblocks;
  const int qblk = blockIdx.y % p.qblocks;
  const int b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int64_t seq_len = load_index(p.seq_len, b, p.len_w);
  const int64_t start = (int64_t)split * p.chunk;
  if (start >= seq_len) return;  // empty split: nothing to contribute (whole CTA exits together)
  const int64_t end = (start + p.chunk < seq_len) ? start + p.chunk : seq_len;
  const int ntiles = (int)((end - start + kTile - 1) / kTile);
  const int64_t req = load_index(p.req_idx, b, p.req_w);

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
    // ===================== producer warp: gather rows through the table =====================
    const int32_t* trow = p.table + req * p.tab_bs;
    const char* kbase = reinterpret_cast<const char*>(p.k) + (int64_t)kvh * p.k_hs * 2;
    const char* vbase = reinterpret_cast<const char*>(p.v) + (int64_t)kvh * p.v_hs * 2;
    int64_t rows_next[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int64_t pos = start + lane + 32 * i;
      pos = pos < end ? pos : end - 1;
      rows_next[i] = trow[pos * p.tab_ss];
    }
    for (int t = 0; t < ntiles; ++t) {
      const int s = t % kStages;
      const uint32_t par = ((t / kStages) & 1) ^ 1;
      int64_t rows[2] = {rows_next[0], rows_next[1]};
      if (t + 1 < ntiles) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          int64_t pos = start + (int64_t)(t + 1) * kTile + lane + 32 * i;
          pos = pos < end ? pos : end - 1;  // clamp: padded rows duplicate a valid row (finite data)
          rows_next[i] = trow[pos * p.tab_ss];
        }
      }
      mbar_wait(&empty_bar[s], par);
      if (lane == 0) mbar_arrive_expect_tx(&full_bar[s], 2u * kTile * ROW_BYTES);
      __syncwarp();
      uint8_t* ks = stage_base + s * stage_bytes<D>();
      uint8_t* vs = ks + tile_bytes<D>();
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int r = lane + 32 * i;
        bulk_g2s(ks + r * PITCH, kbase + rows[i] * p.k_ts * 2, ROW_BYTES, &full_bar[s]);
        bulk_g2s(vs + r * PITCH, vbase + rows[i] * p.v_ts * 2, ROW_BYTES, &full_bar[s]);
      }
    }
  } else {
    // ===================== consumer warps =====================
    const int g = lane >> 2, tq = lane & 3;
    // Q fragments (A operand, m16 x k16 per step); rows = query heads of this KV group.
    uint32_t qf[D / 16][4];
    {
      const int h0 = qblk * 16 + g, h1 = h0 + 8;
      const T* qb = reinterpret_cast<const T*>(p.q) + (int64_t)b * p.q_bs;
      const T* q0 = qb + (int64_t)(kvh * p.G + h0) * p.q_hs;
      const T* q1 = qb + (int64_t)(kvh * p.G + h1) * p.q_hs;
      const bool v0 = h0 < p.G, v1 = h1 < p.G;
#pragma unroll
      for (int kk = 0; kk < D / 16; ++kk) {
        const int c = kk * 16 + 2 * tq;
        qf[kk][0] = v0 ? *reinterpret_cast<const uint32_t*>(q0 + c) : 0u;
        qf[kk][1] = v1 ? *reinterpret_cast<const uint32_t*>(q1 + c) : 0u;
        qf[kk][2] = v0 ? *reinterpret_cast<const uint32_t*>(q0 + c + 8) : 0u;
        qf[kk][3] = v1 ? *reinterpret_cast<const uint32_t*>(q1 + c + 8) : 0u;
      }
    }
    float o[D / 8][4];
#pragma unroll
    for (int i = 0; i < D / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m_run[2] = {-INFINITY, -INFINITY};
    float l_run[2] = {0.f, 0.f};

    for (int t = 0; t < ntiles; ++t) {
      const int s = t % kStages;
      mbar_wait(&full_bar[s], (t / kStages) & 1);
      const uint32_t ks = smem_u32(stage_base + s * stage_bytes<D>());
      const uint32_t vs = ks + tile_bytes<D>();
      const int tok0 = warp * 16;  // this warp's rows inside the tile

      // ---- S = Q K^T  (16 heads x 16 tokens) ----
      float sc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      {
        const int mj = lane >> 3, r = lane & 7;
        const uint32_t a0 = ks + (tok0 + (mj >> 1) * 8 + r) * PITCH + (mj & 1) * 16;
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          uint32_t bf[4];
          ldmatrix_x4(bf, a0 + kk * 32);
          mma_m16n8k16<T>(sc[0], qf[kk], bf[0], bf[1]);
          mma_m16n8k16<T>(sc[1], qf[kk], bf[2], bf[3]);
        }
      }
      // ---- scale, mask, online softmax (log2 domain) ----
      const int64_t tbase = start + (int64_t)t * kTile + tok0;
      float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int j = 0; j < 2; ++j) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int64_t pos = tbase + j * 8 + 2 * tq + (e & 1);
          float v = sc[j][e] * p.scale_log2;
          v = pos < end ? v : -INFINITY;
          sc[j][e] = v;
          mx[e >> 1] = fmaxf(mx[e >> 1], v);
        }
      }
      float alpha[2], mnew[2];
#pragma unroll
      for (int r2 = 0; r2 < 2; ++r2) {
        mx[r2] = fmaxf(mx[r2], __shfl_xor_sync(0xffffffffu, mx[r2], 1));
        mx[r2] = fmaxf(mx[r2], __shfl_xor_sync(0xffffffffu, mx[r2], 2));
        mnew[r2] = fmaxf(m_run[r2], mx[r2]);
        // all-masked rows keep m = -inf; use 0 as the subtraction base to avoid (-inf)-(-inf)
        const float base = mnew[r2] == -INFINITY ? 0.f : mnew[r2];
        alpha[r2] = fast_exp2(m_run[r2] - base);
        m_run[r2] = mnew[r2];
        mnew[r2] = base;
      }
      float ps[2] = {0.f, 0.f};
#pragma unroll
      for (int j = 0; j < 2; ++j) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float pv = fast_exp2(sc[j][e] - mnew[e >> 1]);
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
      // ---- O += P V  (P: 16 heads x 16 tokens as the A operand) ----
      uint32_t pf[4];
      pf[0] = pack2<T>(sc[0][0], sc[0][1]);
      pf[1] = pack2<T>(sc[0][2], sc[0][3]);
      pf[2] = pack2<T>(sc[1][0], sc[1][1]);
      pf[3] = pack2<T>(sc[1][2], sc[1][3]);
      {
        const int mj = lane >> 3, r = lane & 7;
        const uint32_t a0 = vs + (tok0 + (mj & 1) * 8 + r) * PITCH + (mj >> 1) * 16;
#pragma unroll
        for (int dn = 0; dn < D / 16; ++dn) {
          uint32_t bf[4];
          ldmatrix_x4_trans(bf, a0 + dn * 32);
          mma_m16n8k16<T>(o[2 * dn], pf, bf[0], bf[1]);
          mma_m16n8k16<T>(o[2 * dn + 1], pf, bf[2], bf[3]);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);
    }

    // quad-reduce the row sums (each lane only summed its own 4 columns per tile)
#pragma unroll
    for (int r2 = 0; r2 < 2; ++r2) {
      l_run[r2] += __shfl_xor_sync(0xffffffffu, l_run[r2], 1);
      l_run[r2] += __shfl_xor_sync(0xffffffffu, l_run[r2], 2);
    }
    // All bulk copies of this CTA have landed and been consumed once every consumer warp
    // leaves its loop; the barrier below orders that before the ring is reused as scratch.
    asm volatile("bar.sync 1, %0;" ::"n"(kConsumerWarps * 32));

    float* mo = reinterpret_cast<float*>(stage_base) + warp * (16 * (D + 8) + 32);
    float* mm = mo + 16 * (D + 8);
    if (tq == 0) {
      mm[g] = m_run[0]; mm[g + 8] = m_run[1];
      mm[16 + g] = l_run[0]; mm[16 + g + 8] = l_run[1];
    }
#pragma unroll
    for (int i = 0; i < D / 8; ++i) {
      const int c = i * 8 + 2 * tq;
      *reinterpret_cast<float2*>(mo + g * (D + 8) + c) = make_float2(o[i][0], o[i][1]);
      *reinterpret_cast<float2*>(mo + (g + 8) * (D + 8) + c) = make_float2(o[i][2], o[i][3]);
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kConsumerWarps * 32));

    // ---- merge the 4 per-warp partial states, write final or per-split partial ----
    const float* base = reinterpret_cast<const float*>(stage_base);
    constexpr int WS = 16 * (D + 8) + 32;
    const int rows_valid = (p.G - qblk * 16) < 16 ? (p.G - qblk * 16) : 16;
    for (int idx = threadIdx.x; idx < rows_valid * D; idx += kConsumerWarps * 32) {
      const int r = idx / D, d = idx % D;
      float M = -INFINITY;
#pragma unroll
      for (int w = 0; w < kConsumerWarps; ++w) M = fmaxf(M, base[w * WS + 16 * (D + 8) + r]);
      float acc = 0.f, L = 0.f;
#pragma unroll
      for (int w = 0; w < kConsumerWarps; ++w) {
        const float mw = base[w * WS + 16 * (D + 8) + r];
        const float f = mw == -INFINITY ? 0.f : fast_exp2(mw - M);
        acc += f * base[w * WS + r * (D + 8) + d];
        L += f * base[w * WS + 16 * (D + 8) + 16 + r];
      }
      const int head = kvh * p.G + qblk * 16 + r;
      const float val = acc / L;
      if (p.splits == 1) {
        reinterpret_cast<T*>(p.out)[(int64_t)b * p.o_bs + (int64_t)head * p.o_hs + d] = from_f32<T>(val);
      } else {
        const int64_t off = ((int64_t)b * p.Hq + head) * p.splits + split;
        p.ws_o[off * D + d] = val;
        if (d == 0) p.ws_lse[off] = M + log2f(L);
      }
    }
  }
}

// Stage 2: log-sum-exp merge of the per-split partials (reference: flashdecoding.py:224-313).
template <typename T, int D>
__global__ void flash_decode_combine_kernel(const DecodeParams p) {
  const int b = blockIdx.x, head = blockIdx.y;
  const int64_t seq_len = load_index(p.seq_len, b, p.len_w);
  int nsp = (int)((seq_len + p.chunk - 1) / p.chunk);
  nsp = nsp < p.splits ? nsp : p.splits;
  const int64_t off = ((int64_t)b * p.Hq + head) * p.splits;
  float M = -INFINITY;
  for (int s = 0; s < nsp; ++s) M = fmaxf(M, p.ws_lse[off + s]);
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float acc = 0.f, L = 0.f;
    for (int s = 0; s < nsp; ++s) {
      const float w = fast_exp2(p.ws_lse[off + s] - M);
      acc += w * p.ws_o[(off + s) * D + d];
      L += w;
    }
    const float val = nsp > 0 ? acc / L : 0.f;
    reinterpret_cast<T*>(p.out)[(int64_t)b * p.o_bs + (int64_t)head * p.o_hs + d] = from_f32<T>(val);
  }
}

struct SplitPlan { int splits; int chunk; int qblocks; };

// Pure function of host-known sizes => identical inside and outside graph capture.
SplitPlan plan_splits(int64_t batch, int Hq, int Hkv, int64_t max_len) {
  SplitPlan pl;
  const int G = Hq / Hkv;
  pl.qblocks = (G + 15) / 16;
  const int64_t pairs = batch * Hkv * pl.qblocks;
  const int64_t max_tiles = max_len > 0 ? (max_len + kTile - 1) / kTile : 1;
  const int64_t target = 4LL * b2l_num_sms();  // ~2 resident CTAs/SM x 2 waves
  int64_t splits = (target + pairs - 1) / pairs;
  if (splits > max_tiles) splits = max_tiles;
  if (splits < 1) splits = 1;
  const int64_t tiles_per = (max_tiles + splits - 1) / splits;
  pl.splits = (int)((max_tiles + tiles_per - 1) / tiles_per);
  pl.chunk = (int)(tiles_per * kTile);
  return pl;
}

template <typename T, int D>
int launch_decode(const DecodeParams& p, int64_t batch, cudaStream_t st) {
  constexpr int smem = smem_bytes<D>();
  static bool attr_set = false;
  if (!attr_set) {
    B2L_CUDA_CHECK(cudaFuncSetAttribute(flash_decode_split_kernel<T, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  dim3 grid((unsigned)p.splits, (unsigned)(p.Hkv * p.qblocks), (unsigned)batch);
  flash_decode_split_kernel<T, D><<<grid, kThreads, smem, st>>>(p);
  B2L_LAUNCH_CHECK();
  if (p.splits > 1) {
    dim3 g2((unsigned)batch, (unsigned)p.Hq);
    flash_decode_combine_kernel<T, D><<<g2, D >= 128 ? 128 : 64, 0, st>>>(p);
    B2L_LAUNCH_CHECK();
  }
  return B2L_OK;
}

}  // namespace

extern "C" int64_t b2l_flash_decoding_workspace_bytes(int64_t batch, int num_q_heads, int num_kv_heads, int head_dim,
                                                       int64_t max_seq_len) {
  if (batch <= 0 || num_q_heads <= 0 || num_kv_heads <= 0 || num_q_heads % num_kv_heads != 0) return 0;
  SplitPlan pl = plan_splits(batch, num_q_heads, num_kv_heads, max_seq_len);
  if (pl.splits <= 1) return 16;
  return (int64_t)batch * num_q_heads * pl.splits * (head_dim + 1) * 4 + 16;
}

extern "C" int b2l_flash_decoding(void* out, const void* q, const void* k_cache, const void* v_cache,
                                  const void* b_req_tokens_table, const void* b_req_idx, int req_idx_width,
                                  const void* b_seq_len, int seq_len_width, int64_t batch, int num_q_heads,
                                  int num_kv_heads, int head_dim, int64_t max_seq_len, float qk_scale,
                                  int64_t q_batch_stride, int64_t q_head_stride, int64_t k_token_stride,
                                  int64_t k_head_stride, int64_t v_token_stride, int64_t v_head_stride,
                                  int64_t table_batch_stride, int64_t table_seq_stride, int64_t o_batch_stride,
                                  int64_t o_head_stride, void* workspace, int64_t workspace_bytes, int dtype,
                                  void* stream) {
  B2L_REQUIRE(out && q && k_cache && v_cache && b_req_tokens_table && b_req_idx && b_seq_len,
              "flash_decoding: null pointer");
  B2L_REQUIRE(num_q_heads > 0 && num_kv_heads > 0 && num_q_heads % num_kv_heads == 0,
              "flash_decoding: num_q_heads (%d) must be a multiple of num_kv_heads (%d)", num_q_heads, num_kv_heads);
  B2L_REQUIRE((req_idx_width == 4 || req_idx_width == 8) && (seq_len_width == 4 || seq_len_width == 8),
              "flash_decoding: index widths must be 4 or 8");
  B2L_REQUIRE(k_token_stride % 8 == 0 && v_token_stride % 8 == 0 && k_head_stride % 8 == 0 && v_head_stride % 8 == 0 &&
                  (uintptr_t)k_cache % 16 == 0 && (uintptr_t)v_cache % 16 == 0,
              "flash_decoding: K/V rows must be 16-byte aligned");
  B2L_REQUIRE(q_head_stride % 2 == 0 && q_batch_stride % 2 == 0 && (uintptr_t)q % 4 == 0,
              "flash_decoding: q rows must be 4-byte aligned");
  if (batch == 0) return B2L_OK;
  SplitPlan pl = plan_splits(batch, num_q_heads, num_kv_heads, max_seq_len);
  DecodeParams p;
  p.q = q; p.q_bs = q_batch_stride; p.q_hs = q_head_stride;
  p.k = k_cache; p.k_ts = k_token_stride; p.k_hs = k_head_stride;
  p.v = v_cache; p.v_ts = v_token_stride; p.v_hs = v_head_stride;
  p.table = (const int32_t*)b_req_tokens_table; p.tab_bs = table_batch_stride; p.tab_ss = table_seq_stride;
  p.req_idx = b_req_idx; p.req_w = req_idx_width;
  p.seq_len = b_seq_len; p.len_w = seq_len_width;
  p.out = out; p.o_bs = o_batch_stride; p.o_hs = o_head_stride;
  p.Hq = num_q_heads; p.Hkv = num_kv_heads; p.G = num_q_heads / num_kv_heads; p.qblocks = pl.qblocks;
  p.chunk = pl.chunk; p.splits = pl.splits;
  p.scale_log2 = qk_scale * 1.4426950408889634f;
  p.ws_o = nullptr; p.ws_lse = nullptr;
  if (pl.splits > 1) {
    const int64_t need = b2l_flash_decoding_workspace_bytes(batch, num_q_heads, num_kv_heads, head_dim, max_seq_len);
    B2L_REQUIRE(workspace && workspace_bytes >= need, "flash_decoding: workspace too small (%lld < %lld)",
                (long long)workspace_bytes, (long long)need);
    B2L_REQUIRE((uintptr_t)workspace % 16 == 0, "flash_decoding: workspace must be 16-byte aligned");
    p.ws_o = (float*)workspace;
    p.ws_lse = p.ws_o + (int64_t)batch * num_q_heads * pl.splits * head_dim;
  }
  cudaStream_t st = (cudaStream_t)stream;
#define DISPATCH_D(TT)                                                         \
  switch (head_dim) {                                                          \
    case 32: return launch_decode<TT, 32>(p, batch, st);                       \
    case 64: return launch_decode<TT, 64>(p, batch, st);                       \
    case 128: return launch_decode<TT, 128>(p, batch, st);                     \
    case 256: return launch_decode<TT, 256>(p, batch, st);                     \
    default: B2L_UNSUPPORTED("flash_decoding: head_dim %d (supported: 32, 64, 128, 256)", head_dim); \
  }
  if (dtype == B2L_F16) { DISPATCH_D(__half) }
  if (dtype == B2L_BF16) { DISPATCH_D(__nv_bfloat16) }
#undef DISPATCH_D
  B2L_UNSUPPORTED("flash_decoding: dtype %d", dtype);
}
