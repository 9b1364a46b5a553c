// flash_decoding: single-token decode attention over the token-granular paged KV cache.
//
// Replaces lite_llama/kernels/flashdecoding.py:23-380 (stage-1 + stage-2 Triton kernels).
// Same contract: history of batch row b is the cache rows
//   table[b_req_idx[b], 0 .. b_seq_len[b]-1]            (flashdecoding.py:66-69)
// K/V are strided views of one [max_tokens, 2*Hkv, D] pool (models/base.py:119-123).
//
// B200 design (DESIGN.md "flash_decoding"):
//  * one CTA per (split, kv-head block, batch row) streams WHOLE TOKEN ROWS: in the fused pool a
//    token's K heads and V heads are one contiguous 2*Hkv*D*2-byte run (4 KiB for Llama-3-8B), so a
//    single 1-D bulk async copy (cp.async.bulk -> SASS UBLKCP, the TMA engine) per token brings in
//    everything every head needs.  (First version: one CTA per kv head with 256-byte copies reached
//    only 29 % of HBM peak -- the copy engine's request rate, not bytes, was the limit.)
//  * a dedicated producer warp walks the token table and issues the copies into a 3-4 stage
//    shared-memory ring guarded by mbarriers; consumer warps never touch the LSU for K/V;
//  * one consumer warp per (kv head, block of 16 query heads): the GQA group is stacked on the M
//    dimension of the MMA tile, so each cached element is read from HBM exactly once (the reference
//    launches one program per q head and re-reads it G times);
//  * QK^T and PV on tensor cores with fp32 accumulation (fp16 products are exact in fp32, which
//    keeps the reference's "fp32 product" precision, flashdecoding.py:115-120), fp32 online softmax;
//  * split-KV sized from host-known quantities only (graph-safe); fp32 partials + log-sum-exp,
//    merged by a second small kernel (reference stage 2, flashdecoding.py:224-313).
#include "common.cuh"

namespace {

constexpr int kTT = 16;            // tokens per pipeline stage (one m16n8k16 k-step of PV)
constexpr int kMaxConsumers = 8;
constexpr int kMaxStages = 4;

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
  int* tickets;   // one per (batch row, kv-head block), zero on entry and on exit; null = separate combine kernel
  int Hq, Hkv, G, qblocks;
  int HB;         // kv heads per CTA (divides Hkv)
  int copy_mode;  // 0: one copy per token (K|V fused run), 1: one K + one V copy, 2: one copy per head
  int stages;
  int chunk;      // tokens per split (multiple of kTT)
  int splits;
  float scale_log2;
};

__host__ __device__ inline int row_pitch(int HB, int D) { return 2 * HB * D * 2 + 16; }  // +16: ldmatrix conflict-free

template <typename T, int D>
__global__ void __launch_bounds__((kMaxConsumers + 1) * 32) flash_decode_split_kernel(const DecodeParams p) {
  pdl_prologue();   // everything below reads what earlier kernels of the step wrote (q, KV rows, lengths, table)
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);   // [kMaxStages]
  uint64_t* empty_bar = full_bar + kMaxStages;              // [kMaxStages]
  uint8_t* stage_base = smem + 128;

  const int split = blockIdx.x;
  const int hblk = blockIdx.y;
  const int b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nconsumers = p.HB * p.qblocks;      // blockDim.x == (nconsumers + 1) * 32

  const int64_t seq_len = load_index(p.seq_len, b, p.len_w);
  const int64_t start = (int64_t)split * p.chunk;
  if (start >= seq_len) return;  // empty split: nothing to contribute (whole CTA exits together)
  const int64_t end = (start + p.chunk < seq_len) ? start + p.chunk : seq_len;
  const int ntiles = (int)((end - start + kTT - 1) / kTT);
  const int64_t req = load_index(p.req_idx, b, p.req_w);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], nconsumers);
    }
    mbar_fence_init();
  }
  __syncthreads();

  const int PITCH = row_pitch(p.HB, D);
  const int stage_bytes = kTT * PITCH;
  constexpr uint32_t HEAD_BYTES = D * 2;

  if (warp == nconsumers) {
    // ===================== producer warp: gather token rows through the table =====================
    const int32_t* trow = p.table + req * p.tab_bs;
    const int h0 = hblk * p.HB;
    const char* kbase = reinterpret_cast<const char*>(p.k) + (int64_t)h0 * p.k_hs * 2;
    const char* vbase = reinterpret_cast<const char*>(p.v) + (int64_t)h0 * p.v_hs * 2;
    const int tok = lane & 15;          // token of the tile this lane serves
    const int half = lane >> 4;         // lanes 16..31 take the V side in modes 1/2
    const uint32_t side_bytes = p.HB * HEAD_BYTES;
    int64_t pos = start + tok;
    pos = pos < end ? pos : end - 1;    // clamp: padded rows duplicate a valid row (finite data, masked later)
    int64_t row_next = trow[pos * p.tab_ss];
    int s = 0; uint32_t phase = 0;
    for (int t = 0; t < ntiles; ++t) {
      const int64_t row = row_next;
      if (t + 1 < ntiles) {
        int64_t pn = start + (int64_t)(t + 1) * kTT + tok;
        pn = pn < end ? pn : end - 1;
        row_next = trow[pn * p.tab_ss];
      }
      mbar_wait(&empty_bar[s], phase ^ 1);
      if (lane == 0) mbar_arrive_expect_tx(&full_bar[s], (uint32_t)kTT * 2u * side_bytes);
      __syncwarp();
      uint8_t* dst = stage_base + s * stage_bytes + tok * PITCH;
      if (p.copy_mode == 0) {
        if (half == 0) bulk_g2s(dst, kbase + row * p.k_ts * 2, 2u * side_bytes, &full_bar[s]);
      } else if (p.copy_mode == 1) {
        if (half == 0) bulk_g2s(dst, kbase + row * p.k_ts * 2, side_bytes, &full_bar[s]);
        else bulk_g2s(dst + side_bytes, vbase + row * p.v_ts * 2, side_bytes, &full_bar[s]);
      } else {
        const char* src = half == 0 ? kbase + row * p.k_ts * 2 : vbase + row * p.v_ts * 2;
        const int64_t hs = (half == 0 ? p.k_hs : p.v_hs) * 2;
        for (int h = 0; h < p.HB; ++h)
          bulk_g2s(dst + half * side_bytes + h * HEAD_BYTES, src + h * hs, HEAD_BYTES, &full_bar[s]);
      }
      if (++s == p.stages) { s = 0; phase ^= 1; }
    }
  } else {
    // ===================== consumer warps: one (kv head, 16-query-head block) each =====================
    const int hl = warp / p.qblocks;            // kv head inside this CTA's block
    const int qblk = warp % p.qblocks;
    const int kvh = hblk * p.HB + hl;
    const int g = lane >> 2, tq = lane & 3;
    // Q fragments (A operand, m16 x k16 per step); rows = query heads of this KV group.
    uint32_t qf[D / 16][4];
    const int hq0 = qblk * 16 + g, hq1 = hq0 + 8;
    const bool v0 = hq0 < p.G, v1 = hq1 < p.G;
    {
      const T* qb = reinterpret_cast<const T*>(p.q) + (int64_t)b * p.q_bs;
      const T* q0 = qb + (int64_t)(kvh * p.G + (v0 ? hq0 : 0)) * p.q_hs;
      const T* q1 = qb + (int64_t)(kvh * p.G + (v1 ? hq1 : 0)) * p.q_hs;
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

    const uint32_t k_off = hl * HEAD_BYTES;
    const uint32_t v_off = (p.HB + hl) * HEAD_BYTES;
    const int mj = lane >> 3, r8 = lane & 7;
    // ldmatrix lane addresses (relative to the stage base)
    const uint32_t ka = ((mj >> 1) * 8 + r8) * PITCH + k_off + (mj & 1) * 16;
    const uint32_t va = ((mj & 1) * 8 + r8) * PITCH + v_off + (mj >> 1) * 16;

    int s = 0; uint32_t phase = 0;
    for (int t = 0; t < ntiles; ++t) {
      mbar_wait(&full_bar[s], phase);
      const uint32_t sb = smem_u32(stage_base + s * stage_bytes);

      // ---- S = Q K^T  (16 heads x 16 tokens) ----
      float sc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
      for (int kk = 0; kk < D / 16; ++kk) {
        uint32_t bf[4];
        ldmatrix_x4(bf, sb + ka + kk * 32);
        mma_m16n8k16<T>(sc[0], qf[kk], bf[0], bf[1]);
        mma_m16n8k16<T>(sc[1], qf[kk], bf[2], bf[3]);
      }
      // ---- scale, mask, online softmax (log2 domain) ----
      const int64_t tbase = start + (int64_t)t * kTT;
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
      float alpha[2], base[2];
#pragma unroll
      for (int r2 = 0; r2 < 2; ++r2) {
        mx[r2] = fmaxf(mx[r2], __shfl_xor_sync(0xffffffffu, mx[r2], 1));
        mx[r2] = fmaxf(mx[r2], __shfl_xor_sync(0xffffffffu, mx[r2], 2));
        const float mnew = fmaxf(m_run[r2], mx[r2]);
        base[r2] = mnew == -INFINITY ? 0.f : mnew;   // avoid (-inf) - (-inf)
        alpha[r2] = fast_exp2(m_run[r2] - base[r2]);
        m_run[r2] = mnew;
      }
      float ps[2] = {0.f, 0.f};
#pragma unroll
      for (int j = 0; j < 2; ++j) {
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
      // ---- O += P V  (P: 16 heads x 16 tokens as the A operand) ----
      uint32_t pf[4];
      pf[0] = pack2<T>(sc[0][0], sc[0][1]);
      pf[1] = pack2<T>(sc[0][2], sc[0][3]);
      pf[2] = pack2<T>(sc[1][0], sc[1][1]);
      pf[3] = pack2<T>(sc[1][2], sc[1][3]);
#pragma unroll
      for (int dn = 0; dn < D / 16; ++dn) {
        uint32_t bf[4];
        ldmatrix_x4_trans(bf, sb + va + dn * 32);
        mma_m16n8k16<T>(o[2 * dn], pf, bf[0], bf[1]);
        mma_m16n8k16<T>(o[2 * dn + 1], pf, bf[2], bf[3]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);
      if (++s == p.stages) { s = 0; phase ^= 1; }
    }

    // quad-reduce the row sums (each lane only summed its own 4 columns per tile)
#pragma unroll
    for (int r2 = 0; r2 < 2; ++r2) {
      l_run[r2] += __shfl_xor_sync(0xffffffffu, l_run[r2], 1);
      l_run[r2] += __shfl_xor_sync(0xffffffffu, l_run[r2], 2);
    }
    // this warp saw every token of the split for its heads: write the result rows directly
#pragma unroll
    for (int r2 = 0; r2 < 2; ++r2) {
      const bool ok = r2 == 0 ? v0 : v1;
      if (!ok) continue;
      const int head = kvh * p.G + (r2 == 0 ? hq0 : hq1);
      const float inv = 1.0f / l_run[r2];
      if (p.splits == 1) {
        T* orow = reinterpret_cast<T*>(p.out) + (int64_t)b * p.o_bs + (int64_t)head * p.o_hs;
#pragma unroll
        for (int i = 0; i < D / 8; ++i)
          *reinterpret_cast<uint32_t*>(orow + i * 8 + 2 * tq) = pack2<T>(o[i][2 * r2] * inv, o[i][2 * r2 + 1] * inv);
      } else {
        const int64_t off = ((int64_t)b * p.Hq + head) * p.splits + split;
        float* orow = p.ws_o + off * D;
#pragma unroll
        for (int i = 0; i < D / 8; ++i)
          *reinterpret_cast<float2*>(orow + i * 8 + 2 * tq) = make_float2(o[i][2 * r2] * inv, o[i][2 * r2 + 1] * inv);
        if (tq == 0) p.ws_lse[off] = m_run[r2] + log2f(l_run[r2]);
      }
    }
  }

  // ---- stage 2 folded in (reference: a second kernel, flashdecoding.py:224-313): the CTA that finishes a
  // (batch row, kv-head block) LAST merges its splits.  Deterministic: the merge always runs over all splits in index
  // order, whoever executes it.  The ticket resets itself, so graph replays start clean. ----
  if (p.splits > 1 && p.tickets != nullptr) {
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    int nsp = (int)((seq_len + p.chunk - 1) / p.chunk);          // splits that have tokens (the others returned above)
    nsp = nsp < p.splits ? nsp : p.splits;
    if (threadIdx.x == 0) {
      int* t = p.tickets + (int64_t)b * gridDim.y + hblk;
      const int old = atomicAdd(t, 1);
      s_last = (old == nsp - 1);
      if (s_last) *t = 0;
    }
    __syncthreads();
    if (s_last) {
      __threadfence();
      const int heads_here = p.HB * p.G;
      const int head0 = hblk * p.HB * p.G;
      for (int idx = threadIdx.x; idx < heads_here * D; idx += blockDim.x) {
        const int head = head0 + idx / D, d = idx % D;
        const int64_t off = ((int64_t)b * p.Hq + head) * p.splits;
        float M = -INFINITY;
        for (int sp = 0; sp < nsp; ++sp) M = fmaxf(M, __ldcg(p.ws_lse + off + sp));
        float acc = 0.f, L = 0.f;
        for (int sp = 0; sp < nsp; ++sp) {
          const float w = fast_exp2(__ldcg(p.ws_lse + off + sp) - M);
          acc += w * __ldcg(p.ws_o + (off + sp) * D + d);
          L += w;
        }
        reinterpret_cast<T*>(p.out)[(int64_t)b * p.o_bs + (int64_t)head * p.o_hs + d] = from_f32<T>(acc / L);
      }
    }
  }
}

// Stage 2: log-sum-exp merge of the per-split partials (reference: flashdecoding.py:224-313).
template <typename T, int D>
__global__ void flash_decode_combine_kernel(const DecodeParams p) {
  pdl_prologue();
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

struct Plan { int splits, chunk, qblocks, HB, stages, smem; };

// Pure function of host-known sizes => identical inside and outside graph capture.
Plan make_plan(int64_t batch, int Hq, int Hkv, int D, int64_t max_len) {
  Plan pl;
  const int G = Hq / Hkv;
  pl.qblocks = (G + 15) / 16;
  int hb_cap = kMaxConsumers / pl.qblocks;
  if (hb_cap < 1) hb_cap = 1;
  // a 3-stage ring must fit the opt-in shared memory
  const int smem_cap = b2l_max_smem_optin() - 256;
  while (hb_cap > 1 && 3 * kTT * row_pitch(hb_cap, D) > smem_cap) --hb_cap;
  int HB = 1;
  for (int h = 1; h <= hb_cap; ++h)
    if (Hkv % h == 0) HB = h;
  pl.HB = HB;
  const int stage = kTT * row_pitch(HB, D);
  int stages = kMaxStages;
  while (stages > 2 && stages * stage > smem_cap) --stages;
  // prefer two resident CTAs per SM when a 3-deep ring fits in half the shared memory
  if (stages * stage > smem_cap / 2 && 3 * stage <= smem_cap / 2) stages = 3;
  pl.stages = stages;
  pl.smem = stages * stage + 128;
  const int64_t pairs = batch * (Hkv / HB);
  const int64_t tiles = max_len > 0 ? (max_len + kTT - 1) / kTT : 1;
  const int64_t target = 6LL * b2l_num_sms();
  int64_t splits = (target + pairs - 1) / pairs;
  const int64_t max_splits = (tiles + 3) / 4;   // at least 4 tiles (64 tokens) per split
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  const int64_t tiles_per = (tiles + splits - 1) / splits;
  pl.splits = (int)((tiles + tiles_per - 1) / tiles_per);
  pl.chunk = (int)(tiles_per * kTT);
  return pl;
}

template <typename T, int D>
int launch_decode(const DecodeParams& p, const Plan& pl, int64_t batch, cudaStream_t st) {
  static int configured[64] = {0};
  B2L_CUDA_CHECK(b2l_ensure_dynamic_smem(flash_decode_split_kernel<T, D>, pl.smem, configured));
  dim3 grid((unsigned)p.splits, (unsigned)(p.Hkv / p.HB), (unsigned)batch);
  const int threads = (p.HB * p.qblocks + 1) * 32;
  b2l_launch(flash_decode_split_kernel<T, D>, dim3(grid), dim3(threads), (size_t)(pl.smem), st, p);
  B2L_LAUNCH_CHECK();
  if (p.splits > 1 && p.tickets == nullptr) {
    dim3 g2((unsigned)batch, (unsigned)p.Hq);
    b2l_launch(flash_decode_combine_kernel<T, D>, dim3(g2), dim3(D >= 128 ? 128 : 64), (size_t)(0), st, p);
    B2L_LAUNCH_CHECK();
  }
  return B2L_OK;
}

}  // namespace

extern "C" int64_t b2l_flash_decoding_workspace_bytes(int64_t batch, int num_q_heads, int num_kv_heads, int head_dim,
                                                       int64_t max_seq_len) {
  if (batch <= 0 || num_q_heads <= 0 || num_kv_heads <= 0 || num_q_heads % num_kv_heads != 0) return 0;
  Plan pl = make_plan(batch, num_q_heads, num_kv_heads, head_dim, max_seq_len);
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
                                  int64_t o_head_stride, void* workspace, int64_t workspace_bytes, void* tickets,
                                  int64_t num_tickets, int dtype, void* stream) {
  B2L_REQUIRE(out && q && k_cache && v_cache && b_req_tokens_table && b_req_idx && b_seq_len,
              "flash_decoding: null pointer");
  B2L_REQUIRE(num_q_heads > 0 && num_kv_heads > 0 && num_q_heads % num_kv_heads == 0,
              "flash_decoding: num_q_heads (%d) must be a multiple of num_kv_heads (%d)", num_q_heads, num_kv_heads);
  B2L_REQUIRE((req_idx_width == 4 || req_idx_width == 8) && (seq_len_width == 4 || seq_len_width == 8),
              "flash_decoding: index widths must be 4 or 8");
  B2L_REQUIRE(k_token_stride % 8 == 0 && v_token_stride % 8 == 0 && k_head_stride % 8 == 0 && v_head_stride % 8 == 0 &&
                  (uintptr_t)k_cache % 16 == 0 && (uintptr_t)v_cache % 16 == 0,
              "flash_decoding: K/V rows must be 16-byte aligned");
  B2L_REQUIRE(q_head_stride % 2 == 0 && q_batch_stride % 2 == 0 && (uintptr_t)q % 4 == 0 && o_head_stride % 2 == 0 &&
                  o_batch_stride % 2 == 0 && (uintptr_t)out % 4 == 0,
              "flash_decoding: q/out rows must be 4-byte aligned");
  B2L_REQUIRE(head_dim == 32 || head_dim == 64 || head_dim == 128 || head_dim == 256,
              "flash_decoding: head_dim %d (supported: 32, 64, 128, 256)", head_dim);
  if (batch == 0) return B2L_OK;
  Plan pl = make_plan(batch, num_q_heads, num_kv_heads, head_dim, max_seq_len);
  DecodeParams p;
  p.q = q; p.q_bs = q_batch_stride; p.q_hs = q_head_stride;
  p.k = k_cache; p.k_ts = k_token_stride; p.k_hs = k_head_stride;
  p.v = v_cache; p.v_ts = v_token_stride; p.v_hs = v_head_stride;
  p.table = (const int32_t*)b_req_tokens_table; p.tab_bs = table_batch_stride; p.tab_ss = table_seq_stride;
  p.req_idx = b_req_idx; p.req_w = req_idx_width;
  p.seq_len = b_seq_len; p.len_w = seq_len_width;
  p.out = out; p.o_bs = o_batch_stride; p.o_hs = o_head_stride;
  p.Hq = num_q_heads; p.Hkv = num_kv_heads; p.G = num_q_heads / num_kv_heads; p.qblocks = pl.qblocks;
  p.HB = pl.HB; p.stages = pl.stages; p.chunk = pl.chunk; p.splits = pl.splits;
  p.scale_log2 = qk_scale * 1.4426950408889634f;
  // how contiguous is one token's K|V for this CTA's head block?
  const bool heads_packed = k_head_stride == head_dim && v_head_stride == head_dim;
  const bool fused_run = heads_packed && pl.HB == num_kv_heads && k_token_stride == v_token_stride &&
                         (const char*)v_cache == (const char*)k_cache + (int64_t)num_kv_heads * head_dim * 2;
  p.copy_mode = fused_run ? 0 : (heads_packed ? 1 : 2);
  p.ws_o = nullptr; p.ws_lse = nullptr;
  // the split CTAs merge their own results when the caller provides zeroed tickets (one per batch row x head block)
  p.tickets = (tickets && num_tickets >= batch * (num_kv_heads / pl.HB)) ? (int*)tickets : nullptr;
  if (pl.splits > 1) {
    const int64_t need = b2l_flash_decoding_workspace_bytes(batch, num_q_heads, num_kv_heads, head_dim, max_seq_len);
    B2L_REQUIRE(workspace && workspace_bytes >= need, "flash_decoding: workspace too small (%lld < %lld)",
                (long long)workspace_bytes, (long long)need);
    B2L_REQUIRE((uintptr_t)workspace % 16 == 0, "flash_decoding: workspace must be 16-byte aligned");
    p.ws_o = (float*)workspace;
    p.ws_lse = p.ws_o + (int64_t)batch * num_q_heads * pl.splits * head_dim;
  }
  cudaStream_t st = (cudaStream_t)stream;
#define DISPATCH_D(TT)                                                    \
  switch (head_dim) {                                                     \
    case 32: return launch_decode<TT, 32>(p, pl, batch, st);              \
    case 64: return launch_decode<TT, 64>(p, pl, batch, st);              \
    case 128: return launch_decode<TT, 128>(p, pl, batch, st);            \
    default: return launch_decode<TT, 256>(p, pl, batch, st);             \
  }
  if (dtype == B2L_F16) { DISPATCH_D(__half) }
  if (dtype == B2L_BF16) { DISPATCH_D(__nv_bfloat16) }
#undef DISPATCH_D
  B2L_UNSUPPORTED("flash_decoding: dtype %d", dtype);
}
