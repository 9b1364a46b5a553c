// MoE routing glue as single kernels (the reference spends ~17 tiny torch launches per layer here):
//   * moe_align_block_size -- lite_llama/kernels/fused_moe.py:45-99 (vLLM's sorted-slot protocol)
//   * router top-k         -- lite_llama/models/qwen3_moe.py:85-100 (fp32 softmax over all experts,
//                             top-k, renormalise, cast)
// Both are integer/latency work: one CTA, shared-memory histogram + stable counting sort.
#include "common.cuh"

namespace {

constexpr int kAlignWarps = 16;
constexpr int kAlignThreads = kAlignWarps * 32;

// sorted_token_ids: [num_slots + E*(block-1)] filled with the sentinel num_slots, then every expert's
// slots (ascending slot id = stable order) at its block-padded start; expert_ids[b] = expert of block b
// (E-1 past the end, like the reference's searchsorted+clamp); num_post[0] = padded total.
__global__ void __launch_bounds__(kAlignThreads)
moe_align_kernel(const void* __restrict__ topk_ids, int id_width, int num_slots, int block_size, int num_experts,
                 int32_t* __restrict__ sorted_ids, int32_t* __restrict__ expert_ids, int32_t* __restrict__ num_post,
                 int max_padded, int max_blocks) {
  pdl_prologue();
  extern __shared__ int smem_i[];
  int* hist = smem_i;                                  // [kAlignWarps][E] per-warp chunk histograms
  int* starts = hist + kAlignWarps * num_experts;      // [E] padded start of each expert
  int* counts = starts + num_experts;                  // [E]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < kAlignWarps * num_experts; i += blockDim.x) hist[i] = 0;
  for (int i = threadIdx.x; i < max_padded; i += blockDim.x) sorted_ids[i] = num_slots;
  __syncthreads();

  // contiguous chunk of slots per warp (keeps the sort stable)
  const int chunk = ((num_slots + kAlignWarps - 1) / kAlignWarps + 31) / 32 * 32;
  const int c0 = warp * chunk;
  const int c1 = c0 + chunk < num_slots ? c0 + chunk : num_slots;
  for (int i = c0 + lane; i < c1; i += 32) {
    const int e = (int)load_index(topk_ids, i, id_width);
    if (e >= 0 && e < num_experts) atomicAdd(&hist[warp * num_experts + e], 1);   // ids outside [0, E) are dropped
  }
  __syncthreads();
  // per expert: total count, and turn hist[w][e] into the exclusive prefix over warps
  for (int e = threadIdx.x; e < num_experts; e += blockDim.x) {
    int run = 0;
    for (int w = 0; w < kAlignWarps; ++w) {
      const int c = hist[w * num_experts + e];
      hist[w * num_experts + e] = run;
      run += c;
    }
    counts[e] = run;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int pos = 0, blk = 0;
    for (int e = 0; e < num_experts; ++e) {
      starts[e] = pos;
      const int padded = (counts[e] + block_size - 1) / block_size * block_size;
      for (int b = 0; b < padded / block_size; ++b) expert_ids[blk++] = e;
      pos += padded;
    }
    for (; blk < max_blocks; ++blk) expert_ids[blk] = num_experts - 1;
    num_post[0] = pos;
  }
  __syncthreads();
  // stable placement: 32 slots at a time, rank inside the group by match_any
  for (int base = c0; base < c1; base += 32) {
    const int i = base + lane;
    bool valid = i < c1;
    int e = valid ? (int)load_index(topk_ids, i, id_width) : -1 - lane;         // unique dummy keys
    if (valid && (e < 0 || e >= num_experts)) { valid = false; e = -64 - lane; } // out-of-range id: no slot, no smem access
    const unsigned same = __match_any_sync(0xffffffffu, e);
    const int rank = __popc(same & ((1u << lane) - 1u));
    if (valid) {
      const int off = hist[warp * num_experts + e] + rank;
      sorted_ids[starts[e] + off] = i;
    }
    __syncwarp();
    // the highest lane of each key group advances the warp-private cursor
    if (valid && (same & ~((2u << lane) - 1u)) == 0) hist[warp * num_experts + e] += __popc(same);
    __syncwarp();
  }
}

// One warp per token: softmax over E logits (fp32), k rounds of arg-max (lowest index wins ties),
// optional renormalisation, weights cast to the activation dtype.
template <typename T, int MAXV>
__global__ void moe_route_kernel(const T* __restrict__ logits, int64_t row_stride, int tokens, int num_experts,
                                 int top_k, int renorm, T* __restrict__ out_w, int64_t* __restrict__ out_ids) {
  pdl_prologue();
  const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= tokens) return;
  const int lane = threadIdx.x & 31;
  const T* row = logits + (int64_t)t * row_stride;
  // E <= 32*MAXV values kept in registers
  float v[MAXV];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int e = lane + 32 * i;
    v[i] = e < num_experts ? to_f32(row[e]) : -INFINITY;
    mx = fmaxf(mx, v[i]);
  }
  mx = warp_reduce_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int e = lane + 32 * i;
    v[i] = e < num_experts ? expf(v[i] - mx) : 0.f;
    sum += v[i];
  }
  sum = warp_reduce_sum(sum);
  const float inv = 1.0f / sum;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) v[i] = (lane + 32 * i) < num_experts ? v[i] * inv : -1.f;
  float picked = 0.f;
  float my_w = 0.f;
  int my_id = 0;
  for (int k = 0; k < top_k; ++k) {
    float best = -2.f;
    int best_e = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int e = lane + 32 * i;
      if (v[i] > best) { best = v[i]; best_e = e; }   // ascending e inside a lane: first max = lowest index
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oe = __shfl_xor_sync(0xffffffffu, best_e, o);
      if (ob > best || (ob == best && oe < best_e)) { best = ob; best_e = oe; }
    }
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
      if (lane + 32 * i == best_e) v[i] = -1.f;          // remove the winner
    picked += best;
    if (lane == k) { my_w = best; my_id = best_e; }
  }
  if (lane < top_k) {
    const float w = renorm ? my_w / picked : my_w;
    out_w[(int64_t)t * top_k + lane] = from_f32<T>(w);
    out_ids[(int64_t)t * top_k + lane] = my_id;
  }
}

}  // namespace

extern "C" int b2l_moe_align_block_size(const void* topk_ids, int id_width, int64_t num_slots, int block_size,
                                        int num_experts, void* sorted_token_ids, void* expert_ids,
                                        void* num_tokens_post_padded, void* stream) {
  B2L_REQUIRE(topk_ids && sorted_token_ids && expert_ids && num_tokens_post_padded, "moe_align_block_size: null pointer");
  B2L_REQUIRE(id_width == 4 || id_width == 8, "moe_align_block_size: topk_ids must be int32 or int64");
  B2L_REQUIRE(block_size > 0 && num_experts > 0 && num_experts <= 2048 && num_slots >= 0 && num_slots < (1 << 30),
              "moe_align_block_size: bad sizes (slots=%lld block=%d experts=%d)", (long long)num_slots, block_size,
              num_experts);
  const int max_padded = (int)num_slots + num_experts * (block_size - 1);
  const int max_blocks = (max_padded + block_size - 1) / block_size;
  const int smem = (kAlignWarps + 2) * num_experts * (int)sizeof(int);
  B2L_REQUIRE(smem <= 200 * 1024, "moe_align_block_size: too many experts for one CTA");
  static int configured[64] = {0};
  if (smem > 48 * 1024) B2L_CUDA_CHECK(b2l_ensure_dynamic_smem(moe_align_kernel, smem, configured));
  b2l_launch(moe_align_kernel, dim3(1), dim3(kAlignThreads), (size_t)(smem), (cudaStream_t)stream, 
      topk_ids, id_width, (int)num_slots, block_size, num_experts, (int32_t*)sorted_token_ids, (int32_t*)expert_ids,
      (int32_t*)num_tokens_post_padded, max_padded, max_blocks);
  B2L_LAUNCH_CHECK();
  return B2L_OK;
}

extern "C" int b2l_moe_route_topk(void* topk_weights, void* topk_ids_i64, const void* router_logits,
                                  int64_t row_stride, int64_t tokens, int num_experts, int top_k, int renormalize,
                                  int dtype, void* stream) {
  B2L_REQUIRE(topk_weights && topk_ids_i64 && router_logits, "moe_route_topk: null pointer");
  B2L_REQUIRE(num_experts > 0 && num_experts <= 1024 && top_k > 0 && top_k <= 32 && top_k <= num_experts,
              "moe_route_topk: experts=%d top_k=%d unsupported", num_experts, top_k);
  if (tokens == 0) return B2L_OK;
  const int warps = 4;
  const unsigned grid = (unsigned)cdiv64(tokens, warps);
  cudaStream_t st = (cudaStream_t)stream;
#define ROUTE_LAUNCH(TT, MV)                                                                              \
  b2l_launch(moe_route_kernel<TT, MV>, dim3(grid), dim3(warps * 32), (size_t)(0), st, (const TT*)router_logits, row_stride, (int)tokens, \
                                                       num_experts, top_k, renormalize, (TT*)topk_weights, \
                                                       (int64_t*)topk_ids_i64)
#define ROUTE_DISPATCH(TT)                                  \
  do {                                                      \
    if (num_experts <= 128) ROUTE_LAUNCH(TT, 4);            \
    else if (num_experts <= 256) ROUTE_LAUNCH(TT, 8);       \
    else ROUTE_LAUNCH(TT, 32);                              \
  } while (0)
  if (dtype == B2L_F16) ROUTE_DISPATCH(__half);
  else if (dtype == B2L_BF16) ROUTE_DISPATCH(__nv_bfloat16);
  else B2L_UNSUPPORTED("moe_route_topk: dtype %d", dtype);
#undef ROUTE_DISPATCH
#undef ROUTE_LAUNCH
  B2L_LAUNCH_CHECK();
  return B2L_OK;
}
