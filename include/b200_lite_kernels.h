/*
 * b200_lite_kernels.h -- C ABI of libb200lite.so, the B200 (sm_100a) kernel layer that stands in
 * for `lite_llama.kernels` (reference: /root/reference/lite_llama/kernels/__init__.py:11-39).
 *
 * The reference has no FFI: its boundary is Python name binding (`from ..kernels import ...`,
 * lite_llama/models/base.py:31-38, lite_llama/executor/model_runner.py:22,
 * lite_llama/models/quantization/methods/*.py).  Each entry point below is what a ctypes binding
 * for one of those names calls; lite_llama_b200/kernels/ holds that binding with the reference's
 * exact Python signatures, INTEGRATION.md shows how to bind them inside the unmodified reference.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless stated otherwise; the caller owns all memory;
 *   - no entry point allocates, synchronises or touches the default stream: work is enqueued on
 *     `stream` (a cudaStream_t) and is CUDA-graph capturable;
 *   - strides are in ELEMENTS of the tensor's dtype; the innermost dimension must be contiguous;
 *   - `dtype`: B2L_F16 / B2L_BF16 activation dtype (the reference is fp16-only);
 *   - `*_width`: 4 or 8, byte width of an index vector (callers pass int32 or int64 vectors for
 *     b_seq_len / b_req_idx / b_start_loc);
 *   - return 0 on success, negative on error (B2L_ERR_*); b2l_last_error() returns a thread-local
 *     message.  The Python shim turns non-zero returns into ValueError, mirroring the reference's
 *     error behaviour (e.g. kernels/quantization/w4a16.py:174-184);
 *   - workspaces: `b2l_*_workspace_bytes` are pure functions of shapes (+ SM count), so the caller
 *     can size a torch.empty scratch buffer outside graph capture.  Split-K GEMMs additionally take a
 *     small int32 `counters` array that must be ZERO on entry; kernels leave it zero on exit.
 *     The QUANTISED GEMMs' workspace (b2l_w4a16_matmul*, b2l_w8a16_matmul) carries stream-K partial sums as
 *     tagged words (flag-in-data): it must be ZERO on entry as well and is left zero on exit, so keep ONE persistent,
 *     once-zeroed buffer per device (and one stream of launches), exactly like `counters`.
 */
#ifndef B200_LITE_KERNELS_H
#define B200_LITE_KERNELS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2L_F16 0
#define B2L_BF16 1

#define B2L_OK 0
#define B2L_ERR_INVALID (-1)
#define B2L_ERR_CUDA (-2)
#define B2L_ERR_UNSUPPORTED (-3)

/* quantised-weight kinds for b2l_quant_gemm_workspace_bytes */
#define B2L_Q_INT4 0
#define B2L_Q_FP8 1
#define B2L_Q_INT8 2

const char* b2l_last_error(void);
int b2l_version(void);
int b2l_device_info(int* sm_count, int* cc_major, int* cc_minor, int* max_smem_optin);
/* number of kernels this library has launched in this process (monotonic; bench.py's gpu_launches) */
int64_t b2l_launch_count(void);
/* Programmatic dependent launch (every kernel is launched so that the next one's set-up and weight prefetch
 * overlap it; see csrc/common.cuh).  Returns the previous setting.  Also: environment B2L_PDL=0 at load.
 * No reference counterpart: the reference launches every Triton kernel fully serialised. */
int b2l_set_pdl(int enabled);
/* Device-side hang watchdog (csrc/common.cuh): every barrier wait of the warp-specialised kernels traps after
 * B2L_HANG_MS milliseconds (environment, default 4000, 0 = off) instead of parking for ever, after writing a
 * host-pinned record.  out16 receives: [0] 0xB2D00DED when a record exists, [1] FNV-1a hash of the source file's
 * basename, [2] line of the wait, [3..5] blockIdx, [6] threadIdx.x, [7] shared address of the barrier, [8] parity,
 * [9] ms waited, [10..12] gridDim, [13] blockDim.x.  Readable after the CUDA context has died.  No reference
 * counterpart (Triton kernels have no cross-warp pipelines to dead-lock). */
int b2l_hang_report(int* out16);

/* skip_rmsnorm(X, residual, weight, eps) -- lite_llama/kernels/skip_rmsnorm.py:192-234.
 * residual (nullable) is updated in place with x + residual (rounded to dtype); y gets
 * round(sum_f32 * rrms) * weight. */
int b2l_skip_rmsnorm(void* y, const void* x, void* residual, const void* weight, int64_t rows, int64_t n,
                     int64_t x_row_stride, int64_t r_row_stride, int64_t y_row_stride, float eps, int dtype,
                     void* stream);

/* rope_emb_forward(q, k, cos, sin, batch_size, seq_len) -- lite_llama/kernels/rope_emb.py:86-134.
 * In place on q [B*S, n_qh, D] and k [B*S, n_kh, D]; cos/sin [B, S, >=D/2] (first D/2 entries read). */
int b2l_rope_emb(void* q, void* k, const void* cos, const void* sin, int64_t batch, int64_t seq_len, int n_qh,
                 int n_kh, int head_dim, int64_t q_tok_stride, int64_t q_head_stride, int64_t k_tok_stride,
                 int64_t k_head_stride, int64_t cs_batch_stride, int64_t cs_seq_stride, int dtype, void* stream);

/* swiglu_forward(a, b) -- lite_llama/kernels/swiglu.py:45-65; also fused_moe's _silu_and_mul
 * (lite_llama/kernels/fused_moe.py:296-313) by pointing b at the second half of each row. */
int b2l_swiglu(void* c, const void* a, const void* b, int64_t rows, int64_t n, int64_t a_row_stride,
               int64_t b_row_stride, int64_t c_row_stride, int dtype, void* stream);

/* update_kv_buffer(KV_Values, Select_Index, KV_Buffer) -- lite_llama/kernels/update_kv_buffer.py:55-89.
 * Heads [0, heads_a) of every destination row come from src_a, heads [heads_a, heads_a+heads_b) from
 * src_b (so K and V need no torch.cat, lite_llama/models/base.py:76); pass heads_b = 0 for one source.
 * 2-byte element dtype. */
int b2l_update_kv_buffer(const void* src_a, const void* src_b, const void* select_index, int index_width,
                         void* kv_buffer, int64_t tokens, int heads_a, int heads_b, int head_dim,
                         int64_t a_tok_stride, int64_t a_head_stride, int64_t b_tok_stride, int64_t b_head_stride,
                         int64_t d_tok_stride, int64_t d_head_stride, void* stream);

/* update_kv_index(req_to_token_indexs, b_req_idx, b_seq_len, select_index)
 * -- lite_llama/kernels/update_kv_index.py:51-88.  table is int32. */
int b2l_update_kv_index(void* table, const void* b_req_idx, int req_width, const void* b_seq_len, int len_width,
                        const void* select_index, int sel_width, int64_t n, int64_t stride_b, int64_t stride_s,
                        void* stream);

/* flash_decoding(q, k_cache, v_cache, qk_scale, b_req_tokens_table, b_req_idx, b_seq_len,
 * max_actual_seq_len) -- lite_llama/kernels/flashdecoding.py:316-380. */
int64_t b2l_flash_decoding_workspace_bytes(int64_t batch, int num_q_heads, int num_kv_heads, int head_dim,
                                           int64_t max_seq_len);
int b2l_flash_decoding(void* out, const void* q, const void* k_cache, const void* v_cache,
                       const void* b_req_tokens_table, const void* b_req_idx, int req_idx_width,
                       const void* b_seq_len, int seq_len_width, int64_t batch, int num_q_heads, int num_kv_heads,
                       int head_dim, int64_t max_seq_len, float qk_scale, int64_t q_batch_stride,
                       int64_t q_head_stride, int64_t k_token_stride, int64_t k_head_stride, int64_t v_token_stride,
                       int64_t v_head_stride, int64_t table_batch_stride, int64_t table_seq_stride,
                       int64_t o_batch_stride, int64_t o_head_stride, void* workspace, int64_t workspace_bytes,
                       void* tickets, int64_t num_tickets, int dtype, void* stream);

/* flash_attention2_no_pad(q, k, v, sm_scale, b_start_loc, b_seq_len, max_seq_len)
 * -- lite_llama/kernels/flashattention2_nopad.py:175-231.  sm_scale already carries log2(e).  total_tokens = q.shape[0]
 * (rows of q / k / v; bounds the TMA descriptors of the tcgen05 kernel, csrc/flash_attention_tc.cu, which takes
 * head_dim 128 with heads packed inside a token row; other shapes run the mma.sync kernel). */
int b2l_flash_attention2_no_pad(void* out, const void* q, const void* k, const void* v, const void* b_start_loc,
                                int start_loc_width, const void* b_seq_len, int seq_len_width, int64_t batch,
                                int num_q_heads, int num_kv_heads, int head_dim, int64_t max_seq_len,
                                int64_t total_tokens, float sm_scale_log2, int64_t q_tok_stride, int64_t q_head_stride,
                                int64_t k_tok_stride, int64_t k_head_stride, int64_t v_tok_stride,
                                int64_t v_head_stride, int64_t o_tok_stride, int64_t o_head_stride, int dtype,
                                void* stream);

/* rope_emb_forward + torch.cat + update_kv_buffer in ONE launch (lite_llama/kernels/rope_emb.py:86-134,
 * lite_llama/models/base.py:76-79, lite_llama/kernels/update_kv_buffer.py:55-89): rotates q [T, n_qh, D] and
 * k [T, n_kh, D] in place (same arithmetic as b2l_rope_emb), stores the rotated k rows and the v rows [T, n_kh, D] into
 * kv_buffer [max_tokens, 2 * n_kh, D] at rows select_index[t] (K heads first).  Strides in elements. */
int b2l_rope_kv_write(void* q, void* k, const void* v, const void* cos, const void* sin, const void* select_index,
                      int index_width, void* kv_buffer, int64_t batch, int64_t seq_len, int n_qh, int n_kh,
                      int head_dim, int64_t q_tok_stride, int64_t q_head_stride, int64_t k_tok_stride,
                      int64_t k_head_stride, int64_t v_tok_stride, int64_t v_head_stride, int64_t cs_batch_stride,
                      int64_t cs_seq_stride, int64_t d_tok_stride, int64_t d_head_stride, int dtype, void* stream);

/* Greedy sampling: out[r] = argmax_j x[r, j] (int64; ties -> lowest index).  Reference: temperature-0 path of the
 * Sampler, torch.argmax over the [batch, vocab] logits (lite_llama/engine/sampler.py, engine/llm_engine.py:172). */
int b2l_argmax_rows(void* out_index_i64, const void* x, int64_t rows, int64_t n, int64_t row_stride, int dtype,
                    void* stream);

/* Dense projection y = x @ w.T (+ bias): the reference's F.linear call sites
 * (lite_llama/models/quantization/methods/unquantized.py:21-22, lite_llama/models/base.py:489). */
int64_t b2l_linear_workspace_bytes(int64_t m, int64_t n, int64_t k, int elem_bytes);
int64_t b2l_linear_num_counters(int64_t m, int64_t n);
int b2l_linear(void* y, const void* x, const void* w, const void* bias, int64_t m, int64_t n, int64_t k,
               int64_t x_row_stride, int64_t w_row_stride, int64_t y_row_stride, void* workspace,
               int64_t workspace_bytes, void* counters, int64_t num_counters, int dtype, void* stream);

/* w4a16_matmul(x, qweight, scales, zeros, *, group_size, bias)
 * -- lite_llama/kernels/quantization/w4a16.py:152-207.  fp16 activations only (w4a16.py:174). */
int64_t b2l_quant_gemm_workspace_bytes(int kind, int64_t m, int64_t n, int64_t k, int64_t group_k);
int64_t b2l_quant_gemm_num_counters(int64_t m, int64_t n);
/* launch plan of a quantised GEMM, for tests and tools (no GPU work): out[10] = {stream-K CTAs (0 = tile x split
 * grid), workspace slots per tile, raw blocks per tile, weight-row tiles, splits, k-blocks per split, weight ring
 * depth, activation ring depth, shared-memory bytes, scale-table columns}.  The trailing argument is unused. */
int b2l_quant_gemm_plan(int kind, int64_t m, int64_t n, int64_t k, int64_t group_k, int* out, void* reserved);
int b2l_w4a16_matmul(void* y, const void* x, const void* qweight, const void* scales, const void* zeros,
                     const void* bias, int bias_is_f32, int64_t m, int64_t n, int64_t k, int group_size,
                     int64_t x_row_stride, int64_t qw_row_stride, int64_t s_row_stride, int64_t s_col_stride,
                     int64_t y_row_stride, void* workspace, int64_t workspace_bytes, void* counters,
                     int64_t num_counters, void* stream);

/* Load-time repack of the three W4A16 tensors into the kernel-native layout (the reference reserves this seam:
 * lite_llama/models/quantization/_layout/__init__.py:1-6; its own tensors stay as w4a16.py:9-12 defines them).
 * Bit-exact permutations, one pass each:
 *   qweight_packed int32 [N, K/8]: nibble position p of a word holds the K index [0,2,4,6,1,3,5,7][p] of that word
 *   table_packed   fp32 [ceil(N/128)][groups][128]: table_packed[t][g][r] = table[128 t + r][g], rows past N zero
 * b2l_w4a16_matmul_packed computes exactly b2l_w4a16_matmul's result from the repacked tensors; zeros_foldable = 1
 * promises that every zero point is an integer in [0, 1023] (checked once by the packer's caller).
 * epilogue = 1 fuses swiglu_forward (lite_llama/kernels/swiglu.py:45-65; models/base.py:259-264 runs gate_proj,
 * up_proj and swiglu as three launches): the packed rows must be interleaved (row 2i = gate_i, row 2i+1 = up_i);
 * y is [m, n/2] = silu(gate) * up with both projections rounded to fp16 first, bit-identical to the unfused path. */
int b2l_w4a16_pack_qweight(void* out, const void* in, int64_t n, int64_t k8, int64_t in_row_stride,
                           int64_t out_row_stride, void* stream);
int b2l_w4a16_pack_table(void* out, const void* in, int64_t n, int64_t groups, int64_t row_stride,
                         int64_t col_stride, void* stream);
int b2l_w4a16_matmul_packed(void* y, const void* x, const void* qweight_packed, const void* scales_packed,
                            const void* zeros_packed, const void* bias, int bias_is_f32, int64_t m, int64_t n,
                            int64_t k, int group_size, int zeros_foldable, int64_t x_row_stride,
                            int64_t qw_row_stride, int64_t y_row_stride, void* workspace, int64_t workspace_bytes,
                            void* counters, int64_t num_counters, int epilogue, void* stream);

/* w8a16_matmul(x, qweight, scales, *, group_n, group_k, bias)
 * -- lite_llama/kernels/quantization/w8a16.py:155-216.  qweight uint8 (fp8-e4m3 bits) or int8. */
int b2l_w8a16_matmul(void* y, const void* x, const void* qweight, int weight_is_fp8, const void* scales,
                     const void* bias, int bias_is_f32, int64_t m, int64_t n, int64_t k, int64_t group_n,
                     int64_t group_k, int64_t x_row_stride, int64_t qw_row_stride, int64_t s_row_stride,
                     int64_t s_col_stride, int64_t y_row_stride, void* workspace, int64_t workspace_bytes,
                     void* counters, int64_t num_counters, void* stream);

/* smoothquant_matmul(x, qweight, weight_scales, *, bias) -- lite_llama/kernels/quantization/w8a8.py:151-217
 * is two calls: per-token dynamic quantisation (w8a8.py:35-68), then the int8 GEMM (w8a8.py:75-130). */
int b2l_quantize_activations_int8(void* q, void* scales, const void* x, int64_t m, int64_t k, int64_t x_row_stride,
                                  int64_t q_row_stride, int dtype, void* stream);
/* skip_rmsnorm (skip_rmsnorm.py:192-234) followed by the per-token quantisation above (w8a8.py:35-68) in one
 * launch, for a norm whose only consumer is an int8 GEMM (models/base.py:307-312 under SmoothQuant): q / scales
 * are bit-identical to b2l_skip_rmsnorm + b2l_quantize_activations_int8; the fp16 norm output is never written.
 * Needs n % 8 == 0, n <= 16384, 16-byte aligned rows. */
int b2l_skip_rmsnorm_quant_int8(void* q, void* scales, const void* x, void* residual, const void* weight,
                                int64_t rows, int64_t n, int64_t x_row_stride, int64_t r_row_stride,
                                int64_t q_row_stride, float eps, int dtype, void* stream);
int b2l_w8a8_gemm(void* y, const void* xq, const void* wq, const void* a_scale, const void* b_scale,
                  const void* bias_f32, int64_t m, int64_t n, int64_t k, int64_t xq_row_stride,
                  int64_t wq_row_stride, int64_t y_row_stride, void* workspace, int64_t workspace_bytes,
                  void* counters, int64_t num_counters, int out_dtype, void* stream);

/* fused_moe(hidden_states, w1, w2, topk_weights, topk_ids, *, w1_scale, w2_scale, group_n, group_k)
 * -- lite_llama/kernels/fused_moe.py:352-438 -- is moe_align_block_size (:45-99), grouped GEMM1
 * (b2l_moe_gemm), silu_and_mul (b2l_swiglu on the two halves, :296-313), grouped GEMM2 with the routing
 * weight folded in (b2l_moe_gemm, :201-205) and moe_sum (:316-333).
 *
 * moe_align_block_size(topk_ids, block_size, num_experts): sorted_token_ids int32
 * [num_slots + E*(block-1)] (sentinel = num_slots), expert_ids int32 [ceil(that / block)],
 * num_tokens_post_padded int32 [1]; topk_ids int32 or int64. */
int b2l_moe_align_block_size(const void* topk_ids, int id_width, int64_t num_slots, int block_size, int num_experts,
                             void* sorted_token_ids, void* expert_ids, void* num_tokens_post_padded, void* stream);
/* C[slot, :] = A[slot / top_k, :] @ W[expert(block)].T (* topk_weights[slot] when non-null).  fp16 activations.
 * weight_kind: 0 = fp16 [E, N, K], B2L_Q_FP8 / B2L_Q_INT8 = 8-bit with fp32 scales
 * [E, ceil(N/group_n), ceil(K/group_k)] (element strides s_stride_*).  block_size in {16, 32, 64}. */
int b2l_moe_gemm(void* c, const void* a, const void* w, int weight_kind, const void* scales,
                 const void* topk_weights, const void* sorted_token_ids, const void* expert_ids,
                 const void* num_tokens_post_padded, int64_t em, int64_t num_valid_slots, int top_k, int block_size,
                 int num_experts, int64_t n, int64_t k, int64_t a_row_stride, int64_t w_row_stride,
                 int64_t c_row_stride, int64_t group_n, int64_t group_k, int64_t s_stride_e, int64_t s_stride_n,
                 int64_t s_stride_k, void* stream);
int b2l_moe_sum(void* out, const void* in, int64_t tokens, int64_t n, int top_k, int dtype, void* stream);
/* Router (lite_llama/models/qwen3_moe.py:94-100): fp32 softmax over all experts of the [tokens, E] logits,
 * top-k (ties -> lowest index), optional renormalisation; weights in the activation dtype, ids int64. */
int b2l_moe_route_topk(void* topk_weights, void* topk_ids_i64, const void* router_logits, int64_t row_stride,
                       int64_t tokens, int num_experts, int top_k, int renormalize, int dtype, void* stream);

/* ---- tensor-parallel exchange over NVLink peer memory (csrc/tp_comm.cu) ----
 * Replaces all_reduce_tp (lite_llama/distributed/parallel_state.py:208-213: in-place NCCL SUM of the [tokens, hidden]
 * partial after every row-parallel projection, lite_llama/models/linear.py:160-161, models/qwen3_moe.py:110) TOGETHER
 * WITH the skip_rmsnorm that follows every one of them (lite_llama/models/base.py:307,312,486): one kernel pushes this
 * rank's partial rows into every peer's receive buffer as 16-byte {6 values, epoch} packets (no fence, no separate
 * flag), polls the peers' packets in its own buffer, sums in rank order (fp32, rounded to
 * the activation dtype like the all-reduce output) and applies residual += sum; y = rmsnorm(residual) * weight.
 * One process per GPU: b2l_tp_comm_create allocates this rank's buffer (capacity max_rows x max_row_elems 16-bit
 * elements per peer) and returns an opaque handle (0 on error); b2l_tp_comm_get_handle exports its CUDA IPC handle
 * (b2l_tp_comm_handle_bytes() bytes); the caller all-gathers the handles (torch.distributed) and passes the rank-major
 * concatenation to b2l_tp_comm_open.  weight == NULL: plain all-reduce, y = sum.  Every rank must issue the same call
 * sequence.  Graph-capturable (the double-buffer epoch lives in device memory). */
int64_t b2l_tp_comm_create(int rank, int world, int max_rows, int64_t max_row_elems);
int b2l_tp_comm_handle_bytes(void);
int b2l_tp_comm_get_handle(int64_t comm, void* out, int64_t out_bytes);
int b2l_tp_comm_open(int64_t comm, const void* handles, int64_t handles_bytes);
int b2l_tp_comm_destroy(int64_t comm);
int b2l_tp_allreduce_skip_rmsnorm(int64_t comm, void* y, const void* x_partial, void* residual, const void* weight,
                                  int64_t rows, int64_t n, int64_t x_row_stride, int64_t r_row_stride,
                                  int64_t y_row_stride, float eps, int dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200_LITE_KERNELS_H */
