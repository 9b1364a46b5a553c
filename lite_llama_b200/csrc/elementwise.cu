// Bandwidth-bound fused ops of the decoder layer: skip+RMSNorm, RoPE, SiLU*gate,
// KV-cache scatter, token-table update, per-token int8 activation quantisation,
// and the two small MoE glue ops. All are HBM-bound: 128-bit accesses, one pass.
//
// Reference semantics (cited per op) are reproduced including their rounding
// points, see SURVEY.md §7 "bit-level parity traps".
#include "common.cuh"

// Mirrors of the approximate ops Triton emits for the reference kernels on sm_100a
// (probed: sqrt.approx.ftz.f32, div.full.f32, ex2.approx) so results round the same way.
__device__ __forceinline__ float sqrt_approx(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float div_full(float a, float b) {
  float r;
  asm("div.full.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// --------------------------------------------------------------------------- //
// skip_rmsnorm  (reference: lite_llama/kernels/skip_rmsnorm.py:126-234)
//   r += x (stored back in the activation dtype), y = round(r_f32 * rrms) * w
// --------------------------------------------------------------------------- //
template <typename T, int VPT_MAX>
__global__ void __launch_bounds__(512) skip_rmsnorm_block_kernel(T* __restrict__ y, const T* __restrict__ x,
                                                                 T* __restrict__ res, const T* __restrict__ w,
                                                                 int64_t n, int64_t x_stride, int64_t r_stride,
                                                                 int64_t y_stride, float eps) {
  pdl_prologue();
  // One CTA per row; each thread keeps up to VPT_MAX 16-byte chunks in registers.
  const int64_t row = blockIdx.x;
  const T* xr = x + row * x_stride;
  T* rr = res ? res + row * r_stride : nullptr;
  T* yr = y + row * y_stride;
  const int nchunks = (int)(n / 8);
  float vals[VPT_MAX][8];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < VPT_MAX; ++i) {
    const int c = threadIdx.x + i * blockDim.x;
    if (c < nchunks) {
      uint4 xv = ldg_128(xr + c * 8);
      const T* xe = reinterpret_cast<const T*>(&xv);
      if (rr) {
        uint4 rv = ldg_128(rr + c * 8);
        const T* re = reinterpret_cast<const T*>(&rv);
        uint4 ov;
        T* oe = reinterpret_cast<T*>(&ov);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float s = to_f32(xe[j]) + to_f32(re[j]);
          vals[i][j] = s;
          oe[j] = from_f32<T>(s);
        }
        stg_128(rr + c * 8, ov);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) vals[i][j] = to_f32(xe[j]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) ss += vals[i][j] * vals[i][j];
    }
  }
  __shared__ float red[16];
  ss = warp_reduce_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
  const int nw = blockDim.x >> 5;
  for (int i = 0; i < nw; ++i) tot += red[i];
  const float rrms = div_full(1.0f, sqrt_approx(tot / (float)n + eps));
#pragma unroll
  for (int i = 0; i < VPT_MAX; ++i) {
    const int c = threadIdx.x + i * blockDim.x;
    if (c < nchunks) {
      uint4 wv = ldg_128(w + c * 8);
      const T* we = reinterpret_cast<const T*>(&wv);
      uint4 ov;
      T* oe = reinterpret_cast<T*>(&ov);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        // round to the output dtype BEFORE the weight multiply (skip_rmsnorm.py:188)
        T nrm = from_f32<T>(vals[i][j] * rrms);
        oe[j] = from_f32<T>(to_f32(nrm) * to_f32(we[j]));
      }
      stg_128(yr + c * 8, ov);
    }
  }
}

// Generic fallback without a residual: any n, one warp per row, two passes over x.
template <typename T>
__global__ void skip_rmsnorm_warp_kernel(T* __restrict__ y, const T* __restrict__ x, T* __restrict__ res,
                                         const T* __restrict__ w, int64_t rows, int64_t n, int64_t x_stride,
                                         int64_t r_stride, int64_t y_stride, float eps) {
  pdl_prologue();
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const T* xr = x + row * x_stride;
  T* rr = res ? res + row * r_stride : nullptr;
  T* yr = y + row * y_stride;
  // pass 1 only reads: the residual is overwritten in pass 2, where x + residual is formed again (same fp32 sum)
  float ss = 0.f;
  for (int64_t c = lane; c < n; c += 32) {
    const float s = to_f32(xr[c]) + (rr ? to_f32(rr[c]) : 0.f);
    ss += s * s;
  }
  ss = warp_reduce_sum(ss);
  const float rrms = div_full(1.0f, sqrt_approx(ss / (float)n + eps));
  for (int64_t c = lane; c < n; c += 32) {
    float s = to_f32(xr[c]);
    if (rr) {
      s += to_f32(rr[c]);
      rr[c] = from_f32<T>(s);
    }
    T nrm = from_f32<T>(s * rrms);
    yr[c] = from_f32<T>(to_f32(nrm) * to_f32(w[c]));
  }
}

// Warp-per-row kernel that keeps the fp32 sums in registers (n <= 32*KEEP), exact for
// the residual case. Used for small rows (per-head norm, small hidden sizes).
template <typename T, int KEEP>
__global__ void skip_rmsnorm_warp_reg_kernel(T* __restrict__ y, const T* __restrict__ x, T* __restrict__ res,
                                             const T* __restrict__ w, int64_t rows, int64_t n, int64_t x_stride,
                                             int64_t r_stride, int64_t y_stride, float eps) {
  pdl_prologue();
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const T* xr = x + row * x_stride;
  T* rr = res ? res + row * r_stride : nullptr;
  T* yr = y + row * y_stride;
  float v[KEEP];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < KEEP; ++i) {
    const int64_t c = lane + 32 * i;
    v[i] = 0.f;
    if (c < n) {
      float s = to_f32(xr[c]);
      if (rr) {
        s += to_f32(rr[c]);
        rr[c] = from_f32<T>(s);
      }
      v[i] = s;
      ss += s * s;
    }
  }
  ss = warp_reduce_sum(ss);
  const float rrms = div_full(1.0f, sqrt_approx(ss / (float)n + eps));
#pragma unroll
  for (int i = 0; i < KEEP; ++i) {
    const int64_t c = lane + 32 * i;
    if (c < n) {
      T nrm = from_f32<T>(v[i] * rrms);
      yr[c] = from_f32<T>(to_f32(nrm) * to_f32(w[c]));
    }
  }
}

template <typename T>
static int launch_skip_rmsnorm(void* y, const void* x, void* res, const void* w, int64_t rows, int64_t n,
                               int64_t xs, int64_t rs, int64_t ys, float eps, cudaStream_t st) {
  if (rows == 0) return B2L_OK;
  const bool aligned = (n % 8 == 0) && (xs % 8 == 0) && (ys % 8 == 0) && (res == nullptr || rs % 8 == 0) &&
                       ((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0) && ((uintptr_t)w % 16 == 0) &&
                       (res == nullptr || (uintptr_t)res % 16 == 0);
  if (aligned && n >= 1024 && n <= 8 * 256 * 8) {
    // one CTA per row
    // decode rows are latency-bound (32 rows of 4096): one 16-byte chunk per thread puts every load of the row in
    // flight at once (512 threads for hidden 4096); wider rows keep 256 threads and several chunks per thread
    int threads = 256;
    int64_t chunks = n / 8;
    if (chunks <= 128) threads = 128;
    else if (chunks <= 512 && chunks > 256) threads = 512;
    int vpt = (int)cdiv64(chunks, threads);
    T* yy = (T*)y; const T* xx = (const T*)x; T* rr = (T*)res; const T* ww = (const T*)w;
    if (vpt <= 1)
      b2l_launch(skip_rmsnorm_block_kernel<T, 1>, dim3((unsigned)rows), dim3(threads), (size_t)(0), st, yy, xx, rr, ww, n, xs, rs, ys, eps);
    else if (vpt <= 2)
      b2l_launch(skip_rmsnorm_block_kernel<T, 2>, dim3((unsigned)rows), dim3(threads), (size_t)(0), st, yy, xx, rr, ww, n, xs, rs, ys, eps);
    else if (vpt <= 4)
      b2l_launch(skip_rmsnorm_block_kernel<T, 4>, dim3((unsigned)rows), dim3(threads), (size_t)(0), st, yy, xx, rr, ww, n, xs, rs, ys, eps);
    else
      b2l_launch(skip_rmsnorm_block_kernel<T, 8>, dim3((unsigned)rows), dim3(threads), (size_t)(0), st, yy, xx, rr, ww, n, xs, rs, ys, eps);
    B2L_LAUNCH_CHECK();
    return B2L_OK;
  }
  const int warps = 4;
  const unsigned grid = (unsigned)cdiv64(rows, warps);
  if (n <= 32 * 4)
    b2l_launch(skip_rmsnorm_warp_reg_kernel<T, 4>, dim3(grid), dim3(warps * 32), (size_t)(0), st, (T*)y, (const T*)x, (T*)res, (const T*)w, rows, n,
                                                                 xs, rs, ys, eps);
  else if (n <= 32 * 32)
    b2l_launch(skip_rmsnorm_warp_reg_kernel<T, 32>, dim3(grid), dim3(warps * 32), (size_t)(0), st, (T*)y, (const T*)x, (T*)res, (const T*)w, rows,
                                                                  n, xs, rs, ys, eps);
  else if (n <= 32 * 128)
    b2l_launch(skip_rmsnorm_warp_reg_kernel<T, 128>, dim3(grid), dim3(warps * 32), (size_t)(0), st, (T*)y, (const T*)x, (T*)res, (const T*)w, rows,
                                                                   n, xs, rs, ys, eps);
  else  // any width / alignment, with or without a residual: one warp per row, two passes
    b2l_launch(skip_rmsnorm_warp_kernel<T>, dim3(grid), dim3(warps * 32), (size_t)(0), st, (T*)y, (const T*)x, (T*)res,
               (const T*)w, rows, n, xs, rs, ys, eps);
  B2L_LAUNCH_CHECK();
  return B2L_OK;
}

extern "C" int b2l_skip_rmsnorm(void* y, const void* x, void* residual, const void* weight, int64_t rows, int64_t n,
                                int64_t x_row_stride, int64_t r_row_stride, int64_t y_row_stride, float eps,
                                int dtype, void* stream) {
  B2L_REQUIRE(y && x && weight, "skip_rmsnorm: null pointer");
  B2L_REQUIRE(rows >= 0 && n > 0, "skip_rmsnorm: bad shape rows=%lld n=%lld", (long long)rows, (long long)n);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == B2L_F16)
    return launch_skip_rmsnorm<__half>(y, x, residual, weight, rows, n, x_row_stride, r_row_stride, y_row_stride, eps, st);
  if (dtype == B2L_BF16)
    return launch_skip_rmsnorm<__nv_bfloat16>(y, x, residual, weight, rows, n, x_row_stride, r_row_stride, y_row_stride,
                                              eps, st);
  B2L_UNSUPPORTED("skip_rmsnorm: dtype %d", dtype);
}

// --------------------------------------------------------------------------- //
// rope_emb_forward (reference: lite_llama/kernels/rope_emb.py:14-134)
//   half-split rotation, in place, arithmetic in the activation dtype:
//     new1 = fma(x1, cos, -(x2*sin));  new2 = fma(x2, cos, x1*sin)
//   (exactly the mul+fma contraction Triton emits for this kernel on sm_100a)
// --------------------------------------------------------------------------- //
__device__ __forceinline__ void rope_pair(__half2& a, __half2& b, __half2 c, __half2 s) {
  __half2 t1 = __hmul2(b, s);
  __half2 t2 = __hmul2(a, s);
  __half2 n1 = __hfma2(a, c, __hneg2(t1));
  __half2 n2 = __hfma2(b, c, t2);
  a = n1; b = n2;
}
__device__ __forceinline__ void rope_pair(__nv_bfloat162& a, __nv_bfloat162& b, __nv_bfloat162 c, __nv_bfloat162 s) {
  __nv_bfloat162 t1 = __hmul2(b, s);
  __nv_bfloat162 t2 = __hmul2(a, s);
  __nv_bfloat162 n1 = __hfma2(a, c, __hneg2(t1));
  __nv_bfloat162 n2 = __hfma2(b, c, t2);
  a = n1; b = n2;
}

// One CTA per token; threads sweep (head, pair-chunk). cos/sin rows are indexed
// [batch, pos] with explicit strides; only the first head_dim/2 entries are read.
template <typename T>
__global__ void rope_kernel(T* __restrict__ q, T* __restrict__ k, const T* __restrict__ cosp,
                            const T* __restrict__ sinp, int seq_len, int n_qh, int n_kh, int hd, int64_t q_tok_stride,
                            int64_t q_head_stride, int64_t k_tok_stride, int64_t k_head_stride, int64_t cs_b_stride,
                            int64_t cs_s_stride) {
  pdl_prologue();
  using T2 = typename Vec2<T>::type;
  const int64_t tok = blockIdx.x;
  const int64_t b = tok / seq_len, s = tok % seq_len;
  const T* cr = cosp + b * cs_b_stride + s * cs_s_stride;
  const T* sr = sinp + b * cs_b_stride + s * cs_s_stride;
  const int half = hd / 2;
  const int pairs = half / 2;  // half2 units per half-row
  const int total = (n_qh + n_kh) * pairs;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int h = i / pairs, p = i % pairs;
    T* base = (h < n_qh) ? q + tok * q_tok_stride + (int64_t)h * q_head_stride
                         : k + tok * k_tok_stride + (int64_t)(h - n_qh) * k_head_stride;
    T2* p1 = reinterpret_cast<T2*>(base) + p;
    T2* p2 = reinterpret_cast<T2*>(base + half) + p;
    T2 a = *p1, bb = *p2;
    T2 c = reinterpret_cast<const T2*>(cr)[p];
    T2 sn = reinterpret_cast<const T2*>(sr)[p];
    rope_pair(a, bb, c, sn);
    *p1 = a;
    *p2 = bb;
  }
}

// Scalar variant for head dims whose half is odd (not 4-byte pair aligned).
template <typename T>
__global__ void rope_kernel_scalar(T* __restrict__ q, T* __restrict__ k, const T* __restrict__ cosp,
                                   const T* __restrict__ sinp, int seq_len, int n_qh, int n_kh, int hd,
                                   int64_t q_tok_stride, int64_t q_head_stride, int64_t k_tok_stride,
                                   int64_t k_head_stride, int64_t cs_b_stride, int64_t cs_s_stride) {
  pdl_prologue();
  const int64_t tok = blockIdx.x;
  const int64_t b = tok / seq_len, s = tok % seq_len;
  const T* cr = cosp + b * cs_b_stride + s * cs_s_stride;
  const T* sr = sinp + b * cs_b_stride + s * cs_s_stride;
  const int half = hd / 2;
  const int total = (n_qh + n_kh) * half;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int h = i / half, p = i % half;
    T* base = (h < n_qh) ? q + tok * q_tok_stride + (int64_t)h * q_head_stride
                         : k + tok * k_tok_stride + (int64_t)(h - n_qh) * k_head_stride;
    T a = base[p], bb = base[p + half], c = cr[p], sn = sr[p];
    T t1 = __hmul(bb, sn), t2 = __hmul(a, sn);
    base[p] = __hfma(a, c, __hneg(t1));
    base[p + half] = __hfma(bb, c, t2);
  }
}

extern "C" int b2l_rope_emb(void* q, void* k, const void* cos, const void* sin, int64_t batch, int64_t seq_len,
                            int n_qh, int n_kh, int head_dim, int64_t q_tok_stride, int64_t q_head_stride,
                            int64_t k_tok_stride, int64_t k_head_stride, int64_t cs_batch_stride,
                            int64_t cs_seq_stride, int dtype, void* stream) {
  B2L_REQUIRE(q && k && cos && sin, "rope_emb: null pointer");
  B2L_REQUIRE(batch >= 0 && seq_len >= 0 && n_qh > 0 && n_kh > 0 && head_dim >= 2, "rope_emb: bad shape");
  const int64_t tokens = batch * seq_len;
  if (tokens == 0) return B2L_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int half = head_dim / 2;
  const bool vec = (half % 2 == 0) && (q_tok_stride % 2 == 0) && (q_head_stride % 2 == 0) && (k_tok_stride % 2 == 0) &&
                   (k_head_stride % 2 == 0) && (cs_batch_stride % 2 == 0) && (cs_seq_stride % 2 == 0) &&
                   ((uintptr_t)q % 4 == 0) && ((uintptr_t)k % 4 == 0) && ((uintptr_t)cos % 4 == 0) &&
                   ((uintptr_t)sin % 4 == 0);
  const int work = (n_qh + n_kh) * (vec ? half / 2 : half);
  int threads = work >= 256 ? 256 : (work >= 128 ? 128 : 64);
#define ROPE_ARGS (int)seq_len, n_qh, n_kh, head_dim, q_tok_stride, q_head_stride, k_tok_stride, k_head_stride, \
                  cs_batch_stride, cs_seq_stride
  if (dtype == B2L_F16) {
    if (vec)
      b2l_launch(rope_kernel<__half>, dim3((unsigned)tokens), dim3(threads), (size_t)(0), st, (__half*)q, (__half*)k, (const __half*)cos,
                                                               (const __half*)sin, ROPE_ARGS);
    else
      b2l_launch(rope_kernel_scalar<__half>, dim3((unsigned)tokens), dim3(threads), (size_t)(0), st, (__half*)q, (__half*)k, (const __half*)cos,
                                                                      (const __half*)sin, ROPE_ARGS);
  } else if (dtype == B2L_BF16) {
    if (vec)
      b2l_launch(rope_kernel<__nv_bfloat16>, dim3((unsigned)tokens), dim3(threads), (size_t)(0), st, 
          (__nv_bfloat16*)q, (__nv_bfloat16*)k, (const __nv_bfloat16*)cos, (const __nv_bfloat16*)sin, ROPE_ARGS);
    else
      b2l_launch(rope_kernel_scalar<__nv_bfloat16>, dim3((unsigned)tokens), dim3(threads), (size_t)(0), st, 
          (__nv_bfloat16*)q, (__nv_bfloat16*)k, (const __nv_bfloat16*)cos, (const __nv_bfloat16*)sin, ROPE_ARGS);
  } else {
    B2L_UNSUPPORTED("rope_emb: dtype %d", dtype);
  }
#undef ROPE_ARGS
  B2L_LAUNCH_CHECK();
  return B2L_OK;
}

// --------------------------------------------------------------------------- //
// rope + KV-cache write in one launch: the reference rotates q and k (rope_emb.py:86-134), then
// concatenates k|v and scatters the rows into the pool (models/base.py:76-79, update_kv_buffer.py:55-89) -- three
// launches and two copies per layer.  Here one CTA per token rotates q in place, rotates k in place AND stores the
// rotated row into the pool, and copies v into the pool.  Same arithmetic as rope_kernel (bit-identical results).
// --------------------------------------------------------------------------- //
template <typename T>
__global__ void rope_kv_write_kernel(T* __restrict__ q, T* __restrict__ k, const T* __restrict__ v,
                                     const T* __restrict__ cosp, const T* __restrict__ sinp,
                                     const void* __restrict__ select_index, int idx_width, T* __restrict__ pool,
                                     int seq_len, int n_qh, int n_kh, int hd, int64_t q_ts, int64_t q_hs, int64_t k_ts,
                                     int64_t k_hs, int64_t v_ts, int64_t v_hs, int64_t cs_b, int64_t cs_s, int64_t d_ts,
                                     int64_t d_hs) {
  pdl_prologue();
  using T2 = typename Vec2<T>::type;
  const int64_t tok = blockIdx.x;
  const int64_t b = tok / seq_len, s = tok % seq_len;
  const T* cr = cosp + b * cs_b + s * cs_s;
  const T* sr = sinp + b * cs_b + s * cs_s;
  const int64_t dest = load_index(select_index, tok, idx_width);
  const int half = hd / 2;
  const int pairs = half / 2;                       // half2 units per half-row
  const int rot = (n_qh + n_kh) * pairs;
  const int vcopy = n_kh * half;                    // half2 units of the token's V heads
  for (int i = threadIdx.x; i < rot + vcopy; i += blockDim.x) {
    if (i < rot) {
      const int h = i / pairs, p = i % pairs;
      const bool is_k = h >= n_qh;
      T* base = is_k ? k + tok * k_ts + (int64_t)(h - n_qh) * k_hs : q + tok * q_ts + (int64_t)h * q_hs;
      T2* p1 = reinterpret_cast<T2*>(base) + p;
      T2* p2 = reinterpret_cast<T2*>(base + half) + p;
      T2 a = *p1, bb = *p2;
      rope_pair(a, bb, reinterpret_cast<const T2*>(cr)[p], reinterpret_cast<const T2*>(sr)[p]);
      *p1 = a;
      *p2 = bb;
      if (is_k) {
        T* drow = pool + dest * d_ts + (int64_t)(h - n_qh) * d_hs;
        reinterpret_cast<T2*>(drow)[p] = a;
        reinterpret_cast<T2*>(drow + half)[p] = bb;
      }
    } else {
      const int j = i - rot;
      const int h = j / half, c = j % half;
      const T2 val = reinterpret_cast<const T2*>(v + tok * v_ts + (int64_t)h * v_hs)[c];
      reinterpret_cast<T2*>(pool + dest * d_ts + (int64_t)(n_kh + h) * d_hs)[c] = val;
    }
  }
}

extern "C" int b2l_rope_kv_write(void* q, void* k, const void* v, const void* cos, const void* sin,
                                 const void* select_index, int index_width, void* kv_buffer, int64_t batch,
                                 int64_t seq_len, int n_qh, int n_kh, int head_dim, int64_t q_tok_stride,
                                 int64_t q_head_stride, int64_t k_tok_stride, int64_t k_head_stride,
                                 int64_t v_tok_stride, int64_t v_head_stride, int64_t cs_batch_stride,
                                 int64_t cs_seq_stride, int64_t d_tok_stride, int64_t d_head_stride, int dtype,
                                 void* stream) {
  B2L_REQUIRE(q && k && v && cos && sin && select_index && kv_buffer, "rope_kv_write: null pointer");
  B2L_REQUIRE(batch >= 0 && seq_len >= 0 && n_qh > 0 && n_kh > 0 && head_dim >= 4 && head_dim % 4 == 0,
              "rope_kv_write: bad shape (head_dim must be a multiple of 4)");
  B2L_REQUIRE(index_width == 4 || index_width == 8, "rope_kv_write: index width %d", index_width);
  const int64_t strides[] = {q_tok_stride, q_head_stride, k_tok_stride, k_head_stride, v_tok_stride, v_head_stride,
                             cs_batch_stride, cs_seq_stride, d_tok_stride, d_head_stride};
  for (int64_t st : strides) B2L_REQUIRE(st % 2 == 0, "rope_kv_write: strides must be even (4-byte accesses)");
  B2L_REQUIRE(((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)cos | (uintptr_t)sin | (uintptr_t)kv_buffer) % 4 == 0,
              "rope_kv_write: pointers must be 4-byte aligned");
  const int64_t tokens = batch * seq_len;
  if (tokens == 0) return B2L_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int work = (n_qh + n_kh) * (head_dim / 4) + n_kh * (head_dim / 2);
  const int threads = work >= 256 ? 256 : (work >= 128 ? 128 : 64);
  if (dtype == B2L_F16)
    b2l_launch(rope_kv_write_kernel<__half>, dim3((unsigned)tokens), dim3(threads), (size_t)0, st, (__half*)q, (__half*)k,
               (const __half*)v, (const __half*)cos, (const __half*)sin, select_index, index_width, (__half*)kv_buffer,
               (int)seq_len, n_qh, n_kh, head_dim, q_tok_stride, q_head_stride, k_tok_stride, k_head_stride, v_tok_stride,
               v_head_stride, cs_batch_stride, cs_seq_stride, d_tok_stride, d_head_stride);
  else if (dtype == B2L_BF16)
    b2l_launch(rope_kv_write_kernel<__nv_bfloat16>, dim3((unsigned)tokens), dim3(threads), (size_t)0, st,
               (__nv_bfloat16*)q, (__nv_bfloat16*)k, (const __nv_bfloat16*)v, (const __nv_bfloat16*)cos,
               (const __nv_bfloat16*)sin, select_index, index_width, (__nv_bfloat16*)kv_buffer, (int)seq_len, n_qh, n_kh,
               head_dim, q_tok_stride, q_head_stride, k_tok_stride, k_head_stride, v_tok_stride, v_head_stride,
               cs_batch_stride, cs_seq_stride, d_tok_stride, d_head_stride);
  else
    B2L_UNSUPPORTED("rope_kv_write: dtype %d", dtype);
  B2L_LAUNCH_CHECK();
  return B2L_OK;
}

// --------------------------------------------------------------------------- //
// swiglu_forward (reference: lite_llama/kernels/swiglu.py:24-65)
//   c = (a_f32 * sigmoid(a_f32)) * b   -> activation dtype
// Also covers fused_moe's _silu_and_mul_kernel (fused_moe.py:296-313) through
// the `b` pointer offset / row strides (gate and up halves of one row).
// --------------------------------------------------------------------------- //
__device__ __forceinline__ float silu_f32(float x) {
  // sigmoid = 1 / (1 + exp(-x)) with ex2.approx + div.full, as Triton lowers tl.sigmoid
  const float e = ex2_approx(-x * 1.4426950408889634f);
  return x * div_full(1.0f, 1.0f + e);
}

template <typename T>
__global__ void swiglu_kernel(T* __restrict__ c, const T* __restrict__ a, const T* __restrict__ b, int64_t rows,
                              int64_t n, int64_t a_stride, int64_t b_stride, int64_t c_stride, int vec) {
  pdl_prologue();
  // grid.y = rows (grid-stride), grid.x * blockDim.x sweeps the row in 8-element chunks
  for (int64_t row = blockIdx.y; row < rows; row += gridDim.y) {
    const T* ar = a + row * a_stride;
    const T* br = b + row * b_stride;
    T* cr = c + row * c_stride;
    if (vec) {
      const int64_t chunks = n / 8;
      for (int64_t ch = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; ch < chunks;
           ch += (int64_t)gridDim.x * blockDim.x) {
        uint4 av = ldg_stream_128(ar + ch * 8), bv = ldg_stream_128(br + ch * 8), ov;
        const T* ae = reinterpret_cast<const T*>(&av);
        const T* be = reinterpret_cast<const T*>(&bv);
        T* oe = reinterpret_cast<T*>(&ov);
#pragma unroll
        for (int j = 0; j < 8; ++j) oe[j] = from_f32<T>(silu_f32(to_f32(ae[j])) * to_f32(be[j]));
        stg_128(cr + ch * 8, ov);
      }
    } else {
      for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        cr[i] = from_f32<T>(silu_f32(to_f32(ar[i])) * to_f32(br[i]));
    }
  }
}

extern "C" int b2l_swiglu(void* c, const void* a, const void* b, int64_t rows, int64_t n, int64_t a_row_stride,
                          int64_t b_row_stride, int64_t c_row_stride, int dtype, void* stream) {
  B2L_REQUIRE(c && a && b, "swiglu: null pointer");
  B2L_REQUIRE(rows >= 0 && n >= 0, "swiglu: bad shape");
  if (rows == 0 || n == 0) return B2L_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int vec = (n % 8 == 0) && (a_row_stride % 8 == 0) && (b_row_stride % 8 == 0) && (c_row_stride % 8 == 0) &&
                  ((uintptr_t)a % 16 == 0) && ((uintptr_t)b % 16 == 0) && ((uintptr_t)c % 16 == 0);
  const int threads = 256;
  const int64_t per_row = vec ? n / 8 : n;
  unsigned gx = (unsigned)cdiv64(per_row, threads);
  if (gx > 64) gx = 64;
  unsigned gy = (unsigned)(rows > 65535 ? 65535 : rows);
  dim3 grid(gx, gy);
  if (dtype == B2L_F16)
    b2l_launch(swiglu_kernel<__half>, dim3(grid), dim3(threads), (size_t)(0), st, (__half*)c, (const __half*)a, (const __half*)b, rows, n,
                                                   a_row_stride, b_row_stride, c_row_stride, vec);
  else if (dtype == B2L_BF16)
    b2l_launch(swiglu_kernel<__nv_bfloat16>, dim3(grid), dim3(threads), (size_t)(0), st, (__nv_bfloat16*)c, (const __nv_bfloat16*)a,
                                                          (const __nv_bfloat16*)b, rows, n, a_row_stride, b_row_stride,
                                                          c_row_stride, vec);
  else
    B2L_UNSUPPORTED("swiglu: dtype %d", dtype);
  B2L_LAUNCH_CHECK();
  return B2L_OK;
}

// --------------------------------------------------------------------------- //
// update_kv_buffer (reference: lite_llama/kernels/update_kv_buffer.py:15-89)
//   kv_buffer[select_index[t], h, :] = kv_values[t, h, :]
// Two sources (K rows, V rows) may be given so the caller needs no torch.cat
// (lite_llama/models/base.py:76): heads [0, heads_a) come from src_a, the rest from src_b.
// --------------------------------------------------------------------------- //
__global__ void update_kv_buffer_kernel(const char* __restrict__ src_a, const char* __restrict__ src_b,
                                        const void* __restrict__ select_index, int idx_width, char* __restrict__ dst,
                                        int heads_a, int heads_total, int row_bytes, int64_t a_tok_stride,
                                        int64_t a_head_stride, int64_t b_tok_stride, int64_t b_head_stride,
                                        int64_t d_tok_stride, int64_t d_head_stride, int vec) {
  pdl_prologue();
  const int64_t t = blockIdx.x;
  const int64_t dest = load_index(select_index, t, idx_width);
  const int per_head = vec ? row_bytes / 16 : row_bytes / 2;
  const int total = heads_total * per_head;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int h = i / per_head, c = i % per_head;
    const char* s = (h < heads_a) ? src_a + t * a_tok_stride + (int64_t)h * a_head_stride
                                  : src_b + t * b_tok_stride + (int64_t)(h - heads_a) * b_head_stride;
    char* d = dst + dest * d_tok_stride + (int64_t)h * d_head_stride;
    if (vec)
      reinterpret_cast<uint4*>(d)[c] = reinterpret_cast<const uint4*>(s)[c];
    else
      reinterpret_cast<uint16_t*>(d)[c] = reinterpret_cast<const uint16_t*>(s)[c];
  }
}

extern "C" int b2l_update_kv_buffer(const void* src_a, const void* src_b, const void* select_index, int index_width,
                                    void* kv_buffer, int64_t tokens, int heads_a, int heads_b, int head_dim,
                                    int64_t a_tok_stride, int64_t a_head_stride, int64_t b_tok_stride,
                                    int64_t b_head_stride, int64_t d_tok_stride, int64_t d_head_stride,
                                    void* stream) {
  // strides are in ELEMENTS of a 2-byte dtype (fp16/bf16); inner dim must be contiguous.
  B2L_REQUIRE(src_a && select_index && kv_buffer, "update_kv_buffer: null pointer");
  B2L_REQUIRE(heads_b == 0 || src_b, "update_kv_buffer: heads_b > 0 needs src_b");
  B2L_REQUIRE(index_width == 4 || index_width == 8, "update_kv_buffer: index width %d", index_width);
  if (tokens == 0) return B2L_OK;
  const int row_bytes = head_dim * 2;
  auto al = [](const void* p, int64_t s1, int64_t s2) {
    return ((uintptr_t)p % 16 == 0) && ((s1 * 2) % 16 == 0) && ((s2 * 2) % 16 == 0);
  };
  const int vec = (row_bytes % 16 == 0) && al(src_a, a_tok_stride, a_head_stride) &&
                  (heads_b == 0 || al(src_b, b_tok_stride, b_head_stride)) && al(kv_buffer, d_tok_stride, d_head_stride);
  const int heads = heads_a + heads_b;
  const int work = heads * (vec ? row_bytes / 16 : row_bytes / 2);
  const int threads = work >= 256 ? 256 : (work >= 128 ? 128 : 64);
  b2l_launch(update_kv_buffer_kernel, dim3((unsigned)tokens), dim3(threads), (size_t)(0), (cudaStream_t)stream, 
      (const char*)src_a, (const char*)src_b, select_index, index_width, (char*)kv_buffer, heads_a, heads, row_bytes,
      a_tok_stride * 2, a_head_stride * 2, b_tok_stride * 2, b_head_stride * 2, d_tok_stride * 2, d_head_stride * 2,
      vec);
  B2L_LAUNCH_CHECK();
  return B2L_OK;
}

// --------------------------------------------------------------------------- //
// update_kv_index (reference: lite_llama/kernels/update_kv_index.py:15-88)
//   table[b_req_idx[i], b_seq_len[i]-1] = select_index[i]
// --------------------------------------------------------------------------- //
__global__ void update_kv_index_kernel(int32_t* __restrict__ table, const void* __restrict__ req_idx, int req_w,
                                       const void* __restrict__ seq_len, int len_w, const void* __restrict__ sel,
                                       int sel_w, int64_t n, int64_t stride_b, int64_t stride_s) {
  pdl_prologue();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t r = load_index(req_idx, i, req_w);
  const int64_t l = load_index(seq_len, i, len_w);
  table[r * stride_b + (l - 1) * stride_s] = (int32_t)load_index(sel, i, sel_w);
}

extern "C" int b2l_update_kv_index(void* table, const void* b_req_idx, int req_width, const void* b_seq_len,
                                   int len_width, const void* select_index, int sel_width, int64_t n,
                                   int64_t stride_b, int64_t stride_s, void* stream) {
  B2L_REQUIRE(table && b_req_idx && b_seq_len && select_index, "update_kv_index: null pointer");
  B2L_REQUIRE((req_width == 4 || req_width == 8) && (len_width == 4 || len_width == 8) &&
                  (sel_width == 4 || sel_width == 8), "update_kv_index: index widths must be 4 or 8");
  if (n == 0) return B2L_OK;
  b2l_launch(update_kv_index_kernel, dim3((unsigned)cdiv64(n, 128)), dim3(128), (size_t)(0), (cudaStream_t)stream, 
      (int32_t*)table, b_req_idx, req_width, b_seq_len, len_width, select_index, sel_width, n, stride_b, stride_s);
  B2L_LAUNCH_CHECK();
  return B2L_OK;
}

// --------------------------------------------------------------------------- //
// Per-token dynamic int8 activation quantisation
// (reference: lite_llama/kernels/quantization/w8a8.py:35-68)
//   scale = absmax/127 (1.0 if 0);  q = trunc_toward_zero(x / scale)   [div.full.f32 + cvt.rzi]
// One CTA per row, row kept in registers (single HBM pass instead of the reference's two).
// --------------------------------------------------------------------------- //
template <typename T>
__global__ void __launch_bounds__(256) quantize_act_int8_kernel(int8_t* __restrict__ q, float* __restrict__ scales,
                                                                const T* __restrict__ x, int64_t k, int64_t x_stride,
                                                                int64_t q_stride) {
  pdl_prologue();
  const int64_t row = blockIdx.x;
  const T* xr = x + row * x_stride;
  int8_t* qr = q + row * q_stride;
  float amax = 0.f;
  for (int64_t i = threadIdx.x; i < k; i += blockDim.x) amax = fmaxf(amax, fabsf(to_f32(xr[i])));
  __shared__ float red[8];
  amax = warp_reduce_max(amax);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = amax;
  __syncthreads();
  float m = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) m = fmaxf(m, red[i]);
  float scale = div_full(m, 127.0f);
  scale = scale > 0.f ? scale : 1.0f;
  for (int64_t i = threadIdx.x; i < k; i += blockDim.x) {
    const float v = div_full(to_f32(xr[i]), scale);
    int32_t iv;
    // Triton's fp32 -> int8 cast: cvt.rzi (truncate toward zero), then narrow.
    asm("cvt.rzi.s32.f32 %0, %1;" : "=r"(iv) : "f"(v));
    qr[i] = (int8_t)iv;
  }
  if (threadIdx.x == 0) scales[row] = scale;
}

extern "C" int b2l_quantize_activations_int8(void* q, void* scales, const void* x, int64_t m, int64_t k,
                                             int64_t x_row_stride, int64_t q_row_stride, int dtype, void* stream) {
  B2L_REQUIRE(q && scales && x, "quantize_activations_int8: null pointer");
  if (m == 0) return B2L_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int threads = k >= 2048 ? 256 : 128;
  if (dtype == B2L_F16)
    b2l_launch(quantize_act_int8_kernel<__half>, dim3((unsigned)m), dim3(threads), (size_t)(0), st, (int8_t*)q, (float*)scales, (const __half*)x, k,
                                                                     x_row_stride, q_row_stride);
  else if (dtype == B2L_BF16)
    b2l_launch(quantize_act_int8_kernel<__nv_bfloat16>, dim3((unsigned)m), dim3(threads), (size_t)(0), st, 
        (int8_t*)q, (float*)scales, (const __nv_bfloat16*)x, k, x_row_stride, q_row_stride);
  else
    B2L_UNSUPPORTED("quantize_activations_int8: dtype %d", dtype);
  B2L_LAUNCH_CHECK();
  return B2L_OK;
}

// --------------------------------------------------------------------------- //
// skip_rmsnorm + per-token int8 quantisation in one launch (SmoothQuant decoder: the norm output feeds exactly one
// int8 GEMM -- models/base.py:307-312 -> quantization/methods/w8a8.py -> kernels/quantization/w8a8.py:35-68).
// Same arithmetic as skip_rmsnorm_block_kernel followed by quantize_act_int8_kernel on its fp16 output (the rounded
// norm output is what gets quantised), so q / scales are bit-identical to the two-launch path; y is never written.
// --------------------------------------------------------------------------- //
template <typename T, int VPT_MAX>
__global__ void __launch_bounds__(512) skip_rmsnorm_quant_kernel(int8_t* __restrict__ q, float* __restrict__ scales,
                                                                 const T* __restrict__ x, T* __restrict__ res,
                                                                 const T* __restrict__ w, int64_t n, int64_t x_stride,
                                                                 int64_t r_stride, int64_t q_stride, float eps) {
  pdl_prologue();
  const int64_t row = blockIdx.x;
  const T* xr = x + row * x_stride;
  T* rr = res ? res + row * r_stride : nullptr;
  int8_t* qr = q + row * q_stride;
  const int nchunks = (int)(n / 8);
  float vals[VPT_MAX][8];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < VPT_MAX; ++i) {
    const int c = threadIdx.x + i * blockDim.x;
    if (c < nchunks) {
      uint4 xv = ldg_128(xr + c * 8);
      const T* xe = reinterpret_cast<const T*>(&xv);
      if (rr) {
        uint4 rv = ldg_128(rr + c * 8);
        const T* re = reinterpret_cast<const T*>(&rv);
        uint4 ov;
        T* oe = reinterpret_cast<T*>(&ov);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float s = to_f32(xe[j]) + to_f32(re[j]);
          vals[i][j] = s;
          oe[j] = from_f32<T>(s);
        }
        stg_128(rr + c * 8, ov);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) vals[i][j] = to_f32(xe[j]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) ss += vals[i][j] * vals[i][j];
    }
  }
  __shared__ float red[16];
  __shared__ float red_max[16];
  ss = warp_reduce_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
  const int nw = blockDim.x >> 5;
  for (int i = 0; i < nw; ++i) tot += red[i];
  const float rrms = div_full(1.0f, sqrt_approx(tot / (float)n + eps));
  // the norm output, rounded to the activation dtype exactly like the stand-alone kernel, stays in registers
  float amax = 0.f;
#pragma unroll
  for (int i = 0; i < VPT_MAX; ++i) {
    const int c = threadIdx.x + i * blockDim.x;
    if (c < nchunks) {
      uint4 wv = ldg_128(w + c * 8);
      const T* we = reinterpret_cast<const T*>(&wv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        T nrm = from_f32<T>(vals[i][j] * rrms);
        const float out = to_f32(from_f32<T>(to_f32(nrm) * to_f32(we[j])));
        vals[i][j] = out;
        amax = fmaxf(amax, fabsf(out));
      }
    }
  }
  amax = warp_reduce_max(amax);
  if ((threadIdx.x & 31) == 0) red_max[threadIdx.x >> 5] = amax;
  __syncthreads();
  float m = 0.f;
  for (int i = 0; i < nw; ++i) m = fmaxf(m, red_max[i]);
  float scale = div_full(m, 127.0f);
  scale = scale > 0.f ? scale : 1.0f;
#pragma unroll
  for (int i = 0; i < VPT_MAX; ++i) {
    const int c = threadIdx.x + i * blockDim.x;
    if (c < nchunks) {
      uint32_t lo = 0u, hi = 0u;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        int32_t iv;
        asm("cvt.rzi.s32.f32 %0, %1;" : "=r"(iv) : "f"(div_full(vals[i][j], scale)));   // Triton's fp32 -> int8 cast
        const uint32_t b = (uint32_t)(uint8_t)(int8_t)iv;
        if (j < 4) lo |= b << (8 * j); else hi |= b << (8 * (j - 4));
      }
      *reinterpret_cast<uint2*>(qr + c * 8) = make_uint2(lo, hi);
    }
  }
  if (threadIdx.x == 0) scales[row] = scale;
}

template <typename T>
static int launch_skip_rmsnorm_quant(void* q, void* scales, const void* x, void* res, const void* w, int64_t rows,
                                     int64_t n, int64_t xs, int64_t rs, int64_t qs, float eps, cudaStream_t st) {
  if (rows == 0) return B2L_OK;
  const bool aligned = (n % 8 == 0) && (xs % 8 == 0) && (res == nullptr || rs % 8 == 0) && (qs % 8 == 0) &&
                       ((uintptr_t)x % 16 == 0) && ((uintptr_t)w % 16 == 0) && (res == nullptr || (uintptr_t)res % 16 == 0) &&
                       ((uintptr_t)q % 8 == 0);
  B2L_REQUIRE(aligned && n / 8 <= 8 * 256, "skip_rmsnorm_quant_int8: needs n %% 8 == 0, n <= 16384 and 16-byte aligned rows (n=%lld)",
              (long long)n);
  int threads = 256;
  const int64_t chunks = n / 8;
  if (chunks <= 128) threads = 128;
  else if (chunks <= 512 && chunks > 256) threads = 512;
  const int vpt = (int)cdiv64(chunks, threads);
  int8_t* qq = (int8_t*)q; float* ss = (float*)scales; const T* xx = (const T*)x; T* rr = (T*)res; const T* ww = (const T*)w;
  if (vpt <= 1)
    b2l_launch(skip_rmsnorm_quant_kernel<T, 1>, dim3((unsigned)rows), dim3(threads), (size_t)(0), st, qq, ss, xx, rr, ww, n, xs, rs, qs, eps);
  else if (vpt <= 2)
    b2l_launch(skip_rmsnorm_quant_kernel<T, 2>, dim3((unsigned)rows), dim3(threads), (size_t)(0), st, qq, ss, xx, rr, ww, n, xs, rs, qs, eps);
  else if (vpt <= 4)
    b2l_launch(skip_rmsnorm_quant_kernel<T, 4>, dim3((unsigned)rows), dim3(threads), (size_t)(0), st, qq, ss, xx, rr, ww, n, xs, rs, qs, eps);
  else
    b2l_launch(skip_rmsnorm_quant_kernel<T, 8>, dim3((unsigned)rows), dim3(threads), (size_t)(0), st, qq, ss, xx, rr, ww, n, xs, rs, qs, eps);
  B2L_LAUNCH_CHECK();
  return B2L_OK;
}

extern "C" int b2l_skip_rmsnorm_quant_int8(void* q, void* scales, const void* x, void* residual, const void* weight,
                                           int64_t rows, int64_t n, int64_t x_row_stride, int64_t r_row_stride,
                                           int64_t q_row_stride, float eps, int dtype, void* stream) {
  B2L_REQUIRE(q && scales && x && weight, "skip_rmsnorm_quant_int8: null pointer");
  B2L_REQUIRE(rows >= 0 && n > 0, "skip_rmsnorm_quant_int8: bad shape rows=%lld n=%lld", (long long)rows, (long long)n);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == B2L_F16)
    return launch_skip_rmsnorm_quant<__half>(q, scales, x, residual, weight, rows, n, x_row_stride, r_row_stride,
                                             q_row_stride, eps, st);
  if (dtype == B2L_BF16)
    return launch_skip_rmsnorm_quant<__nv_bfloat16>(q, scales, x, residual, weight, rows, n, x_row_stride, r_row_stride,
                                                    q_row_stride, eps, st);
  B2L_UNSUPPORTED("skip_rmsnorm_quant_int8: dtype %d", dtype);
}

// --------------------------------------------------------------------------- //
// moe_sum (reference: lite_llama/kernels/fused_moe.py:316-333)
//   out[m, :] = sum_k in[m*top_k + k, :]   (fp32 accumulate, in slot order)
// --------------------------------------------------------------------------- //
template <typename T>
__global__ void moe_sum_kernel(T* __restrict__ out, const T* __restrict__ in, int64_t tokens, int64_t n, int top_k) {
  pdl_prologue();
  // grid.y is capped at 65535: tokens are swept with a grid stride (MoE prefill can exceed the cap)
  for (int64_t m = blockIdx.y; m < tokens; m += gridDim.y) {
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (int64_t)gridDim.x * blockDim.x) {
      float acc = 0.f;
      const T* base = in + m * top_k * n + c;
      for (int i = 0; i < top_k; ++i) acc += to_f32(base[(int64_t)i * n]);
      out[m * n + c] = from_f32<T>(acc);
    }
  }
}

extern "C" int b2l_moe_sum(void* out, const void* in, int64_t tokens, int64_t n, int top_k, int dtype, void* stream) {
  B2L_REQUIRE(out && in, "moe_sum: null pointer");
  if (tokens == 0 || n == 0) return B2L_OK;
  dim3 grid((unsigned)cdiv64(n, 256), (unsigned)(tokens > 65535 ? 65535 : tokens));
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == B2L_F16)
    b2l_launch(moe_sum_kernel<__half>, dim3(grid), dim3(256), (size_t)(0), st, (__half*)out, (const __half*)in, tokens, n, top_k);
  else if (dtype == B2L_BF16)
    b2l_launch(moe_sum_kernel<__nv_bfloat16>, dim3(grid), dim3(256), (size_t)(0), st, (__nv_bfloat16*)out, (const __nv_bfloat16*)in, tokens, n, top_k);
  else
    B2L_UNSUPPORTED("moe_sum: dtype %d", dtype);
  B2L_LAUNCH_CHECK();
  return B2L_OK;
}


// --------------------------------------------------------------------------- //
// greedy sampling: argmax over the vocabulary, one CTA per row (reference: Sampler's temperature-0 path,
// lite_llama/engine/sampler.py -> torch.argmax on the [batch, vocab] logits, engine/llm_engine.py:172).
// torch's generic reduce kernel takes ~45 us for 32 x 128256 fp16 (one 32-thread-wide CTA per row); this streams each
// row with 1024 threads and 128-bit loads.  Ties resolve to the LOWEST index (first occurrence, like torch.argmax).
// --------------------------------------------------------------------------- //
template <typename T>
__global__ void __launch_bounds__(1024) argmax_rows_kernel(int64_t* __restrict__ out, const T* __restrict__ x, int64_t n,
                                                           int64_t row_stride, int vec) {
  pdl_prologue();
  const T* xr = x + (int64_t)blockIdx.x * row_stride;
  float best = -INFINITY;
  int64_t best_i = 0x7fffffffffffffffll;
  auto consider = [&](float v, int64_t i) {
    if (v > best || (v == best && i < best_i) || (v != v && best == best)) { best = v; best_i = i; }   // NaN wins, like torch
  };
  if (vec) {
    const int64_t chunks = n / 8;
    for (int64_t c = threadIdx.x; c < chunks; c += blockDim.x) {
      const uint4 v = ldg_stream_128(xr + c * 8);
      const T* e = reinterpret_cast<const T*>(&v);
#pragma unroll
      for (int j = 0; j < 8; ++j) consider(to_f32(e[j]), c * 8 + j);
    }
    for (int64_t i = chunks * 8 + threadIdx.x; i < n; i += blockDim.x) consider(to_f32(xr[i]), i);
  } else {
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) consider(to_f32(xr[i]), i);
  }
  // warp, then block reduction of (value, index) with the same ordering
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int64_t oi = __shfl_xor_sync(0xffffffffu, best_i, o);
    consider(ov, oi);
  }
  __shared__ float sv[32];
  __shared__ int64_t si[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sv[warp] = best; si[warp] = best_i; }
  __syncthreads();
  if (warp == 0) {
    const int nw = blockDim.x >> 5;
    best = lane < nw ? sv[lane] : -INFINITY;
    best_i = lane < nw ? si[lane] : 0x7fffffffffffffffll;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int64_t oi = __shfl_xor_sync(0xffffffffu, best_i, o);
      consider(ov, oi);
    }
    if (lane == 0) out[blockIdx.x] = best_i == 0x7fffffffffffffffll ? 0 : best_i;
  }
}

extern "C" int b2l_argmax_rows(void* out_index_i64, const void* x, int64_t rows, int64_t n, int64_t row_stride, int dtype,
                               void* stream) {
  B2L_REQUIRE(out_index_i64 && x, "argmax_rows: null pointer");
  B2L_REQUIRE(rows >= 0 && n > 0, "argmax_rows: bad shape rows=%lld n=%lld", (long long)rows, (long long)n);
  if (rows == 0) return B2L_OK;
  const int vec = (row_stride % 8 == 0) && ((uintptr_t)x % 16 == 0);
  cudaStream_t st = (cudaStream_t)stream;
  const int threads = n >= 8192 ? 1024 : 256;
  if (dtype == B2L_F16)
    b2l_launch(argmax_rows_kernel<__half>, dim3((unsigned)rows), dim3(threads), (size_t)0, st, (int64_t*)out_index_i64,
               (const __half*)x, n, row_stride, vec);
  else if (dtype == B2L_BF16)
    b2l_launch(argmax_rows_kernel<__nv_bfloat16>, dim3((unsigned)rows), dim3(threads), (size_t)0, st,
               (int64_t*)out_index_i64, (const __nv_bfloat16*)x, n, row_stride, vec);
  else
    B2L_UNSUPPORTED("argmax_rows: dtype %d", dtype);
  B2L_LAUNCH_CHECK();
  return B2L_OK;
}
