// Tensor-parallel exchange over NVLink peer memory, fused with the op that follows it.
//
// Reference: every row-parallel projection (o_proj, down_proj, the MoE block) ends in an in-place NCCL SUM
// all-reduce of the [tokens, hidden] partial (lite_llama/models/linear.py:160-161,
// lite_llama/distributed/parallel_state.py:208-213, lite_llama/models/qwen3_moe.py:110), and every one of those
// all-reduces is immediately followed by the fused residual-add + RMSNorm of the next block
// (lite_llama/models/base.py:307,312,486; lite_llama/kernels/skip_rmsnorm.py:192-234).  At decode the message is
// 8 KiB per token -- 256 KiB at batch 32 -- so the collective is pure latency: a ring/tree with
// several hops and a separate kernel for the norm costs more than the bytes.
//
// Here the two are ONE kernel over peer memory (one process per GPU, buffers exchanged once through CUDA IPC):
//   push    each CTA owns one token row: it stores its rank's partial row straight into slot [rank] of every
//           peer's receive buffer (posted NVLink writes, no round trip), fences, and raises a per-row flag on the peer;
//   gather  it waits for the same-row flags of all peers, then sums the world_size partials from its OWN receive
//           buffer in rank order (fp32; every rank adds in the same order -> bit-identical results on all ranks),
//           rounds to the activation dtype exactly where the reference's all-reduce output would be,
//   norm    and runs skip_rmsnorm on it (same arithmetic and rounding points as elementwise.cu) while the row is
//           still in registers: residual += sum; y = round(residual_f32 * rrms) * w.
// Receive buffers and flags are double-buffered by an epoch each rank keeps per row in its own memory, so the
// kernel needs no trailing barrier and is CUDA-graph capturable (no host state changes between replays): a peer can
// only reach epoch e+2 (which reuses the parity of e) after this rank raised its flag for e+1, i.e. after it finished
// reading epoch e.
#include "common.cuh"

namespace {

constexpr int kMaxWorld = 8;
constexpr int kThreads = 256;
constexpr int kMaxChunksPerThread = 4;      // 16-byte chunks per thread: rows up to 256*4*8 = 8192 elements

struct CommState {
  int rank, world;
  int max_rows;
  int64_t row_bytes;                         // capacity of one row slot
  unsigned char* local;                      // this rank's allocation
  unsigned char* peer[kMaxWorld];            // peer[r] = mapped base of rank r's allocation (peer[rank] = local)
  cudaIpcMemHandle_t handle;
  bool opened;
};

// layout of one rank's allocation:
//   [recv: 2 parities][world slots][max_rows][row_bytes]   data pushed by the peers
//   [flags: 2 parities][max_rows][kMaxWorld] uint32        raised by the peers
//   [epoch: max_rows] uint32                               private to the owner
__host__ __device__ inline int64_t recv_bytes(int world, int max_rows, int64_t row_bytes) {
  return 2ll * world * max_rows * row_bytes;
}
__host__ __device__ inline int64_t flags_bytes(int max_rows) { return 2ll * max_rows * kMaxWorld * 4; }
__host__ __device__ inline int64_t total_bytes(int world, int max_rows, int64_t row_bytes) {
  return recv_bytes(world, max_rows, row_bytes) + flags_bytes(max_rows) + (int64_t)max_rows * 4;
}

struct FusedParams {
  unsigned char* peer[kMaxWorld];
  int rank, world, max_rows;
  int64_t row_bytes;
  void* y; int64_t y_stride;
  const void* x; int64_t x_stride;           // this rank's partial [rows, n]
  void* res; int64_t r_stride;               // residual (in place), nullable
  const void* w;                             // norm weight [n], null = plain all-reduce (y = sum)
  int n;
  float eps;
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float sqrt_approx_f(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float div_full_f(float a, float b) {
  float r;
  asm("div.full.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}

template <typename T>
__global__ void __launch_bounds__(kThreads) tp_allreduce_rmsnorm_kernel(const FusedParams p) {
  pdl_prologue();                              // the partial rows are the previous kernel's output
  const int row = blockIdx.x;
  const int tid = threadIdx.x;
  const int nchunks = p.n / 8;
  unsigned char* mine_base = p.peer[p.rank];
  const int64_t rbytes = recv_bytes(p.world, p.max_rows, p.row_bytes);
  unsigned int* my_flags = reinterpret_cast<unsigned int*>(mine_base + rbytes);
  unsigned int* my_epoch = reinterpret_cast<unsigned int*>(mine_base + rbytes + flags_bytes(p.max_rows));
  const unsigned int e = my_epoch[row] + 1u;   // same value on every rank: all ranks run the same call sequence
  const int par = (int)(e & 1u);

  // ---- push this rank's partial row into slot [rank] of every peer ----
  const T* xr = reinterpret_cast<const T*>(p.x) + (int64_t)row * p.x_stride;
  uint4 mine[kMaxChunksPerThread];
#pragma unroll
  for (int i = 0; i < kMaxChunksPerThread; ++i) {
    const int c = tid + i * kThreads;
    mine[i] = make_uint4(0u, 0u, 0u, 0u);
    if (c < nchunks) mine[i] = ldg_128(xr + c * 8);
  }
  const int64_t slot_off = (((int64_t)par * p.world + p.rank) * p.max_rows + row) * p.row_bytes;
  for (int d = 1; d < p.world; ++d) {
    const int peer = (p.rank + d) % p.world;   // staggered so that the ranks do not all hit the same peer first
    unsigned char* dst = p.peer[peer] + slot_off;
#pragma unroll
    for (int i = 0; i < kMaxChunksPerThread; ++i) {
      const int c = tid + i * kThreads;
      if (c < nchunks) *reinterpret_cast<uint4*>(dst + (int64_t)c * 16) = mine[i];
    }
  }
  __syncthreads();                             // every thread's stores precede the flag writers' fence
  if (tid < p.world && tid != p.rank) {
    __threadfence_system();
    unsigned int* pf = reinterpret_cast<unsigned int*>(p.peer[tid] + rbytes) +
                       ((int64_t)par * p.max_rows + row) * kMaxWorld + p.rank;
    st_release_sys(pf, e);
    // ---- wait for that peer's row ----
    const unsigned int* wf = my_flags + ((int64_t)par * p.max_rows + row) * kMaxWorld + tid;
    if (ld_acquire_sys(wf) != e) {
      const unsigned long long t0 = b2l_globaltimer();
      while (ld_acquire_sys(wf) != e) {
        const unsigned long long lim = b2l_hang_ctl.limit_ns;
        const unsigned long long dt = b2l_globaltimer() - t0;
        if (lim != 0ull && dt > lim) b2l_hang_trap((uint32_t)tid, e, b2l_file_hash(__FILE__), __LINE__, dt);
      }
    }
  }
  __syncthreads();

  // ---- sum the partials in rank order (fp32), round like the reference's fp16 all-reduce output ----
  float vals[kMaxChunksPerThread][8];
  float ss = 0.f;
  T* rr = p.res ? reinterpret_cast<T*>(p.res) + (int64_t)row * p.r_stride : nullptr;
  T* yr = reinterpret_cast<T*>(p.y) + (int64_t)row * p.y_stride;
#pragma unroll
  for (int i = 0; i < kMaxChunksPerThread; ++i) {
    const int c = tid + i * kThreads;
    if (c < nchunks) {
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
      for (int r = 0; r < p.world; ++r) {
        uint4 v;
        if (r == p.rank) {
          v = mine[i];
        } else {
          const unsigned char* src = mine_base + (((int64_t)par * p.world + r) * p.max_rows + row) * p.row_bytes;
          v = __ldcg(reinterpret_cast<const uint4*>(src + (int64_t)c * 16));
        }
        const T* ve = reinterpret_cast<const T*>(&v);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += to_f32(ve[j]);
      }
      uint4 rv = make_uint4(0u, 0u, 0u, 0u);
      if (rr) rv = ldg_128(rr + c * 8);
      const T* re = reinterpret_cast<const T*>(&rv);
      uint4 ov;
      T* oe = reinterpret_cast<T*>(&ov);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float o16 = to_f32(from_f32<T>(acc[j]));        // the all-reduce result in the activation dtype
        const float s = rr ? o16 + to_f32(re[j]) : o16;        // skip_rmsnorm.py: r += x
        vals[i][j] = s;
        oe[j] = from_f32<T>(s);
        ss += s * s;
      }
      if (rr) stg_128(rr + c * 8, ov);
      else if (!p.w) stg_128(yr + c * 8, ov);                  // plain all-reduce
    }
  }
  if (tid == 0) my_epoch[row] = e;
  if (!p.w) return;

  // ---- RMSNorm of the row (same arithmetic as skip_rmsnorm_block_kernel) ----
  __shared__ float red[kThreads / 32];
  ss = warp_reduce_sum(ss);
  if ((tid & 31) == 0) red[tid >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < kThreads / 32; ++i) tot += red[i];
  const float rrms = div_full_f(1.0f, sqrt_approx_f(tot / (float)p.n + p.eps));
  const T* w = reinterpret_cast<const T*>(p.w);
#pragma unroll
  for (int i = 0; i < kMaxChunksPerThread; ++i) {
    const int c = tid + i * kThreads;
    if (c < nchunks) {
      uint4 wv = ldg_128(w + c * 8);
      const T* we = reinterpret_cast<const T*>(&wv);
      uint4 ov;
      T* oe = reinterpret_cast<T*>(&ov);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        T nrm = from_f32<T>(vals[i][j] * rrms);
        oe[j] = from_f32<T>(to_f32(nrm) * to_f32(we[j]));
      }
      stg_128(yr + c * 8, ov);
    }
  }
}

}  // namespace

// ---- host side: one communicator per process (one process per GPU) ----
extern "C" int64_t b2l_tp_comm_create(int rank, int world, int max_rows, int64_t max_row_elems) {
  if (world < 2 || world > kMaxWorld || rank < 0 || rank >= world || max_rows <= 0 || max_row_elems <= 0 ||
      max_row_elems % 8 != 0 || max_row_elems > (int64_t)kThreads * kMaxChunksPerThread * 8) {
    b2l_set_error("tp_comm_create: bad arguments (world %d, rank %d, rows %d, row elems %lld; world <= %d, row elems <= %d)",
                  world, rank, max_rows, (long long)max_row_elems, kMaxWorld, kThreads * kMaxChunksPerThread * 8);
    return 0;
  }
  CommState* cs = new CommState();
  memset(cs, 0, sizeof(*cs));
  cs->rank = rank; cs->world = world; cs->max_rows = max_rows; cs->row_bytes = max_row_elems * 2;
  const int64_t bytes = total_bytes(world, max_rows, cs->row_bytes);
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, (size_t)bytes);
  if (e == cudaSuccess) e = cudaMemset(p, 0, (size_t)bytes);
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&cs->handle, p);
  if (e != cudaSuccess) {
    b2l_set_error("tp_comm_create: %s", cudaGetErrorString(e));
    if (p) cudaFree(p);
    delete cs;
    return 0;
  }
  cs->local = (unsigned char*)p;
  cs->peer[rank] = cs->local;
  return (int64_t)(intptr_t)cs;
}

extern "C" int b2l_tp_comm_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }

extern "C" int b2l_tp_comm_get_handle(int64_t comm, void* out, int64_t out_bytes) {
  CommState* cs = (CommState*)(intptr_t)comm;
  B2L_REQUIRE(cs && out && out_bytes >= (int64_t)sizeof(cudaIpcMemHandle_t), "tp_comm_get_handle: bad arguments");
  memcpy(out, &cs->handle, sizeof(cudaIpcMemHandle_t));
  return B2L_OK;
}

// handles: world * handle_bytes, rank-major (the all-gathered result of b2l_tp_comm_get_handle)
extern "C" int b2l_tp_comm_open(int64_t comm, const void* handles, int64_t handles_bytes) {
  CommState* cs = (CommState*)(intptr_t)comm;
  B2L_REQUIRE(cs && handles && handles_bytes >= (int64_t)cs->world * (int64_t)sizeof(cudaIpcMemHandle_t),
              "tp_comm_open: bad arguments");
  for (int r = 0; r < cs->world; ++r) {
    if (r == cs->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles + (size_t)r * sizeof(h), sizeof(h));
    void* p = nullptr;
    B2L_CUDA_CHECK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    cs->peer[r] = (unsigned char*)p;
  }
  cs->opened = true;
  return B2L_OK;
}

extern "C" int b2l_tp_comm_destroy(int64_t comm) {
  CommState* cs = (CommState*)(intptr_t)comm;
  if (!cs) return B2L_OK;
  for (int r = 0; r < cs->world; ++r)
    if (r != cs->rank && cs->peer[r]) cudaIpcCloseMemHandle(cs->peer[r]);
  if (cs->local) cudaFree(cs->local);
  delete cs;
  return B2L_OK;
}

// Fused all-reduce(SUM over ranks) + residual add + RMSNorm of [rows, n] partials; weight == null gives the plain
// all-reduce (y = sum, residual ignored).  Every rank of the communicator must make the same sequence of calls.
extern "C" int b2l_tp_allreduce_skip_rmsnorm(int64_t comm, void* y, const void* x_partial, void* residual,
                                             const void* weight, int64_t rows, int64_t n, int64_t x_row_stride,
                                             int64_t r_row_stride, int64_t y_row_stride, float eps, int dtype,
                                             void* stream) {
  CommState* cs = (CommState*)(intptr_t)comm;
  B2L_REQUIRE(cs && cs->opened, "tp_allreduce_skip_rmsnorm: communicator not opened");
  B2L_REQUIRE(y && x_partial, "tp_allreduce_skip_rmsnorm: null pointer");
  B2L_REQUIRE(rows >= 0 && rows <= cs->max_rows, "tp_allreduce_skip_rmsnorm: %lld rows > communicator capacity %d",
              (long long)rows, cs->max_rows);
  B2L_REQUIRE(n > 0 && n % 8 == 0 && n * 2 <= cs->row_bytes, "tp_allreduce_skip_rmsnorm: row of %lld elements (capacity %lld, multiple of 8)",
              (long long)n, (long long)(cs->row_bytes / 2));
  B2L_REQUIRE(x_row_stride % 8 == 0 && y_row_stride % 8 == 0 && (residual == nullptr || r_row_stride % 8 == 0) &&
                  (uintptr_t)x_partial % 16 == 0 && (uintptr_t)y % 16 == 0 && (uintptr_t)residual % 16 == 0 &&
                  (uintptr_t)weight % 16 == 0,
              "tp_allreduce_skip_rmsnorm: rows must be 16-byte aligned");
  B2L_REQUIRE(dtype == B2L_F16 || dtype == B2L_BF16, "tp_allreduce_skip_rmsnorm: dtype %d", dtype);
  if (rows == 0) return B2L_OK;
  FusedParams p;
  for (int r = 0; r < kMaxWorld; ++r) p.peer[r] = r < cs->world ? cs->peer[r] : nullptr;
  p.rank = cs->rank; p.world = cs->world; p.max_rows = cs->max_rows; p.row_bytes = cs->row_bytes;
  p.y = y; p.y_stride = y_row_stride; p.x = x_partial; p.x_stride = x_row_stride;
  p.res = weight ? residual : nullptr; p.r_stride = r_row_stride; p.w = weight; p.n = (int)n; p.eps = eps;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == B2L_F16)
    b2l_launch(tp_allreduce_rmsnorm_kernel<__half>, dim3((unsigned)rows), dim3(kThreads), 0, st, p);
  else
    b2l_launch(tp_allreduce_rmsnorm_kernel<__nv_bfloat16>, dim3((unsigned)rows), dim3(kThreads), 0, st, p);
  B2L_LAUNCH_CHECK();
  return B2L_OK;
}
