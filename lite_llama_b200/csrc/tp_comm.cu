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
//           peer's receive buffer as 16-byte packets {6 values, epoch} (posted NVLink writes: no round trip, no fence,
//           no separate flag -- every packet carries its own validity word);
//   gather  it polls the same-row packets of all peers in its OWN receive buffer and sums the world_size partials
//           in rank order (fp32; every rank adds in the same order -> bit-identical results on all ranks),
//           rounds to the activation dtype exactly where the reference's all-reduce output would be,
//   norm    and runs skip_rmsnorm on it (same arithmetic and rounding points as elementwise.cu) while the row is
//           still in registers: residual += sum; y = round(residual_f32 * rrms) * w.
// Receive buffers are double-buffered by an epoch each rank keeps per row in its own memory, so the kernel needs
// no trailing barrier and is CUDA-graph capturable (no host state changes between replays): a peer can only reach
// epoch e+2 (which reuses the parity of e) after it received this rank's packets of e+1, which this rank sends after
// it finished reading epoch e.
#include "common.cuh"

namespace {

constexpr int kMaxWorld = 8;
constexpr int kThreads = 768;                 // one packet per thread for hidden 4096: every poll of a row is in flight at once
constexpr int kMaxIters = 2;                 // packets per thread: rows up to 768 * 2 * 6 = 9216 elements
constexpr int kPktElems = 6;                 // a 16-byte packet = 6 activation values (3 words) + the epoch word

struct CommState {
  int rank, world;
  int max_rows;
  int64_t row_bytes;                         // capacity of one row slot (packets * 16)
  int64_t max_row_elems;
  unsigned char* local;                      // this rank's allocation
  unsigned char* peer[kMaxWorld];            // peer[r] = mapped base of rank r's allocation (peer[rank] = local)
  cudaIpcMemHandle_t handle;
  bool opened;
};

// layout of one rank's allocation:
//   [recv: 2 parities][world slots][max_rows][row_bytes]   packets pushed by the peers
//   [epoch: max_rows] uint32                               private to the owner
__host__ __device__ inline int64_t recv_bytes(int world, int max_rows, int64_t row_bytes) {
  return 2ll * world * max_rows * row_bytes;
}
__host__ __device__ inline int64_t total_bytes(int world, int max_rows, int64_t row_bytes) {
  return recv_bytes(world, max_rows, row_bytes) + (int64_t)max_rows * 4;
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

// Flag-in-data packets (the idea of NCCL's LL protocols): every 16-byte store carries its own validity word, so the
// sender needs no fence and no separate flag write, and the receiver sees a packet one NVLink traversal after it was
// stored.  A naturally aligned 16-byte store of one thread is delivered as a unit.
__device__ __forceinline__ void st_pkt(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t flag) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(flag) : "memory");
}
__device__ __forceinline__ uint4 ld_pkt(const void* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
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
template <typename T> __device__ __forceinline__ void unpack2(uint32_t w, float& a, float& b) {
  const T* e = reinterpret_cast<const T*>(&w);
  a = to_f32(e[0]); b = to_f32(e[1]);
}
template <typename T> __device__ __forceinline__ uint32_t pack2t(float a, float b) {
  uint32_t w;
  T* e = reinterpret_cast<T*>(&w);
  e[0] = from_f32<T>(a); e[1] = from_f32<T>(b);
  return w;
}

template <typename T>
__global__ void __launch_bounds__(kThreads) tp_allreduce_rmsnorm_kernel(const FusedParams p) {
  pdl_prologue();                              // the partial rows are the previous kernel's output
  const int row = blockIdx.x;
  const int tid = threadIdx.x;
  const int nwords = p.n / 2;                  // 32-bit words per row
  const int npkts = (nwords + 2) / 3;
  unsigned char* mine_base = p.peer[p.rank];
  const int64_t rbytes = recv_bytes(p.world, p.max_rows, p.row_bytes);
  unsigned int* my_epoch = reinterpret_cast<unsigned int*>(mine_base + rbytes);
  const unsigned int e = my_epoch[row] + 1u;   // same value on every rank: all ranks run the same call sequence
  const int par = (int)(e & 1u);

  // ---- push this rank's partial row, packet by packet, into slot [rank] of every peer ----
  const uint32_t* xr = reinterpret_cast<const uint32_t*>(reinterpret_cast<const T*>(p.x) + (int64_t)row * p.x_stride);
  uint32_t mine[kMaxIters][3];
#pragma unroll
  for (int i = 0; i < kMaxIters; ++i) {
    const int c = tid + i * kThreads;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int wi = 3 * c + k;
      mine[i][k] = (c < npkts && wi < nwords) ? __ldg(xr + wi) : 0u;
    }
  }
  const int64_t slot_off = (((int64_t)par * p.world + p.rank) * p.max_rows + row) * p.row_bytes;
  for (int d = 1; d < p.world; ++d) {
    const int peer = (p.rank + d) % p.world;   // staggered so that the ranks do not all hit the same peer first
    unsigned char* dst = p.peer[peer] + slot_off;
#pragma unroll
    for (int i = 0; i < kMaxIters; ++i) {
      const int c = tid + i * kThreads;
      if (c < npkts) st_pkt(dst + (int64_t)c * 16, mine[i][0], mine[i][1], mine[i][2], e);
    }
  }

  // ---- gather: poll the peers' packets in this rank's own buffer, sum in rank order (fp32) ----
  float vals[kMaxIters][kPktElems];
  float ss = 0.f;
  T* rr = p.res ? reinterpret_cast<T*>(p.res) + (int64_t)row * p.r_stride : nullptr;
  T* yr = reinterpret_cast<T*>(p.y) + (int64_t)row * p.y_stride;
#pragma unroll
  for (int i = 0; i < kMaxIters; ++i) {
    const int c = tid + i * kThreads;
    if (c < npkts) {
      float acc[kPktElems];
#pragma unroll
      for (int j = 0; j < kPktElems; ++j) acc[j] = 0.f;
      // all peers' packets of this chunk are requested before the first one is examined (the polls are L2 round trips:
      // one after the other they cost (world - 1) x ~0.7 us per chunk, which is what made the 4-rank exchange 11 us)
      uint4 pv[kMaxWorld];
#pragma unroll
      for (int r = 0; r < kMaxWorld; ++r) {
        if (r < p.world && r != p.rank)
          pv[r] = ld_pkt(mine_base + (((int64_t)par * p.world + r) * p.max_rows + row) * p.row_bytes + (int64_t)c * 16);
      }
#pragma unroll
      for (int r = 0; r < kMaxWorld; ++r) {
        if (r >= p.world) continue;
        uint32_t w3[3];
        if (r == p.rank) {
          w3[0] = mine[i][0]; w3[1] = mine[i][1]; w3[2] = mine[i][2];
        } else {
          uint4 v = pv[r];
          if (v.w != e) {
            const unsigned char* src = mine_base + (((int64_t)par * p.world + r) * p.max_rows + row) * p.row_bytes + (int64_t)c * 16;
            const unsigned long long t0 = b2l_globaltimer();
            do {
              v = ld_pkt(src);
              const unsigned long long lim = b2l_hang_ctl.limit_ns;
              const unsigned long long dt = b2l_globaltimer() - t0;
              if (lim != 0ull && dt > lim && v.w != e) b2l_hang_trap((uint32_t)r, e, b2l_file_hash(__FILE__), __LINE__, dt);
            } while (v.w != e);
          }
          w3[0] = v.x; w3[1] = v.y; w3[2] = v.z;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          float a, b;
          unpack2<T>(w3[k], a, b);
          acc[2 * k] += a; acc[2 * k + 1] += b;
        }
      }
      uint32_t* rw = rr ? reinterpret_cast<uint32_t*>(rr) : nullptr;
      uint32_t* yw = reinterpret_cast<uint32_t*>(yr);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int wi = 3 * c + k;
        float s0 = 0.f, s1 = 0.f;
        if (wi < nwords) {
          // the all-reduce result in the activation dtype (what the reference's NCCL call leaves in the tensor) ...
          const float o0 = to_f32(from_f32<T>(acc[2 * k])), o1 = to_f32(from_f32<T>(acc[2 * k + 1]));
          s0 = o0; s1 = o1;
          if (rw) {                                              // ... then skip_rmsnorm.py: r += x
            float r0, r1;
            unpack2<T>(rw[wi], r0, r1);
            s0 += r0; s1 += r1;
            rw[wi] = pack2t<T>(s0, s1);
          } else if (!p.w) {
            yw[wi] = pack2t<T>(s0, s1);                          // plain all-reduce
          }
        }
        vals[i][2 * k] = s0; vals[i][2 * k + 1] = s1;
        ss += s0 * s0 + s1 * s1;
      }
    }
  }
  if (tid == 0) my_epoch[row] = e;
  if (!p.w) return;

  // ---- RMSNorm of the row (skip_rmsnorm.py:126-189: fp32 mean of squares, round before the weight multiply) ----
  __shared__ float red[kThreads / 32];
  ss = warp_reduce_sum(ss);
  if ((tid & 31) == 0) red[tid >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < kThreads / 32; ++i) tot += red[i];
  const float rrms = div_full_f(1.0f, sqrt_approx_f(tot / (float)p.n + p.eps));
  const uint32_t* ww = reinterpret_cast<const uint32_t*>(p.w);
  uint32_t* yw = reinterpret_cast<uint32_t*>(yr);
#pragma unroll
  for (int i = 0; i < kMaxIters; ++i) {
    const int c = tid + i * kThreads;
    if (c < npkts) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int wi = 3 * c + k;
        if (wi < nwords) {
          float w0, w1;
          unpack2<T>(__ldg(ww + wi), w0, w1);
          const float n0 = to_f32(from_f32<T>(vals[i][2 * k] * rrms)), n1 = to_f32(from_f32<T>(vals[i][2 * k + 1] * rrms));
          yw[wi] = pack2t<T>(n0 * w0, n1 * w1);
        }
      }
    }
  }
}

}  // namespace

// ---- host side: one communicator per process (one process per GPU) ----
extern "C" int64_t b2l_tp_comm_create(int rank, int world, int max_rows, int64_t max_row_elems) {
  if (world < 2 || world > kMaxWorld || rank < 0 || rank >= world || max_rows <= 0 || max_row_elems <= 0 ||
      max_row_elems % 2 != 0 || max_row_elems > (int64_t)kThreads * kMaxIters * kPktElems) {
    b2l_set_error("tp_comm_create: bad arguments (world %d, rank %d, rows %d, row elems %lld; world <= %d, row elems <= %d)",
                  world, rank, max_rows, (long long)max_row_elems, kMaxWorld, kThreads * kMaxIters * kPktElems);
    return 0;
  }
  CommState* cs = new CommState();
  memset(cs, 0, sizeof(*cs));
  cs->rank = rank; cs->world = world; cs->max_rows = max_rows; cs->max_row_elems = max_row_elems;
  cs->row_bytes = ((max_row_elems / 2 + 2) / 3) * 16;
  const int64_t bytes = total_bytes(world, max_rows, cs->row_bytes);
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, (size_t)bytes);
  if (e == cudaSuccess) e = cudaMemset(p, 0, (size_t)bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
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
  B2L_REQUIRE(n > 0 && n % 2 == 0 && n <= cs->max_row_elems, "tp_allreduce_skip_rmsnorm: row of %lld elements (capacity %lld, even)",
              (long long)n, (long long)cs->max_row_elems);
  B2L_REQUIRE(x_row_stride % 2 == 0 && y_row_stride % 2 == 0 && (residual == nullptr || r_row_stride % 2 == 0) &&
                  (uintptr_t)x_partial % 4 == 0 && (uintptr_t)y % 4 == 0 && (uintptr_t)residual % 4 == 0 &&
                  (uintptr_t)weight % 4 == 0,
              "tp_allreduce_skip_rmsnorm: rows must be 4-byte aligned");
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
