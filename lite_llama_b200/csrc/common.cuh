// Shared device/host helpers for the B200 (sm_100a) kernel layer.
//
// Everything in this library is launched on a caller-supplied stream, never
// allocates and never synchronises, so every entry point is CUDA-graph
// capturable (reference contract: lite_llama/executor/cuda_graph.py:117-119).
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>

// --------------------------------------------------------------------------- //
// Error plumbing (thread-local last-error string, negative return codes)
// --------------------------------------------------------------------------- //
#define B2L_OK 0
#define B2L_ERR_INVALID (-1)
#define B2L_ERR_CUDA (-2)
#define B2L_ERR_UNSUPPORTED (-3)

void b2l_set_error(const char* fmt, ...);

#define B2L_REQUIRE(cond, ...)          \
  do {                                  \
    if (!(cond)) {                      \
      b2l_set_error(__VA_ARGS__);       \
      return B2L_ERR_INVALID;           \
    }                                   \
  } while (0)

#define B2L_UNSUPPORTED(...)            \
  do {                                  \
    b2l_set_error(__VA_ARGS__);         \
    return B2L_ERR_UNSUPPORTED;         \
  } while (0)

#define B2L_CUDA_CHECK(expr)                                                        \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      b2l_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return B2L_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)

// every kernel launch of this library goes through B2L_LAUNCH_CHECK: count them (bench.py reports it)
void b2l_count_launch();

#define B2L_LAUNCH_CHECK()                                                          \
  do {                                                                              \
    b2l_count_launch();                                                             \
    cudaError_t _e = cudaGetLastError();                                            \
    if (_e != cudaSuccess) {                                                        \
      b2l_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return B2L_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)

// dtype enum shared with include/b200_lite_kernels.h
#define B2L_F16 0
#define B2L_BF16 1

int b2l_pdl_enabled();        // programmatic dependent launch switch (b2l_set_pdl / env B2L_PDL; default on)
int b2l_num_sms();            // cached cudaDevAttrMultiProcessorCount of the current device
int b2l_max_smem_optin();     // cached cudaDevAttrMaxSharedMemoryPerBlockOptin

static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// --------------------------------------------------------------------------- //
// Device helpers
// --------------------------------------------------------------------------- //
#ifdef __CUDACC__

template <typename T> struct Vec2;
template <> struct Vec2<__half> { using type = __half2; };
template <> struct Vec2<__nv_bfloat16> { using type = __nv_bfloat162; };

__device__ __forceinline__ float to_f32(__half v) { return __half2float(v); }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// 128-bit global accesses. `nc` + no L1 allocation for streamed, read-once data.
__device__ __forceinline__ uint4 ldg_stream_128(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint4 ldg_128(const void* p) {
  return *reinterpret_cast<const uint4*>(p);
}
__device__ __forceinline__ void stg_128(void* p, const uint4& v) {
  *reinterpret_cast<uint4*>(p) = v;
}

// Index loads that accept both int32 and int64 device vectors
// (reference callers pass either: SURVEY.md §0.7).
__device__ __forceinline__ int64_t load_index(const void* p, int64_t i, int width) {
  return width == 8 ? reinterpret_cast<const int64_t*>(p)[i]
                    : (int64_t)reinterpret_cast<const int32_t*>(p)[i];
}

__device__ __forceinline__ float warp_reduce_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_reduce_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ex2.approx: exp2(-inf) = +0, what the online-softmax bookkeeping relies on.
__device__ __forceinline__ float fast_exp2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// One lane of a fully converged warp (always the same one while the mask is full).  Role loops are run
// by the whole warp and only the issuing instructions are predicated on this: the loop state stays
// warp-uniform, so ptxas keeps descriptors/addresses in uniform registers instead of emitting a
// per-thread broadcast loop around every tcgen05/TMA instruction (measured: 2.3k cycles per k-block).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier ------------------------------------------------------------- //
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
// The suspend-time hint lets the hardware park a waiting warp until the phase completes instead of returning
// "not yet" every few cycles: a polling loop costs issue slots the dequant warps need (ncu: the TMA producer's
// poll loop alone was 11% of all executed instructions of the quantised GEMM).
constexpr uint32_t kMbarSuspendHintNs = 0x989680u;
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(kMbarSuspendHintNs) : "memory");
  return ok != 0;
}
// Pure polling (mbarrier.test_wait never suspends the thread): lowest wake-up latency, costs issue slots.
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// try_wait with the default (short, implementation-defined) suspend window instead of the long hint above
__device__ __forceinline__ bool mbar_try_wait_short(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}

// ---- device-side hang watchdog -------------------------------------------- //
// A pipeline bug in a warp-specialised kernel shows up as a warp parked on an mbarrier for ever: the GPU stays
// "busy", the host blocks in the next synchronisation and nothing says which kernel, role or barrier it was.
// Every barrier wait of this library therefore goes through b2l_wait_slow() once its first probe fails: after
// `limit_ns` of wall time (default 4 s, env B2L_HANG_MS; decode kernels take microseconds) the waiting thread
// writes {file hash, line, block, thread, barrier address, parity} into a host-pinned record and traps, which
// turns the hang into a CUDA error the host reports with the record (b2l_hang_report()).  The fast path (the
// probe succeeds) costs nothing; the slow path reads %globaltimer once per failed probe (each probe already
// suspends the warp for up to 10 ms).
struct B2lHangCtl {
  unsigned int* rec;              // host-pinned, 16 words; null = watchdog disarmed
  unsigned long long limit_ns;    // 0 = never trap
  int taken;                      // first trapping thread of this translation unit writes the record
};
static __device__ B2lHangCtl b2l_hang_ctl = {nullptr, 0ull, 0};
constexpr unsigned int b2l_file_hash(const char* s) {
  unsigned int h = 2166136261u;
  // hash of the basename only, so that build location does not matter
  const char* base = s;
  for (const char* c = s; *c; ++c)
    if (*c == '/') base = c + 1;
  for (const char* c = base; *c; ++c) h = (h ^ (unsigned char)*c) * 16777619u;
  return h;
}
__device__ __forceinline__ unsigned long long b2l_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
static __device__ __noinline__ void b2l_hang_trap(uint32_t bar_addr, uint32_t parity, unsigned int file_hash, int line,
                                           unsigned long long waited_ns) {
  unsigned int* rec = b2l_hang_ctl.rec;
  if (rec != nullptr && atomicCAS(&b2l_hang_ctl.taken, 0, 1) == 0) {
    rec[1] = file_hash; rec[2] = (unsigned)line;
    rec[3] = blockIdx.x; rec[4] = blockIdx.y; rec[5] = blockIdx.z;
    rec[6] = threadIdx.x; rec[7] = bar_addr; rec[8] = parity;
    rec[9] = (unsigned)(waited_ns / 1000000ull);
    rec[10] = gridDim.x; rec[11] = gridDim.y; rec[12] = gridDim.z; rec[13] = blockDim.x;
    __threadfence_system();
    rec[0] = 0xB2D00DEDu;          // "record valid"
    __threadfence_system();
  } else {
    // give the record's writer time to reach host memory before this thread kills the context
    const unsigned long long t0 = b2l_globaltimer();
    while (b2l_globaltimer() - t0 < 2000000ull) {
    }
  }
  __trap();
}
// mode 0: try_wait + long suspend hint, 1: test_wait polling, 2: try_wait with the default suspend window
template <int MODE>
__device__ __forceinline__ void b2l_wait_slow(uint64_t* bar, uint32_t parity, unsigned int file_hash, int line) {
  const unsigned long long t0 = b2l_globaltimer();
  for (;;) {
    const bool ok = MODE == 0 ? mbar_try_wait(bar, parity)
                  : MODE == 1 ? mbar_test_wait(bar, parity) : mbar_try_wait_short(bar, parity);
    if (ok) return;
    const unsigned long long lim = b2l_hang_ctl.limit_ns;
    const unsigned long long dt = b2l_globaltimer() - t0;
    if (lim != 0ull && dt > lim) b2l_hang_trap(smem_u32(bar), parity, file_hash, line, dt);
  }
}
__device__ __forceinline__ void mbar_wait_at(uint64_t* bar, uint32_t parity, unsigned int file_hash, int line) {
  if (mbar_try_wait(bar, parity)) return;
  b2l_wait_slow<0>(bar, parity, file_hash, line);
}
__device__ __forceinline__ void mbar_wait_spin_at(uint64_t* bar, uint32_t parity, unsigned int file_hash, int line) {
  if (mbar_test_wait(bar, parity)) return;
  b2l_wait_slow<1>(bar, parity, file_hash, line);
}
__device__ __forceinline__ void mbar_wait_short_at(uint64_t* bar, uint32_t parity, unsigned int file_hash, int line) {
  if (mbar_try_wait_short(bar, parity)) return;
  b2l_wait_slow<2>(bar, parity, file_hash, line);
}
#define mbar_wait(bar, parity) mbar_wait_at((bar), (parity), b2l_file_hash(__FILE__), __LINE__)
#define mbar_wait_spin(bar, parity) mbar_wait_spin_at((bar), (parity), b2l_file_hash(__FILE__), __LINE__)
#define mbar_wait_short(bar, parity) mbar_wait_short_at((bar), (parity), b2l_file_hash(__FILE__), __LINE__)

// ---- async-proxy fences --------------------------------------------------- //
// Generic-proxy smem writes (st.shared) must be made visible to the async proxy
// (TMA / tcgen05 operand reads) before the consumer is signalled.
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- bulk (TMA engine) copies -------------------------------------------- //
// 1-D bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// 2-D tiled TMA load (SASS: UTMALDG). c0 = innermost coordinate.
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tmap, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_2d_hint(void* smem_dst, const void* tmap, int c0, int c1, uint64_t* bar,
                                                 uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(hint) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const void* tmap, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
// L2 eviction-priority policies for TMA loads (same encodings CUTLASS uses).
static constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;
static constexpr uint64_t kEvictLast = 0x14F0000000000000ull;
static constexpr uint64_t kEvictNormal = 0x1000000000000000ull;

// ---- tcgen05 / TMEM ------------------------------------------------------- //
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread retire.
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 covers fp16 and bf16 inputs.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// int8 x int8 -> int32 (sm_100a has kind::i8; it is gone on sm_103).
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// Shared-memory matrix descriptor, K-major operand in the canonical 128B-swizzled
// layout: rows of 128 bytes, 8-row groups 1024 B apart (what a TMA box of
// {64 x 2-byte elements, rows} with CU_TENSOR_MAP_SWIZZLE_128B produces).
__device__ __forceinline__ uint64_t umma_desc_sw128_kmajor(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);   // start address, 16-byte units
  d |= (uint64_t)0 << 16;                          // LBO: unused for swizzled K-major
  d |= (uint64_t)(1024u >> 4) << 32;               // SBO: 8 rows * 128 B
  d |= (uint64_t)1 << 46;                          // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
  return d;
}
// MN-major operand (the MN index is the contiguous one), 128B swizzle. Canonical
// layout in 16-byte units: ((8,n),(8,k)):((1,LBO),(8,SBO)) -- an "atom" is 64
// MN-contiguous elements x 8 k-rows = 1024 B; atoms step by LBO along MN and by
// SBO along K.
__device__ __forceinline__ uint64_t umma_desc_sw128_mnmajor(uint32_t smem_addr, uint32_t lbo_bytes,
                                                            uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// Instruction descriptor (upper 32 bits of the "idesc" operand), dense, no negate.
//   c_fmt: 0 = f16, 1 = f32, 2 = s32; ab_fmt: kind::f16 0 = f16, 1 = bf16; kind::i8 0 = u8, 1 = s8
__host__ __device__ constexpr uint32_t umma_idesc(uint32_t c_fmt, uint32_t a_fmt, uint32_t b_fmt, uint32_t a_mn_major,
                                                  uint32_t b_mn_major, uint32_t M, uint32_t N) {
  return (c_fmt << 4) | (a_fmt << 7) | (b_fmt << 10) | (a_mn_major << 15) | (b_mn_major << 16) | ((N >> 3) << 17) |
         ((M >> 4) << 24);
}

// TMEM -> registers: each lane of the warp reads its own TMEM lane (row),
// N consecutive 32-bit columns.
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(taddr) : "memory");
}

// ---- legacy-path helpers used by the gather-bound attention kernels ------- //
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
template <typename T>
__device__ __forceinline__ void mma_m16n8k16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1);
template <>
__device__ __forceinline__ void mma_m16n8k16<__half>(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <>
__device__ __forceinline__ void mma_m16n8k16<__nv_bfloat16>(float (&d)[4], const uint32_t (&a)[4], uint32_t b0,
                                                            uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <typename T> __device__ __forceinline__ uint32_t pack2(float lo, float hi);
template <> __device__ __forceinline__ uint32_t pack2<__half>(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// Programmatic dependent launch (PDL).  A decode step is ~350 kernels of 2-40 us each, so the launch gap
// between two kernels (grid drain + launch latency, ~2 us even inside a CUDA graph) is a quarter of the step.
// Every kernel of this library is launched with the "programmatic stream serialization" attribute and
//   * calls pdl_launch_dependents() first thing, so the NEXT kernel is launched while this one runs: its CTAs
//     become resident wherever an SM has room, set up barriers/TMEM and -- the GEMMs -- start streaming
//     their (static) weights;
//   * calls pdl_wait_prior() before its first access to anything an earlier kernel may have written, or to
//     anything it writes itself that an earlier kernel may still read.  The wait returns when the previous
//     kernel has completed and flushed; since that kernel waited for ITS predecessor before finishing,
//     completion is transitive over the whole stream.  Every kernel must therefore execute the wait on
//     every path.  Without the launch attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_wait_prior() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// for kernels whose first instruction already reads the previous kernel's output
__device__ __forceinline__ void pdl_prologue() {
  pdl_launch_dependents();
  pdl_wait_prior();
}

// Host side of the hang watchdog: every translation unit has its own copy of b2l_hang_ctl (no relocatable device
// code), armed once per device from the first launch of one of its kernels outside a stream capture.
unsigned int* b2l_hang_record_ptr();          // core.cu: the process-wide host-pinned record
unsigned long long b2l_hang_limit_ns();       // core.cu: env B2L_HANG_MS (default 4000, 0 = off)
static inline void b2l_hang_arm(cudaStream_t st) {
  static bool armed[64] = {false};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || armed[dev]) return;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) {
    (void)cudaGetLastError();
    return;   // a synchronous symbol copy is not capturable: arm at the next eager launch
  }
  B2lHangCtl h;
  h.rec = b2l_hang_record_ptr();
  h.limit_ns = b2l_hang_limit_ns();
  h.taken = 0;
  if (cudaMemcpyToSymbolAsync(b2l_hang_ctl, &h, sizeof(h), 0, cudaMemcpyHostToDevice, st) != cudaSuccess)
    (void)cudaGetLastError();
  armed[dev] = true;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute: remember the largest value set per device
// (a process that drives several GPUs must set it on each; round-1 ADVICE).  `cache` is a function-local static array.
template <typename Kern>
static inline cudaError_t b2l_ensure_dynamic_smem(Kern kern, int bytes, int (&cache)[64]) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  dev &= 63;
  if (bytes > cache[dev]) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return e;
    cache[dev] = bytes;
  }
  return cudaSuccess;
}

// Host side: launch `kern` on `st`, with the PDL attribute unless the library switch is off.
template <typename... KArgs, typename... Args>
static inline cudaError_t b2l_launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args&&... args) {
  b2l_hang_arm(st);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = b2l_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

#endif  // __CUDACC__
