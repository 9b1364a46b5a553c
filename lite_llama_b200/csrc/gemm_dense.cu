// Dense projections on tcgen05: fp16/bf16 linear (q/kv/o/gate/up/down/lm_head, the reference's
// F.linear call sites: lite_llama/models/quantization/methods/unquantized.py:21-22,
// lite_llama/models/base.py:489) and the int8 x int8 GEMM of SmoothQuant-W8A8
// (lite_llama/kernels/quantization/w8a8.py:75-130).  See gemm_tc.cuh for the tile design.
#include "gemm_tc.cuh"
#include <type_traits>

namespace gemm {

struct DenseParams {
  void* y; int64_t ldy;          // output [M, N], row stride in elements
  const void* bias;              // [N] in the output dtype (f16 path) or fp32 (int8 path), nullable
  const float* a_scale;          // int8 path: per-token activation scales [M]
  const float* b_scale;          // int8 path: per-channel weight scales [N]
  uint32_t* ws;                  // split-K partials
  int* counters;                 // split-K tickets, one per output tile, zero on entry
  int M, N;
  int splits, kb_per_split, num_kb;
  int stages;
};

// KIND 0: fp16/bf16 operands, fp32 accumulate.  KIND 1: int8 operands, int32 accumulate.
template <int KIND, typename T, int BT>
__global__ void __launch_bounds__(192, 1)
gemm_swapab_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x,
                   const DenseParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  Control* ctl = reinterpret_cast<Control*>(smem);
  uint8_t* a_ring = smem + 1024;
  uint8_t* b_ring = a_ring + (size_t)p.stages * kABytes;
  constexpr int kBBytes = BT * kStageKBytes;
  constexpr int kKPerStage = KIND == 0 ? 64 : 128;     // elements along K per stage
  constexpr int kTmemCols = BT < 32 ? 32 : BT;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tile = blockIdx.x, m_tile = blockIdx.y, split = blockIdx.z;
  const int kb0 = split * p.kb_per_split;
  const int kb1 = kb0 + p.kb_per_split < p.num_kb ? kb0 + p.kb_per_split : p.num_kb;
  const int nkb = kb1 - kb0;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_x);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&ctl->full[s], 1);
      mbar_init(&ctl->empty[s], 1);
    }
    mbar_init(&ctl->acc_full[0], 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(&ctl->tmem_base, kTmemCols);
    tmem_relinquish();
    // The next kernel may set itself up while this one streams weights (common.cuh).  Only the TMEM-owning warp
    // triggers, and only AFTER the allocation: a dependent CTA that became co-resident earlier could take the SM's
    // TMEM columns and then park in griddepcontrol.wait for this grid -- while this CTA blocks in tcgen05.alloc.
    pdl_launch_dependents();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;

  if (warp == 0) {
    // ===================== TMA producer (whole warp loops, one elected lane issues) =====================
    // Weights are static, activations are the previous kernel's output: under programmatic dependent launch
    // the first ring-full of weight tiles is requested BEFORE waiting for the previous kernel to finish.
    const int pre = nkb < p.stages ? nkb : p.stages;
    if (elect_one()) {
      for (int i = 0; i < pre; ++i) {
        mbar_arrive_expect_tx(&ctl->full[i], kABytes + kBBytes);
        tma_load_2d_hint(a_ring + (size_t)i * kABytes, &tmap_w, (kb0 + i) * kKPerStage, n_tile * kBlockW, &ctl->full[i],
                         kEvictFirst);
      }
    }
    __syncwarp();
    pdl_wait_prior();
    if (elect_one()) {
      for (int i = 0; i < pre; ++i)
        tma_load_2d_hint(b_ring + (size_t)i * kBBytes, &tmap_x, (kb0 + i) * kKPerStage, m_tile * BT, &ctl->full[i],
                         kEvictLast);
    }
    __syncwarp();
    PipeState ps(p.stages);
    for (int i = 0; i < pre; ++i) ps.advance();
    for (int i = pre; i < nkb; ++i) {
      mbar_wait(&ctl->empty[ps.stage], ps.phase ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&ctl->full[ps.stage], kABytes + kBBytes);
        const int kc = (kb0 + i) * kKPerStage;
        // weights are streamed exactly once: evict-first; activations are re-read by every CTA: evict-last
        tma_load_2d_hint(a_ring + (size_t)ps.stage * kABytes, &tmap_w, kc, n_tile * kBlockW, &ctl->full[ps.stage],
                         kEvictFirst);
        tma_load_2d_hint(b_ring + (size_t)ps.stage * kBBytes, &tmap_x, kc, m_tile * BT, &ctl->full[ps.stage],
                         kEvictLast);
      }
      __syncwarp();
      ps.advance();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp loops, one elected lane issues) =====================
    constexpr uint32_t idesc = KIND == 0
        ? umma_idesc(1u, std::is_same<T, __nv_bfloat16>::value ? 1u : 0u,
                     std::is_same<T, __nv_bfloat16>::value ? 1u : 0u, 0u, 0u, kBlockW, BT)
        : umma_idesc(2u, 1u, 1u, 0u, 0u, kBlockW, BT);
    const uint64_t a_desc0 = umma_desc_sw128_kmajor(smem_u32(a_ring));
    const uint64_t b_desc0 = umma_desc_sw128_kmajor(smem_u32(b_ring));
    PipeState ps(p.stages);
    for (int i = 0; i < nkb; ++i) {
      mbar_wait(&ctl->full[ps.stage], ps.phase);
      tc_fence_after();
      const uint64_t adesc = a_desc0 + (uint64_t)((ps.stage * kABytes) >> 4);
      const uint64_t bdesc = b_desc0 + (uint64_t)((ps.stage * kBBytes) >> 4);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          // one MMA consumes 32 bytes of K per row: advance the start address by 2 (16-byte units)
          if (KIND == 0)
            umma_f16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (i | k) != 0 ? 1u : 0u);
          else
            umma_i8(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (i | k) != 0 ? 1u : 0u);
        }
        tc_commit(&ctl->empty[ps.stage]);  // frees the smem stage once these MMAs retire
        if (i + 1 == nkb) tc_commit(&ctl->acc_full[0]);
      }
      __syncwarp();
      ps.advance();
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int quad = warp & 3;                  // TMEM lane quadrant this warp may access
    const int epi_tid = quad * 32 + lane;       // == accumulator lane == local weight row
    const int n = n_tile * kBlockW + epi_tid;
    const int tile_id = m_tile * gridDim.x + n_tile;
    mbar_wait(&ctl->acc_full[0], 0);
    tc_fence_after();
    pdl_wait_prior();   // activation scales are upstream data; y / the split-K workspace may still be read upstream
    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
    constexpr int CH = 16;
    T* y = reinterpret_cast<T*>(p.y);

    auto finish = [&](int col0, const uint32_t (&v)[CH]) {
      if (n >= p.N) return;
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        const int m = m_tile * BT + col0 + j;
        if (m < p.M) {
          float r;
          if (KIND == 0) {
            r = __uint_as_float(v[j]);
            if (p.bias) r += to_f32(reinterpret_cast<const T*>(p.bias)[n]);
          } else {
            // w8a8.py:117-123: acc.float * a_scale[m] * b_scale[n] (+ bias)
            r = (float)(int)v[j] * p.a_scale[m] * p.b_scale[n];
            if (p.bias) r += reinterpret_cast<const float*>(p.bias)[n];
          }
          y[(int64_t)m * p.ldy + n] = from_f32<T>(r);
        }
      }
    };

    if (p.splits == 1) {
#pragma unroll 1
      for (int c = 0; c < BT; c += CH) {
        uint32_t v[CH];
        tmem_ld_32x32b_x16(taddr + c, v);
        tc_wait_ld();
        finish(c, v);
      }
    } else {
      uint32_t* mine = p.ws + ((size_t)tile_id * p.splits + split) * (BT * kBlockW);
#pragma unroll 1
      for (int c = 0; c < BT; c += CH) {
        uint32_t v[CH];
        tmem_ld_32x32b_x16(taddr + c, v);
        tc_wait_ld();
#pragma unroll
        for (int j = 0; j < CH; ++j) mine[(c + j) * kBlockW + epi_tid] = v[j];
      }
      if (splitk_ticket(p.counters, tile_id, p.splits, epi_tid, &ctl->is_last)) {
        const uint32_t* base = p.ws + (size_t)tile_id * p.splits * (BT * kBlockW);
#pragma unroll 1
        for (int c = 0; c < BT; c += CH) {
          uint32_t v[CH];
#pragma unroll
          for (int j = 0; j < CH; ++j) v[j] = 0u;
          // four splits' worth of loads in flight before the first addition (one L2 round trip per batch instead of
          // one per split); the sums still run in split order, the padding zeros of a short batch change nothing
          for (int s0 = 0; s0 < p.splits; s0 += 4) {
            uint32_t w[4][CH];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const bool on = s0 + u < p.splits;
              const uint32_t* src = base + (size_t)(s0 + u) * (BT * kBlockW) + c * kBlockW + epi_tid;
#pragma unroll
              for (int j = 0; j < CH; ++j) w[u][j] = on ? __ldcg(src + j * kBlockW) : 0u;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
#pragma unroll
              for (int j = 0; j < CH; ++j) {
                if (KIND == 0)
                  v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(w[u][j]));
                else
                  v[j] = (uint32_t)((int)v[j] + (int)w[u][j]);
              }
            }
          }
          finish(c, v);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

template <int KIND, typename T, int BT>
static int launch_bt(const CUtensorMap& tw, const CUtensorMap& tx, const DenseParams& p, const SplitPlan& pl,
                     cudaStream_t st) {
  auto kern = gemm_swapab_kernel<KIND, T, BT>;
  static int configured_smem[64] = {0};
  B2L_CUDA_CHECK(b2l_ensure_dynamic_smem(kern, pl.smem_bytes, configured_smem));
  dim3 grid((unsigned)pl.n_tiles, (unsigned)pl.m_tiles, (unsigned)pl.splits);
  b2l_launch(kern, dim3(grid), dim3(192), (size_t)(pl.smem_bytes), st, tw, tx, p);
  B2L_LAUNCH_CHECK();
  return B2L_OK;
}

template <int KIND, typename T>
static int launch(const CUtensorMap& tw, const CUtensorMap& tx, const DenseParams& p, const SplitPlan& pl,
                  cudaStream_t st) {
  switch (pl.bt) {
    case 16: return launch_bt<KIND, T, 16>(tw, tx, p, pl, st);
    case 32: return launch_bt<KIND, T, 32>(tw, tx, p, pl, st);
    case 64: return launch_bt<KIND, T, 64>(tw, tx, p, pl, st);
    case 128: return launch_bt<KIND, T, 128>(tw, tx, p, pl, st);
    default: return launch_bt<KIND, T, 256>(tw, tx, p, pl, st);
  }
}

}  // namespace gemm

using namespace gemm;

extern "C" int64_t b2l_linear_workspace_bytes(int64_t m, int64_t n, int64_t k, int elem_bytes) {
  if (m <= 0 || n <= 0 || k <= 0) return 0;
  SplitPlan pl = make_plan(m, n, k, elem_bytes == 1 ? 128 : 64, 0, 1024, 4);
  if (pl.splits <= 1) return 0;
  return (int64_t)pl.n_tiles * pl.m_tiles * pl.splits * pl.bt * kBlockW * 4;
}

// number of int32 split-K tickets a launch may touch (they must be zero on entry; the kernel re-zeroes them)
extern "C" int64_t b2l_linear_num_counters(int64_t m, int64_t n) {
  if (m <= 0 || n <= 0) return 0;
  const int bt = pick_bt(m) > 64 ? 64 : pick_bt(m);   // the plan may shrink the token tile down to 64 (gemm_tc.cuh)
  return cdiv64(n, kBlockW) * cdiv64(m, bt);
}

extern "C" int b2l_linear(void* y, const void* x, const void* w, const void* bias, int64_t m, int64_t n, int64_t k,
                          int64_t x_row_stride, int64_t w_row_stride, int64_t y_row_stride, void* workspace,
                          int64_t workspace_bytes, void* counters, int64_t num_counters, int dtype, void* stream) {
  B2L_REQUIRE(y && x && w, "linear: null pointer");
  B2L_REQUIRE(m >= 0 && n > 0 && k > 0, "linear: bad shape m=%lld n=%lld k=%lld", (long long)m, (long long)n, (long long)k);
  B2L_REQUIRE(dtype == B2L_F16 || dtype == B2L_BF16, "linear: dtype %d", dtype);
  B2L_REQUIRE(k % 8 == 0 && x_row_stride % 8 == 0 && w_row_stride % 8 == 0,
              "linear: K and the row strides must be multiples of 8 elements (TMA 16-byte rule); k=%lld", (long long)k);
  if (m == 0) return B2L_OK;
  SplitPlan pl = make_plan(m, n, k, 64, 0, 1024, 4);
  DenseParams p;
  p.y = y; p.ldy = y_row_stride; p.bias = bias; p.a_scale = nullptr; p.b_scale = nullptr;
  p.M = (int)m; p.N = (int)n; p.splits = pl.splits; p.kb_per_split = pl.kb_per_split; p.num_kb = pl.num_kb;
  p.stages = pl.stages; p.ws = (uint32_t*)workspace; p.counters = (int*)counters;
  if (pl.splits > 1) {
    const int64_t need = b2l_linear_workspace_bytes(m, n, k, 2);
    B2L_REQUIRE(workspace && workspace_bytes >= need, "linear: workspace too small (%lld < %lld)",
                (long long)workspace_bytes, (long long)need);
    B2L_REQUIRE(counters && num_counters >= (int64_t)pl.n_tiles * pl.m_tiles, "linear: split-K counters too small");
  }
  CUtensorMap tw, tx;
  const TmaElem el = dtype == B2L_F16 ? TMA_F16 : TMA_BF16;
  int rc = b2l_make_tmap_2d(&tw, w, el, (uint64_t)n, (uint64_t)k, (uint64_t)w_row_stride * 2, kBlockW, 64, true);
  if (rc != B2L_OK) return rc;
  rc = b2l_make_tmap_2d(&tx, x, el, (uint64_t)m, (uint64_t)k, (uint64_t)x_row_stride * 2, (uint32_t)pl.bt, 64, true);
  if (rc != B2L_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == B2L_F16) return launch<0, __half>(tw, tx, p, pl, st);
  return launch<0, __nv_bfloat16>(tw, tx, p, pl, st);
}

// int8 x int8 -> int32 -> out dtype, with per-token / per-channel scales (SmoothQuant W8A8 GEMM stage).
extern "C" int b2l_w8a8_gemm(void* y, const void* xq, const void* wq, const void* a_scale, const void* b_scale,
                             const void* bias_f32, int64_t m, int64_t n, int64_t k, int64_t xq_row_stride,
                             int64_t wq_row_stride, int64_t y_row_stride, void* workspace, int64_t workspace_bytes,
                             void* counters, int64_t num_counters, int out_dtype, void* stream) {
  B2L_REQUIRE(y && xq && wq && a_scale && b_scale, "w8a8_gemm: null pointer");
  B2L_REQUIRE(m >= 0 && n > 0 && k > 0, "w8a8_gemm: bad shape");
  B2L_REQUIRE(out_dtype == B2L_F16 || out_dtype == B2L_BF16, "w8a8_gemm: out dtype %d", out_dtype);
  B2L_REQUIRE(k % 16 == 0 && xq_row_stride % 16 == 0 && wq_row_stride % 16 == 0,
              "w8a8_gemm: K and the row strides must be multiples of 16 bytes; k=%lld", (long long)k);
  if (m == 0) return B2L_OK;
  SplitPlan pl = make_plan(m, n, k, 128, 0, 1024, 4);
  DenseParams p;
  p.y = y; p.ldy = y_row_stride; p.bias = bias_f32; p.a_scale = (const float*)a_scale; p.b_scale = (const float*)b_scale;
  p.M = (int)m; p.N = (int)n; p.splits = pl.splits; p.kb_per_split = pl.kb_per_split; p.num_kb = pl.num_kb;
  p.stages = pl.stages; p.ws = (uint32_t*)workspace; p.counters = (int*)counters;
  if (pl.splits > 1) {
    const int64_t need = b2l_linear_workspace_bytes(m, n, k, 1);
    B2L_REQUIRE(workspace && workspace_bytes >= need, "w8a8_gemm: workspace too small (%lld < %lld)",
                (long long)workspace_bytes, (long long)need);
    B2L_REQUIRE(counters && num_counters >= (int64_t)pl.n_tiles * pl.m_tiles, "w8a8_gemm: split-K counters too small");
  }
  CUtensorMap tw, tx;
  int rc = b2l_make_tmap_2d(&tw, wq, TMA_U8, (uint64_t)n, (uint64_t)k, (uint64_t)wq_row_stride, kBlockW, 128, true);
  if (rc != B2L_OK) return rc;
  rc = b2l_make_tmap_2d(&tx, xq, TMA_U8, (uint64_t)m, (uint64_t)k, (uint64_t)xq_row_stride, (uint32_t)pl.bt, 128, true);
  if (rc != B2L_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (out_dtype == B2L_F16) return launch<1, __half>(tw, tx, p, pl, st);
  return launch<1, __nv_bfloat16>(tw, tx, p, pl, st);
}
