// Weight-only quantised GEMMs on tcgen05: W4A16 (int4 groups + fp32 scale/zero) and W8A16
// (fp8-e4m3 or int8 bytes + fp32 block scales).  Replaces
//   lite_llama/kernels/quantization/w4a16.py:28-207   (scalar outer-product Triton kernel)
//   lite_llama/kernels/quantization/w8a16.py:48-216
// and consumes exactly their weight layouts:
//   int4 : qweight int32 [N, K/8], nibble j of word i = K index 8i+j; scales, zeros fp32 [N, K/g]
//   8-bit: qweight uint8(e4m3)/int8 [N, K]; scales fp32 [ceil(N/gn), ceil(K/gk)]
//
// Pipeline per CTA (see gemm_tc.cuh for the tile orientation):
//   TMA producer   raw quantised weight tile + activation tile  -> deep "raw" ring (HBM latency hiding)
//   8 dequant warps raw tile -> exact fp16 integers/e4m3 values -> UMMA operand ring (128B-swizzled, K-major)
//   MMA issuer     tcgen05.mma per 16 K-elements into one of two TMEM accumulator buffers;
//                  a buffer is closed at every scale-group boundary ("segment")
//   4 epilogue warps drain each closed buffer:  acc += scale * (P - zero * xsum)   in fp32
// Numerics: the tensor core multiplies the *unscaled* integer (or exact e4m3) weight values, so products
// and the per-group partial sums are exact in fp32; scale and zero point are applied in fp32 afterwards,
// exactly where the reference applies them (w4a16.py:104, w8a16.py:116).  For int4, xsum[m] = sum_k x[m,k]
// over the segment comes from a second MMA against an all-ones tile, so it costs no CUDA-core work.
#include "gemm_tc.cuh"
#include <type_traits>

namespace gemm {

constexpr int kDequantWarps = 8;
constexpr int kQuantThreads = (2 + 4 + kDequantWarps) * 32;   // 448
constexpr int kOpStages = 3;                                  // dequantised-operand ring depth
constexpr int kMaxRawStages = 8;

enum QuantKind { Q_INT4 = 0, Q_FP8 = 1, Q_INT8 = 2 };

struct QuantParams {
  void* y; int64_t ldy;
  const void* bias; int bias_is_f32;
  const float* scales; int64_t s_stride_n, s_stride_k;
  const float* zeros;                      // int4 only (same strides as scales)
  int group_n, group_k;                    // int4: group_n = 1, group_k = group_size
  uint32_t* ws; int* counters;
  int M, N, K;
  int splits, kb_per_split, num_kb;
  int raw_stages;
  int max_groups;                          // scale groups one CTA's k-range can touch (sizes the smem scale table)
};

struct __align__(8) QControl {
  uint64_t raw_full[kMaxRawStages];
  uint64_t raw_empty[kMaxRawStages];
  uint64_t op_full[kOpStages];
  uint64_t op_empty[kOpStages];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint32_t tmem_base;
  int is_last;
};
static_assert(sizeof(QControl) <= 1024, "control block too large");

// A raw stage holds 128 weight rows x 128 BYTES of quantised K (TMA needs full 128-byte rows to stream
// HBM efficiently: with 32/64-byte rows the first version reached 5-10 % of peak) = RW operand k-blocks
// of 64 elements (int4: 256 elements -> 4, 8-bit: 128 -> 2), plus the RW activation sub-tiles.
template <int QK> __host__ __device__ constexpr int raw_sub() { return QK == Q_INT4 ? 4 : 2; }
template <int QK, int BT> __host__ __device__ constexpr int raw_stage_bytes() {
  return kABytes + raw_sub<QK>() * BT * kStageKBytes;
}

// ---- dequantisers: 8 packed weights -> 8 fp16 values in K order (one 16-byte operand chunk) ----
__device__ __forceinline__ uint32_t lop3_and_or(uint32_t a, uint32_t mask, uint32_t orv) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(r) : "r"(a), "r"(mask), "r"(orv));  // (a & mask) | orv
  return r;
}
__device__ __forceinline__ uint4 dequant_int4_word(uint32_t w) {
  // byte j holds nibbles (2j, 2j+1).  Replicate the byte into both halves, keep nibble 2j in the low
  // half's mantissa and nibble 2j+1 (still shifted left by 4) in the high half's, OR in the fp16
  // exponent of 1024: low = 1024 + n0, high = 1024 + 16*n1; one HFMA2 maps both back to exact integers.
  const __half2 mul = __halves2half2(__float2half(1.0f), __float2half(0.0625f));
  const __half2 add = __halves2half2(__float2half(-1024.0f), __float2half(-64.0f));
  uint32_t out[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t rep = __byte_perm(w, w, 0x1111 * j);
    const uint32_t h = lop3_and_or(rep, 0x00F0000Fu, 0x64006400u);
    const __half2 v = __hfma2(*reinterpret_cast<const __half2*>(&h), mul, add);
    out[j] = *reinterpret_cast<const uint32_t*>(&v);
  }
  return make_uint4(out[0], out[1], out[2], out[3]);
}
__device__ __forceinline__ uint32_t cvt_e4m3x2_f16x2(uint16_t v) {
  uint32_t r;
  asm("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(r) : "h"(v));
  return r;
}
__device__ __forceinline__ uint4 dequant_fp8_8(uint32_t lo, uint32_t hi) {
  return make_uint4(cvt_e4m3x2_f16x2((uint16_t)(lo & 0xFFFFu)), cvt_e4m3x2_f16x2((uint16_t)(lo >> 16)),
                    cvt_e4m3x2_f16x2((uint16_t)(hi & 0xFFFFu)), cvt_e4m3x2_f16x2((uint16_t)(hi >> 16)));
}
__device__ __forceinline__ uint4 dequant_int8_8(uint32_t lo, uint32_t hi) {
  // b -> u = b + 128 (xor 0x80), fp16(0x6400 | u) = 1024 + u, minus 1152 = b   (all exact)
  const __half2 off = __float2half2_rn(-1152.0f);
  uint32_t out[4];
  const uint32_t ul = lo ^ 0x80808080u, uh = hi ^ 0x80808080u;
  const uint32_t src[2] = {ul, uh};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t sel = (j & 1) ? 0x4342u : 0x4140u;   // [0x64, byte(2i+1), 0x64, byte(2i)]
    const uint32_t h = __byte_perm(src[j >> 1], 0x64646464u, sel);
    const __half2 v = __hadd2(*reinterpret_cast<const __half2*>(&h), off);
    out[j] = *reinterpret_cast<const uint32_t*>(&v);
  }
  return make_uint4(out[0], out[1], out[2], out[3]);
}

template <int QK, int BT>
__global__ void __launch_bounds__(kQuantThreads, 1)
gemm_quant_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x,
                  const QuantParams p) {
  using T = __half;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  QControl* ctl = reinterpret_cast<QControl*>(smem);
  uint8_t* ones_tile = smem + 1024;                               // int4: 128 x 64 fp16 ones (16 KiB)
  uint8_t* op_ring = ones_tile + (QK == Q_INT4 ? kABytes : 0);    // kOpStages x 16 KiB
  uint8_t* raw_ring = op_ring + kOpStages * kABytes;              // raw_stages x (X tile, raw W tile)
  constexpr int kXBytes = BT * kStageKBytes;          // one activation sub-tile (64 K elements)
  constexpr int RW = raw_sub<QK>();
  constexpr int kRawStage = raw_stage_bytes<QK, BT>(); // [raw W 16 KiB (128B-swizzled)] [RW x X sub-tile]
  // per-row scale (and zero-point) table of this CTA's k-range: [group][128 rows] fp32, filled once with
  // cp.async by the epilogue threads (each thread only ever touches its own column -> no barrier needed)
  float* scale_tab = reinterpret_cast<float*>(raw_ring + (size_t)p.raw_stages * kRawStage);
  float* zero_tab = scale_tab + (size_t)p.max_groups * kBlockW;
  constexpr int kAccCols = QK == Q_INT4 ? 2 * BT : BT;            // per buffer: P (+ xsum)
  constexpr int kTmemCols = 2 * kAccCols < 32 ? 32 : 2 * kAccCols;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tile = blockIdx.x, m_tile = blockIdx.y, split = blockIdx.z;
  const int kb0 = split * p.kb_per_split;
  const int kb1 = kb0 + p.kb_per_split < p.num_kb ? kb0 + p.kb_per_split : p.num_kb;
  const int nkb = kb1 - kb0;
  const int seg_len = p.group_k;                  // K elements per scale group (multiple of 16)
  const int k_begin = kb0 * 64;
  int k_end = kb1 * 64; if (k_end > p.K) k_end = p.K;
  // segments: maximal runs of 16-element MMA steps inside one scale group and inside [k_begin, k_end)
  const int steps_total = (k_end - k_begin + 15) / 16;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_x);
    for (int s = 0; s < p.raw_stages; ++s) {
      mbar_init(&ctl->raw_full[s], 1);
      mbar_init(&ctl->raw_empty[s], kDequantWarps + 1);
    }
    for (int s = 0; s < kOpStages; ++s) {
      mbar_init(&ctl->op_full[s], kDequantWarps);
      mbar_init(&ctl->op_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&ctl->acc_full[s], 1);
      mbar_init(&ctl->acc_empty[s], 4);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(&ctl->tmem_base, kTmemCols);
    tmem_relinquish();
  }
  if (QK == Q_INT4 && warp >= 6) {
    // fill the all-ones operand tile (layout-independent: every element is 1.0)
    const uint32_t one2 = 0x3C003C00u;
    uint4* dst = reinterpret_cast<uint4*>(ones_tile);
    for (int i = threadIdx.x - 6 * 32; i < kABytes / 16; i += kDequantWarps * 32) dst[i] = make_uint4(one2, one2, one2, one2);
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;

  if (warp == 0) {
    // ===================== TMA producer (whole warp loops, one elected lane issues) =====================
    PipeState rs(p.raw_stages);
    const int nrb = (nkb + RW - 1) / RW;
    for (int rb = 0; rb < nrb; ++rb) {
      mbar_wait(&ctl->raw_empty[rs.stage], rs.phase ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&ctl->raw_full[rs.stage], kRawStage);   // OOB parts are zero-filled but still counted
        uint8_t* st = raw_ring + (size_t)rs.stage * kRawStage;
        const int kb = kb0 + rb * RW;                                  // first 64-element k-block of this raw block
        // raw weight tile: 128 rows x 128 bytes (int4: 32 int32 words = 256 K; 8-bit: 128 K), 128B-swizzled
        tma_load_2d_hint(st, &tmap_w, QK == Q_INT4 ? kb * 8 : kb * 64, n_tile * kBlockW, &ctl->raw_full[rs.stage],
                         kEvictFirst);
#pragma unroll
        for (int sub = 0; sub < RW; ++sub)
          tma_load_2d_hint(st + kABytes + sub * kXBytes, &tmap_x, (kb + sub) * 64, m_tile * BT,
                           &ctl->raw_full[rs.stage], kEvictLast);
      }
      __syncwarp();
      rs.advance();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp loops, one elected lane issues) =====================
    constexpr uint32_t idesc = umma_idesc(1u, 0u, 0u, 0u, 0u, kBlockW, BT);
    const uint64_t ones_desc = umma_desc_sw128_kmajor(smem_u32(ones_tile));
    const uint64_t op_desc0 = umma_desc_sw128_kmajor(smem_u32(op_ring));
    const uint64_t raw_desc0 = umma_desc_sw128_kmajor(smem_u32(raw_ring + kABytes));
    PipeState rs(p.raw_stages), os(kOpStages);
    const int seg_steps = seg_len / 16;
    int steps_left = steps_total;                       // 16-element MMA steps still to issue
    int seg_left = ((k_begin / seg_len + 1) * seg_len - k_begin) / 16;   // steps until the open segment closes
    if (seg_left > steps_left) seg_left = steps_left;
    int seg = 0;
    bool seg_open = false;
    for (int i = 0; i < nkb; ++i) {
      const int sub = i % RW;
      if (sub == 0) mbar_wait(&ctl->raw_full[rs.stage], rs.phase);   // activation sub-tiles (written by TMA)
      mbar_wait(&ctl->op_full[os.stage], os.phase);                  // dequantised weight tile
      tc_fence_after();
      // descriptors differ from the ring bases only in the 16-byte-unit start address field
      const uint64_t adesc = op_desc0 + (uint64_t)((os.stage * kABytes) >> 4);
      const uint64_t bdesc = raw_desc0 + (uint64_t)((rs.stage * kRawStage + sub * kXBytes) >> 4);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (steps_left > 0) {
          const int buf = seg & 1;
          if (!seg_open) {
            // buffer `buf` was last drained for segment seg-2
            mbar_wait(&ctl->acc_empty[buf], ((seg >> 1) & 1) ^ 1);
            tc_fence_after();
          }
          const uint32_t d = tmem_base + buf * kAccCols;
          --steps_left;
          --seg_left;
          const bool close = seg_left == 0;
          if (elect_one()) {
            umma_f16(d, adesc + 2 * k, bdesc + 2 * k, idesc, seg_open ? 1u : 0u);
            if (QK == Q_INT4) umma_f16(d + BT, ones_desc + 2 * k, bdesc + 2 * k, idesc, seg_open ? 1u : 0u);
            if (close) tc_commit(&ctl->acc_full[buf]);
          }
          seg_open = !close;
          if (close) {
            ++seg;
            seg_left = seg_steps < steps_left ? seg_steps : steps_left;
          }
        }
      }
      const bool raw_done = (sub == RW - 1) || (i + 1 == nkb);
      if (elect_one()) {
        tc_commit(&ctl->op_empty[os.stage]);
        if (raw_done) tc_commit(&ctl->raw_empty[rs.stage]);
      }
      __syncwarp();
      os.advance();
      if (raw_done) rs.advance();
    }
  } else if (warp < 6) {
    // ===================== epilogue (warps 2..5) =====================
    const int quad = warp & 3;
    const int epi_tid = quad * 32 + lane;
    const int n = n_tile * kBlockW + epi_tid;
    const bool n_ok = n < p.N;
    const int tile_id = m_tile * gridDim.x + n_tile;
    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
    float acc[BT];
#pragma unroll
    for (int j = 0; j < BT; ++j) acc[j] = 0.f;

    // stage this row's scales / zero points for every group of the CTA's k-range (asynchronously:
    // a per-segment global load would put a DRAM round trip on the drain's critical path)
    const int g_first = k_begin / seg_len;
    {
      const int g_last = (k_begin + steps_total * 16 - 1) / seg_len;
      if (n_ok) {
        const float* srow = p.scales + (int64_t)(n / p.group_n) * p.s_stride_n;
        const float* zrow = QK == Q_INT4 ? p.zeros + (int64_t)n * p.s_stride_n : nullptr;
        for (int g = g_first; g <= g_last; ++g) {
          const int slot = (g - g_first) * kBlockW + epi_tid;
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(scale_tab + slot)),
                       "l"(srow + (int64_t)g * p.s_stride_k) : "memory");
          if (QK == Q_INT4)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(zero_tab + slot)),
                         "l"(zrow + (int64_t)g * p.s_stride_k) : "memory");
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }

    // enumerate segments exactly like the MMA warp does
    int k_pos = k_begin;
    int seg = 0;
    while (k_pos < k_begin + steps_total * 16) {
      int k_next = (k_pos / seg_len + 1) * seg_len;
      const int k_lim = k_begin + steps_total * 16;
      if (k_next > k_lim) k_next = k_lim;
      const int g = k_pos / seg_len;
      float sc = 0.f, zp = 0.f;
      if (n_ok) {
        sc = scale_tab[(g - g_first) * kBlockW + epi_tid];
        if (QK == Q_INT4) zp = zero_tab[(g - g_first) * kBlockW + epi_tid];
      }
      const int buf = seg & 1;
      mbar_wait(&ctl->acc_full[buf], (seg >> 1) & 1);
      tc_fence_after();
      const uint32_t ta = taddr + buf * kAccCols;
#pragma unroll
      for (int c = 0; c < BT; c += 16) {
        uint32_t pv[16];
        tmem_ld_32x32b_x16(ta + c, pv);
        if (QK == Q_INT4) {
          uint32_t xs[16];
          tmem_ld_32x32b_x16(ta + BT + c, xs);
          tc_wait_ld();
#pragma unroll
          for (int j = 0; j < 16; ++j)
            acc[c + j] = fmaf(sc, fmaf(-zp, __uint_as_float(xs[j]), __uint_as_float(pv[j])), acc[c + j]);
        } else {
          tc_wait_ld();
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[c + j] = fmaf(sc, __uint_as_float(pv[j]), acc[c + j]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl->acc_empty[buf]);
      k_pos = k_next;
      ++seg;
    }

    T* y = reinterpret_cast<T*>(p.y);
    auto store_final = [&]() {
      if (!n_ok) return;
      float bv = 0.f;
      if (p.bias) bv = p.bias_is_f32 ? reinterpret_cast<const float*>(p.bias)[n]
                                     : __half2float(reinterpret_cast<const __half*>(p.bias)[n]);
#pragma unroll
      for (int j = 0; j < BT; ++j) {
        const int m = m_tile * BT + j;
        if (m < p.M) y[(int64_t)m * p.ldy + n] = __float2half_rn(acc[j] + bv);
      }
    };
    if (p.splits == 1) {
      store_final();
    } else {
      uint32_t* mine = p.ws + ((size_t)tile_id * p.splits + split) * (BT * kBlockW);
#pragma unroll
      for (int j = 0; j < BT; ++j) mine[j * kBlockW + epi_tid] = __float_as_uint(acc[j]);
      if (splitk_ticket(p.counters, tile_id, p.splits, epi_tid, &ctl->is_last)) {
        const uint32_t* base = p.ws + (size_t)tile_id * p.splits * (BT * kBlockW);
#pragma unroll
        for (int j = 0; j < BT; ++j) acc[j] = 0.f;
        for (int s = 0; s < p.splits; ++s) {
          const uint32_t* src = base + (size_t)s * (BT * kBlockW) + epi_tid;
#pragma unroll
          for (int j = 0; j < BT; ++j) acc[j] += __uint_as_float(__ldcg(src + j * kBlockW));
        }
        store_final();
      }
    }
  } else {
    // ===================== dequant warps (6..13) =====================
    const int dt = threadIdx.x - 6 * 32;          // 0..255
    // 8 consecutive threads = 8 consecutive rows: their 16-byte accesses hit 8 different swizzled chunks,
    // so both the raw-tile loads and the operand-tile stores are bank-conflict free
    const int row = dt & 127, hsel = dt >> 7;
    const int rx = row & 7;
    PipeState rs(p.raw_stages), os(kOpStages);
    for (int i = 0; i < nkb; ++i) {
      const int sub = i % RW;
      if (sub == 0) mbar_wait(&ctl->raw_full[rs.stage], rs.phase);
      mbar_wait(&ctl->op_empty[os.stage], os.phase ^ 1);
      const uint8_t* rawrow = raw_ring + (size_t)rs.stage * kRawStage + row * 128;
      uint8_t* arow = op_ring + (size_t)os.stage * kABytes + row * 128;
      if (QK == Q_INT4) {
        // this k-block's 64 nibbles of the row = logical 16-byte chunks 2*sub, 2*sub+1; this thread takes one
        const uint4 w4 = *reinterpret_cast<const uint4*>(rawrow + (((sub * 2 + hsel) ^ rx) << 4));
        const uint32_t ws[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int chunk = hsel * 4 + j;
          *reinterpret_cast<uint4*>(arow + ((chunk ^ (row & 7)) << 4)) = dequant_int4_word(ws[j]);
        }
      } else {
        // this k-block's 64 bytes of the row = logical chunks 4*sub .. 4*sub+3; this thread takes two
        const int c0 = sub * 4 + hsel * 2;
        const uint4 lo4 = *reinterpret_cast<const uint4*>(rawrow + ((c0 ^ rx) << 4));
        const uint4 hi4 = *reinterpret_cast<const uint4*>(rawrow + (((c0 + 1) ^ rx) << 4));
        const uint32_t ws[8] = {lo4.x, lo4.y, lo4.z, lo4.w, hi4.x, hi4.y, hi4.z, hi4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int chunk = hsel * 4 + j;
          const uint4 v = QK == Q_FP8 ? dequant_fp8_8(ws[2 * j], ws[2 * j + 1]) : dequant_int8_8(ws[2 * j], ws[2 * j + 1]);
          *reinterpret_cast<uint4*>(arow + ((chunk ^ (row & 7)) << 4)) = v;
        }
      }
      fence_proxy_async_smem();   // generic-proxy stores -> visible to the tensor core's async proxy
      __syncwarp();
      const bool raw_done = (sub == RW - 1) || (i + 1 == nkb);
      if (lane == 0) {
        mbar_arrive(&ctl->op_full[os.stage]);
        if (raw_done) mbar_arrive(&ctl->raw_empty[rs.stage]);
      }
      os.advance();
      if (raw_done) rs.advance();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

constexpr int kMaxGroupsCta = 34;   // scale-table rows per CTA (x 128 rows x 4 B, twice for int4)

static inline int groups_touched(int kb_per_split, int seg_len) {
  // worst case over split offsets: a k-range of kb_per_split*64 elements starting anywhere on a 64 boundary
  const int span = kb_per_split * 64;
  return (span + seg_len - 1) / seg_len + 1;
}

static inline SplitPlan quant_plan(int qk, int64_t M, int64_t N, int64_t K, int seg_len) {
  // token tile capped at 64: the epilogue keeps the running fp32 accumulator in registers
  SplitPlan pl;
  const int64_t mt = M > 64 ? 64 : M;
  pl.bt = pick_bt(mt);
  pl.n_tiles = (int)cdiv64(N, kBlockW);
  pl.m_tiles = (int)cdiv64(M, pl.bt);
  pl.num_kb = (int)cdiv64(K, 64);
  const int tiles = pl.n_tiles * pl.m_tiles;
  const int sms = b2l_num_sms();
  int splits = 1;
  if (tiles < sms) {
    splits = sms / tiles;
    int max_splits = pl.num_kb / 8;
    if (max_splits < 1) max_splits = 1;
    if (splits > max_splits) splits = max_splits;
    if (splits > 32) splits = 32;
  }
  const int rw = qk == Q_INT4 ? 4 : 2;     // k-blocks per raw stage: splits must start on a raw-block boundary
  pl.kb_per_split = ((pl.num_kb + splits - 1) / splits + rw - 1) / rw * rw;
  // keep the per-CTA scale table bounded: more splits when groups are fine-grained -- but keep them BALANCED
  // (equal k-ranges), an uneven last split would make one CTA of every tile the straggler
  while (pl.kb_per_split > rw && groups_touched(pl.kb_per_split, seg_len) > kMaxGroupsCta) {
    ++splits;
    pl.kb_per_split = ((pl.num_kb + splits - 1) / splits + rw - 1) / rw * rw;
  }
  pl.splits = (pl.num_kb + pl.kb_per_split - 1) / pl.kb_per_split;
  const int max_groups = groups_touched(pl.kb_per_split, seg_len);
  const int raw_stage = kABytes + rw * pl.bt * kStageKBytes;
  const int fixed = 1024 + (qk == Q_INT4 ? kABytes : 0) + kOpStages * kABytes + 1024 +
                    max_groups * kBlockW * 4 * (qk == Q_INT4 ? 2 : 1);
  int stages = (b2l_max_smem_optin() - fixed) / raw_stage;
  if (stages > kMaxRawStages) stages = kMaxRawStages;
  if (stages > pl.kb_per_split / rw + 1) stages = pl.kb_per_split / rw + 1;
  if (stages < 2) stages = 2;
  pl.stages = stages;
  pl.smem_bytes = fixed + stages * raw_stage;
  return pl;
}

template <int QK, int BT>
static int launch_q_bt(const CUtensorMap& tw, const CUtensorMap& tx, const QuantParams& p, const SplitPlan& pl,
                       cudaStream_t st) {
  auto kern = gemm_quant_kernel<QK, BT>;
  static int configured_smem = 0;
  if (pl.smem_bytes > configured_smem) {
    B2L_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, pl.smem_bytes));
    configured_smem = pl.smem_bytes;
  }
  dim3 grid((unsigned)pl.n_tiles, (unsigned)pl.m_tiles, (unsigned)pl.splits);
  kern<<<grid, kQuantThreads, pl.smem_bytes, st>>>(tw, tx, p);
  B2L_LAUNCH_CHECK();
  return B2L_OK;
}

template <int QK>
static int launch_q(const CUtensorMap& tw, const CUtensorMap& tx, const QuantParams& p, const SplitPlan& pl,
                    cudaStream_t st) {
  switch (pl.bt) {
    case 16: return launch_q_bt<QK, 16>(tw, tx, p, pl, st);
    case 32: return launch_q_bt<QK, 32>(tw, tx, p, pl, st);
    default: return launch_q_bt<QK, 64>(tw, tx, p, pl, st);
  }
}

static int quant_common(int qk, QuantParams& p, SplitPlan& pl, void* y, const void* x, const void* qweight,
                        int64_t m, int64_t n, int64_t k, int64_t x_row_stride, int64_t w_row_stride_elems,
                        int64_t y_row_stride, void* workspace, int64_t workspace_bytes, void* counters,
                        int64_t num_counters, CUtensorMap* tw, CUtensorMap* tx, const char* what, int seg_len) {
  pl = quant_plan(qk, m, n, k, seg_len);
  p.max_groups = groups_touched(pl.kb_per_split, seg_len);
  p.y = y; p.ldy = y_row_stride; p.M = (int)m; p.N = (int)n; p.K = (int)k;
  p.splits = pl.splits; p.kb_per_split = pl.kb_per_split; p.num_kb = pl.num_kb; p.raw_stages = pl.stages;
  p.ws = (uint32_t*)workspace; p.counters = (int*)counters;
  if (pl.splits > 1) {
    const int64_t need = (int64_t)pl.n_tiles * pl.m_tiles * pl.splits * pl.bt * kBlockW * 4;
    B2L_REQUIRE(workspace && workspace_bytes >= need, "%s: workspace too small (%lld < %lld)", what,
                (long long)workspace_bytes, (long long)need);
    B2L_REQUIRE(counters && num_counters >= (int64_t)pl.n_tiles * pl.m_tiles, "%s: split-K counters too small", what);
  }
  int rc;
  if (qk == Q_INT4)
    rc = b2l_make_tmap_2d(tw, qweight, TMA_I32, (uint64_t)n, (uint64_t)(k / 8), (uint64_t)w_row_stride_elems * 4,
                          kBlockW, 32, true);
  else
    rc = b2l_make_tmap_2d(tw, qweight, TMA_U8, (uint64_t)n, (uint64_t)k, (uint64_t)w_row_stride_elems, kBlockW, 128,
                          true);
  if (rc != B2L_OK) return rc;
  return b2l_make_tmap_2d(tx, x, TMA_F16, (uint64_t)m, (uint64_t)k, (uint64_t)x_row_stride * 2, (uint32_t)pl.bt, 64,
                          true);
}

}  // namespace gemm

using namespace gemm;

static inline int seg_len_for(int kind, int64_t k, int64_t group_k) {
  if (kind == Q_INT4) return (int)group_k;
  int64_t gk = group_k < k ? group_k : k;
  if (gk % 16 != 0) gk = k;
  return (int)((gk + 15) / 16 * 16);
}

extern "C" int64_t b2l_quant_gemm_workspace_bytes(int kind, int64_t m, int64_t n, int64_t k, int64_t group_k) {
  if (m <= 0 || n <= 0 || k <= 0 || group_k <= 0) return 0;
  SplitPlan pl = quant_plan(kind, m, n, k, seg_len_for(kind, k, group_k));
  if (pl.splits <= 1) return 0;
  return (int64_t)pl.n_tiles * pl.m_tiles * pl.splits * pl.bt * kBlockW * 4;
}
extern "C" int64_t b2l_quant_gemm_num_counters(int64_t m, int64_t n) {
  if (m <= 0 || n <= 0) return 0;
  return cdiv64(n, kBlockW) * cdiv64(m, pick_bt(m > 64 ? 64 : m));
}

extern "C" int b2l_w4a16_matmul(void* y, const void* x, const void* qweight, const void* scales, const void* zeros,
                                const void* bias, int bias_is_f32, int64_t m, int64_t n, int64_t k, int group_size,
                                int64_t x_row_stride, int64_t qw_row_stride, int64_t s_row_stride,
                                int64_t s_col_stride, int64_t y_row_stride, void* workspace, int64_t workspace_bytes,
                                void* counters, int64_t num_counters, void* stream) {
  B2L_REQUIRE(y && x && qweight && scales && zeros, "w4a16_matmul: null pointer");
  B2L_REQUIRE(m >= 0 && n > 0 && k > 0 && k % 8 == 0, "w4a16_matmul: bad shape m=%lld n=%lld k=%lld", (long long)m,
              (long long)n, (long long)k);
  B2L_REQUIRE(group_size > 0 && k % group_size == 0, "w4a16_matmul: K (%lld) must be a multiple of group_size (%d)",
              (long long)k, group_size);
  B2L_REQUIRE(group_size % 16 == 0, "w4a16_matmul: group_size (%d) must be a multiple of 16", group_size);
  B2L_REQUIRE(x_row_stride % 8 == 0 && qw_row_stride % 4 == 0, "w4a16_matmul: row strides must be 16-byte multiples");
  if (m == 0) return B2L_OK;
  QuantParams p; SplitPlan pl; CUtensorMap tw, tx;
  int rc = quant_common(Q_INT4, p, pl, y, x, qweight, m, n, k, x_row_stride, qw_row_stride, y_row_stride, workspace,
                        workspace_bytes, counters, num_counters, &tw, &tx, "w4a16_matmul", group_size);
  if (rc != B2L_OK) return rc;
  p.bias = bias; p.bias_is_f32 = bias_is_f32;
  p.scales = (const float*)scales; p.zeros = (const float*)zeros; p.s_stride_n = s_row_stride; p.s_stride_k = s_col_stride;
  p.group_n = 1; p.group_k = group_size;
  return launch_q<Q_INT4>(tw, tx, p, pl, (cudaStream_t)stream);
}

extern "C" int b2l_w8a16_matmul(void* y, const void* x, const void* qweight, int weight_is_fp8, const void* scales,
                                const void* bias, int bias_is_f32, int64_t m, int64_t n, int64_t k, int64_t group_n,
                                int64_t group_k, int64_t x_row_stride, int64_t qw_row_stride, int64_t s_row_stride,
                                int64_t s_col_stride, int64_t y_row_stride, void* workspace, int64_t workspace_bytes,
                                void* counters, int64_t num_counters, void* stream) {
  B2L_REQUIRE(y && x && qweight && scales, "w8a16_matmul: null pointer");
  B2L_REQUIRE(m >= 0 && n > 0 && k > 0 && k % 16 == 0, "w8a16_matmul: K must be a multiple of 16, got %lld", (long long)k);
  B2L_REQUIRE(group_n >= 1 && group_k >= 1, "w8a16_matmul: bad scale block %lldx%lld", (long long)group_n, (long long)group_k);
  int64_t gk = group_k < k ? group_k : k;
  // a scale group must contain whole 16-element MMA steps; per-row scales (gk >= K) become one segment
  if (gk % 16 != 0) gk = k;
  B2L_REQUIRE(group_k >= k || group_k % 16 == 0, "w8a16_matmul: group_k (%lld) must be a multiple of 16 unless it covers K",
              (long long)group_k);
  B2L_REQUIRE(x_row_stride % 8 == 0 && qw_row_stride % 16 == 0, "w8a16_matmul: row strides must be 16-byte multiples");
  if (m == 0) return B2L_OK;
  QuantParams p; SplitPlan pl; CUtensorMap tw, tx;
  const int qk = weight_is_fp8 ? Q_FP8 : Q_INT8;
  int rc = quant_common(qk, p, pl, y, x, qweight, m, n, k, x_row_stride, qw_row_stride, y_row_stride, workspace,
                        workspace_bytes, counters, num_counters, &tw, &tx, "w8a16_matmul", seg_len_for(qk, k, group_k));
  if (rc != B2L_OK) return rc;
  p.bias = bias; p.bias_is_f32 = bias_is_f32;
  p.scales = (const float*)scales; p.zeros = nullptr; p.s_stride_n = s_row_stride; p.s_stride_k = s_col_stride;
  p.group_n = (int)(group_n > n ? n : group_n);
  // K must be rounded up to whole 16-steps inside the kernel; scale groups beyond K are never indexed
  p.group_k = (int)((gk + 15) / 16 * 16);
  if (weight_is_fp8) return launch_q<Q_FP8>(tw, tx, p, pl, (cudaStream_t)stream);
  return launch_q<Q_INT8>(tw, tx, p, pl, (cudaStream_t)stream);
}
