// attention_tc.cu — tcgen05 / TMEM FlashAttention forward for the two dense attention cores of
// the sampling path:
//   * VAE mid-block AttentionBlock: 1 head, d = 512, N = h*w tokens per frame (184 320 at
//     320x576 -> 70 TFLOP per frame, 71 % of VAE-decode FLOPs);
//   * UNet spatial self-attention: 8 heads, d = 128, N = 2880.
//
// One CTA = 128 query rows x DVT output columns.  TMEM cannot hold S (128 x 64 fp32, double
// buffered = 128 columns) plus a 128 x 512 fp32 output accumulator (512 columns), so for d = 512
// the output is split in two DVT = 256 halves handled by different CTAs (QK^T is recomputed:
// 1.5x the minimal MMA work; K tiles are shared through L2).
//
// Warp roles: warp 0 TMA producer (Q once; then a ring of 8 KB [64 x 64] K / V chunk tiles in
// MMA consumption order), warp 1 single-thread MMA issuer (S_j = Q K_j^T with both operands
// K-major; O += P_j V_j with P K-major from smem and V MN-major exactly as TMA wrote it),
// warp 2 TMEM allocator, warps 4-7 softmax: one thread per query row (no shuffles), tcgen05.ld
// of the score row, exp2 with a lazily updated row maximum (O and l are rescaled only when the
// maximum grows by more than 2^8, so the TMEM round trip for the correction is rare), P written
// as fp16 into a 128B-swizzled smem tile.  QK^T of tile j+1 is issued before P_j V_j so the
// tensor pipe works while the softmax of tile j runs.
#include "uav_common.cuh"

#include <atomic>
#include <string.h>

namespace uav {
extern std::atomic<uint64_t> g_launches;

constexpr int TC_BM = 128;          // query rows per CTA
constexpr int TC_BN = 64;           // kv rows per tile
constexpr int TC_CHUNK_BYTES = 8192;    // [64 rows][64 fp16]
constexpr int TC_QSLAB_BYTES = 16384;   // [128 rows][64 fp16]
constexpr int TC_PBUF_BYTES = 16384;    // [128 rows][64 fp16]
constexpr int TC_THREADS = 256;
constexpr float TC_RESCALE_THRESHOLD = 8.0f;  // log2 units

struct alignas(64) FaTcParams {
  CUtensorMap map_q, map_k, map_v;
  __half* out;
  int64_t ldo, bso;   // output token stride / batch stride (elements)
  int nq, nk, heads, kv_batch_div;
  float scale_log2;
};

template <int DQK, int DVT>
struct FaTcCfg {
  static constexpr int QSLABS = DQK / 64;
  static constexpr int VCHUNKS = DVT / 64;
  static constexpr int Q_BYTES = QSLABS * TC_QSLAB_BYTES;
  static constexpr int RING_RAW = (232448 - 1024 - 1024 - Q_BYTES - 2 * TC_PBUF_BYTES) / TC_CHUNK_BYTES;
  static constexpr int STAGES = RING_RAW > 12 ? 12 : RING_RAW;
  static constexpr int SMEM_BYTES = Q_BYTES + 2 * TC_PBUF_BYTES + STAGES * TC_CHUNK_BYTES + 1024 + 1024;
  static constexpr int TMEM_COLS = (128 + DVT) <= 256 ? 256 : 512;
};

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]),
        "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]),
        "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
        "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// MN-major operand tile as written by a TMA box {64 contiguous elements, 64 rows}, SWIZZLE_128B:
// row = one K index (128 bytes = 64 MN elements), 8-row swizzle atoms of 1024 bytes (SBO).
__device__ __forceinline__ uint64_t umma_desc_sw128_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;            // LBO: stride between 64-element MN blocks (N = 64: unused)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;    // SBO: stride between 8-row K groups
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

template <int DQK, int DVT>
__global__ void __launch_bounds__(TC_THREADS, 1)
    fa_tc_kernel(const __grid_constant__ FaTcParams p) {
  using Cfg = FaTcCfg<DQK, DVT>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int QSLABS = Cfg::QSLABS;
  constexpr int VCHUNKS = Cfg::VCHUNKS;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sq = smem;
  uint8_t* sp = sq + Cfg::Q_BYTES;             // 2 P buffers
  uint8_t* ring = sp + 2 * TC_PBUF_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + STAGES * TC_CHUNK_BYTES);
  uint64_t* full_bar = bars;                 // [STAGES]
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]
  uint64_t* q_bar = bars + 2 * STAGES;       // [1]
  uint64_t* s_full = q_bar + 1;              // [2]
  uint64_t* s_free = s_full + 2;             // [2]
  uint64_t* p_full = s_free + 2;             // [2]
  uint64_t* p_free = p_full + 2;             // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p_free + 2);

  const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * TC_BM;
  const int dv_off = blockIdx.y * DVT;          // which slice of the value / output columns
  const int bh = blockIdx.z;
  const int b = bh / p.heads, h = bh % p.heads;
  const int bkv = b / p.kv_batch_div;
  const int ntiles = (p.nk + TC_BN - 1) / TC_BN;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&p.map_q);
    tma_prefetch_desc(&p.map_k);
    tma_prefetch_desc(&p.map_v);
  }
  if (warp_idx == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(q_bar, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], 4);
      mbar_init(&p_full[i], 4);
      mbar_init(&p_free[i], 1);
    }
    fence_barrier_init();
  }
  if (warp_idx == 2) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_o = tmem_base + 128;

  if (warp_idx == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      mbar_expect_tx(q_bar, Cfg::Q_BYTES);
      for (int c = 0; c < QSLABS; ++c)
        tma_load_3d(&p.map_q, q_bar, sq + c * TC_QSLAB_BYTES, h * DQK + c * 64, q0, b);
      int stage = 0;
      uint32_t phase = 0;
      auto load_chunk = [&](const CUtensorMap* map, int col, int row) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_expect_tx(&full_bar[stage], TC_CHUNK_BYTES);
        tma_load_3d(map, &full_bar[stage], ring + stage * TC_CHUNK_BYTES, col, row, bkv);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      };
      auto load_k = [&](int j) {
        for (int c = 0; c < QSLABS; ++c) load_chunk(&p.map_k, h * DQK + c * 64, j * TC_BN);
      };
      auto load_v = [&](int j) {
        for (int c = 0; c < VCHUNKS; ++c)
          load_chunk(&p.map_v, h * (DVT * (int)gridDim.y) + dv_off + c * 64, j * TC_BN);
      };
      // same order as the MMA warp consumes: K0, {K(j+1), V(j)}...
      load_k(0);
      for (int j = 0; j < ntiles; ++j) {
        if (j + 1 < ntiles) load_k(j + 1);
        load_v(j);
      }
    }
  } else if (warp_idx == 1) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      constexpr uint32_t idesc_qk = umma_idesc(0, TC_BM, TC_BN);
      constexpr uint32_t idesc_pv = umma_idesc(0, TC_BM, 64) | (1u << 16);  // B is MN-major
      int stage = 0;
      uint32_t phase = 0;
      auto issue_qk = [&](int j) {
        const uint32_t sb = j & 1;
        mbar_wait(&s_free[sb], ((j >> 1) & 1) ^ 1);
        tc_fence_after();
        for (int c = 0; c < QSLABS; ++c) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t adesc = umma_desc_sw128(smem_u32(sq + c * TC_QSLAB_BYTES));
          const uint64_t bdesc = umma_desc_sw128(smem_u32(ring + stage * TC_CHUNK_BYTES));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(tmem_base + sb * 64, adesc + 2 * k, bdesc + 2 * k, idesc_qk, (c | k) != 0);
          umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&s_full[sb]);
      };
      mbar_wait(q_bar, 0);
      tc_fence_after();
      issue_qk(0);
      for (int j = 0; j < ntiles; ++j) {
        if (j + 1 < ntiles) issue_qk(j + 1);
        const uint32_t pb = j & 1;
        mbar_wait(&p_full[pb], (j >> 1) & 1);
        tc_fence_after();
        for (int c = 0; c < VCHUNKS; ++c) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t adesc = umma_desc_sw128(smem_u32(sp + pb * TC_PBUF_BYTES));
          const uint64_t bdesc = umma_desc_sw128_mn(smem_u32(ring + stage * TC_CHUNK_BYTES));
#pragma unroll
          for (int k = 0; k < 4; ++k)  // 16 kv rows per step: A +32 B (K-major), B +16 rows * 128 B
            umma_f16(tmem_o + c * 64, adesc + 2 * k, bdesc + 128 * k, idesc_pv, (j | k) != 0);
          umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&p_free[pb]);
      }
    }
  } else if (warp_idx >= 4) {
    // =============================== softmax / correction / epilogue ===============================
    const int quad = warp_idx & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < ntiles; ++j) {
      const uint32_t sb = j & 1;
      mbar_wait(&s_full[sb], (j >> 1) & 1);
      tc_fence_after();
      uint32_t s0[32], s1[32];
      tmem_ld_32x32(tmem_base + lane_addr + sb * 64, s0);
      tmem_ld_32x32(tmem_base + lane_addr + sb * 64 + 32, s1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[sb]);

      const int kbase = j * TC_BN;
      float x[64];
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 64; ++c) {
        float v = __uint_as_float(c < 32 ? s0[c & 31] : s1[c & 31]) * p.scale_log2;
        if (kbase + c >= p.nk) v = -INFINITY;
        x[c] = v;
        mx = fmaxf(mx, v);
      }
      float alpha = 1.f;
      bool need = false;
      if (j == 0) {
        m_run = mx;
      } else if (mx > m_run + TC_RESCALE_THRESHOLD) {
        alpha = exp2f(m_run - mx);
        m_run = mx;
        need = true;
      }
      if (__any_sync(0xffffffffu, need)) {
        // all PVs up to tile j-1 must have landed in O before it is rescaled
        mbar_wait(&p_free[(j - 1) & 1], ((j - 1) >> 1) & 1);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < DVT / 32; ++c) {
          uint32_t o[32];
          tmem_ld_32x32(tmem_o + lane_addr + c * 32, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st_32x32(tmem_o + lane_addr + c * 32, o);
        }
        tmem_st_wait();
        l_run *= alpha;
      }
      float rs = 0.f;
      uint32_t pk[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const float p0 = exp2f(x[2 * c] - m_run), p1 = exp2f(x[2 * c + 1] - m_run);
        rs += p0 + p1;
        __half2 hh = __floats2half2_rn(p0, p1);
        pk[c] = *reinterpret_cast<uint32_t*>(&hh);
      }
      l_run += rs;
      // P buffer sb is free once the PV of tile j-2 completed
      mbar_wait(&p_free[sb], ((j >> 1) & 1) ^ 1);
      uint8_t* prow = sp + sb * TC_PBUF_BYTES + row * 128;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const uint4 v = make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
        *reinterpret_cast<uint4*>(prow + ((g ^ (row & 7)) << 4)) = v;
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[sb]);
    }
    // ---- epilogue: O / l -> global ----
    mbar_wait(&p_free[(ntiles - 1) & 1], ((ntiles - 1) >> 1) & 1);
    tc_fence_after();
    const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
    const bool row_ok = q0 + row < p.nq;
    __half* orow = p.out + static_cast<int64_t>(b) * p.bso + static_cast<int64_t>(q0 + row) * p.ldo +
                   static_cast<int64_t>(h) * (DVT * gridDim.y) + dv_off;
#pragma unroll 1
    for (int c = 0; c < DVT / 32; ++c) {
      uint32_t o[32];
      tmem_ld_32x32(tmem_o + lane_addr + c * 32, o);
      tmem_ld_wait();
      if (row_ok) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 v;
          uint32_t* vw = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            __half2 hh = __floats2half2_rn(__uint_as_float(o[g * 8 + 2 * i]) * inv,
                                           __uint_as_float(o[g * 8 + 2 * i + 1]) * inv);
            vw[i] = *reinterpret_cast<uint32_t*>(&hh);
          }
          stg16(orow + c * 32 + g * 8, v);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

static uav_status_t make_map_3d(CUtensorMap* map, const void* base, int64_t cols, int64_t rows,
                                int64_t batch, int64_t ld, int64_t bs, uint32_t box_rows) {
  PFN_encodeTiled encode = get_encode_tiled();
  UAV_REQUIRE(encode != nullptr, "attention_tc: cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)bs * 2};
  cuuint32_t box[3] = {64, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides,
                      box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  UAV_REQUIRE(r == CUDA_SUCCESS, "attention_tc: cuTensorMapEncodeTiled failed with %d", (int)r);
  return UAV_OK;
}

template <int DQK, int DVT>
static uav_status_t launch_fa_tc(FaTcParams& p, int64_t batch, int dv_splits, cudaStream_t stream) {
  using Cfg = FaTcCfg<DQK, DVT>;
  static bool configured = false;
  if (!configured) {
    UAV_CHECK_CUDA(cudaFuncSetAttribute(fa_tc_kernel<DQK, DVT>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        Cfg::SMEM_BYTES));
    configured = true;
  }
  dim3 grid((p.nq + TC_BM - 1) / TC_BM, dv_splits, (unsigned)(batch * p.heads));
  fa_tc_kernel<DQK, DVT><<<grid, TC_THREADS, Cfg::SMEM_BYTES, stream>>>(p);
  UAV_CHECK_CUDA(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return UAV_OK;
}

// entry used by uav_attention (attention.cu) for head_dim 128 and 512
uav_status_t attention_tc(const void* q, const void* k, const void* v, void* out, int64_t batch,
                          int heads, int head_dim, int64_t nq, int64_t nk, int64_t ldq, int64_t ldk,
                          int64_t ldv, int64_t ldo, int64_t kv_batch_div, float scale,
                          cudaStream_t stream) {
  UAV_REQUIRE((reinterpret_cast<uintptr_t>(q) & 15) == 0 && (reinterpret_cast<uintptr_t>(k) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(v) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
              "attention_tc: pointers must be 16-byte aligned");
  FaTcParams p;
  memset(&p, 0, sizeof(p));
  const int64_t C = (int64_t)heads * head_dim;
  uav_status_t st;
  if ((st = make_map_3d(&p.map_q, q, C, nq, batch, ldq, nq * ldq, TC_BM)) != UAV_OK) return st;
  if ((st = make_map_3d(&p.map_k, k, C, nk, batch / kv_batch_div, ldk, nk * ldk, TC_BN)) != UAV_OK) return st;
  if ((st = make_map_3d(&p.map_v, v, C, nk, batch / kv_batch_div, ldv, nk * ldv, TC_BN)) != UAV_OK) return st;
  p.out = reinterpret_cast<__half*>(out);
  p.ldo = ldo;
  p.bso = nq * ldo;
  p.nq = (int)nq;
  p.nk = (int)nk;
  p.heads = heads;
  p.kv_batch_div = (int)kv_batch_div;
  p.scale_log2 = scale * 1.4426950408889634f;
  if (head_dim == 512) {
    UAV_REQUIRE(heads == 1, "attention_tc: head_dim 512 supports a single head");
    return launch_fa_tc<512, 256>(p, batch, 2, stream);
  }
  if (head_dim == 128) return launch_fa_tc<128, 128>(p, batch, 1, stream);
  set_last_error("attention_tc: head_dim %d unsupported", head_dim);
  return UAV_ERR_UNSUPPORTED;
}

}  // namespace uav
