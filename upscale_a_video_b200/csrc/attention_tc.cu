// attention_tc.cu — tcgen05 / TMEM FlashAttention forward for the two dense attention cores of
// the sampling path:
//   * VAE mid-block AttentionBlock: 1 head, d = 512, N = h*w tokens per frame (184 320 at
//     320x576 -> 70 TFLOP per frame, 71 % of VAE-decode FLOPs);
//   * UNet spatial self-attention: 8 heads, d = 128, N = 2880.
//
// One CTA = 128 query rows x DVT output columns, kv tiles of 128 rows.  Tensor memory (512 columns):
// S double buffered (2 x 128 fp32 columns) + the O accumulator (DVT <= 256 columns); P is written back as packed fp16
// INTO the S buffer it was computed from and consumed from there as the A operand of the P V MMAs (tcgen05.mma with
// A in tensor memory), so P costs neither shared memory nor shared-memory bandwidth.  For d = 512 the output does not
// fit (2 x 128 + 512 columns), so two CTAs each own a 256-wide half of O (QK^T is recomputed: 1.5x the minimal MMA
// work; K tiles are shared through L2).
//
// Shared memory: Q (128 x DQK fp16, up to 128 KB) + a ring of 16 KB slots.  A kv tile streams QSLABS K slots
// ([128 kv][64 d], K-major) followed by DVT/64 V slots ([128 kv][64 dv], exactly as TMA writes them = MN-major B
// operand); the ring length is chosen so that the V slots of a tile are always contiguous.
//
// Why this shape (measured on the first version: 64-row kv tiles, P through shared memory, 64 KB ring): M=128 x N=64
// MMAs read 4 KB of A + 2 KB of B from shared memory per 32 tensor-pipe cycles = 192 B/clk against the 128 B/clk the
// SM can deliver, and the ring held less than one kv tile (96 KB) -> 50 % tensor-pipe utilisation.  N=128 score MMAs
// halve the A re-reads, P from tensor memory removes them for P V, and the freed 32 KB deepen the ring.
//
// Warp roles: warp 0 TMA producer, warp 1 single-thread MMA issuer, warp 2 TMEM allocator, warps 4-7 softmax: one
// thread per query row (no shuffles), two passes over the score row in tensor memory (row max, then exp2 / pack /
// tcgen05.st), lazily updated row maximum (O and l are rescaled only when the maximum grows by more than 2^8, so the
// TMEM round trip for the correction is rare).  QK^T of tile j+1 is issued before P_j V_j so the tensor pipe works
// while the softmax of tile j runs; all buffer reuse hazards are covered by the in-order tensor pipe.
#include "uav_common.cuh"

#include <atomic>
#include <string.h>

namespace uav {
extern std::atomic<uint64_t> g_launches;

constexpr int TC_BM = 128;            // query rows per CTA
constexpr int TC_BN = 128;            // kv rows per tile
constexpr int TC_SLOT_BYTES = 16384;  // [128 rows][64 fp16]
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
  static constexpr int VBLKS = DVT / 64;
  static constexpr int Q_BYTES = QSLABS * TC_SLOT_BYTES;
  static constexpr int RING_RAW = (232448 - 1024 - 1024 - Q_BYTES) / TC_SLOT_BYTES;
  // the V slots of a tile must never wrap around the ring: with QSLABS + VBLKS slots per tile this holds when the ring
  // length divides the per-tile slot count or is a multiple of it (the V group then always starts at the same offsets)
  static constexpr int PER_TILE = QSLABS + VBLKS;
  static constexpr int STAGES = (RING_RAW >= PER_TILE) ? (RING_RAW / PER_TILE) * PER_TILE
                                                       : (PER_TILE % 6 == 0 && RING_RAW >= 6 ? 6 : 4);
  static_assert(STAGES <= RING_RAW, "ring does not fit");
  static_assert((STAGES % PER_TILE == 0) || (PER_TILE % STAGES == 0 && (QSLABS % STAGES) + VBLKS <= STAGES),
                "V slots of a tile would wrap around the ring");
  static constexpr int SMEM_BYTES = Q_BYTES + STAGES * TC_SLOT_BYTES + 1024 + 1024;
  static constexpr int TMEM_COLS = 512;  // S0 | S1 | O
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
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]),
        "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]),
        "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]: A = 128 rows (lanes) x 16 fp16 packed in 8 consecutive 32-bit columns
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// MN-major operand tile as written by a TMA box {64 contiguous elements, 128 rows}, SWIZZLE_128B:
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
  constexpr int VBLKS = Cfg::VBLKS;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sq = smem;
  uint8_t* ring = sq + Cfg::Q_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + STAGES * TC_SLOT_BYTES);
  uint64_t* full_bar = bars;                 // [STAGES]
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]
  uint64_t* q_bar = bars + 2 * STAGES;       // [1]
  uint64_t* s_full = q_bar + 1;              // [2]  QK^T of a tile landed in S[buf]
  uint64_t* p_full = s_full + 2;             // [2]  softmax wrote P into S[buf]
  uint64_t* pv_done = p_full + 2;            // [2]  P V of a tile landed in O
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);

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
      mbar_init(&p_full[i], 4);
      mbar_init(&pv_done[i], 1);
    }
    fence_barrier_init();
  }
  if (warp_idx == 2) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_o = tmem_base + 2 * TC_BN;

  if (warp_idx == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      mbar_expect_tx(q_bar, Cfg::Q_BYTES);
      for (int c = 0; c < QSLABS; ++c)
        tma_load_3d(&p.map_q, q_bar, sq + c * TC_SLOT_BYTES, h * DQK + c * 64, q0, b);
      int stage = 0;
      uint32_t phase = 0;
      auto load_slot = [&](const CUtensorMap* map, int col, int row) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_expect_tx(&full_bar[stage], TC_SLOT_BYTES);
        tma_load_3d(map, &full_bar[stage], ring + stage * TC_SLOT_BYTES, col, row, bkv);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      };
      auto load_k = [&](int j) {
        for (int c = 0; c < QSLABS; ++c) load_slot(&p.map_k, h * DQK + c * 64, j * TC_BN);
      };
      auto load_v = [&](int j) {
        for (int c = 0; c < VBLKS; ++c)
          load_slot(&p.map_v, h * (DVT * (int)gridDim.y) + dv_off + c * 64, j * TC_BN);
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
      // S[buf] = Q K_j^T.  S[buf] last held P_{j-2}, whose P V MMAs precede this in the (in-order) tensor pipe.
      auto issue_qk = [&](int j) {
        const uint32_t sb = j & 1;
        for (int c = 0; c < QSLABS; ++c) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t adesc = umma_desc_sw128(smem_u32(sq + c * TC_SLOT_BYTES));
          const uint64_t bdesc = umma_desc_sw128(smem_u32(ring + stage * TC_SLOT_BYTES));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(tmem_base + sb * TC_BN, adesc + 2 * k, bdesc + 2 * k, idesc_qk, (c | k) != 0);
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
        const uint32_t sb = j & 1;
        mbar_wait(&p_full[sb], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t tmem_p = tmem_base + sb * TC_BN;  // packed fp16 P: 64 columns
        for (int c = 0; c < VBLKS; ++c) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t bdesc = umma_desc_sw128_mn(smem_u32(ring + stage * TC_SLOT_BYTES));
#pragma unroll
          for (int k = 0; k < TC_BN / 16; ++k)  // 16 kv rows per step: A +8 columns, B +16 rows * 128 B
            umma_f16_ts(tmem_o + c * 64, tmem_p + 8 * k, bdesc + 128 * k, idesc_pv, (j | k) != 0);
          umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&pv_done[sb]);
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
      const uint32_t tmem_s = tmem_base + lane_addr + sb * TC_BN;
      mbar_wait(&s_full[sb], (j >> 1) & 1);
      tc_fence_after();
      const int kbase = j * TC_BN;
      const int valid = p.nk - kbase;  // columns >= valid are padding
      // ---- the whole score row of this thread in registers: ONE tensor-memory read per tile ----
      // (round 1 read S twice — row max, then exp — in 32-column blocks with the load latency exposed before each, through
      // the precise exp2f and with a bounds check per element: 3.4k clk per tile, more than the MMAs of the tile (1.0k at
      // d = 128, 3.1k at d = 512: ncu 29.5 % / 51.9 % tensor pipe).  Now: 4 loads in flight, 1 wait, ~4.5 instructions
      // per score (FMNMX, FFMA, MUFU.EX2, FADD, half a pack), masking only on the ragged last tile.)
      uint32_t sc[TC_BN / 32][32];
#pragma unroll
      for (int blk = 0; blk < TC_BN / 32; ++blk) tmem_ld_32x32(tmem_s + blk * 32, sc[blk]);
      tmem_ld_wait();
      const bool full = valid >= TC_BN;  // warp-uniform
      float mx = -INFINITY;
      if (full) {
#pragma unroll
        for (int blk = 0; blk < TC_BN / 32; ++blk)
#pragma unroll
          for (int c = 0; c < 32; ++c) mx = fmaxf(mx, __uint_as_float(sc[blk][c]));
      } else {
#pragma unroll
        for (int blk = 0; blk < TC_BN / 32; ++blk)
#pragma unroll
          for (int c = 0; c < 32; ++c)
            if (blk * 32 + c < valid) mx = fmaxf(mx, __uint_as_float(sc[blk][c]));
      }
      mx *= p.scale_log2;  // scale > 0
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
        mbar_wait(&pv_done[(j - 1) & 1], ((j - 1) >> 1) & 1);
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
      // ---- P = exp2(S * scale - m), packed fp16 written over the first half of the S buffer ----
      float rs = 0.f;
      const float neg_m = -m_run;
#pragma unroll
      for (int blk = 0; blk < TC_BN / 32; ++blk) {
        uint32_t pk[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          float p0 = ex2_ftz(fmaf(__uint_as_float(sc[blk][2 * c]), p.scale_log2, neg_m));
          float p1 = ex2_ftz(fmaf(__uint_as_float(sc[blk][2 * c + 1]), p.scale_log2, neg_m));
          if (!full) {
            if (blk * 32 + 2 * c >= valid) p0 = 0.f;
            if (blk * 32 + 2 * c + 1 >= valid) p1 = 0.f;
          }
          rs += p0 + p1;
          __half2 hh = __floats2half2_rn(p0, p1);
          pk[c] = *reinterpret_cast<uint32_t*>(&hh);
        }
        tmem_st_32x16(tmem_s + blk * 16, pk);  // columns [16 blk, +16): all of S is already in registers
      }
      l_run += rs;
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[sb]);
    }
    // ---- epilogue: O / l -> global ----
    mbar_wait(&pv_done[(ntiles - 1) & 1], ((ntiles - 1) >> 1) & 1);
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
  static uint64_t configured = 0;  // per-device bit: cudaFuncSetAttribute applies to the current device only
  const uint64_t dev_bit = 1ull << (current_device() & 63);
  if (!(configured & dev_bit)) {
    UAV_CHECK_CUDA(cudaFuncSetAttribute(fa_tc_kernel<DQK, DVT>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        Cfg::SMEM_BYTES));
    configured |= dev_bit;
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
