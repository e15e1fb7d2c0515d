// igemm.cu — the tcgen05 implicit-GEMM kernel behind every convolution and Linear on the
// Upscale-A-Video sampling path.
//
//   out[pixel][n] = epilogue( sum_{tap, c} A[pixel + off(tap)][c] * W[n][tap][c] )
//
// Design (B200-first, not a cuDNN translation):
//  * activations stay channels-last, so an M-tile of 128 output pixels is a rectangular box
//    of the input tensor and ONE TMA box load per filter tap (shifted coordinates, hardware
//    zero fill outside the image = the convolution's zero padding) lands directly in the
//    128B-swizzled K-major layout tcgen05.mma consumes — no im2col buffer, no layout copies
//    (the reference pays two permute copies per conv: resnet.py:97-99).
//  * persistent CTAs (one per SM), warp-specialised: warp 0 = activation TMA producer, warp 3 = weight TMA
//    producer, warp 1 = MMA issuer (single thread), warp 2 = TMEM allocator, warps 4-11 = epilogue
//    (tcgen05.ld -> bias/temb/act/residual -> swizzled smem staging -> TMA store).  Two TMEM accumulators so
//    the epilogue of tile i overlaps the main loop of tile i+1; smem rings of {A 128x64, B BLOCK_Nx64} fp16 tiles.
//  * large launches (>= 2 M-tiles per SM, BLOCK_N >= 128) run as CTA PAIRS: tcgen05.mma.cta_group::2 with
//    M=256 across the two SMs of a TPC, each CTA holding its 128 activation rows and half of the weight tile
//    (5 x 32 KB stages instead of 3 x 48 KB; half the L2->SM weight traffic).  Measured on B200 (interleaved,
//    thermally settled): +12..29 % on the single-tap GEMMs (Linear 512->512 / 512->1536 / 2048->512, temporal
//    conv), parity on the MMA-bound 3x3 convs.  Smaller launches use cta_group::1 (M=128).
//  * stride-2 convs read a 5-D "phase" view (2C, W/2, 2, H/2, NB) of the same buffer, the
//    temporal (k,1,1) conv a (C, HW, T, B) view, Conv3d a (C, W, H, T, B) view, Linear a
//    (K, M) view: all the same kernel, only the tensor map and the tap table differ.
#include <stdarg.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <mutex>

#include "uav_common.cuh"

namespace uav {

// ---------------------------------------------------------------------------------------
// host-side globals shared by all translation units
// ---------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
std::atomic<uint64_t> g_launches{0};

int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev;
}

int num_sms() {
  static int n[64] = {0};  // per device (a process may drive several GPUs)
  const int dev = current_device() & 63;
  if (n[dev] == 0) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    n[dev] = v > 0 ? v : 148;
  }
  return n[dev];
}

PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

// ---------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------
constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // fp16 elements: one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int MAX_TAPS = 49;  // up to 7 x 7 (RAFT motion encoder)
constexpr int NUM_THREADS = 384;      // warps 0-3: TMA / MMA / TMEM alloc / idle; warps 4-11: epilogue
constexpr int NUM_EPI_THREADS = 256;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr int SLAB_BYTES = BLOCK_M * 128;  // 128 rows x 64 fp16 output columns, 128B-swizzled

struct alignas(64) IgemmParams {
  CUtensorMap map_a;
  CUtensorMap map_b;
  CUtensorMap map_out;           // only valid when tma_store != 0
  CUtensorMap map_res;           // residual operand as a TMA tensor (same box geometry as map_out); valid when res_mode != 0
  int32_t tap_off[MAX_TAPS][5];  // coordinate offset of each tap (dim0 = channel offset)
  int32_t num_taps;
  int32_t kblocks_per_tap;
  int32_t k_per_tap;
  uint32_t box[5];       // box[0] = 64, box[1..4] = M-tile extents (product 128)
  uint32_t tiles[5];     // tiles along dims 1..4
  uint32_t out_dims[5];  // output extents along dims 1..4
  uint32_t n_tiles, num_tiles;
  int32_t stages_a, stages_b;  // depth of the activation / weight smem rings (runtime split of the same smem)
  uint32_t num_pairs;  // cluster mode: ceil(m_tiles / 2) * n_tiles work items, two M-tiles each
  int32_t N;      // rows of B
  int32_t n_out;  // output columns (N, or N/2 with GEGLU)
  int32_t tma_store;
  // residual operand of the TMA-store epilogue: 1 = loaded into the output staging tile itself as soon as the previous
  // store released it (multi-tap convolutions: the epilogue waits for the accumulator anyway, so the load hides behind the
  // main loop and costs no shared memory); 2 = loaded ONE TILE AHEAD into a dedicated buffer carved from the end of the
  // ring (single-tap GEMMs: HBM / epilogue bound, an exposed load latency per tile would be the critical path)
  int32_t res_mode;
  int32_t staging_tiles;  // 1, or 2 (GEGLU): output tiles alternate between two staging buffers
  const float* bias;
  const __half* rowvec;
  int64_t rows_per_vec, ld_rowvec;
  const __half* residual;
  int64_t ld_res;
  int32_t act, out_dtype;
  int64_t ld_out;
  void* out;
  float out_scale;     // applied before the residual add (1 = off)
  float* gn_partial;   // [n_out / 8][gn_blocks][2] GroupNorm statistics of the output, or nullptr
  int64_t gn_blocks;
  // LayerNorm folded into the GEMM (see uav_epilogue_t): per-row {sum, sumsq} slots of the INPUT rows written by its
  // producer, the column sums of the gamma-scaled weight, and (as a producer) the slots of the OUTPUT rows
  const float2* ln_in;
  const float* ln_colsum;
  int32_t ln_slots;
  float ln_inv_c, ln_eps;
  float2* ln_out;
  int32_t ln_out_slots;
};

template <int BLOCK_N, bool GEGLU>
struct IgemmCfg {
  static constexpr int OUT_TILE_N = GEGLU ? BLOCK_N / 2 : BLOCK_N;
  static constexpr int B_STAGE_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int TILE_STAGING_BYTES = (OUT_TILE_N >= 64) ? (OUT_TILE_N / 64) * SLAB_BYTES : 0;
  // GEGLU tiles are 128 output columns (32 KB): two staging tiles, so the TMA store of tile i drains while the epilogue of
  // tile i + 1 fills the other one (with a single tile every epilogue starts by waiting for the previous store)
  static constexpr int STAGING_TILES = GEGLU ? 2 : 1;
  static constexpr int STAGING_BYTES = TILE_STAGING_BYTES * STAGING_TILES;
  static constexpr int AUX_BYTES = 2048;  // barriers + tmem slot + bias tile (256 floats)
  // everything left of the 227 KB after the output staging tile is one region shared by the A and B rings
  static constexpr int RING_BYTES = ((232448 - 1024 - STAGING_BYTES - AUX_BYTES) / 1024) * 1024;
  static constexpr int STAGES_RAW = RING_BYTES / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;  // balanced depth
  static constexpr int TMEM_COLS = (2 * BLOCK_N <= 32)    ? 32
                                   : (2 * BLOCK_N <= 64)  ? 64
                                   : (2 * BLOCK_N <= 128) ? 128
                                   : (2 * BLOCK_N <= 256) ? 256
                                                          : 512;
  static constexpr int SMEM_BYTES = RING_BYTES + STAGING_BYTES + AUX_BYTES + 1024 /*align*/;
};

__device__ __forceinline__ uint32_t pack_half2(float a, float b) { return pack_half2_sat(a, b); }
__device__ __forceinline__ void epi_bar_sync() {
  asm volatile("bar.sync 1, %0;" ::"n"(NUM_EPI_THREADS) : "memory");
}
// pointwise activations of the fused epilogue (GEGLU is handled separately: it pairs two accumulator columns)
__device__ __forceinline__ float apply_act(float x, int act) {
  switch (act) {
    case UAV_ACT_SILU: return silu_f(x);
    case UAV_ACT_RELU: return fmaxf(x, 0.f);
    case UAV_ACT_SIGMOID: return rcp_ftz(1.0f + ex2_ftz(-1.4426950408889634f * x));
    case UAV_ACT_TANH: return 1.0f - 2.0f * rcp_ftz(1.0f + ex2_ftz(2.8853900817779268f * x));
    case UAV_ACT_GELU: return gelu_erf_f(x);
    case UAV_ACT_QUICK_GELU: return x * rcp_ftz(1.0f + ex2_ftz(-2.4554669595930156f * x));  // x * sigmoid(1.702 x)
    default: return x;
  }
}
// x = x * alpha + residual
__device__ __forceinline__ void fma_half8(float (&x)[8], float alpha, const uint4& q) {
  const __half2* h2 = reinterpret_cast<const __half2*>(&q);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __half22float2(h2[j]);
    x[2 * j] = fmaf(x[2 * j], alpha, f.x);
    x[2 * j + 1] = fmaf(x[2 * j + 1], alpha, f.y);
  }
}
// sum of 8 values per lane over the 32 lanes of a warp in 7 shuffles (transpose-reduce): afterwards lane L holds the
// total of value index ((L >> 4) & 1) * 4 + ((L >> 3) & 1) * 2 + ((L >> 2) & 1) (every lane of its group of 4)
__device__ __forceinline__ float warp_reduce8(float (&v)[8], int lane) {
  bool hi = (lane & 16) != 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = hi ? v[i] : v[i + 4];
    const float keep = hi ? v[i + 4] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
  hi = (lane & 8) != 0;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = hi ? v[i] : v[i + 2];
    const float keep = hi ? v[i + 2] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  hi = (lane & 4) != 0;
  {
    const float send = hi ? v[0] : v[1];
    const float keep = hi ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 2);
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
  return v[0];
}
__device__ __forceinline__ void add_half8(float (&x)[8], const uint4& q) {
  const __half2* h2 = reinterpret_cast<const __half2*>(&q);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __half22float2(h2[j]);
    x[2 * j] += f.x;
    x[2 * j + 1] += f.y;
  }
}

// TMA_EPI: smem-staged TMA-store epilogue (aligned fp16 output, >= 64-column tiles) vs per-row direct stores.
// AUX: the epilogue has a row vector / residual / SiLU on top of the bias (compiled out otherwise: the hot
// bias-only GEMMs get a small loop body that stays in the instruction cache).
// CL: 1 = independent CTAs; 2 = CTA pairs (tcgen05 cta_group::2) computing a 256 x BLOCK_N tile: each CTA holds its
// own 128 activation rows and HALF of the weight tile (rows [rank * BLOCK_N/2, +BLOCK_N/2)); the leader (rank 0) issues
// M=256 MMAs that read both CTAs' shared memory and write both CTAs' tensor memory.  A weight stage shrinks to half
// (deeper rings in the same shared memory: 5 x 32 KB instead of 3 x 48 KB at BLOCK_N = 256, which is what hides the
// HBM latency of the single-tap GEMMs) and the L2->SM weight traffic halves.  Barrier protocol: the full barriers
// live in the leader (both producers arrive.expect_tx on them and both CTAs' TMA loads complete_tx there), the empty /
// accumulator-full barriers are per CTA and signalled by multicast commits, the accumulator-empty barrier lives in the
// leader and collects the epilogue warps of both CTAs.
template <int BLOCK_N, bool GEGLU, bool TMA_EPI, bool AUX, int CL>
__global__ void __launch_bounds__(NUM_THREADS, 1)
    igemm_kernel(const __grid_constant__ IgemmParams p) {
  using Cfg = IgemmCfg<BLOCK_N, GEGLU>;
  static_assert(!TMA_EPI || Cfg::OUT_TILE_N >= 64, "TMA-store epilogue needs >= 64-column output tiles");
  constexpr int OUT_TILE_N = Cfg::OUT_TILE_N;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment (SWIZZLE_128B atoms) in the shared address space
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  // two independent rings in the same RING_BYTES region: activations (A) come from HBM on the
  // K <= 1024 GEMMs and need more bytes in flight than the weight tiles (B), which are L2 hits
  const int SA = p.stages_a, SB = p.stages_b;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + SA * A_STAGE_BYTES;
  constexpr int B_STAGE = Cfg::B_STAGE_BYTES / CL;  // bytes of one weight stage in THIS CTA
  uint8_t* staging0 = smem + Cfg::RING_BYTES;
  uint8_t* aux = staging0 + Cfg::STAGING_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(aux);
  constexpr int MAXS = 10;
  uint64_t* fulla_bar = bars;                     // [MAXS]
  uint64_t* emptya_bar = bars + MAXS;             // [MAXS]
  uint64_t* fullb_bar = bars + 2 * MAXS;          // [MAXS]
  uint64_t* emptyb_bar = bars + 3 * MAXS;         // [MAXS]
  uint64_t* tfull_bar = bars + 4 * MAXS;          // [2]
  uint64_t* tempty_bar = bars + 4 * MAXS + 2;     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4 * MAXS + 4);
  uint64_t* res_bar = bars + 4 * MAXS + 5;        // [1] residual tile landed in shared memory
  float* sbias = reinterpret_cast<float*>(aux + 512);  // [BLOCK_N]
  // residual tile: the staging tile itself (mode 1) or the last STAGING_BYTES of the ring region (mode 2)
  uint8_t* res_tile_ahead = smem + Cfg::RING_BYTES - Cfg::TILE_STAGING_BYTES;

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr bool tma_store = TMA_EPI;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&p.map_a);
    tma_prefetch_desc(&p.map_b);
    if (tma_store) tma_prefetch_desc(&p.map_out);
    if (tma_store && p.res_mode != 0) tma_prefetch_desc(&p.map_res);
  }
  if (warp_idx == 1 && lane == 0) {
    for (int s = 0; s < SA; ++s) {
      mbar_init(&fulla_bar[s], 1);  // CL == 2: the leader's producer expects the bytes of both CTAs
      mbar_init(&emptya_bar[s], 1);
    }
    for (int s = 0; s < SB; ++s) {
      mbar_init(&fullb_bar[s], 1);
      mbar_init(&emptyb_bar[s], 1);
    }
    mbar_init(&tfull_bar[0], 1);
    mbar_init(&tfull_bar[1], 1);
    mbar_init(&tempty_bar[0], CL * (tma_store ? 8 : 4));
    mbar_init(&tempty_bar[1], CL * (tma_store ? 8 : 4));
    mbar_init(res_bar, 1);
    fence_barrier_init();
  }
  if (warp_idx == 2) {
    if (CL == 2) tmem_alloc_2sm(tmem_slot, Cfg::TMEM_COLS);
    else tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  if (CL == 2) cluster_sync_all();  // peer barriers initialised / tensor memory allocated in both CTAs
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_kb = p.num_taps * p.kblocks_per_tap;
  // work items: a tile (CL == 1) or a pair of M-tiles sharing one N-tile (CL == 2)
  const uint32_t cta_rank = (CL == 2) ? cluster_ctarank() : 0u;
  const uint32_t work0 = (CL == 2) ? (blockIdx.x >> 1) : blockIdx.x;
  const uint32_t work_stride = (CL == 2) ? (gridDim.x >> 1) : gridDim.x;
  const uint32_t num_work = (CL == 2) ? p.num_pairs : p.num_tiles;
  // CL == 2: the leader's barriers as seen through the cluster window (identity mapping for the leader itself)
  const uint32_t fulla_leader = (CL == 2) ? mapa_shared(smem_u32(fulla_bar), 0) : 0u;
  const uint32_t fullb_leader = (CL == 2) ? mapa_shared(smem_u32(fullb_bar), 0) : 0u;
  const uint32_t tempty_leader = (CL == 2) ? mapa_shared(smem_u32(tempty_bar), 0) : 0u;

  if (warp_idx == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (uint32_t w = work0; w < num_work; w += work_stride) {
        uint32_t idx = (w / p.n_tiles) * CL + cta_rank;  // M-tile (may be a ghost tile past the end: all OOB)
        const int c1 = (idx % p.tiles[1]) * p.box[1];
        idx /= p.tiles[1];
        const int c2 = (idx % p.tiles[2]) * p.box[2];
        idx /= p.tiles[2];
        const int c3 = (idx % p.tiles[3]) * p.box[3];
        idx /= p.tiles[3];
        const int c4 = idx * p.box[4];
        for (int tap = 0; tap < p.num_taps; ++tap) {
          const int o0 = p.tap_off[tap][0], o1 = p.tap_off[tap][1], o2 = p.tap_off[tap][2],
                    o3 = p.tap_off[tap][3], o4 = p.tap_off[tap][4];
          for (int kc = 0; kc < p.kblocks_per_tap; ++kc) {
            mbar_wait(&emptya_bar[stage], phase ^ 1);
            if (CL == 2) {
              // the leader expects the activation bytes of both CTAs; each CTA's box lands in its own shared memory
              if (cta_rank == 0) mbar_expect_tx(&fulla_bar[stage], 2 * A_STAGE_BYTES);
              tma_load_5d_2sm(&p.map_a, fulla_leader + 8 * stage, smem_a + stage * A_STAGE_BYTES,
                              kc * BLOCK_K + o0, c1 + o1, c2 + o2, c3 + o3, c4 + o4);
            } else {
              mbar_expect_tx(&fulla_bar[stage], A_STAGE_BYTES);
              tma_load_5d(&p.map_a, &fulla_bar[stage], smem_a + stage * A_STAGE_BYTES,
                          kc * BLOCK_K + o0, c1 + o1, c2 + o2, c3 + o3, c4 + o4);
            }
            if (++stage == SA) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp_idx == 3) {
    // =============================== TMA producer: weights ===============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (uint32_t w = work0; w < num_work; w += work_stride) {
        const uint32_t n_tile = w % p.n_tiles;
        for (int tap = 0; tap < p.num_taps; ++tap) {
          for (int kc = 0; kc < p.kblocks_per_tap; ++kc) {
            mbar_wait(&emptyb_bar[stage], phase ^ 1);
            const int kcoord = tap * p.k_per_tap + kc * BLOCK_K;
            uint8_t* sb = smem_b + stage * B_STAGE;
            if (CL == 2) {
              // this CTA's half of the weight tile (GEGLU: rank 0 = value rows, rank 1 = gate rows)
              const int brow = GEGLU ? (cta_rank == 0 ? n_tile * (BLOCK_N / 2) : p.N / 2 + n_tile * (BLOCK_N / 2))
                                     : (n_tile * BLOCK_N + cta_rank * (BLOCK_N / 2));
              if (cta_rank == 0) mbar_expect_tx(&fullb_bar[stage], 2 * B_STAGE);
              tma_load_2d_2sm(&p.map_b, fullb_leader + 8 * stage, sb, kcoord, brow);
              if (++stage == SB) {
                stage = 0;
                phase ^= 1;
              }
              continue;
            }
            mbar_expect_tx(&fullb_bar[stage], Cfg::B_STAGE_BYTES);
            if (GEGLU) {
              tma_load_2d(&p.map_b, &fullb_bar[stage], sb, kcoord, n_tile * (BLOCK_N / 2));
              tma_load_2d(&p.map_b, &fullb_bar[stage], sb + Cfg::B_STAGE_BYTES / 2, kcoord,
                          p.N / 2 + n_tile * (BLOCK_N / 2));
            } else {
              tma_load_2d(&p.map_b, &fullb_bar[stage], sb, kcoord, n_tile * BLOCK_N);
            }
            if (++stage == SB) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp_idx == 1) {
    // =============================== MMA issuer ===============================
    if (lane == 0 && cta_rank == 0) {
      constexpr uint32_t idesc = umma_idesc(0 /*f16*/, BLOCK_M * CL, BLOCK_N);
      int sa = 0, sb = 0;
      uint32_t pha = 0, phb = 0;
      uint32_t it = 0;
      for (uint32_t w = work0; w < num_work; w += work_stride, ++it) {
        const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
        if (CL == 2) mbar_wait_cluster(&tempty_bar[acc], acc_phase ^ 1);
        else mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&fulla_bar[sa], pha);
          mbar_wait(&fullb_bar[sb], phb);
          tc_fence_after();
          const uint64_t adesc = umma_desc_sw128(smem_u32(smem_a + sa * A_STAGE_BYTES));
          const uint64_t bdesc = umma_desc_sw128(smem_u32(smem_b + sb * B_STAGE));
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // advance 16 elements = 32 bytes along K inside the 128B swizzle row: +2 (>>4)
            if (CL == 2) umma_f16_2sm(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
            else umma_f16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          }
          if (CL == 2) {
            umma_commit_2sm(&emptya_bar[sa], (uint16_t)3);
            umma_commit_2sm(&emptyb_bar[sb], (uint16_t)3);
            if (kb == num_kb - 1) umma_commit_2sm(&tfull_bar[acc], (uint16_t)3);
          } else {
            umma_commit(&emptya_bar[sa]);
            umma_commit(&emptyb_bar[sb]);
            if (kb == num_kb - 1) umma_commit(&tfull_bar[acc]);
          }
          if (++sa == SA) {
            sa = 0;
            pha ^= 1;
          }
          if (++sb == SB) {
            sb = 0;
            phb ^= 1;
          }
        }
      }
    }
  } else if (warp_idx >= 4) {
    // =============================== epilogue ===============================
    const int quad = warp_idx & 3;        // TMEM lane quadrant this warp may access
    const int half = (warp_idx - 4) >> 2;  // column half handled by this warp (TMA-store path)
    const int et = threadIdx.x - 128;      // 0..255
    const uint32_t row = quad * 32 + lane;
    uint32_t r = row;
    const uint32_t l1 = r % p.box[1];
    r /= p.box[1];
    const uint32_t l2 = r % p.box[2];
    r /= p.box[2];
    const uint32_t l3 = r % p.box[3];
    r /= p.box[3];
    const uint32_t l4 = r;
    uint32_t res_phase = 0;
    // one thread: TMA-load the residual tile (64-column slabs, the layout the epilogue stores) and arm res_bar
    auto issue_res_load = [&](uint8_t* res_smem, int nb, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t c4) {
      uint32_t bytes = 0;
#pragma unroll
      for (int sl = 0; sl < (OUT_TILE_N >= 64 ? OUT_TILE_N / 64 : 0); ++sl)
        if (nb + sl * 64 < p.n_out) bytes += SLAB_BYTES;
      mbar_expect_tx(res_bar, bytes);
#pragma unroll
      for (int sl = 0; sl < (OUT_TILE_N >= 64 ? OUT_TILE_N / 64 : 0); ++sl)
        if (nb + sl * 64 < p.n_out)
          tma_load_5d(&p.map_res, res_bar, res_smem + sl * SLAB_BYTES, nb + sl * 64, (int)c1, (int)c2, (int)c3, (int)c4);
    };
    if (tma_store || half == 0) {
      uint32_t it = 0;
      for (uint32_t w = work0; w < num_work; w += work_stride, ++it) {
        const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
        const uint32_t n_tile = w % p.n_tiles;
        uint32_t idx = (w / p.n_tiles) * CL + cta_rank;
        const uint32_t t1 = (idx % p.tiles[1]) * p.box[1];
        idx /= p.tiles[1];
        const uint32_t t2 = (idx % p.tiles[2]) * p.box[2];
        idx /= p.tiles[2];
        const uint32_t t3 = (idx % p.tiles[3]) * p.box[3];
        idx /= p.tiles[3];
        const uint32_t t4 = idx * p.box[4];
        const uint32_t o1 = t1 + l1, o2 = t2 + l2, o3 = t3 + l3, o4 = t4 + l4;
        const bool row_ok = o1 < p.out_dims[1] && o2 < p.out_dims[2] && o3 < p.out_dims[3] &&
                            o4 < p.out_dims[4];
        const int64_t out_row =
            ((static_cast<int64_t>(o4) * p.out_dims[3] + o3) * p.out_dims[2] + o2) *
                p.out_dims[1] + o1;
        const __half* rv = (p.rowvec != nullptr && row_ok)
                               ? p.rowvec + (out_row / p.rows_per_vec) * p.ld_rowvec
                               : nullptr;
        const __half* res =
            (p.residual != nullptr && row_ok) ? p.residual + out_row * p.ld_res : nullptr;
        const int n_base = n_tile * OUT_TILE_N;
        const uint32_t taddr =
            tmem_base + acc * BLOCK_N + (static_cast<uint32_t>(quad * 32) << 16);

        if constexpr (TMA_EPI) {
          {
            uint8_t* staging = staging0 + (p.staging_tiles == 2 ? (it & 1) * Cfg::TILE_STAGING_BYTES : 0);
            uint8_t* res_smem = (p.res_mode == 2) ? res_tile_ahead : staging;
            // ---- stage the bias tile, make sure the previous TMA store released the staging ----
            if (GEGLU) {
              const int j = et & (OUT_TILE_N - 1);
              const int n = (et < OUT_TILE_N) ? (n_base + j) : (p.N / 2 + n_base + j);
              sbias[et] = (p.bias != nullptr && n_base + j < p.n_out) ? __ldg(p.bias + n) : 0.f;
            } else if (et < OUT_TILE_N) {
              sbias[et] = (p.bias != nullptr && n_base + et < p.n_out) ? __ldg(p.bias + n_base + et)
                                                                       : 0.f;
            }
            constexpr int COLS_PER_HALF = OUT_TILE_N / 2;
            constexpr int CHUNKS = COLS_PER_HALF / 32;
            if (et == 0) {  // the store that last used THIS staging tile must have read it
              if (p.staging_tiles == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
              else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            if constexpr (AUX) {
              // mode 1: the staging tile is free now -> fetch this tile's residual into it (in-place epilogue); the first
              // tile of mode 2 is fetched here too (later ones are issued one tile ahead, right after the store below)
              if (et == 0 && (p.res_mode == 1 || (p.res_mode == 2 && it == 0))) issue_res_load(res_smem, n_base, t1, t2, t3, t4);
            }
            epi_bar_sync();
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            if constexpr (AUX) {
              if (p.res_mode != 0) {
                mbar_wait(res_bar, res_phase);
                res_phase ^= 1;
              }
            }
            // LayerNorm of the input rows folded into this GEMM: acc was computed on the RAW rows x against W' = W * gamma, so
            // LN(x) W^T + b = rstd * (acc - mean * colsum(W')) + b'  with the row statistics emitted by x's producer
            float ln_a = 1.0f, ln_b = 0.0f;
            float ln_s = 0.f, ln_q = 0.f;  // statistics of THIS tile's output row (for the next LayerNorm)
            if constexpr (AUX) {
              if (p.ln_in != nullptr) {
                float s_ = 0.f, q_ = 0.f;
                if (row_ok) {
                  const float2* lp = p.ln_in + out_row * p.ln_slots;
                  for (int i = 0; i < p.ln_slots; ++i) {
                    const float2 t = __ldg(lp + i);
                    s_ += t.x;
                    q_ += t.y;
                  }
                }
                const float mean = s_ * p.ln_inv_c;
                const float var = fmaxf(fmaf(-mean, mean, q_ * p.ln_inv_c), 0.f);
                ln_a = rsqrtf(var + p.ln_eps);
                ln_b = -mean * ln_a;
              }
            }
#pragma unroll 1
            for (int c = 0; c < CHUNKS; ++c) {
              const int col0 = half * COLS_PER_HALF + c * 32;  // column inside the output tile
              const int n0 = n_base + col0;
              const bool cols_ok = n0 < p.n_out;  // n_out % 8 == 0; groups of 8 checked below
              // issue the global loads first so their latency overlaps the TMEM load
              uint4 vq[4];
              if constexpr (AUX) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                  vq[g] = make_uint4(0, 0, 0, 0);
                  if (cols_ok && n0 + g * 8 < p.n_out) {
                    if (rv != nullptr) vq[g] = ldg16(rv + n0 + g * 8);
                  }
                }
              }
              uint32_t a[32];
              tmem_ld_32x32(taddr + col0, a);
              float v[32];
              if constexpr (GEGLU) {
                uint32_t gt[32];
                tmem_ld_32x32(taddr + BLOCK_N / 2 + col0, gt);
                tmem_ld_wait();
                if (AUX && p.ln_in != nullptr) {
                  const float* csh = p.ln_colsum + n0;            // value rows of W'
                  const float* csg = p.ln_colsum + p.N / 2 + n0;  // gate rows
#pragma unroll
                  for (int j4 = 0; j4 < 8; ++j4) {
                    const float4 bh = *reinterpret_cast<const float4*>(sbias + col0 + j4 * 4);
                    const float4 bg = *reinterpret_cast<const float4*>(sbias + OUT_TILE_N + col0 + j4 * 4);
                    const float4 ch = cols_ok ? __ldg(reinterpret_cast<const float4*>(csh + j4 * 4)) : make_float4(0, 0, 0, 0);
                    const float4 cg = cols_ok ? __ldg(reinterpret_cast<const float4*>(csg + j4 * 4)) : make_float4(0, 0, 0, 0);
                    v[j4 * 4 + 0] = fmaf(__uint_as_float(a[j4 * 4 + 0]), ln_a, fmaf(ch.x, ln_b, bh.x)) *
                                    gelu_erf_f(fmaf(__uint_as_float(gt[j4 * 4 + 0]), ln_a, fmaf(cg.x, ln_b, bg.x)));
                    v[j4 * 4 + 1] = fmaf(__uint_as_float(a[j4 * 4 + 1]), ln_a, fmaf(ch.y, ln_b, bh.y)) *
                                    gelu_erf_f(fmaf(__uint_as_float(gt[j4 * 4 + 1]), ln_a, fmaf(cg.y, ln_b, bg.y)));
                    v[j4 * 4 + 2] = fmaf(__uint_as_float(a[j4 * 4 + 2]), ln_a, fmaf(ch.z, ln_b, bh.z)) *
                                    gelu_erf_f(fmaf(__uint_as_float(gt[j4 * 4 + 2]), ln_a, fmaf(cg.z, ln_b, bg.z)));
                    v[j4 * 4 + 3] = fmaf(__uint_as_float(a[j4 * 4 + 3]), ln_a, fmaf(ch.w, ln_b, bh.w)) *
                                    gelu_erf_f(fmaf(__uint_as_float(gt[j4 * 4 + 3]), ln_a, fmaf(cg.w, ln_b, bg.w)));
                  }
                } else {
#pragma unroll
                  for (int j4 = 0; j4 < 8; ++j4) {  // bias tiles as 128-bit broadcast loads (2 scalar LDS per output before)
                    const float4 bh = *reinterpret_cast<const float4*>(sbias + col0 + j4 * 4);
                    const float4 bg = *reinterpret_cast<const float4*>(sbias + OUT_TILE_N + col0 + j4 * 4);
                    v[j4 * 4 + 0] = (__uint_as_float(a[j4 * 4 + 0]) + bh.x) * gelu_erf_f(__uint_as_float(gt[j4 * 4 + 0]) + bg.x);
                    v[j4 * 4 + 1] = (__uint_as_float(a[j4 * 4 + 1]) + bh.y) * gelu_erf_f(__uint_as_float(gt[j4 * 4 + 1]) + bg.y);
                    v[j4 * 4 + 2] = (__uint_as_float(a[j4 * 4 + 2]) + bh.z) * gelu_erf_f(__uint_as_float(gt[j4 * 4 + 2]) + bg.z);
                    v[j4 * 4 + 3] = (__uint_as_float(a[j4 * 4 + 3]) + bh.w) * gelu_erf_f(__uint_as_float(gt[j4 * 4 + 3]) + bg.w);
                  }
                }
              } else {
                tmem_ld_wait();
                if (AUX && p.ln_in != nullptr) {
#pragma unroll
                  for (int j4 = 0; j4 < 8; ++j4) {
                    const float4 bb = *reinterpret_cast<const float4*>(sbias + col0 + j4 * 4);
                    const float4 cc = (n0 + j4 * 4 < p.n_out) ? __ldg(reinterpret_cast<const float4*>(p.ln_colsum + n0 + j4 * 4))
                                                              : make_float4(0, 0, 0, 0);
                    v[j4 * 4 + 0] = fmaf(__uint_as_float(a[j4 * 4 + 0]), ln_a, fmaf(cc.x, ln_b, bb.x));
                    v[j4 * 4 + 1] = fmaf(__uint_as_float(a[j4 * 4 + 1]), ln_a, fmaf(cc.y, ln_b, bb.y));
                    v[j4 * 4 + 2] = fmaf(__uint_as_float(a[j4 * 4 + 2]), ln_a, fmaf(cc.z, ln_b, bb.z));
                    v[j4 * 4 + 3] = fmaf(__uint_as_float(a[j4 * 4 + 3]), ln_a, fmaf(cc.w, ln_b, bb.w));
                  }
                } else {
#pragma unroll
                  for (int j4 = 0; j4 < 8; ++j4) {
                    const float4 bb = *reinterpret_cast<const float4*>(sbias + col0 + j4 * 4);
                    v[j4 * 4 + 0] = __uint_as_float(a[j4 * 4 + 0]) + bb.x;
                    v[j4 * 4 + 1] = __uint_as_float(a[j4 * 4 + 1]) + bb.y;
                    v[j4 * 4 + 2] = __uint_as_float(a[j4 * 4 + 2]) + bb.z;
                    v[j4 * 4 + 3] = __uint_as_float(a[j4 * 4 + 3]) + bb.w;
                  }
                }
              }
              const int slab = col0 >> 6;
              const int chunk_base = (col0 & 63) >> 3;  // 16-byte chunk index inside the 128B row
              uint8_t* srow = staging + slab * SLAB_BYTES + row * 128;
              float st[8];  // {sum, sum of squares} of this row's four 8-column groups (GroupNorm statistics)
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                float x[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) x[j] = v[g * 8 + j];
                if constexpr (AUX) {
                  if (rv != nullptr) add_half8(x, vq[g]);
                  if (p.act != UAV_ACT_NONE) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) x[j] = apply_act(x[j], p.act);
                  }
                  if (p.res_mode != 0) {  // residual chunk of this row from shared memory (same swizzle as the store)
                    const uint4 rq = *reinterpret_cast<const uint4*>(res_smem + slab * SLAB_BYTES + row * 128 +
                                                                     (((chunk_base + g) ^ (row & 7)) << 4));
                    fma_half8(x, p.out_scale, rq);
                  } else if (p.out_scale != 1.0f) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) x[j] *= p.out_scale;
                  }
                  if (p.ln_out != nullptr && n0 + g * 8 < p.n_out) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                      ln_s += x[j];
                      ln_q = fmaf(x[j], x[j], ln_q);
                    }
                  }
                  if (p.gn_partial != nullptr) {
                    float sm = 0.f, sq = 0.f;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                      sm += x[j];
                      sq = fmaf(x[j], x[j], sq);
                    }
                    const bool ok = row_ok && (n0 + g * 8 < p.n_out);
                    st[2 * g] = ok ? sm : 0.f;
                    st[2 * g + 1] = ok ? sq : 0.f;
                  }
                }
                uint4 o;
                o.x = pack_half2(x[0], x[1]);
                o.y = pack_half2(x[2], x[3]);
                o.z = pack_half2(x[4], x[5]);
                o.w = pack_half2(x[6], x[7]);
                const int phys = (chunk_base + g) ^ (row & 7);  // CU_TENSOR_MAP_SWIZZLE_128B
                *reinterpret_cast<uint4*>(srow + phys * 16) = o;
              }
              if constexpr (AUX) {
                if (p.gn_partial != nullptr) {  // warp-uniform
                  const float tot = warp_reduce8(st, lane);
                  const int vi = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
                  const int oct = (n0 >> 3) + (vi >> 1);
                  const int64_t blk = static_cast<int64_t>((w / p.n_tiles) * CL + cta_rank) * 4 + quad;
                  if ((lane & 3) == 0 && blk < p.gn_blocks && oct * 8 < p.n_out)
                    p.gn_partial[(static_cast<int64_t>(oct) * p.gn_blocks + blk) * 2 + (vi & 1)] = tot;
                }
              }
            }
            if constexpr (AUX) {
              if (p.ln_out != nullptr && row_ok)
                p.ln_out[out_row * p.ln_out_slots + n_tile * 2 + half] = make_float2(ln_s, ln_q);
            }
            // accumulator fully read: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (CL == 2) mbar_arrive_cluster_relaxed(tempty_leader + 8 * acc);
              else mbar_arrive(&tempty_bar[acc]);
            }
            // publish the staged tile to the async proxy and store it with TMA (clips OOB rows)
            fence_proxy_async();
            epi_bar_sync();
            if (et == 0) {
#pragma unroll
              for (int sl = 0; sl < OUT_TILE_N / 64; ++sl) {
                if (n_base + sl * 64 < p.n_out) {
                  asm volatile(
                      "cp.async.bulk.tensor.5d.global.shared::cta.bulk_group"
                      " [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(
                          reinterpret_cast<uint64_t>(&p.map_out)),
                      "r"(smem_u32(staging + sl * SLAB_BYTES)), "r"(n_base + sl * 64), "r"(t1),
                      "r"(t2), "r"(t3), "r"(t4)
                      : "memory");
                }
              }
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
              if constexpr (AUX) {
                // mode 2: every thread has consumed this tile's residual (barrier above) -> fetch the next tile's
                if (p.res_mode == 2 && w + work_stride < num_work) {
                  const uint32_t wn = w + work_stride;
                  uint32_t i2 = (wn / p.n_tiles) * CL + cta_rank;
                  const uint32_t u1 = (i2 % p.tiles[1]) * p.box[1];
                  i2 /= p.tiles[1];
                  const uint32_t u2 = (i2 % p.tiles[2]) * p.box[2];
                  i2 /= p.tiles[2];
                  const uint32_t u3 = (i2 % p.tiles[3]) * p.box[3];
                  i2 /= p.tiles[3];
                  issue_res_load(res_tile_ahead, (wn % p.n_tiles) * OUT_TILE_N, u1, u2, u3, i2 * p.box[4]);
                }
              }
            }
          }
        } else {
          // ---- direct path (small N, fp32 output, unaligned): 4 warps, per-row stores ----
          constexpr int CHUNK = (OUT_TILE_N >= 32) ? 32 : 16;
          const bool vec_ok = (p.ld_out % 8 == 0) && (p.n_out % 8 == 0) &&
                              (p.residual == nullptr || p.ld_res % 8 == 0) &&
                              (p.rowvec == nullptr || p.ld_rowvec % 8 == 0);
          mbar_wait(&tfull_bar[acc], acc_phase);
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < OUT_TILE_N / CHUNK; ++c) {
            const int n0 = n_base + c * CHUNK;
            uint32_t a[CHUNK];
            float v[CHUNK];
            if constexpr (CHUNK == 32) {
              tmem_ld_32x32(taddr + c * 32, a);
              if constexpr (GEGLU) {
                uint32_t gt[32];
                tmem_ld_32x32(taddr + BLOCK_N / 2 + c * 32, gt);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  float hh = __uint_as_float(a[j]), gg = __uint_as_float(gt[j]);
                  if (p.bias != nullptr && n0 + j < p.n_out) {
                    hh += __ldg(p.bias + n0 + j);
                    gg += __ldg(p.bias + p.N / 2 + n0 + j);
                  }
                  v[j] = hh * gelu_erf_f(gg);
                }
              } else {
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  v[j] = __uint_as_float(a[j]);
                  if (p.bias != nullptr && n0 + j < p.n_out) v[j] += __ldg(p.bias + n0 + j);
                }
              }
            } else {
              tmem_ld_32x16(taddr + c * 16, a);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                v[j] = __uint_as_float(a[j]);
                if (p.bias != nullptr && n0 + j < p.n_out) v[j] += __ldg(p.bias + n0 + j);
              }
            }
            if (!row_ok || n0 >= p.n_out) continue;
            if (vec_ok) {
#pragma unroll
              for (int g8 = 0; g8 < CHUNK / 8; ++g8) {
                const int n = n0 + g8 * 8;
                if (n < p.n_out) {
                  float x[8];
#pragma unroll
                  for (int j = 0; j < 8; ++j) x[j] = v[g8 * 8 + j];
                  if (rv != nullptr) add_half8(x, ldg16(rv + n));
                  if (p.act != UAV_ACT_NONE) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) x[j] = apply_act(x[j], p.act);
                  }
                  if (res != nullptr) fma_half8(x, p.out_scale, ldg16(res + n));
                  else if (p.out_scale != 1.0f) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) x[j] *= p.out_scale;
                  }
                  if (p.out_dtype == UAV_F16) {
                    uint4 o;
                    o.x = pack_half2(x[0], x[1]);
                    o.y = pack_half2(x[2], x[3]);
                    o.z = pack_half2(x[4], x[5]);
                    o.w = pack_half2(x[6], x[7]);
                    stg16(reinterpret_cast<__half*>(p.out) + out_row * p.ld_out + n, o);
                  } else {
                    float* op = reinterpret_cast<float*>(p.out) + out_row * p.ld_out + n;
                    *reinterpret_cast<float4*>(op) = make_float4(x[0], x[1], x[2], x[3]);
                    *reinterpret_cast<float4*>(op + 4) = make_float4(x[4], x[5], x[6], x[7]);
                  }
                }
              }
            } else {
#pragma unroll
              for (int j = 0; j < CHUNK; ++j) {
                const int n = n0 + j;
                if (n < p.n_out) {
                  float x = v[j];
                  if (rv != nullptr) x += __half2float(rv[n]);
                  x = apply_act(x, p.act) * p.out_scale;
                  if (res != nullptr) x += __half2float(res[n]);
                  if (p.out_dtype == UAV_F16)
                    reinterpret_cast<__half*>(p.out)[out_row * p.ld_out + n] =
                        __float2half_rn(fminf(fmaxf(x, -65504.f), 65504.f));
                  else
                    reinterpret_cast<float*>(p.out)[out_row * p.ld_out + n] = x;
                }
              }
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
              if (CL == 2) mbar_arrive_cluster_relaxed(tempty_leader + 8 * acc);
              else mbar_arrive(&tempty_bar[acc]);
            }
        }
      }
      if (TMA_EPI && et == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
  }

  // teardown
  tc_fence_before();
  __syncthreads();
  if (CL == 2) cluster_sync_all();  // the leader may still read / signal this CTA's shared and tensor memory
  if (warp_idx == 2) {
    tc_fence_after();
    if (CL == 2) tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
    else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------
// host: descriptor + launch
// ---------------------------------------------------------------------------------------
struct IgemmDesc {
  const void* a;
  int rank_a;                // always encoded as rank 5
  uint64_t a_dims[5];        // dim0 = channels
  uint64_t a_strides[5];     // elements; a_strides[0] == 1
  uint32_t box[5];           // box[0] = 64
  uint32_t tiles[5];
  uint32_t out_dims[5];
  int num_taps;
  int32_t tap_off[MAX_TAPS][5];
  int k_per_tap;
  const void* w;
  int64_t N;
  void* out;
  const uav_epilogue_t* epi;
  // optional strided output view (elements) for dims 1..4; 0 = dense (derived from ld_out / out_dims).
  // Only the TMA-store epilogue understands it (used by the fused nearest-x2 upsample + 3x3 conv).
  uint64_t out_strides[5];
};

template <int BLOCK_N, bool GEGLU, bool TMA_EPI, bool AUX, int CL>
static uav_status_t launch_instance3(IgemmParams& p, cudaStream_t stream) {
  using Cfg = IgemmCfg<BLOCK_N, GEGLU>;
  static uint64_t configured = 0;  // per-device bit: cudaFuncSetAttribute applies to the current device only
  const uint64_t dev_bit = 1ull << (current_device() & 63);
  auto kern = igemm_kernel<BLOCK_N, GEGLU, TMA_EPI, AUX, CL>;
  if (!(configured & dev_bit)) {
    UAV_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured |= dev_bit;
  }
  const uint32_t sms = (uint32_t)num_sms();
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  if (CL == 2) {
    const uint32_t clusters = p.num_pairs < sms / 2 ? p.num_pairs : sms / 2;
    cfg.gridDim = dim3(2 * clusters);
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  } else {
    cfg.gridDim = dim3(p.num_tiles < sms ? p.num_tiles : sms);
  }
  UAV_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return UAV_OK;
}

template <int BLOCK_N, bool GEGLU, bool TMA_EPI, bool AUX>
static uav_status_t launch_instance2(IgemmParams& p, bool cluster, cudaStream_t stream) {
  if constexpr (BLOCK_N >= 128) {
    if (cluster) return launch_instance3<BLOCK_N, GEGLU, TMA_EPI, AUX, 2>(p, stream);
  }
  return launch_instance3<BLOCK_N, GEGLU, TMA_EPI, AUX, 1>(p, stream);
}

template <int BLOCK_N, bool GEGLU>
static uav_status_t launch_instance(IgemmParams& p, bool cluster, cudaStream_t stream) {
  const bool aux = p.rowvec != nullptr || p.residual != nullptr || (p.act != UAV_ACT_NONE && p.act != UAV_ACT_GEGLU) ||
                   p.out_scale != 1.0f || p.gn_partial != nullptr || p.ln_in != nullptr || p.ln_out != nullptr;
  if constexpr (IgemmCfg<BLOCK_N, GEGLU>::OUT_TILE_N >= 64) {
    if (p.tma_store) {
      return aux ? launch_instance2<BLOCK_N, GEGLU, true, true>(p, cluster, stream)
                 : launch_instance2<BLOCK_N, GEGLU, true, false>(p, cluster, stream);
    }
  }
  return launch_instance2<BLOCK_N, GEGLU, false, true>(p, cluster, stream);
}

static uav_status_t launch_igemm(const IgemmDesc& d, cudaStream_t stream) {
  const uav_epilogue_t* e = d.epi;
  UAV_REQUIRE(e != nullptr, "igemm: epilogue descriptor is NULL");
  UAV_REQUIRE(d.a && d.w && d.out, "igemm: null pointer");
  UAV_REQUIRE(d.k_per_tap > 0 && d.k_per_tap % 8 == 0,
              "igemm: input channels (%d) must be a positive multiple of 8", d.k_per_tap);
  UAV_REQUIRE((reinterpret_cast<uintptr_t>(d.a) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(d.w) & 15) == 0,
              "igemm: operands must be 16-byte aligned");
  const bool geglu = e->act == UAV_ACT_GEGLU;
  UAV_REQUIRE(!geglu || (d.N % 256 == 0), "igemm: GEGLU needs N %% 256 == 0 (N=%lld)",
              (long long)d.N);
  PFN_encodeTiled encode = get_encode_tiled();
  UAV_REQUIRE(encode != nullptr, "igemm: cuTensorMapEncodeTiled entry point unavailable");

  IgemmParams p;
  memset(&p, 0, sizeof(p));
  int block_n;
  if (geglu) block_n = 256;
  else if (d.N > 128) block_n = 256;
  else if (d.N > 64) block_n = 128;
  else if (d.N > 32) block_n = 64;
  else if (d.N > 16) block_n = 32;
  else block_n = 16;

  // CTA pairs (cta_group::2) when there are enough M-tiles to keep every SM busy
  uint64_t m_tiles_pre = 1;
  for (int i = 1; i < 5; ++i) m_tiles_pre *= d.tiles[i];
  static const bool cluster_enabled = !(getenv("UAV_IGEMM_CLUSTER") && getenv("UAV_IGEMM_CLUSTER")[0] == '0');
  const bool use_cluster = cluster_enabled && block_n >= 128 && m_tiles_pre >= 2ull * (uint64_t)num_sms();

  // A map
  {
    cuuint64_t dims[5], strides[4];
    cuuint32_t box[5], estr[5] = {1, 1, 1, 1, 1};
    for (int i = 0; i < 5; ++i) {
      dims[i] = d.a_dims[i];
      box[i] = d.box[i];
      UAV_REQUIRE(dims[i] >= 1 && box[i] >= 1 && box[i] <= 256, "igemm: bad A dim/box %d", i);
    }
    for (int i = 1; i < 5; ++i) {
      strides[i - 1] = d.a_strides[i] * 2;
      UAV_REQUIRE(strides[i - 1] % 16 == 0, "igemm: A stride %d not 16-byte aligned", i);
    }
    // L2 promotion 256B: a 128-byte k-block fetch also brings the neighbouring 128 bytes of the row into L2, i.e. the
    // next k-block of the same rows (half the DRAM transactions, better page locality for the streaming GEMMs)
    static const int promo = getenv("UAV_IGEMM_A_PROMO") ? atoi(getenv("UAV_IGEMM_A_PROMO")) : 256;
    CUresult r = encode(&p.map_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(d.a),
                        dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B,
                        promo == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                     : (promo == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                                    : CU_TENSOR_MAP_L2_PROMOTION_L2_128B),
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    UAV_REQUIRE(r == CUDA_SUCCESS, "igemm: cuTensorMapEncodeTiled(A) failed with %d", (int)r);
  }
  // B map: [N][K_total] K-major
  const int64_t k_total = (int64_t)d.num_taps * d.k_per_tap;
  {
    cuuint64_t dims[2] = {(cuuint64_t)k_total, (cuuint64_t)d.N};
    cuuint64_t strides[1] = {(cuuint64_t)k_total * 2};
    // cluster mode: each CTA of a pair fetches (and multicasts) half of the weight tile
    cuuint32_t box[2] = {64, (cuuint32_t)((geglu || use_cluster) ? block_n / 2 : block_n)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&p.map_b, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(d.w), dims,
                        strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    UAV_REQUIRE(r == CUDA_SUCCESS, "igemm: cuTensorMapEncodeTiled(B) failed with %d", (int)r);
  }
  UAV_REQUIRE(d.num_taps >= 1 && d.num_taps <= MAX_TAPS, "igemm: bad tap count %d", d.num_taps);
  memcpy(p.tap_off, d.tap_off, sizeof(int32_t) * 5 * d.num_taps);
  p.num_taps = d.num_taps;
  p.k_per_tap = d.k_per_tap;
  p.kblocks_per_tap = (d.k_per_tap + BLOCK_K - 1) / BLOCK_K;
  uint64_t m_tiles = 1;
  uint32_t box_prod = 1;
  for (int i = 0; i < 5; ++i) {
    p.box[i] = d.box[i];
    p.tiles[i] = d.tiles[i];
    p.out_dims[i] = d.out_dims[i];
    if (i >= 1) {
      m_tiles *= d.tiles[i];
      box_prod *= d.box[i];
    }
  }
  UAV_REQUIRE(box_prod == BLOCK_M && d.box[0] == 64, "igemm: M-tile box must cover 128 rows");
  p.N = (int32_t)d.N;
  p.n_out = (int32_t)(geglu ? d.N / 2 : d.N);
  const int out_tile_n = geglu ? block_n / 2 : block_n;
  p.n_tiles = (uint32_t)((p.n_out + out_tile_n - 1) / out_tile_n);
  UAV_REQUIRE(m_tiles * p.n_tiles < (1ull << 31), "igemm: too many tiles");
  p.num_tiles = (uint32_t)(m_tiles * p.n_tiles);
  p.num_pairs = (uint32_t)(((m_tiles + 1) / 2) * p.n_tiles);
  // residual operand of the TMA-store epilogue (decided here because mode 2 takes its buffer out of the ring)
  auto aligned16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  const bool can_tma = out_tile_n >= 64 && e->out_dtype == UAV_F16 && p.n_out % 8 == 0 && e->ld_out % 8 == 0 &&
                       aligned16(d.out) &&
                       (e->residual == nullptr || (e->ld_res % 8 == 0 && aligned16(e->residual))) &&
                       (e->rowvec == nullptr || (e->ld_rowvec % 8 == 0 && aligned16(e->rowvec)));
  {
    // ring split inside the fixed RING_BYTES region.  Convolutions re-read their activation tiles from L2
    // (9 taps) -> balanced rings.  Single-tap GEMMs stream activations from HBM (high latency) against L2-resident
    // weights -> deep A ring, shallow B ring.
    const int b_stage = block_n * BLOCK_K * 2 / (use_cluster ? 2 : 1);  // CTA pairs hold half a weight tile each
    const int staging = out_tile_n >= 64 ? (out_tile_n / 64) * SLAB_BYTES : 0;  // one output tile
    static const bool double_staging = !(getenv("UAV_IGEMM_DOUBLE_STAGING") && getenv("UAV_IGEMM_DOUBLE_STAGING")[0] == '0');
    p.staging_tiles = (geglu && double_staging) ? 2 : 1;
    int ring = ((232448 - 1024 - staging * (geglu ? 2 : 1) - 2048) / 1024) * 1024;  // == IgemmCfg::RING_BYTES
    p.res_mode = 0;
    if (can_tma && e->residual != nullptr) {
      static const int res_knob = getenv("UAV_IGEMM_RES_MODE") ? atoi(getenv("UAV_IGEMM_RES_MODE")) : -1;
      // short-K single-tap GEMMs are HBM / epilogue bound: residual one tile ahead in its own buffer (if >= 3 stages
      // remain).  With K >= 1024 the ring stage it costs is worth more than the early residual (tools/bench_linear512.py,
      // B200: K=2048 N=512 1443 -> 1322 us, K=1024 N=1024 117 -> 108 us in mode 1; K=512 541 vs 548 us, M=184320 167 vs 184)
      const bool ahead = d.num_taps == 1 && d.k_per_tap <= 512 && (ring - staging) / (A_STAGE_BYTES + b_stage) >= 3;
      p.res_mode = res_knob > 0 ? res_knob : (ahead ? 2 : 1);
      if (p.res_mode == 2 && (ring - staging) / (A_STAGE_BYTES + b_stage) < 2) p.res_mode = 1;
      if (p.res_mode == 2) ring -= staging;
    }
    int sb = ring / (A_STAGE_BYTES + b_stage);  // balanced depth
    if (sb > 10) sb = 10;
    int sa = sb;
    // single-tap GEMMs: a shallower weight ring (L2 hits) buys activation stages (HBM latency).  UAV_IGEMM_B_STAGES=n
    // sets the weight depth; UAV_IGEMM_DEEP_A=1 is n = 2 (measured on B200: the weight ring becomes the limiter)
    static const int b_knob = getenv("UAV_IGEMM_B_STAGES") ? atoi(getenv("UAV_IGEMM_B_STAGES"))
                              : ((getenv("UAV_IGEMM_DEEP_A") && getenv("UAV_IGEMM_DEEP_A")[0] == '1') ? 2 : 0);
    if (b_knob >= 2 && d.num_taps == 1 && sb > b_knob) {
      sb = b_knob;
      sa = (ring - sb * b_stage) / A_STAGE_BYTES;
      if (sa > 10) sa = 10;
    }
    p.stages_a = sa;
    p.stages_b = sb;
  }
  p.bias = e->bias;
  p.rowvec = reinterpret_cast<const __half*>(e->rowvec);
  p.rows_per_vec = e->rows_per_vec > 0 ? e->rows_per_vec : 1;
  p.ld_rowvec = e->ld_rowvec;
  p.residual = reinterpret_cast<const __half*>(e->residual);
  p.ld_res = e->ld_res;
  p.act = e->act;
  p.out_dtype = e->out_dtype;
  p.ld_out = e->ld_out;
  p.out = d.out;
  p.out_scale = e->out_scale == 0.0f ? 1.0f : e->out_scale;
  p.gn_partial = reinterpret_cast<float*>(e->gn_partial);
  p.gn_blocks = e->gn_blocks;
  p.ln_in = reinterpret_cast<const float2*>(e->ln_in);
  p.ln_colsum = e->ln_colsum;
  p.ln_slots = e->ln_slots;
  p.ln_inv_c = 1.0f / static_cast<float>(d.k_per_tap);
  p.ln_eps = e->ln_eps;
  p.ln_out = reinterpret_cast<float2*>(e->ln_out);
  p.ln_out_slots = 0;
  UAV_REQUIRE(!(geglu && p.out_scale != 1.0f), "igemm: out_scale is not supported with GEGLU");
  UAV_REQUIRE(p.ld_out >= p.n_out, "igemm: ld_out (%lld) < output columns (%d)",
              (long long)p.ld_out, p.n_out);
  UAV_REQUIRE(e->out_dtype == UAV_F16 || e->out_dtype == UAV_F32, "igemm: bad out_dtype");
  UAV_REQUIRE(e->act >= UAV_ACT_NONE && e->act <= UAV_ACT_QUICK_GELU, "igemm: bad activation");
  if (p.num_tiles == 0) return UAV_OK;

  // TMA-store epilogue (smem-staged, fully coalesced, clips partial tiles) whenever the output
  // is an aligned fp16 tensor with at least 64-column tiles; otherwise per-row direct stores.
  p.tma_store = can_tma ? 1 : 0;
  if (p.ln_in != nullptr || p.ln_out != nullptr) {
    UAV_REQUIRE(can_tma && d.num_taps == 1 && d.out_strides[1] == 0,
                "igemm: LayerNorm folding needs a Linear (single tap) with the dense fp16 TMA-store epilogue");
    UAV_REQUIRE(p.ln_in == nullptr || (p.ln_colsum != nullptr && p.ln_slots >= 1 && p.ln_slots <= 16 &&
                                       (reinterpret_cast<uintptr_t>(p.ln_colsum) & 15) == 0),
                "igemm: ln_in needs ln_colsum (16-byte aligned) and 1..16 slots");
    UAV_REQUIRE(p.ln_out == nullptr || !geglu, "igemm: ln_out is not supported with GEGLU");
    p.ln_out_slots = (int32_t)p.n_tiles * 2;
    UAV_REQUIRE(p.ln_out == nullptr || e->ln_out_slots == p.ln_out_slots,
                "igemm: ln_out_slots is %d, this launch writes %d slots per row", e->ln_out_slots, p.ln_out_slots);
  }
  if (p.gn_partial != nullptr) {
    UAV_REQUIRE(can_tma && !geglu && d.out_strides[1] == 0,
                "igemm: GroupNorm statistics need the dense fp16 TMA-store epilogue (n_out >= 33, 16-byte aligned) without GEGLU");
    UAV_REQUIRE(p.gn_blocks == (int64_t)m_tiles * 4, "igemm: gn_blocks is %lld, this launch produces %lld blocks",
                (long long)p.gn_blocks, (long long)m_tiles * 4);
  }
  UAV_REQUIRE(d.out_strides[1] == 0 || (can_tma && p.residual == nullptr && p.rowvec == nullptr),
              "igemm: a strided output view needs the TMA-store epilogue without residual / row vector");
  if (can_tma) {
    cuuint64_t dims[5], strides[4];
    cuuint32_t box[5], estr[5] = {1, 1, 1, 1, 1};
    dims[0] = (cuuint64_t)p.n_out;
    box[0] = 64;
    uint64_t stride_el = (uint64_t)p.ld_out;
    for (int i = 1; i < 5; ++i) {
      dims[i] = d.out_dims[i];
      box[i] = d.box[i];
      strides[i - 1] = (d.out_strides[i] ? d.out_strides[i] : stride_el) * 2;
      stride_el *= d.out_dims[i];
    }
    CUresult r = encode(&p.map_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, d.out, dims, strides, box,
                        estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    UAV_REQUIRE(r == CUDA_SUCCESS, "igemm: cuTensorMapEncodeTiled(out) failed with %d", (int)r);
    if (p.res_mode != 0) {
      uint64_t rs = (uint64_t)p.ld_res;
      for (int i = 1; i < 5; ++i) {
        strides[i - 1] = rs * 2;
        rs *= d.out_dims[i];
      }
      r = encode(&p.map_res, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<__half*>(p.residual), dims, strides, box, estr,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      UAV_REQUIRE(r == CUDA_SUCCESS, "igemm: cuTensorMapEncodeTiled(residual) failed with %d", (int)r);
    }
  }

  if (geglu) return launch_instance<256, true>(p, use_cluster, stream);
  switch (block_n) {
    case 256: return launch_instance<256, false>(p, use_cluster, stream);
    case 128: return launch_instance<128, false>(p, use_cluster, stream);
    case 64: return launch_instance<64, false>(p, false, stream);
    case 32: return launch_instance<32, false>(p, false, stream);
    default: return launch_instance<16, false>(p, false, stream);
  }
}

// choose the (tw, th) rectangle with tw * th == 128 that wastes the fewest rows
static void pick_tile_2d(int64_t W, int64_t H, uint32_t* tw, uint32_t* th) {
  int64_t best = -1;
  for (uint32_t w = 128; w >= 1; w >>= 1) {
    const uint32_t h = 128 / w;
    const int64_t cover = ((W + w - 1) / w) * w * ((H + h - 1) / h) * h;
    if (best < 0 || cover < best) {
      best = cover;
      *tw = w;
      *th = h;
    }
  }
}

}  // namespace uav

using namespace uav;

extern "C" {

int uav_ln_partial_slots(int64_t n_out) {
  // two column halves per N-tile of the TMA-store epilogue (tile width 256 for N > 128, 128 for N > 64, else 64)
  if (n_out <= 0) return 0;
  const int tile = n_out > 128 ? 256 : (n_out > 64 ? 128 : 64);
  return (int)((n_out + tile - 1) / tile) * 2;
}

int64_t uav_gn_partial_blocks(int64_t w, int64_t h, int64_t images) {
  if (w <= 0 || h <= 0 || images <= 0) return 0;
  if (h == 1) return ((w + 127) / 128) * images * 4;
  uint32_t tw, th;
  pick_tile_2d(w, h, &tw, &th);
  return ((w + tw - 1) / tw) * ((h + th - 1) / th) * images * 4;
}

const char* uav_version(void) { return "uav_b200 0.2 (sm_100a)"; }
const char* uav_last_error_string(void) { return g_err; }
uint64_t uav_launch_count(void) { return g_launches.load(); }

uav_status_t uav_linear(const void* a, int64_t M, int64_t K, int64_t lda, const void* w,
                        int64_t N, void* out, const uav_epilogue_t* epi, uav_stream_t stream) {
  UAV_REQUIRE(M >= 0 && K > 0 && N > 0 && lda >= K, "uav_linear: bad shape");
  if (M == 0) return UAV_OK;
  IgemmDesc d;
  memset(&d, 0, sizeof(d));
  d.a = a;
  const uint64_t dims[5] = {(uint64_t)K, (uint64_t)M, 1, 1, 1};
  const uint64_t strides[5] = {1, (uint64_t)lda, (uint64_t)lda * M, (uint64_t)lda * M,
                               (uint64_t)lda * M};
  const uint32_t box[5] = {64, 128, 1, 1, 1};
  for (int i = 0; i < 5; ++i) {
    d.a_dims[i] = dims[i];
    d.a_strides[i] = strides[i];
    d.box[i] = box[i];
    d.tiles[i] = 1;
    d.out_dims[i] = 1;
  }
  d.tiles[1] = (uint32_t)((M + 127) / 128);
  d.out_dims[1] = (uint32_t)M;
  d.num_taps = 1;
  d.k_per_tap = (int)K;
  d.w = w;
  d.N = N;
  d.out = out;
  d.epi = epi;
  return launch_igemm(d, (cudaStream_t)stream);
}

uav_status_t uav_conv2d(const void* x, int64_t NB, int64_t H, int64_t W, int64_t Cin,
                        int64_t ld_in, const void* w, int64_t Cout, int ksize, int stride,
                        int pad_mode, void* out, const uav_epilogue_t* epi,
                        uav_stream_t stream) {
  UAV_REQUIRE(NB > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && ld_in >= Cin,
              "uav_conv2d: bad shape");
  UAV_REQUIRE(ksize == 1 || ksize == 3, "uav_conv2d: ksize must be 1 or 3");
  UAV_REQUIRE(stride == 1 || stride == 2, "uav_conv2d: stride must be 1 or 2");
  UAV_REQUIRE(pad_mode == 0 || (pad_mode == 1 && stride == 2 && ksize == 3),
              "uav_conv2d: pad_mode 1 needs a stride-2 3x3 conv");
  // a 1x1 stride-1 convolution is a plain GEMM over the NB*H*W pixels (always-full 128-row tiles)
  if (ksize == 1 && stride == 1) return uav_linear(x, NB * H * W, Cin, ld_in, w, Cout, out, epi, stream);
  IgemmDesc d;
  memset(&d, 0, sizeof(d));
  d.a = x;
  d.w = w;
  d.N = Cout;
  d.out = out;
  d.epi = epi;
  d.k_per_tap = (int)Cin;
  const int pad = ksize / 2;
  if (stride == 1) {
    uint32_t tw, th;
    pick_tile_2d(W, H, &tw, &th);
    const uint64_t dims[5] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)NB, 1};
    const uint64_t strides[5] = {1, (uint64_t)ld_in, (uint64_t)ld_in * W,
                                 (uint64_t)ld_in * W * H, (uint64_t)ld_in * W * H * NB};
    const uint32_t box[5] = {64, tw, th, 1, 1};
    const uint32_t tiles[5] = {1, (uint32_t)((W + tw - 1) / tw), (uint32_t)((H + th - 1) / th),
                               (uint32_t)NB, 1};
    const uint32_t odims[5] = {1, (uint32_t)W, (uint32_t)H, (uint32_t)NB, 1};
    for (int i = 0; i < 5; ++i) {
      d.a_dims[i] = dims[i];
      d.a_strides[i] = strides[i];
      d.box[i] = box[i];
      d.tiles[i] = tiles[i];
      d.out_dims[i] = odims[i];
    }
    d.num_taps = ksize * ksize;
    for (int ky = 0; ky < ksize; ++ky)
      for (int kx = 0; kx < ksize; ++kx) {
        int32_t* o = d.tap_off[ky * ksize + kx];
        o[0] = 0;
        o[1] = kx - pad;
        o[2] = ky - pad;
        o[3] = 0;
        o[4] = 0;
      }
  } else {
    UAV_REQUIRE(H % 2 == 0 && W % 2 == 0, "uav_conv2d: stride 2 needs even H, W (got %lldx%lld)",
                (long long)H, (long long)W);
    UAV_REQUIRE(ld_in == Cin, "uav_conv2d: stride 2 needs a dense input (ld_in == Cin)");
    const int64_t Wo = W / 2, Ho = H / 2;
    uint32_t tw, th;
    pick_tile_2d(Wo, Ho, &tw, &th);
    // phase view: (2C [px*C + c], W/2, 2 [py], H/2, NB)
    const uint64_t dims[5] = {(uint64_t)(2 * Cin), (uint64_t)Wo, 2, (uint64_t)Ho, (uint64_t)NB};
    const uint64_t strides[5] = {1, (uint64_t)(2 * Cin), (uint64_t)(W * Cin),
                                 (uint64_t)(2 * W * Cin), (uint64_t)(H * W * Cin)};
    const uint32_t box[5] = {64, tw, 1, th, 1};
    const uint32_t tiles[5] = {1, (uint32_t)((Wo + tw - 1) / tw), 1,
                               (uint32_t)((Ho + th - 1) / th), (uint32_t)NB};
    const uint32_t odims[5] = {1, (uint32_t)Wo, 1, (uint32_t)Ho, (uint32_t)NB};
    for (int i = 0; i < 5; ++i) {
      d.a_dims[i] = dims[i];
      d.a_strides[i] = strides[i];
      d.box[i] = box[i];
      d.tiles[i] = tiles[i];
      d.out_dims[i] = odims[i];
    }
    d.num_taps = 9;
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx) {
        // input coordinate = 2*o + k - pad  (pad = 1 for pad_mode 0, 0 for pad_mode 1)
        const int iy = ky - (pad_mode == 0 ? 1 : 0);  // in {-1,0,1} or {0,1,2}
        const int ix = kx - (pad_mode == 0 ? 1 : 0);
        const int py = ((iy % 2) + 2) % 2, px = ((ix % 2) + 2) % 2;
        const int oy = (iy - py) / 2, ox = (ix - px) / 2;  // exact
        int32_t* o = d.tap_off[ky * 3 + kx];
        o[0] = px * (int)Cin;
        o[1] = ox;
        o[2] = py;
        o[3] = oy;
        o[4] = 0;
      }
  }
  return launch_igemm(d, (cudaStream_t)stream);
}

uav_status_t uav_conv2d_taps(const void* x, int64_t NB, int64_t H, int64_t W, int64_t Cin, int64_t ld_in,
                             const void* w, int64_t Cout, int kh, int kw, int pad_top, int pad_left, void* out,
                             const uav_epilogue_t* epi, uav_stream_t stream) {
  UAV_REQUIRE(NB > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && ld_in >= Cin, "uav_conv2d_taps: bad shape");
  UAV_REQUIRE(kh >= 1 && kw >= 1 && kh * kw <= MAX_TAPS, "uav_conv2d_taps: at most %d taps (got %d x %d)", MAX_TAPS, kh,
              kw);
  UAV_REQUIRE(pad_top >= 0 && pad_top < kh && pad_left >= 0 && pad_left < kw, "uav_conv2d_taps: bad padding");
  IgemmDesc d;
  memset(&d, 0, sizeof(d));
  d.a = x;
  d.w = w;
  d.N = Cout;
  d.out = out;
  d.epi = epi;
  d.k_per_tap = (int)Cin;
  uint32_t tw, th;
  pick_tile_2d(W, H, &tw, &th);
  const uint64_t dims[5] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)NB, 1};
  const uint64_t strides[5] = {1, (uint64_t)ld_in, (uint64_t)ld_in * W, (uint64_t)ld_in * W * H,
                               (uint64_t)ld_in * W * H * NB};
  const uint32_t box[5] = {64, tw, th, 1, 1};
  const uint32_t tiles[5] = {1, (uint32_t)((W + tw - 1) / tw), (uint32_t)((H + th - 1) / th), (uint32_t)NB, 1};
  const uint32_t odims[5] = {1, (uint32_t)W, (uint32_t)H, (uint32_t)NB, 1};
  for (int i = 0; i < 5; ++i) {
    d.a_dims[i] = dims[i];
    d.a_strides[i] = strides[i];
    d.box[i] = box[i];
    d.tiles[i] = tiles[i];
    d.out_dims[i] = odims[i];
  }
  d.num_taps = kh * kw;
  for (int ky = 0; ky < kh; ++ky)
    for (int kx = 0; kx < kw; ++kx) {
      int32_t* o = d.tap_off[ky * kw + kx];
      o[0] = 0;
      o[1] = kx - pad_left;
      o[2] = ky - pad_top;
      o[3] = 0;
      o[4] = 0;
    }
  return launch_igemm(d, (cudaStream_t)stream);
}

uav_status_t uav_conv_temporal(const void* x, int64_t B, int64_t T, int64_t HW, int64_t Cin,
                               int64_t ld_in, const void* w, int64_t Cout, int k, void* out,
                               const uav_epilogue_t* epi, uav_stream_t stream) {
  UAV_REQUIRE(B > 0 && T > 0 && HW > 0 && Cin > 0 && Cout > 0 && ld_in >= Cin,
              "uav_conv_temporal: bad shape");
  UAV_REQUIRE(k == 1 || k == 3 || k == 5, "uav_conv_temporal: k must be 1, 3 or 5");
  IgemmDesc d;
  memset(&d, 0, sizeof(d));
  d.a = x;
  d.w = w;
  d.N = Cout;
  d.out = out;
  d.epi = epi;
  d.k_per_tap = (int)Cin;
  const uint64_t dims[5] = {(uint64_t)Cin, (uint64_t)HW, (uint64_t)T, (uint64_t)B, 1};
  const uint64_t strides[5] = {1, (uint64_t)ld_in, (uint64_t)ld_in * HW, (uint64_t)ld_in * HW * T,
                               (uint64_t)ld_in * HW * T * B};
  const uint32_t box[5] = {64, 128, 1, 1, 1};
  const uint32_t tiles[5] = {1, (uint32_t)((HW + 127) / 128), (uint32_t)T, (uint32_t)B, 1};
  const uint32_t odims[5] = {1, (uint32_t)HW, (uint32_t)T, (uint32_t)B, 1};
  for (int i = 0; i < 5; ++i) {
    d.a_dims[i] = dims[i];
    d.a_strides[i] = strides[i];
    d.box[i] = box[i];
    d.tiles[i] = tiles[i];
    d.out_dims[i] = odims[i];
  }
  d.num_taps = k;
  for (int kt = 0; kt < k; ++kt) {
    int32_t* o = d.tap_off[kt];
    o[0] = 0;
    o[1] = 0;
    o[2] = kt - k / 2;
    o[3] = 0;
    o[4] = 0;
  }
  return launch_igemm(d, (cudaStream_t)stream);
}

uav_status_t uav_conv3d(const void* x, int64_t B, int64_t T, int64_t H, int64_t W, int64_t Cin,
                        int64_t ld_in, const void* w, int64_t Cout, void* out,
                        const uav_epilogue_t* epi, uav_stream_t stream) {
  UAV_REQUIRE(B > 0 && T > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && ld_in >= Cin,
              "uav_conv3d: bad shape");
  IgemmDesc d;
  memset(&d, 0, sizeof(d));
  d.a = x;
  d.w = w;
  d.N = Cout;
  d.out = out;
  d.epi = epi;
  d.k_per_tap = (int)Cin;
  uint32_t tw, th;
  pick_tile_2d(W, H, &tw, &th);
  const uint64_t dims[5] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)T, (uint64_t)B};
  const uint64_t strides[5] = {1, (uint64_t)ld_in, (uint64_t)ld_in * W, (uint64_t)ld_in * W * H,
                               (uint64_t)ld_in * W * H * T};
  const uint32_t box[5] = {64, tw, th, 1, 1};
  const uint32_t tiles[5] = {1, (uint32_t)((W + tw - 1) / tw), (uint32_t)((H + th - 1) / th),
                             (uint32_t)T, (uint32_t)B};
  const uint32_t odims[5] = {1, (uint32_t)W, (uint32_t)H, (uint32_t)T, (uint32_t)B};
  for (int i = 0; i < 5; ++i) {
    d.a_dims[i] = dims[i];
    d.a_strides[i] = strides[i];
    d.box[i] = box[i];
    d.tiles[i] = tiles[i];
    d.out_dims[i] = odims[i];
  }
  d.num_taps = 27;
  for (int kt = 0; kt < 3; ++kt)
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx) {
        int32_t* o = d.tap_off[(kt * 3 + ky) * 3 + kx];
        o[0] = 0;
        o[1] = kx - 1;
        o[2] = ky - 1;
        o[3] = kt - 1;
        o[4] = 0;
      }
  return launch_igemm(d, (cudaStream_t)stream);
}

uav_status_t uav_upsample2x_conv3x3(const void* x, int64_t NB, int64_t H, int64_t W, int64_t Cin,
                                    int64_t ld_in, const void* w4, int64_t Cout, void* out,
                                    const uav_epilogue_t* epi, uav_stream_t stream) {
  UAV_REQUIRE(NB > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && ld_in >= Cin && epi != nullptr,
              "uav_upsample2x_conv3x3: bad shape");
  UAV_REQUIRE(epi->residual == nullptr && epi->rowvec == nullptr && epi->out_dtype == UAV_F16 && Cout >= 33,
              "uav_upsample2x_conv3x3: bias-only fp16 epilogue with Cout > 32 required");
  const int64_t ld_out = epi->ld_out;
  uint32_t tw, th;
  pick_tile_2d(W, H, &tw, &th);
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b) {
      IgemmDesc d;
      memset(&d, 0, sizeof(d));
      d.a = x;
      d.w = reinterpret_cast<const __half*>(w4) + static_cast<int64_t>(a * 2 + b) * Cout * 4 * Cin;
      d.N = Cout;
      d.epi = epi;
      d.k_per_tap = (int)Cin;
      const uint64_t dims[5] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)NB, 1};
      const uint64_t strides[5] = {1, (uint64_t)ld_in, (uint64_t)ld_in * W, (uint64_t)ld_in * W * H,
                                   (uint64_t)ld_in * W * H * NB};
      const uint32_t box[5] = {64, tw, th, 1, 1};
      const uint32_t tiles[5] = {1, (uint32_t)((W + tw - 1) / tw), (uint32_t)((H + th - 1) / th), (uint32_t)NB, 1};
      const uint32_t odims[5] = {1, (uint32_t)W, (uint32_t)H, (uint32_t)NB, 1};
      for (int i = 0; i < 5; ++i) {
        d.a_dims[i] = dims[i];
        d.a_strides[i] = strides[i];
        d.box[i] = box[i];
        d.tiles[i] = tiles[i];
        d.out_dims[i] = odims[i];
      }
      // output phase (a, b): pixel (2y + a, 2x + b) of the [NB][2H][2W][ld_out] tensor
      d.out = reinterpret_cast<__half*>(out) + (static_cast<int64_t>(a) * 2 * W + b) * ld_out;
      d.out_strides[1] = 2 * (uint64_t)ld_out;
      d.out_strides[2] = 2 * 2 * (uint64_t)W * ld_out;
      d.out_strides[3] = 4 * (uint64_t)H * W * ld_out;
      d.out_strides[4] = 4 * (uint64_t)H * W * ld_out * NB;
      // source taps of the collapsed 2x2 filter: phase 0 reads {-1, 0}, phase 1 reads {0, +1}
      d.num_taps = 4;
      for (int ty = 0; ty < 2; ++ty)
        for (int tx = 0; tx < 2; ++tx) {
          int32_t* o = d.tap_off[ty * 2 + tx];
          o[0] = 0;
          o[1] = (b == 0 ? -1 : 0) + tx;
          o[2] = (a == 0 ? -1 : 0) + ty;
          o[3] = 0;
          o[4] = 0;
        }
      uav_status_t st = launch_igemm(d, (cudaStream_t)stream);
      if (st != UAV_OK) return st;
    }
  return UAV_OK;
}

}  // extern "C"
