// attention.cu — attention cores of the sampling path.
//
//  * uav_attention: softmax(Q K^T * scale) V without materialising the score matrix
//    (the reference materialises it: attention.py:209-238, 2.1 GB per call at 320x576).
//    FlashAttention-2 style tiling on mma.sync.m16n8k16 (fp16 in, fp32 accumulate, fp32
//    online softmax): used for the UNet spatial self-attention (N = h*w, d = 128), the text
//    cross-attention (Nk = 77, d = 64/128, K/V shared by all frames of a batch item) and the
//    VAE mid-block attention (1 head, d = 512, as four 128-wide V slices).
//    NOTE (round-1 status): this is the HMMA baseline; the tcgen05/TMEM version is the next
//    step for the d=512 VAE attention, which dominates VAE-decode FLOPs.
//  * uav_temporal_attention: the seq = T <= 8 per-pixel attention with rotary embedding on
//    the first 32 dims and the T5-style relative-position bias (attention.py:699-733) as a
//    register-resident warp kernel: one warp per (pixel, head), lane = (frame, quarter of the
//    head dim), K/V exchanged with warp shuffles.  Reads q/k/v in the (b, f, hw, c) layout,
//    so the reference's two "(b f) d c <-> (b d) f c" rearrange copies (attention.py:555,560)
//    do not exist.
#include "uav_common.cuh"

#include <atomic>
#include <stdlib.h>

namespace uav {
extern std::atomic<uint64_t> g_launches;

// ---------------------------------------------------------------------------------------
// flash attention (mma.sync)
// ---------------------------------------------------------------------------------------
constexpr int FA_BM = 64;   // query rows per CTA (4 warps x 16)
constexpr int FA_BN = 64;   // kv rows per iteration
constexpr int FA_THREADS = 128;

struct FaParams {
  const __half* q;
  const __half* k;
  const __half* v;
  __half* o;
  int64_t ldq, ldk, ldv, ldo;        // token stride (elements)
  int64_t bsq, bsk, bsv, bso;        // batch stride (elements)
  int nq, nk, heads, kv_batch_div;
  float scale_log2;                  // softmax scale * log2(e)
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  const uint32_t s = smem_u32(smem);
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() {
  asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// smem tile: rows of DT halfs, 16-byte chunks XOR-swizzled by (row & 7)
template <int DT>
__device__ __forceinline__ __half* tile_ptr(__half* base, int row, int chunk) {
  return base + row * DT + ((chunk ^ (row & 7)) << 3);
}

// DQK: head dim of q/k; DV: width of the V slice processed by this launch
template <int DQK, int DV>
__global__ void __launch_bounds__(FA_THREADS)
    flash_attn_kernel(const FaParams p) {
  extern __shared__ __align__(16) uint8_t fa_smem[];
  __half* sq = reinterpret_cast<__half*>(fa_smem);  // [64][DQK]
  __half* sk = sq + FA_BM * DQK;                    // [2][64][DQK]
  __half* sv = sk + 2 * FA_BN * DQK;                // [2][64][DV]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bh = blockIdx.y;
  const int b = bh / p.heads, h = bh % p.heads;
  const int q0 = blockIdx.x * FA_BM;
  const __half* qg = p.q + b * p.bsq + static_cast<int64_t>(h) * DQK;
  const __half* kg = p.k + (b / p.kv_batch_div) * p.bsk + static_cast<int64_t>(h) * DQK;
  const __half* vg = p.v + (b / p.kv_batch_div) * p.bsv + static_cast<int64_t>(h) * DV;
  __half* og = p.o + b * p.bso + static_cast<int64_t>(h) * DV;

  constexpr int QC = DQK / 8, VC = DV / 8;  // 16B chunks per row
  // ---- async loads: Q tile, then K/V tile 0 ----
  for (int i = tid; i < FA_BM * QC; i += FA_THREADS) {
    const int r = i / QC, c = i % QC;
    const bool ok = q0 + r < p.nq;
    cp_async16(tile_ptr<DQK>(sq, r, c), qg + static_cast<int64_t>(ok ? q0 + r : 0) * p.ldq + c * 8,
               ok);
  }
  auto load_kv = [&](int tile, int buf) {
    const int k0 = tile * FA_BN;
    __half* skb = sk + buf * FA_BN * DQK;
    __half* svb = sv + buf * FA_BN * DV;
    for (int i = tid; i < FA_BN * QC; i += FA_THREADS) {
      const int r = i / QC, c = i % QC;
      const bool ok = k0 + r < p.nk;
      cp_async16(tile_ptr<DQK>(skb, r, c),
                 kg + static_cast<int64_t>(ok ? k0 + r : 0) * p.ldk + c * 8, ok);
    }
    for (int i = tid; i < FA_BN * VC; i += FA_THREADS) {
      const int r = i / VC, c = i % VC;
      const bool ok = k0 + r < p.nk;
      cp_async16(tile_ptr<DV>(svb, r, c),
                 vg + static_cast<int64_t>(ok ? k0 + r : 0) * p.ldv + c * 8, ok);
    }
  };
  load_kv(0, 0);
  cp_async_commit();

  const int ntiles = (p.nk + FA_BN - 1) / FA_BN;
  float o_acc[DV / 8][4];
#pragma unroll
  for (int i = 0; i < DV / 8; ++i) o_acc[i][0] = o_acc[i][1] = o_acc[i][2] = o_acc[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};

  const int g = lane >> 2, t4 = lane & 3;
  const int arow = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;  // ldmatrix A row
  const int achk = lane >> 4;

  for (int tile = 0; tile < ntiles; ++tile) {
    const int buf = tile & 1;
    if (tile + 1 < ntiles) load_kv(tile + 1, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const __half* skb = sk + buf * FA_BN * DQK;
    const __half* svb = sv + buf * FA_BN * DV;

    // ---- S = Q K^T (16 x 64 per warp) ----
    float s[FA_BN / 8][4];
#pragma unroll
    for (int i = 0; i < FA_BN / 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < DQK / 16; ++kk) {
      uint32_t a[4];
      ldmatrix_x4(a, tile_ptr<DQK>(sq, arow, kk * 2 + achk));
#pragma unroll
      for (int nb = 0; nb < FA_BN / 16; ++nb) {
        uint32_t bfr[4];
        const int brow = nb * 16 + (lane & 7) + (lane >> 4) * 8;
        const int bchk = kk * 2 + ((lane >> 3) & 1);
        ldmatrix_x4(bfr, tile_ptr<DQK>(const_cast<__half*>(skb), brow, bchk));
        mma16816(s[nb * 2], a, bfr[0], bfr[1]);
        mma16816(s[nb * 2 + 1], a, bfr[2], bfr[3]);
      }
    }
    // ---- mask + online softmax (rows g and g+8 of this warp's 16) ----
    const int kbase = tile * FA_BN;
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nb = 0; nb < FA_BN / 8; ++nb) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int col = kbase + nb * 8 + t4 * 2 + (e & 1);
        float x = s[nb][e] * p.scale_log2;
        if (col >= p.nk) x = -INFINITY;
        s[nb][e] = x;
        mx[e >> 1] = fmaxf(mx[e >> 1], x);
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffff, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffff, mx[r], 2));
    }
    float corr[2], rs[2] = {0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const float mnew = fmaxf(m_run[r], mx[r]);
      corr[r] = (m_run[r] == -INFINITY) ? 0.f : exp2f(m_run[r] - mnew);
      m_run[r] = mnew;
    }
    uint32_t pf[FA_BN / 8][2];  // P as fp16 pairs (A fragments)
#pragma unroll
    for (int nb = 0; nb < FA_BN / 8; ++nb) {
      const float p0 = exp2f(s[nb][0] - m_run[0]), p1 = exp2f(s[nb][1] - m_run[0]);
      const float p2 = exp2f(s[nb][2] - m_run[1]), p3 = exp2f(s[nb][3] - m_run[1]);
      rs[0] += p0 + p1;
      rs[1] += p2 + p3;
      __half2 h01 = __floats2half2_rn(p0, p1), h23 = __floats2half2_rn(p2, p3);
      pf[nb][0] = *reinterpret_cast<uint32_t*>(&h01);
      pf[nb][1] = *reinterpret_cast<uint32_t*>(&h23);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * corr[r] + rs[r];
#pragma unroll
    for (int i = 0; i < DV / 8; ++i) {
      o_acc[i][0] *= corr[0];
      o_acc[i][1] *= corr[0];
      o_acc[i][2] *= corr[1];
      o_acc[i][3] *= corr[1];
    }
    // ---- O += P V ----
#pragma unroll
    for (int kk = 0; kk < FA_BN / 16; ++kk) {
      const uint32_t a[4] = {pf[kk * 2][0], pf[kk * 2][1], pf[kk * 2 + 1][0], pf[kk * 2 + 1][1]};
#pragma unroll
      for (int nb = 0; nb < DV / 16; ++nb) {
        uint32_t bfr[4];
        const int vrow = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int vchk = nb * 2 + (lane >> 4);
        ldmatrix_x4_trans(bfr, tile_ptr<DV>(const_cast<__half*>(svb), vrow, vchk));
        mma16816(o_acc[nb * 2], a, bfr[0], bfr[1]);
        mma16816(o_acc[nb * 2 + 1], a, bfr[2], bfr[3]);
      }
    }
    __syncthreads();  // all warps done with buf before it is refilled
  }
  cp_async_wait<0>();

  // ---- finalize: O / l -> smem (reuse Q tile region, needs DV <= DQK) -> coalesced stores ----
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    float l = l_run[r];
    l += __shfl_xor_sync(0xffffffff, l, 1);
    l += __shfl_xor_sync(0xffffffff, l, 2);
    l_run[r] = (l > 0.f) ? 1.f / l : 0.f;
  }
  __half* so = sq;  // [64][DV] swizzled like a DV tile
#pragma unroll
  for (int nb = 0; nb < DV / 8; ++nb) {
    const int r0 = warp * 16 + g, r1 = r0 + 8;
    __half2 v0 = __floats2half2_rn(o_acc[nb][0] * l_run[0], o_acc[nb][1] * l_run[0]);
    __half2 v1 = __floats2half2_rn(o_acc[nb][2] * l_run[1], o_acc[nb][3] * l_run[1]);
    *reinterpret_cast<__half2*>(tile_ptr<DV>(so, r0, nb) + t4 * 2) = v0;
    *reinterpret_cast<__half2*>(tile_ptr<DV>(so, r1, nb) + t4 * 2) = v1;
  }
  __syncthreads();
  for (int i = tid; i < FA_BM * VC; i += FA_THREADS) {
    const int r = i / VC, c = i % VC;
    if (q0 + r < p.nq)
      stg16(og + static_cast<int64_t>(q0 + r) * p.ldo + c * 8,
            *reinterpret_cast<const uint4*>(tile_ptr<DV>(so, r, c)));
  }
}

// ---------------------------------------------------------------------------------------
// text cross-attention (nk <= 128, typically 77): K and V of one (batch, head) stay resident in
// shared memory, the CTA streams query tiles past them (double-buffered cp.async), the whole
// score row fits in registers (single-pass softmax), and each warp stages its 16 output rows in
// the Q buffer it has just consumed.  The generic kernel above spends most of its time in the
// per-CTA prologue / epilogue when there are only two KV tiles (measured 1.0 TB/s at h/2).
// ---------------------------------------------------------------------------------------
template <int D, int NB16>
__global__ void __launch_bounds__(FA_THREADS)
    cross_attn_kernel(const FaParams p) {
  constexpr int NKP = NB16 * 16;
  extern __shared__ __align__(16) uint8_t fa_smem[];
  __half* sk = reinterpret_cast<__half*>(fa_smem);   // [NKP][D]
  __half* sv = sk + NKP * D;                          // [NKP][D]
  __half* sq = sv + NKP * D;                          // [2][64][D]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // the head is the FASTEST block index: the CTAs that read the 2*D-byte head slices of the same token rows run side
  // by side, so each 1 KB token row is fetched from DRAM within one burst instead of once per head pass
  const int b = blockIdx.y, h = blockIdx.x % p.heads;
  const int tile_stride = gridDim.x / p.heads;
  const __half* qg = p.q + b * p.bsq + static_cast<int64_t>(h) * D;
  const __half* kg = p.k + (b / p.kv_batch_div) * p.bsk + static_cast<int64_t>(h) * D;
  const __half* vg = p.v + (b / p.kv_batch_div) * p.bsv + static_cast<int64_t>(h) * D;
  __half* og = p.o + b * p.bso + static_cast<int64_t>(h) * D;
  constexpr int QC = D / 8;
  const int ntiles = (p.nq + FA_BM - 1) / FA_BM;

  for (int i = tid; i < NKP * QC; i += FA_THREADS) {
    const int r = i / QC, c = i % QC;
    const bool ok = r < p.nk;
    cp_async16(tile_ptr<D>(sk, r, c), kg + static_cast<int64_t>(ok ? r : 0) * p.ldk + c * 8, ok);
    cp_async16(tile_ptr<D>(sv, r, c), vg + static_cast<int64_t>(ok ? r : 0) * p.ldv + c * 8, ok);
  }
  auto load_q = [&](int tile, int buf) {
    const int q0 = tile * FA_BM;
    __half* sqb = sq + buf * FA_BM * D;
    for (int i = tid; i < FA_BM * QC; i += FA_THREADS) {
      const int r = i / QC, c = i % QC;
      const bool ok = q0 + r < p.nq;
      cp_async16(tile_ptr<D>(sqb, r, c), qg + static_cast<int64_t>(ok ? q0 + r : 0) * p.ldq + c * 8, ok);
    }
  };
  int tile = blockIdx.x / p.heads;
  if (tile < ntiles) load_q(tile, 0);
  cp_async_commit();

  const int g = lane >> 2, t4 = lane & 3;
  const int arow = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
  const int achk = lane >> 4;
  int it = 0;
  for (; tile < ntiles; tile += tile_stride, ++it) {
    const int buf = it & 1;
    const int next = tile + tile_stride;
    if (next < ntiles) load_q(next, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    __half* sqb = sq + buf * FA_BM * D;

    // S = Q K^T  (16 x NKP per warp)
    float s[NB16 * 2][4];
#pragma unroll
    for (int i = 0; i < NB16 * 2; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < D / 16; ++kk) {
      uint32_t a[4];
      ldmatrix_x4(a, tile_ptr<D>(sqb, arow, kk * 2 + achk));
#pragma unroll
      for (int nb = 0; nb < NB16; ++nb) {
        uint32_t bfr[4];
        const int brow = nb * 16 + (lane & 7) + (lane >> 4) * 8;
        const int bchk = kk * 2 + ((lane >> 3) & 1);
        ldmatrix_x4(bfr, tile_ptr<D>(sk, brow, bchk));
        mma16816(s[nb * 2], a, bfr[0], bfr[1]);
        mma16816(s[nb * 2 + 1], a, bfr[2], bfr[3]);
      }
    }
    // row maxima of the raw scores (scale > 0); only 8-column blocks past nk hold padded keys
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nb = 0; nb < NB16 * 2; ++nb) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if ((nb + 1) * 8 > p.nk) {  // block-uniform test, false for all but the tail blocks
          const int col = nb * 8 + t4 * 2 + (e & 1);
          if (col >= p.nk) s[nb][e] = -INFINITY;
        }
        mx[e >> 1] = fmaxf(mx[e >> 1], s[nb][e]);
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffff, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffff, mx[r], 2));
      mx[r] *= p.scale_log2;
    }
    float rs[2] = {0.f, 0.f};
    uint32_t pf[NB16 * 2][2];
#pragma unroll
    for (int nb = 0; nb < NB16 * 2; ++nb) {
      // exp2(s * scale - max): one FFMA + one MUFU per score
      const float p0 = ex2_ftz(fmaf(s[nb][0], p.scale_log2, -mx[0])), p1 = ex2_ftz(fmaf(s[nb][1], p.scale_log2, -mx[0]));
      const float p2 = ex2_ftz(fmaf(s[nb][2], p.scale_log2, -mx[1])), p3 = ex2_ftz(fmaf(s[nb][3], p.scale_log2, -mx[1]));
      rs[0] += p0 + p1;
      rs[1] += p2 + p3;
      __half2 h01 = __floats2half2_rn(p0, p1), h23 = __floats2half2_rn(p2, p3);
      pf[nb][0] = *reinterpret_cast<uint32_t*>(&h01);
      pf[nb][1] = *reinterpret_cast<uint32_t*>(&h23);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      rs[r] += __shfl_xor_sync(0xffffffff, rs[r], 1);
      rs[r] += __shfl_xor_sync(0xffffffff, rs[r], 2);
      rs[r] = 1.f / rs[r];
    }
    // O = P V
    float o_acc[D / 8][4];
#pragma unroll
    for (int i = 0; i < D / 8; ++i) o_acc[i][0] = o_acc[i][1] = o_acc[i][2] = o_acc[i][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < NB16; ++kk) {
      const uint32_t a[4] = {pf[kk * 2][0], pf[kk * 2][1], pf[kk * 2 + 1][0], pf[kk * 2 + 1][1]};
#pragma unroll
      for (int nb = 0; nb < D / 16; ++nb) {
        uint32_t bfr[4];
        const int vrow = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int vchk = nb * 2 + (lane >> 4);
        ldmatrix_x4_trans(bfr, tile_ptr<D>(sv, vrow, vchk));
        mma16816(o_acc[nb * 2], a, bfr[0], bfr[1]);
        mma16816(o_acc[nb * 2 + 1], a, bfr[2], bfr[3]);
      }
    }
    // stage this warp's 16 rows in its own (already consumed) rows of the Q buffer, then 16-byte stores
    __syncwarp();
#pragma unroll
    for (int nb = 0; nb < D / 8; ++nb) {
      const int r0 = warp * 16 + g, r1 = r0 + 8;
      __half2 v0 = __floats2half2_rn(o_acc[nb][0] * rs[0], o_acc[nb][1] * rs[0]);
      __half2 v1 = __floats2half2_rn(o_acc[nb][2] * rs[1], o_acc[nb][3] * rs[1]);
      *reinterpret_cast<__half2*>(tile_ptr<D>(sqb, r0, nb) + t4 * 2) = v0;
      *reinterpret_cast<__half2*>(tile_ptr<D>(sqb, r1, nb) + t4 * 2) = v1;
    }
    __syncwarp();
    const int q0 = tile * FA_BM;
    for (int i = lane; i < 16 * QC; i += 32) {
      const int r = warp * 16 + i / QC, c = i % QC;
      if (q0 + r < p.nq)
        stg16(og + static_cast<int64_t>(q0 + r) * p.ldo + c * 8,
              *reinterpret_cast<const uint4*>(tile_ptr<D>(sqb, r, c)));
    }
    __syncthreads();  // buffer `buf` is refilled by the prefetch issued in the next iteration
  }
  cp_async_wait<0>();
}

template <int D, int NB16>
static uav_status_t launch_cross(const FaParams& p, int batch, cudaStream_t stream) {
  constexpr int smem = (2 * NB16 * 16 * D + 2 * FA_BM * D) * 2;
  static uint64_t configured = 0;  // per-device bit: cudaFuncSetAttribute applies to the current device only
  const uint64_t dev_bit = 1ull << (current_device() & 63);
  if (!(configured & dev_bit)) {
    UAV_CHECK_CUDA(cudaFuncSetAttribute(cross_attn_kernel<D, NB16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured |= dev_bit;
  }
  const int ntiles = (p.nq + FA_BM - 1) / FA_BM;
  // enough CTAs to fill the GPU ~4x over, each streaming several query tiles past its resident K/V
  int gx = (num_sms() * 8 + batch * p.heads - 1) / (batch * p.heads);
  if (gx > ntiles) gx = ntiles;
  if (gx < 1) gx = 1;
  dim3 grid(gx * p.heads, batch);
  cross_attn_kernel<D, NB16><<<grid, FA_THREADS, smem, stream>>>(p);
  UAV_CHECK_CUDA(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return UAV_OK;
}

// ---------------------------------------------------------------------------------------
// temporal attention: one warp per (batch, pixel, head); lane = frame * 4 + quarter
// ---------------------------------------------------------------------------------------
struct TaParams {
  const __half* q;
  const __half* k;
  const __half* v;
  __half* o;
  int64_t ldq, ldk, ldv, ldo;  // token stride (elements); tokens ordered (b, f, hw)
  int B, F, heads;
  int64_t HW;
  float scale;
  const float* rot;   // [F][16][2] cos, sin of frame * freq_pair
  const float* bias;  // [heads][F][F]
};

template <int D>
__global__ void __launch_bounds__(256)
    temporal_attn_kernel(const TaParams p) {
  constexpr int DP = D / 4;   // dims per lane
  constexpr int HP = DP / 2;  // half2 per lane
  const int lane = threadIdx.x & 31;
  const int64_t wid = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int64_t total = static_cast<int64_t>(p.B) * p.HW * p.heads;
  if (wid >= total) return;
  const int h = static_cast<int>(wid % p.heads);
  const int64_t pix = (wid / p.heads) % p.HW;
  const int b = static_cast<int>(wid / (p.heads * p.HW));
  const int i = lane >> 2, part = lane & 3;
  const bool act = i < p.F;
  const int64_t tok = (static_cast<int64_t>(b) * p.F + (act ? i : 0)) * p.HW + pix;
  const int off = h * D + part * DP;

  float qf[DP];
  __half2 kh[HP], vh[HP];
  {
    const __half* qp = p.q + tok * p.ldq + off;
    const __half* kp = p.k + tok * p.ldk + off;
    const __half* vp = p.v + tok * p.ldv + off;
#pragma unroll
    for (int c = 0; c < DP / 8; ++c) {
      const uint4 a = act ? ldg16(qp + c * 8) : make_uint4(0, 0, 0, 0);
      const uint4 bq = act ? ldg16(kp + c * 8) : make_uint4(0, 0, 0, 0);
      const uint4 cq = act ? ldg16(vp + c * 8) : make_uint4(0, 0, 0, 0);
      const __half2* ah = reinterpret_cast<const __half2*>(&a);
      const __half2* bh = reinterpret_cast<const __half2*>(&bq);
      const __half2* ch = reinterpret_cast<const __half2*>(&cq);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(ah[j]);
        qf[c * 8 + 2 * j] = f.x * p.scale;
        qf[c * 8 + 2 * j + 1] = f.y * p.scale;
        kh[c * 4 + j] = bh[j];
        vh[c * 4 + j] = ch[j];
      }
    }
  }
  // rotary on dims [0, 32): interleaved pairs (x0, x1) -> (x0 c - x1 s, x1 c + x0 s)
  if (part * DP < 32 && act) {
#pragma unroll
    for (int j = 0; j < HP; ++j) {
      const int pair = part * HP + j;
      if (pair < 16) {
        const float c = p.rot[(i * 16 + pair) * 2], s = p.rot[(i * 16 + pair) * 2 + 1];
        const float q0 = qf[2 * j], q1 = qf[2 * j + 1];
        qf[2 * j] = q0 * c - q1 * s;
        qf[2 * j + 1] = q1 * c + q0 * s;
        const float2 kf = __half22float2(kh[j]);
        kh[j] = __floats2half2_rn(kf.x * c - kf.y * s, kf.y * c + kf.x * s);
      }
    }
  }
  // scores s[j] = q_i . k_j
  float sc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float acc = 0.f;
    if (j < p.F) {
#pragma unroll
      for (int d = 0; d < HP; ++d) {
        uint32_t w = *reinterpret_cast<uint32_t*>(&kh[d]);
        w = __shfl_sync(0xffffffff, w, j * 4 + part);
        const float2 kf = __half22float2(*reinterpret_cast<__half2*>(&w));
        acc += qf[2 * d] * kf.x + qf[2 * d + 1] * kf.y;
      }
      acc += __shfl_xor_sync(0xffffffff, acc, 1);
      acc += __shfl_xor_sync(0xffffffff, acc, 2);
      acc += act ? p.bias[(h * p.F + i) * p.F + j] : 0.f;
    } else {
      acc = -INFINITY;
    }
    sc[j] = acc;
  }
  float mx = sc[0];
#pragma unroll
  for (int j = 1; j < 8; ++j) mx = fmaxf(mx, sc[j]);
  float den = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = __expf(sc[j] - mx);
    den += sc[j];
  }
  const float inv = 1.f / den;
  float of[DP];
#pragma unroll
  for (int d = 0; d < DP; ++d) of[d] = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (j < p.F) {
      const float pj = sc[j] * inv;
#pragma unroll
      for (int d = 0; d < HP; ++d) {
        uint32_t w = *reinterpret_cast<uint32_t*>(&vh[d]);
        w = __shfl_sync(0xffffffff, w, j * 4 + part);
        const float2 vf = __half22float2(*reinterpret_cast<__half2*>(&w));
        of[2 * d] += pj * vf.x;
        of[2 * d + 1] += pj * vf.y;
      }
    }
  }
  if (act) {
    __half* op = p.o + tok * p.ldo + off;
#pragma unroll
    for (int c = 0; c < DP / 8; ++c) {
      uint4 o;
      uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        __half2 r = __floats2half2_rn(of[c * 8 + 2 * j], of[c * 8 + 2 * j + 1]);
        ow[j] = *reinterpret_cast<uint32_t*>(&r);
      }
      stg16(op + c * 8, o);
    }
  }
}

// ---------------------------------------------------------------------------------------
// temporal attention on mma.sync: one warp per PAIR of heads of one (batch, pixel).
// The two items' 8 frames are stacked into one 16-row tile, so that
//   S  = [Q_a; Q_b] K_a^T , [Q_a; Q_b] K_b^T   (two m16n8k16 n-blocks per k-step; the cross terms are discarded)
//   O  = diag(P_a, P_b) [V_a; V_b]              (ONE k-step: the block-diagonal P is exactly the A fragment
//                                                {P_a, 0, 0, P_b} built from the S accumulators in registers)
// ~125 instructions per (pixel, head) instead of ~700 for the shuffle formulation above (which is kept for an odd
// head count): the kernel becomes a pure 8 B/element stream.  Rotary (first 32 dims, interleaved pairs) is applied to
// the 16-byte chunks on their way into shared memory; scale and the relative-position bias are applied to the fp32
// scores.
// ---------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(128)
    temporal_attn_mma_kernel(const TaParams p) {
  constexpr int CH = D / 8;             // 16-byte chunks per row
  constexpr int LD_ITERS = 16 * CH / 32;  // chunks per lane and tensor
  __shared__ __align__(16) __half smem[4][3][16 * D];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hp = p.heads >> 1;
  const int64_t pw = static_cast<int64_t>(blockIdx.x) * 4 + warp;
  const int64_t total = static_cast<int64_t>(p.B) * p.HW * hp;
  if (pw >= total) return;  // whole warp
  const int h0 = static_cast<int>(pw % hp) * 2;
  const int64_t pix = (pw / hp) % p.HW;
  const int64_t b = pw / (static_cast<int64_t>(hp) * p.HW);
  __half* sQ = smem[warp][0];
  __half* sK = smem[warp][1];
  __half* sV = smem[warp][2];

  // ---- global -> (rotary) -> swizzled smem tiles: row = item * 8 + frame.  All loads are issued before the first
  // use so that 3 * LD_ITERS 16-byte requests per lane are in flight ----
  uint4 v[3][LD_ITERS];
#pragma unroll
  for (int tsr = 0; tsr < 3; ++tsr) {
    const __half* src = tsr == 0 ? p.q : (tsr == 1 ? p.k : p.v);
    const int64_t ld = tsr == 0 ? p.ldq : (tsr == 1 ? p.ldk : p.ldv);
#pragma unroll
    for (int i = 0; i < LD_ITERS; ++i) {
      const int id = lane + 32 * i;
      const int row = id / CH, chunk = id % CH;
      const int item = row >> 3, frame = row & 7;
      v[tsr][i] = make_uint4(0, 0, 0, 0);
      if (frame < p.F) {
        const int64_t tok = (b * p.F + frame) * p.HW + pix;
        v[tsr][i] = ldg16(src + tok * ld + (h0 + item) * D + chunk * 8);
      }
    }
  }
#pragma unroll
  for (int tsr = 0; tsr < 3; ++tsr) {
    __half* dst = smem[warp][tsr];
#pragma unroll
    for (int i = 0; i < LD_ITERS; ++i) {
      const int id = lane + 32 * i;
      const int row = id / CH, chunk = id % CH;
      const int frame = row & 7;
      if (tsr < 2 && chunk < 4 && frame < p.F) {
        // rotary pairs 4*chunk .. 4*chunk+3 of this frame: (x0, x1) -> (x0 c - x1 s, x1 c + x0 s)
        const float4 r0 = __ldg(reinterpret_cast<const float4*>(p.rot + (frame * 16 + chunk * 4) * 2));
        const float4 r1 = __ldg(reinterpret_cast<const float4*>(p.rot + (frame * 16 + chunk * 4) * 2 + 4));
        const float cs[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
        __half2* h = reinterpret_cast<__half2*>(&v[tsr][i]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __half22float2(h[j]);
          const float c = cs[2 * j], sn = cs[2 * j + 1];
          h[j] = __floats2half2_rn(f.x * c - f.y * sn, f.y * c + f.x * sn);
        }
      }
      *reinterpret_cast<uint4*>(tile_ptr<D>(dst, row, chunk)) = v[tsr][i];
    }
  }
  __syncwarp();

  const int g = lane >> 2, t4 = lane & 3;
  // ---- S = Q K^T ----
  float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
  {
    const int arow = (lane & 7) + ((lane >> 3) & 1) * 8, achk = lane >> 4;
    const int brow = (lane & 7) + (lane >> 4) * 8, bchk = (lane >> 3) & 1;
#pragma unroll
    for (int kk = 0; kk < D / 16; ++kk) {
      uint32_t a[4], bfr[4];
      ldmatrix_x4(a, tile_ptr<D>(sQ, arow, kk * 2 + achk));
      ldmatrix_x4(bfr, tile_ptr<D>(sK, brow, kk * 2 + bchk));
      mma16816(s0, a, bfr[0], bfr[1]);  // keys of item a
      mma16816(s1, a, bfr[2], bfr[3]);  // keys of item b
    }
  }
  // ---- softmax over the <= 8 keys of each item: row g of item a lives in s0[0..1], row g of item b in s1[2..3] ----
  uint32_t pa, pb;
  {
    const int j0 = 2 * t4, j1 = j0 + 1;
    const bool row_ok = g < p.F;
    float xa0 = -INFINITY, xa1 = -INFINITY, xb0 = -INFINITY, xb1 = -INFINITY;
    if (row_ok && j0 < p.F) {
      xa0 = s0[0] * p.scale + __ldg(p.bias + (h0 * p.F + g) * p.F + j0);
      xb0 = s1[2] * p.scale + __ldg(p.bias + ((h0 + 1) * p.F + g) * p.F + j0);
    }
    if (row_ok && j1 < p.F) {
      xa1 = s0[1] * p.scale + __ldg(p.bias + (h0 * p.F + g) * p.F + j1);
      xb1 = s1[3] * p.scale + __ldg(p.bias + ((h0 + 1) * p.F + g) * p.F + j1);
    }
    float ma = fmaxf(xa0, xa1), mb = fmaxf(xb0, xb1);
    ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, 1));
    mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 1));
    ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, 2));
    mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 2));
    if (!row_ok) ma = mb = 0.f;  // padded query rows: all keys masked, keep the arithmetic finite
    const float ea0 = __expf(xa0 - ma), ea1 = __expf(xa1 - ma), eb0 = __expf(xb0 - mb), eb1 = __expf(xb1 - mb);
    float da = ea0 + ea1, db = eb0 + eb1;
    da += __shfl_xor_sync(0xffffffffu, da, 1);
    db += __shfl_xor_sync(0xffffffffu, db, 1);
    da += __shfl_xor_sync(0xffffffffu, da, 2);
    db += __shfl_xor_sync(0xffffffffu, db, 2);
    const float ia = da > 0.f ? 1.f / da : 0.f, ib = db > 0.f ? 1.f / db : 0.f;
    __half2 ha = __floats2half2_rn(ea0 * ia, ea1 * ia), hb = __floats2half2_rn(eb0 * ib, eb1 * ib);
    pa = *reinterpret_cast<uint32_t*>(&ha);
    pb = *reinterpret_cast<uint32_t*>(&hb);
  }
  // ---- O = diag(P_a, P_b) [V_a; V_b] ----
  float o[D / 8][4];
#pragma unroll
  for (int i = 0; i < D / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  {
    const uint32_t a[4] = {pa, 0u, 0u, pb};
    const int vrow = (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
    for (int nb = 0; nb < D / 16; ++nb) {
      uint32_t bfr[4];
      ldmatrix_x4_trans(bfr, tile_ptr<D>(sV, vrow, nb * 2 + (lane >> 4)));
      mma16816(o[nb * 2], a, bfr[0], bfr[1]);
      mma16816(o[nb * 2 + 1], a, bfr[2], bfr[3]);
    }
  }
  // ---- O -> smem (Q tile) -> 16-byte stores ----
  __syncwarp();
#pragma unroll
  for (int n = 0; n < D / 8; ++n) {
    *reinterpret_cast<__half2*>(tile_ptr<D>(sQ, g, n) + 2 * t4) = __floats2half2_rn(o[n][0], o[n][1]);
    *reinterpret_cast<__half2*>(tile_ptr<D>(sQ, g + 8, n) + 2 * t4) = __floats2half2_rn(o[n][2], o[n][3]);
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < LD_ITERS; ++i) {
    const int id = lane + 32 * i;
    const int row = id / CH, chunk = id % CH;
    const int item = row >> 3, frame = row & 7;
    if (frame < p.F) {
      const int64_t tok = (b * p.F + frame) * p.HW + pix;
      stg16(p.o + tok * p.ldo + (h0 + item) * D + chunk * 8, *reinterpret_cast<const uint4*>(tile_ptr<D>(sQ, row, chunk)));
    }
  }
}

uav_status_t attention_tc(const void* q, const void* k, const void* v, void* out, int64_t batch,
                          int heads, int head_dim, int64_t nq, int64_t nk, int64_t ldq, int64_t ldk,
                          int64_t ldv, int64_t ldo, int64_t kv_batch_div, float scale,
                          cudaStream_t stream);  // attention_tc.cu (tcgen05 / TMEM)

template <int DQK, int DV>
static uav_status_t launch_fa(const FaParams& p, int batch, cudaStream_t stream) {
  constexpr int smem = (FA_BM * DQK + 2 * FA_BN * DQK + 2 * FA_BN * DV) * 2;
  static uint64_t configured = 0;  // per-device bit: cudaFuncSetAttribute applies to the current device only
  const uint64_t dev_bit = 1ull << (current_device() & 63);
  if (!(configured & dev_bit)) {
    UAV_CHECK_CUDA(cudaFuncSetAttribute(flash_attn_kernel<DQK, DV>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured |= dev_bit;
  }
  dim3 grid((p.nq + FA_BM - 1) / FA_BM, batch * p.heads);
  flash_attn_kernel<DQK, DV><<<grid, FA_THREADS, smem, stream>>>(p);
  UAV_CHECK_CUDA(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return UAV_OK;
}

}  // namespace uav

using namespace uav;

extern "C" {

uav_status_t uav_attention(const void* q, const void* k, const void* v, void* out, int64_t batch,
                           int heads, int head_dim, int64_t nq, int64_t nk, int64_t ldq,
                           int64_t ldk, int64_t ldv, int64_t ldo, int64_t kv_batch_div,
                           float scale, uav_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  UAV_REQUIRE(q && k && v && out, "uav_attention: null pointer");
  UAV_REQUIRE(batch > 0 && heads > 0 && nq > 0 && nk > 0 && kv_batch_div > 0,
              "uav_attention: bad shape");
  UAV_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0,
              "uav_attention: token strides must be multiples of 8");
  UAV_REQUIRE(batch * heads <= 65535, "uav_attention: batch*heads too large");
  UAV_REQUIRE(batch % kv_batch_div == 0, "uav_attention: batch must be a multiple of kv_batch_div");
  // d = 128 (UNet self / cross attention at h/8) and d = 512 (VAE AttentionBlock) run on the tcgen05 / TMEM
  // kernel; UAV_ATTENTION_HMMA=1 selects the mma.sync kernel instead (kept for A/B measurements)
  static const bool no_cross = getenv("UAV_ATTENTION_NOCROSS") != nullptr && getenv("UAV_ATTENTION_NOCROSS")[0] == '1';
  if (!no_cross && nk <= 128 && nq >= 4 * nk && (head_dim == 64 || head_dim == 128)) {
    // short key/value sequence (the 77 prompt tokens): resident-KV streaming kernel
    FaParams pc;
    pc.q = (const __half*)q; pc.k = (const __half*)k; pc.v = (const __half*)v; pc.o = (__half*)out;
    pc.ldq = ldq; pc.ldk = ldk; pc.ldv = ldv; pc.ldo = ldo;
    pc.bsq = nq * ldq; pc.bsk = nk * ldk; pc.bsv = nk * ldv; pc.bso = nq * ldo;
    pc.nq = (int)nq; pc.nk = (int)nk; pc.heads = heads; pc.kv_batch_div = (int)kv_batch_div;
    pc.scale_log2 = scale * 1.4426950408889634f;
    if (head_dim == 64) return nk <= 80 ? launch_cross<64, 5>(pc, (int)batch, stream) : launch_cross<64, 8>(pc, (int)batch, stream);
    return nk <= 80 ? launch_cross<128, 5>(pc, (int)batch, stream) : launch_cross<128, 8>(pc, (int)batch, stream);
  }
  static const bool force_hmma = getenv("UAV_ATTENTION_HMMA") != nullptr && getenv("UAV_ATTENTION_HMMA")[0] == '1';
  if (!force_hmma && (head_dim == 128 || head_dim == 512))
    return attention_tc(q, k, v, out, batch, heads, head_dim, nq, nk, ldq, ldk, ldv, ldo, kv_batch_div, scale,
                        stream);
  FaParams p;
  p.q = (const __half*)q;
  p.k = (const __half*)k;
  p.v = (const __half*)v;
  p.o = (__half*)out;
  p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.ldo = ldo;
  p.bsq = nq * ldq; p.bsk = nk * ldk; p.bsv = nk * ldv; p.bso = nq * ldo;
  p.nq = (int)nq; p.nk = (int)nk; p.heads = heads; p.kv_batch_div = (int)kv_batch_div;
  p.scale_log2 = scale * 1.4426950408889634f;
  if (head_dim == 64) return launch_fa<64, 64>(p, (int)batch, stream);
  if (head_dim == 128) return launch_fa<128, 128>(p, (int)batch, stream);
  if (head_dim == 512) {
    // single wide head (VAE AttentionBlock): four passes over 128-wide V / output slices
    UAV_REQUIRE(heads == 1, "uav_attention: head_dim 512 supports a single head");
    for (int s = 0; s < 4; ++s) {
      FaParams ps = p;
      ps.v = p.v + s * 128;
      ps.o = p.o + s * 128;
      uav_status_t st = launch_fa<512, 128>(ps, (int)batch, stream);
      if (st != UAV_OK) return st;
    }
    return UAV_OK;
  }
  set_last_error("uav_attention: head_dim %d unsupported (64, 128, 512)", head_dim);
  return UAV_ERR_UNSUPPORTED;
}

uav_status_t uav_temporal_attention(const void* q, const void* k, const void* v, void* out,
                                    int64_t B, int64_t F, int64_t HW, int heads, int head_dim,
                                    int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo,
                                    float scale, const float* rot_cos_sin, const float* rel_bias,
                                    uav_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  UAV_REQUIRE(q && k && v && out && rot_cos_sin && rel_bias,
              "uav_temporal_attention: null pointer");
  UAV_REQUIRE(B > 0 && F > 0 && F <= 8 && HW > 0 && heads > 0,
              "uav_temporal_attention: bad shape (F=%lld must be <= 8)", (long long)F);
  UAV_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0,
              "uav_temporal_attention: token strides must be multiples of 8");
  TaParams p;
  p.q = (const __half*)q; p.k = (const __half*)k; p.v = (const __half*)v; p.o = (__half*)out;
  p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.ldo = ldo;
  p.B = (int)B; p.F = (int)F; p.heads = heads; p.HW = HW;
  p.scale = scale; p.rot = rot_cos_sin; p.bias = rel_bias;
  const int64_t warps = B * HW * heads;
  const unsigned grid = (unsigned)((warps + 7) / 8);
  static const bool force_shfl = getenv("UAV_TEMPORAL_SHFL") && getenv("UAV_TEMPORAL_SHFL")[0] == '1';
  if (heads % 2 == 0 && !force_shfl && (head_dim == 64 || head_dim == 128)) {
    // mma.sync formulation: one warp per pair of heads
    const unsigned grid2 = (unsigned)((warps / 2 + 3) / 4);
    if (head_dim == 64) temporal_attn_mma_kernel<64><<<grid2, 128, 0, stream>>>(p);
    else temporal_attn_mma_kernel<128><<<grid2, 128, 0, stream>>>(p);
  } else if (head_dim == 64) temporal_attn_kernel<64><<<grid, 256, 0, stream>>>(p);
  else if (head_dim == 128) temporal_attn_kernel<128><<<grid, 256, 0, stream>>>(p);
  else {
    set_last_error("uav_temporal_attention: head_dim %d unsupported (64, 128)", head_dim);
    return UAV_ERR_UNSUPPORTED;
  }
  UAV_CHECK_CUDA(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return UAV_OK;
}

}  // extern "C"
