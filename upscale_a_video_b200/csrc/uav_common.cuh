// uav_common.cuh — sm_100a PTX wrappers shared by the uav_b200 kernels.
// mbarrier / TMA (cp.async.bulk.tensor) / tcgen05 (alloc, mma, commit, ld) helpers.
// No torch, no CUTLASS: plain CUDA + inline PTX. Bit layouts of the UMMA shared-memory
// and instruction descriptors follow the PTX ISA "tcgen05 matrix descriptors" tables.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda.h>
#include <stdint.h>

#include "../../include/uav_b200.h"

namespace uav {

// ---------------------------------------------------------------------------------------
// error plumbing (host)
// ---------------------------------------------------------------------------------------
void set_last_error(const char* fmt, ...);
#define UAV_CHECK_CUDA(expr)                                                             \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      uav::set_last_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                  \
                          cudaGetErrorString(_e));                                       \
      return UAV_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)
#define UAV_REQUIRE(cond, ...)                                                           \
  do {                                                                                   \
    if (!(cond)) {                                                                       \
      uav::set_last_error(__VA_ARGS__);                                                  \
      return UAV_ERR_INVALID;                                                            \
    }                                                                                    \
  } while (0)

int num_sms();         // of the current device
int current_device();

// ---------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
      "@P1 bra DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "DONE:\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity), "r"(0x989680u)
      : "memory");
}

// ---- TMA ----
__device__ __forceinline__ void tma_prefetch_desc(const void* desc) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(desc)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const void* desc, uint64_t* bar, void* smem, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(const void* desc, uint64_t* bar, void* smem, int c0,
                                            int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(const void* desc, uint64_t* bar, void* smem, int c0,
                                            int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3), "r"(c4)
      : "memory");
}

// ---- CTA pairs (tcgen05 cta_group::2): barriers of the leader CTA are addressed through the cluster window ----
__device__ __forceinline__ uint32_t mapa_shared(uint32_t cta_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(cta_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_expect_tx_cluster(uint32_t cluster_addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.release.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_addr),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// arrive WITHOUT a memory fence.  The release form above compiles to MEMBAR.ALL.CTA + ERRBAR (14 % of the epilogue warps'
// stall samples in the GEGLU GEMM, ncu round 2).  Used to hand a tensor-memory accumulator back to the MMA warp: the only
// accesses that must be ordered before the arrive are the tcgen05.ld reads, which tcgen05.fence::before_thread_sync
// (issued before the arrive) / ::after_thread_sync (after the wait) order on their own.
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// wait with cluster-scope acquire (the arrivals come from the peer CTA)
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], %1, %2;\n\t"
      "@P1 bra DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "DONE:\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity), "r"(0x989680u)
      : "memory");
}
// TMA loads of a CTA pair: the box lands in THIS CTA's shared memory, the complete_tx goes to the mbarrier at
// `bar_cluster_addr`, which may live in the peer (leader) CTA
__device__ __forceinline__ void tma_load_2d_2sm(const void* desc, uint32_t bar_cluster_addr, void* smem, int c0,
                                                int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d_2sm(const void* desc, uint32_t bar_cluster_addr, void* smem, int c0,
                                                int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---- tcgen05 ----
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// whole warp executes; writes TMEM base address to *smem_dst
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; fp16/bf16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when all previously issued tcgen05.mma of this thread completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
// ---- CTA-pair (cta_group::2) variants: one warp of EACH CTA allocates / frees; only the leader issues MMAs ----
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
// D[256 x N] (+)= A[256 x 16] * B[N x 16]^T: rows 0-127 of A/D and rows 0..N/2-1 of B live in the leader CTA,
// the other halves at the same shared / tensor memory addresses of its peer
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: lane i of the warp reads TMEM lane (quadrant*32+i)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
        "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
        "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory descriptor for a K-major tile stored as rows of 128 bytes with the
// 128B swizzle (what a TMA box with 128-byte inner extent and CU_TENSOR_MAP_SWIZZLE_128B
// writes): 8-row atoms of 1024 bytes, SBO = 1024, LBO unused, version 1 (Blackwell).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);        // [0,14) start address >> 4
  d |= static_cast<uint64_t>(1) << 16;                            // [16,30) LBO (ignored)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;                    // [32,46) SBO >> 4
  d |= static_cast<uint64_t>(1) << 46;                            // [46,48) version = 1
  d |= static_cast<uint64_t>(2) << 61;                            // [61,64) SWIZZLE_128B
  return d;
}
// instruction descriptor: A,B K-major, fp32 accumulate; ab_fmt 0=f16 1=bf16 2=tf32
__host__ __device__ constexpr uint32_t umma_idesc(uint32_t ab_fmt, uint32_t M, uint32_t N) {
  return (1u << 4) | (ab_fmt << 7) | (ab_fmt << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// ---- misc math ----
// MUFU.EX2 / MUFU.RCP with flush-to-zero: the non-ftz forms (__expf, __fdividef) wrap every MUFU in denormal
// range fix-ups (FSETP + predicated FMULs: ~8 extra instructions per GELU), which matters in the GEMM epilogues that
// are issue-bound.  <= 2 ulp; results that would be denormal flush to zero (|x| > 87 for SiLU, > 13 for GELU tails).
__device__ __forceinline__ float ex2_ftz(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rcp_ftz(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// x * sigmoid(x): x -> -inf gives x * rcp(inf) = -0, x -> +inf gives x * rcp(1) = x
__device__ __forceinline__ float silu_f(float x) { return x * rcp_ftz(1.0f + ex2_ftz(-1.4426950408889634f * x)); }
// SiLU of two values with ONE reciprocal: 1 / d0 = d1 * (1 / (d0 d1)).  3 MUFU per pair instead of 4 — the GroupNorm+SiLU
// apply pass runs the XU pipe (MUFU + the fp16 pack) at 68 % while streaming at 0.73 of the HBM roofline (ncu, round 2).
// The exponent argument is clamped to 60 so that d0 d1 stays finite (x < -41.6: sigmoid ~ 8.7e-19 either way).
__device__ __forceinline__ void silu2_f(float& a, float& b) {
  const float d0 = 1.0f + ex2_ftz(fminf(-1.4426950408889634f * a, 60.0f));
  const float d1 = 1.0f + ex2_ftz(fminf(-1.4426950408889634f * b, 60.0f));
  const float r = rcp_ftz(d0 * d1);
  a *= r * d1;
  b *= r * d0;
}
// exact-GELU x * Phi(x) with ONE MUFU: Phi(-|x|) = 2^p(|x|), p = degree-8 polynomial fit of log2(erfc(|x| / sqrt 2) / 2)
// on [0, 6] (Chebyshev least squares weighted by Phi, converted to monomials in t = |x| / 3 - 1; tools/fit_gelu.py), then
// gelu(x) = max(x, 0) - |x| * Phi(-|x|) (x >= 0: x (1 - Phi(-x)); x < 0: x Phi(x)).  |x| is clamped to 6 where
// |x| Phi(-|x|) < 6e-9.  11 FP32 instructions + 1 MUFU.EX2 instead of 12 + 2 MUFU (Abramowitz-Stegun erfc with a
// reciprocal, round 1): the GEGLU epilogue is MUFU-throughput bound (16 lanes / clk / SM).  Absolute error of the result
// <= 4.8e-7 in fp32 evaluation (the rounding of x - r near |x| = 8; fit error 1.6e-7), same as the fp32
// 0.5 x (1 + erff(x / sqrt 2)) formula against float64.
__device__ __forceinline__ float gelu_erf_f(float x) {
  const float ax = fminf(fabsf(x), 6.0f);
  const float t = fmaf(ax, 0.3333333333333333f, -1.0f);
  float p = -0.005517261102795601f;
  p = fmaf(p, t, -0.012404678389430046f);
  p = fmaf(p, t, 0.020884426310658455f);
  p = fmaf(p, t, -0.02991572767496109f);
  p = fmaf(p, t, 0.09405267238616943f);
  p = fmaf(p, t, -0.2069002389907837f);
  p = fmaf(p, t, -6.035426616668701f);
  p = fmaf(p, t, -14.20971965789795f);
  p = fmaf(p, t, -9.532933235168457f);
  return fmaf(-ax, ex2_ftz(p), fmaxf(x, 0.0f));
}

// fp32 pair -> packed fp16 with saturation to +-65504 (one F2FP.SATFINITE): an activation that leaves the fp16 range
// clamps instead of becoming inf (which the next GroupNorm would turn into NaN for the whole group)
__device__ __forceinline__ uint32_t pack_half2_sat(float a, float b) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}

// 16-byte global access
__device__ __forceinline__ uint4 ldg16(const void* p) {
  return *reinterpret_cast<const uint4*>(p);
}
__device__ __forceinline__ void stg16(void* p, const uint4& v) {
  *reinterpret_cast<uint4*>(p) = v;
}

// TMA descriptor encode through the driver entry point (no link-time libcuda dependency)
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled get_encode_tiled();

}  // namespace uav
