// univl_b200 — shared device helpers for the sm_100a kernels.
//
// Thin inline-PTX wrappers for the Blackwell primitives the kernels use
// (mbarrier, TMA bulk-tensor copies, tcgen05 MMA / TMEM), plus small math and
// packing helpers.  Everything here is header-only and device-side except the
// error-string plumbing at the bottom.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace univl {

typedef __nv_bfloat16 bf16;

// ----------------------------------------------------------------------------
// error plumbing shared by all translation units (defined in api.cu)
// ----------------------------------------------------------------------------
int set_error(int code, const char* fmt, ...);
int usable_sms();        // SM count minus univl_set_reserved_sms (api.cu)
int usable_sm_pairs();  // TPCs with both SMs free under the same reservation
#define UNIVL_OK 0
#define UNIVL_ERR_ARG -1
#define UNIVL_ERR_CUDA -2
#define UNIVL_ERR_UNSUPPORTED -3

#define UNIVL_CHECK_ARG(cond, ...)                                   \
  do {                                                               \
    if (!(cond)) return univl::set_error(UNIVL_ERR_ARG, __VA_ARGS__); \
  } while (0)

#define UNIVL_CHECK_LAUNCH(name)                                                          \
  do {                                                                                    \
    cudaError_t e__ = cudaGetLastError();                                                 \
    if (e__ != cudaSuccess)                                                               \
      return univl::set_error(UNIVL_ERR_CUDA, "%s launch: %s", name, cudaGetErrorString(e__)); \
  } while (0)

// ----------------------------------------------------------------------------
// programmatic dependent launch (PDL): every hot kernel is launched with the programmatic-stream-serialization
// attribute, calls pdl_trigger() on entry (the NEXT kernel's CTAs may become resident as soon as all of ours have
// started) and pdl_wait() before its first global-memory access (blocks until the previous kernel has completed and
// flushed).  Launch latency and per-CTA set-up of kernel N+1 overlap the tail of kernel N — the text / visual encoder
// layers are ~25 kernels of 5-15 us each.  Opt-in with UNIVL_PDL=1 (without the attribute the device calls are no-ops);
// measured neutral (1806 vs 1816 samples/s) under CUDA-graph replay, so it is off by default.
// ----------------------------------------------------------------------------
bool pdl_enabled();  // api.cu
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_kernel_cluster(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                                cudaStream_t st, int cluster_x, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = (unsigned)cluster_x;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                        Args... args) {
  return launch_kernel_cluster(kern, grid, block, smem, st, 1, args...);
}

// ----------------------------------------------------------------------------
// generic helpers
// ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

// erf-GELU as the reference states it: x * 0.5 * (1 + erf(x / sqrt(2)))  (modules/until_module.py:28-33).
// erf is evaluated with Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7, two orders below bf16 resolution) because the
// GEMM epilogue is instruction-bound: one MUFU.RCP + one MUFU.EX2 + a handful of FMAs instead of libdevice's branchy
// erff.  The same exponential exp(-x^2/2) serves the derivative's Gaussian term.
// Arranged for the fewest issue slots (11 FP + 2 MUFU forward, 15 + 2 for the derivative):
//   w(x) = 1 - Phi(|x|) = 0.5 * t * poly(t) * exp(-x^2/2),  t = 1 / (1 + p |x| / sqrt(2))      (A&S 7.1.26)
//   gelu(x)  = max(x, 0) - |x| * w(x)
//   gelu'(x) = Phi(x) + x * phi(x),   Phi(x) = x >= 0 ? 1 - w : w
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// w = 1 - Phi(|x|) (in [0, 0.5]) and gauss = exp(-x^2 / 2)
__device__ __forceinline__ void erf_exp_terms(float x, float& w, float& gauss) {
  const float t = rcp_approx(fmaf(fabsf(x), 0.3275911f * 0.70710678118654752440f, 1.0f));
  gauss = ex2_approx((x * x) * (-0.5f * 1.44269504088896340736f));
  float poly = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
  poly = fmaf(poly, t, 0.5f * 1.421413741f);
  poly = fmaf(poly, t, 0.5f * -0.284496736f);
  poly = fmaf(poly, t, 0.5f * 0.254829592f);
  w = (poly * t) * gauss;
}
__device__ __forceinline__ float gelu_erf(float x) {
  float w, g;
  erf_exp_terms(x, w, g);
  return fmaf(-fabsf(x), w, fmaxf(x, 0.f));
}
// d/dx gelu_erf(x) = Phi(x) + x * phi(x)
__device__ __forceinline__ float gelu_erf_grad(float x) {
  float w, g;
  erf_exp_terms(x, w, g);
  const float cdf = x >= 0.f ? 1.0f - w : w;
  return fmaf(x * 0.39894228040143267794f, g, cdf);
}

// ----------------------------------------------------------------------------
// Philox4x32-10 counter RNG (dropout masks are regenerated in backward from
// (seed, offset) — never stored).
// ----------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32(uint64_t seed, uint64_t ctr_hi, uint64_t ctr_lo) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32);
  uint32_t c2 = (uint32_t)ctr_hi, c3 = (uint32_t)(ctr_hi >> 32);
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

// Keep-mask for element `idx` of dropout stream (seed, stream): element idx uses
// 32 random bits: word (idx & 3) of philox(seed, stream, idx >> 2).
__device__ __forceinline__ bool dropout_keep(uint64_t seed, uint64_t stream, uint64_t idx,
                                             uint32_t keep_threshold) {
  const uint4 r = philox4x32(seed, stream, idx >> 2);
  const uint32_t w = (idx & 3) == 0 ? r.x : (idx & 3) == 1 ? r.y : (idx & 3) == 2 ? r.z : r.w;
  return w < keep_threshold;
}
// threshold such that P(keep) = 1 - p  (p in [0,1))
__host__ __device__ __forceinline__ uint32_t dropout_threshold(float p) {
  double keep = 1.0 - (double)p;
  if (keep >= 1.0) return 0xFFFFFFFFu;
  return (uint32_t)(keep * 4294967296.0);
}
// 16-bit variant: an element is kept iff its 16 random bits are < threshold16, so one Philox4x32 call (128 bits)
// decides 8 elements; P(keep) is exact to 2^-16, far below the sampling noise of any dropout mask.
__host__ __device__ __forceinline__ uint32_t dropout_threshold16(float p) {
  double keep = 1.0 - (double)p;
  if (keep >= 1.0) return 65536u;
  return (uint32_t)(keep * 65536.0 + 0.5);
}
// 16-bit word w (0..7) of a Philox output
__device__ __forceinline__ uint32_t philox_u16(const uint4& r, int w) {
  const uint32_t x = (w >> 1) == 0 ? r.x : (w >> 1) == 1 ? r.y : (w >> 1) == 2 ? r.z : r.w;
  return (w & 1) ? (x >> 16) : (x & 0xFFFFu);
}
// keep-mask (bit j = keep element idx0 + j) of 8 consecutive elements, idx0 % 8 == 0: ONE Philox call
__device__ __forceinline__ uint32_t dropout_keep8(uint64_t seed, uint64_t stream, uint64_t idx0, uint32_t thr16) {
  const uint4 r = philox4x32(seed, stream, idx0 >> 3);
  return ((r.x & 0xFFFFu) < thr16 ? 1u : 0u) | ((r.x >> 16) < thr16 ? 2u : 0u) | ((r.y & 0xFFFFu) < thr16 ? 4u : 0u) |
         ((r.y >> 16) < thr16 ? 8u : 0u) | ((r.z & 0xFFFFu) < thr16 ? 16u : 0u) | ((r.z >> 16) < thr16 ? 32u : 0u) |
         ((r.w & 0xFFFFu) < thr16 ? 64u : 0u) | ((r.w >> 16) < thr16 ? 128u : 0u);
}

// ----------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
// non-suspending probe (mbarrier.try_wait may put the thread to sleep for a system-dependent time when the phase is not
// complete: fine for a blocking wait, wrong for a scheduler that polls several barriers)
__device__ __forceinline__ uint32_t mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
// Bounded wait: a pipeline bug must surface as a trap (launch failure), never as
// a hung GPU.  try_wait already suspends in hardware, so the spin count stays low
// in healthy runs; 1<<26 polls is many seconds.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) {
      printf("univl: mbarrier wait timed out (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y,
             blockIdx.z, threadIdx.x);
      __trap();
    }
  }
}

// ----------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor) — 2D tiled loads into shared memory
// ----------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}

// multicast variant: the box lands at the same shared-memory offset of every CTA in cta_mask (bit i = cluster rank i) and
// completes tx bytes on the mbarrier at the same offset in each of them
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                               uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%2, %3}], [%4], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---- TMA stores (shared -> global) -----------------------------------------------------------------------------
// bulk tensor store / reduce of one staging box; coordinates {inner (column), outer (row)}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(c0), "r"(c1), "r"(smem_u32(smem_src))
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(c0), "r"(c1), "r"(smem_u32(smem_src))
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all but the most recent N bulk groups of this thread have finished READING their shared-memory source
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ----------------------------------------------------------------------------
// tcgen05 / TMEM
// ----------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// One lane of a fully converged warp (elect.sync).  The MMA-issuing warp runs its loop with ALL lanes (warp-uniform
// control flow keeps addresses / descriptors in uniform registers) and wraps only the tcgen05.mma / commit instructions
// in `if (elect_one())`; an `if (lane == 0)` region around the whole loop makes the compiler re-broadcast every operand
// (ELECT + R2UR per instruction), which measurably throttles the single issuing thread.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
// whole warp; writes the TMEM base address to *smem_slot
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; single thread issues
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued MMAs of this thread retire
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// same, arriving on the barrier at this shared-memory offset in every CTA of cta_mask (cluster ranks)
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(cta_mask)
      : "memory");
}
// 32 lanes x 16 consecutive fp32 columns: thread t of the warp reads TMEM lane (base_lane + t)
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Instruction descriptor, kind::f16, bf16 x bf16 -> fp32 (bit layout: CUTLASS
// cute/arch/mma_sm100_desc.hpp UMMA::InstrDescriptor).
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4)                          // c_format = F32
         | (1u << 7)                        // a_format = BF16
         | (1u << 10)                       // b_format = BF16
         | ((a_mn_major ? 1u : 0u) << 15)   // a_major
         | ((b_mn_major ? 1u : 0u) << 16)   // b_major
         | ((uint32_t)(N >> 3) << 17)       // n_dim
         | ((uint32_t)(M >> 4) << 24);      // m_dim
}

// Shared-memory matrix descriptor, SWIZZLE_128B (UMMA::SmemDescriptor): start
// address, leading/stride byte offsets all in 16-byte units, version=1 (sm_100).
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes,
                                                         uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;  // version
  d |= (uint64_t)2 << 61;  // layout_type = SWIZZLE_128B
  return d;
}

}  // namespace univl
