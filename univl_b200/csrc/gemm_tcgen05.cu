// univl_b200 — tcgen05 / TMEM / TMA GEMM for sm_100a.
//
//   D[M,N] = epilogue( sum_k A(m,k) * B(n,k) )          bf16 operands, fp32 accumulation in TMEM
//
// One kernel template serves the three GEMMs of every nn.Linear on the UniVL hot path
// (reference: modules/module_bert.py:172-174,208,234,247 and the autograd mirror):
//   forward  Y  = X  W^T   A = X  [T,K]  K-major     B = W [N,K]  K-major
//   dgrad    dX = dY W     A = dY [T,N]  K-major     B = W [N,K]  MN-major (contraction over N)
//   wgrad    dW = dY^T X   A = dY [T,N]  MN-major    B = X [T,K]  MN-major (contraction over T)
// so neither weights nor activations are ever transposed in HBM: the operand "major" is a bit in the
// UMMA instruction descriptor and a different TMA box shape.
//
// Structure (per CTA, 192 threads, one 128 x BLOCK_N output tile, optional split-K over blockIdx.z):
//   warp 0      TMA producer   : cp.async.bulk.tensor 2D boxes (128B swizzle) into a STAGES-deep smem ring
//   warp 1      MMA issuer     : one lane issues tcgen05.mma (M=128, N=BLOCK_N, K=16), tcgen05.commit frees stages
//   warps 2..5  epilogue       : tcgen05.ld TMEM -> registers -> fused epilogue -> global
// Synchronisation is mbarrier-only (full/empty per stage, one "accumulator ready" barrier).
#include "common.cuh"
#include "tmap.cuh"

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

namespace univl {

enum GemmEpilogue : int {
  EPI_BIAS_BF16 = 0,      // out(bf16) = alpha*acc + bias
  EPI_BIAS_GELU_BF16 = 1, // aux_out(bf16) = acc + bias ; out(bf16) = gelu_erf(acc + bias)
  EPI_GELU_BWD_BF16 = 2,  // out(bf16) = acc * gelu_erf'(aux_in)
  EPI_ADD_BF16 = 3,       // out(bf16) = alpha*acc + aux_in
  EPI_BIAS_F32 = 4,       // out(f32)  = alpha*acc + bias
  EPI_ATOMIC_F32 = 5,     // out(f32) += alpha*acc      (red.global.add; split-K / gradient accumulation)
};

struct GemmParams {
  int M, N, Kc;
  int k_blocks_per_split;
  int epilogue;
  float alpha;
  void* out;
  long long ldo;
  const float* bias;
  const bf16* aux_in;
  long long ld_aux_in;
  bf16* aux_out;
  long long ld_aux_out;
  int tma_epilogue;  // persistent kernels: 1 = bulk tensor stores through tmap_out / tmap_aux, 0 = manual stores
};

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 192;

template <int BLOCK_N, int STAGES>
struct GemmSmem {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  // full[STAGES], empty[STAGES], tmem_full, tmem slot
  static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 1) * 8 + 16;
  static constexpr int DYN_BYTES = TOTAL + 1024;  // slack for manual 1024B alignment
};

template <int BLOCK_N, int STAGES, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(GEMM_THREADS)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const GemmParams p) {
  using L = GemmSmem<BLOCK_N, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BLOCK_M;
  const int n0 = blockIdx.y * BLOCK_N;
  const int total_kb = (p.Kc + BLOCK_K - 1) / BLOCK_K;
  const int kb_begin = blockIdx.z * p.k_blocks_per_split;
  const int kb_end = min(total_kb, kb_begin + p.k_blocks_per_split);
  const int num_kb = kb_end - kb_begin;  // host guarantees >= 1

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, BLOCK_N);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* sa = smem + s * L::STAGE_BYTES;
        uint8_t* sb = sa + L::A_BYTES;
        mbar_arrive_expect_tx(&full_bar[s], L::STAGE_BYTES);
        const int k_elem = (kb_begin + i) * BLOCK_K;
        if (!A_MN) {
          tma_load_2d(sa, &tmap_a, &full_bar[s], k_elem, m0);  // box {64 k, 128 rows}
        } else {
#pragma unroll
          for (int c = 0; c < BLOCK_M / 64; ++c)  // box {64 mn, 64 k-rows}
            tma_load_2d(sa + c * (BLOCK_K * 128), &tmap_a, &full_bar[s], m0 + c * 64, k_elem);
        }
        if (!B_MN) {
          tma_load_2d(sb, &tmap_b, &full_bar[s], k_elem, n0);  // box {64 k, BLOCK_N rows}
        } else {
#pragma unroll
          for (int c = 0; c < BLOCK_N / 64; ++c)
            tma_load_2d(sb + c * (BLOCK_K * 128), &tmap_b, &full_bar[s], n0 + c * 64, k_elem);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer --------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M, BLOCK_N, A_MN, B_MN);
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after_sync();
        const uint32_t sa = smem_u32(smem + s * L::STAGE_BYTES);
        const uint32_t sb = sa + L::A_BYTES;
#pragma unroll
        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
          // K-major : 8-row groups are 1024 B apart (SBO); a K=16 slice is 32 B inside the swizzled row.
          // MN-major: 64-wide MN chunks are BLOCK_K*128 B apart (LBO); 8-k-row groups 1024 B apart (SBO);
          //           a K=16 slice is two such groups = 2048 B.
          const uint64_t da = A_MN ? make_smem_desc_sw128(sa + k * 2048, BLOCK_K * 128, 1024)
                                   : make_smem_desc_sw128(sa + k * 32, 16, 1024);
          const uint64_t db = B_MN ? make_smem_desc_sw128(sb + k * 2048, BLOCK_K * 128, 1024)
                                   : make_smem_desc_sw128(sb + k * 32, 16, 1024);
          umma_bf16(tmem_base, da, db, idesc, (i > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);  // stage reusable once these MMAs have read it
      }
      umma_commit(tmem_full_bar);  // accumulator complete
    }
  } else {
    // ------------------------------ epilogue ----------------------------------
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int row = m0 + q * 32 + lane;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after_sync();
    const bool row_ok = row < p.M;
    const int epi = p.epilogue;
    const float alpha = p.alpha;
#pragma unroll 1
    for (int c = 0; c < BLOCK_N; c += 16) {
      uint32_t r[16];
      __syncwarp();  // tcgen05.ld is .sync.aligned: reconverge after the divergent stores below
      tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, r);
      tmem_ld_wait();
      const int col = n0 + c;
      if (!row_ok || col >= p.N) continue;
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]) * alpha;
      const bool full = (col + 16 <= p.N);
      if (p.bias != nullptr && (epi == EPI_BIAS_BF16 || epi == EPI_BIAS_GELU_BF16 || epi == EPI_BIAS_F32)) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (full || col + j < p.N) v[j] += __ldg(p.bias + col + j);
      }
      if (epi == EPI_ATOMIC_F32) {
        float* o = reinterpret_cast<float*>(p.out) + (long long)row * p.ldo + col;
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (full || col + j < p.N) atomicAdd(o + j, v[j]);
        continue;
      }
      if (epi == EPI_BIAS_F32) {
        float* o = reinterpret_cast<float*>(p.out) + (long long)row * p.ldo + col;
        if (full && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (col + j < p.N) o[j] = v[j];
        }
        continue;
      }
      // bf16 outputs
      if (epi == EPI_GELU_BWD_BF16 || epi == EPI_ADD_BF16) {
        const bf16* a = p.aux_in + (long long)row * p.ld_aux_in + col;
        float x[16];
        if (full && (reinterpret_cast<uintptr_t>(a) & 15) == 0) {
          const uint4 u0 = *reinterpret_cast<const uint4*>(a);
          const uint4 u1 = *reinterpret_cast<const uint4*>(a + 8);
          const uint32_t w[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float2 f = unpack_bf16x2(w[j]);
            x[2 * j] = f.x;
            x[2 * j + 1] = f.y;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) x[j] = (col + j < p.N) ? __bfloat162float(a[j]) : 0.f;
        }
        if (epi == EPI_GELU_BWD_BF16) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] *= gelu_erf_grad(x[j]);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] += x[j];
        }
      }
      if (epi == EPI_BIAS_GELU_BF16) {
        bf16* ao = p.aux_out + (long long)row * p.ld_aux_out + col;
        if (full && (reinterpret_cast<uintptr_t>(ao) & 15) == 0) {
          uint4 u0, u1;
          u0.x = pack_bf16x2(v[0], v[1]);   u0.y = pack_bf16x2(v[2], v[3]);
          u0.z = pack_bf16x2(v[4], v[5]);   u0.w = pack_bf16x2(v[6], v[7]);
          u1.x = pack_bf16x2(v[8], v[9]);   u1.y = pack_bf16x2(v[10], v[11]);
          u1.z = pack_bf16x2(v[12], v[13]); u1.w = pack_bf16x2(v[14], v[15]);
          *reinterpret_cast<uint4*>(ao) = u0;
          *reinterpret_cast<uint4*>(ao + 8) = u1;
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (col + j < p.N) ao[j] = __float2bfloat16(v[j]);
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = gelu_erf(v[j]);
      }
      bf16* o = reinterpret_cast<bf16*>(p.out) + (long long)row * p.ldo + col;
      if (full && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
        uint4 u0, u1;
        u0.x = pack_bf16x2(v[0], v[1]);   u0.y = pack_bf16x2(v[2], v[3]);
        u0.z = pack_bf16x2(v[4], v[5]);   u0.w = pack_bf16x2(v[6], v[7]);
        u1.x = pack_bf16x2(v[8], v[9]);   u1.y = pack_bf16x2(v[10], v[11]);
        u1.z = pack_bf16x2(v[12], v[13]); u1.w = pack_bf16x2(v[14], v[15]);
        *reinterpret_cast<uint4*>(o) = u0;
        *reinterpret_cast<uint4*>(o + 8) = u1;
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (col + j < p.N) o[j] = __float2bfloat16(v[j]);
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, BLOCK_N);
  }
}

}  // namespace univl

#include "gemm_persistent.cuh"

namespace univl {

// ----------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------

template <int BLOCK_N, int STAGES, bool A_MN, bool B_MN>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int splits,
                       cudaStream_t stream) {
  using L = GemmSmem<BLOCK_N, STAGES>;
  auto kern = gemm_tcgen05_kernel<BLOCK_N, STAGES, A_MN, B_MN>;
  // per launch: the attribute is per device and callers may drive several devices from one process
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYN_BYTES);
  if (e != cudaSuccess) return set_error(UNIVL_ERR_CUDA, "gemm smem attribute: %s", cudaGetErrorString(e));
  dim3 grid((p.M + BLOCK_M - 1) / BLOCK_M, (p.N + BLOCK_N - 1) / BLOCK_N, splits);
  kern<<<grid, GEMM_THREADS, L::DYN_BYTES, stream>>>(ta, tb, p);
  UNIVL_CHECK_LAUNCH("gemm_tcgen05");
  return UNIVL_OK;
}

template <int BLOCK_N, int STAGES, bool A_MN, bool B_MN>
static int launch_gemm_persistent(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to,
                                  const CUtensorMap& tx, const GemmParams& p, int splits, cudaStream_t stream) {
  using L = GemmSmemP<BLOCK_N, STAGES>;
  auto kern = gemm_tcgen05_persistent_kernel<BLOCK_N, STAGES, A_MN, B_MN>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYN_BYTES);
  if (e != cudaSuccess) return set_error(UNIVL_ERR_CUDA, "gemm smem attribute: %s", cudaGetErrorString(e));
  const int sms = usable_sms();
  const long long work =
      (long long)((p.M + BLOCK_M - 1) / BLOCK_M) * ((p.N + BLOCK_N - 1) / BLOCK_N) * (long long)splits;
  if (work > 0x7fffffffLL) return set_error(UNIVL_ERR_ARG, "gemm: too many tiles");
  const int grid = (int)(work < sms ? work : sms);
  e = launch_kernel(kern, dim3(grid), dim3(GEMM_P_THREADS), (size_t)L::DYN_BYTES, stream, ta, tb, to, tx, p, (int)work);
  if (e != cudaSuccess) return set_error(UNIVL_ERR_CUDA, "gemm launch: %s", cudaGetErrorString(e));
  UNIVL_CHECK_LAUNCH("gemm_tcgen05_persistent");
  return UNIVL_OK;
}

template <int BLOCK_N, int STAGES>
static int dispatch_major_p(bool a_mn, bool b_mn, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to,
                            const CUtensorMap& tx, const GemmParams& p, int splits, cudaStream_t stream) {
  if (!a_mn && !b_mn) return launch_gemm_persistent<BLOCK_N, STAGES, false, false>(ta, tb, to, tx, p, splits, stream);
  if (!a_mn && b_mn) return launch_gemm_persistent<BLOCK_N, STAGES, false, true>(ta, tb, to, tx, p, splits, stream);
  if (a_mn && b_mn) return launch_gemm_persistent<BLOCK_N, STAGES, true, true>(ta, tb, to, tx, p, splits, stream);
  return launch_gemm_persistent<BLOCK_N, STAGES, true, false>(ta, tb, to, tx, p, splits, stream);
}

template <int BLOCK_N, int STAGES, bool A_MN, bool B_MN>
static int launch_gemm_2cta(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to,
                            const CUtensorMap& tx, const GemmParams& p, int splits, cudaStream_t stream) {
  using L = GemmSmem2<BLOCK_N, STAGES>;
  auto kern = gemm_tcgen05_2cta_kernel<BLOCK_N, STAGES, A_MN, B_MN>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYN_BYTES);
  if (e != cudaSuccess) return set_error(UNIVL_ERR_CUDA, "gemm smem attribute: %s", cudaGetErrorString(e));
  const int max_pairs = usable_sm_pairs();
  const long long work =
      (long long)((p.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M)) * ((p.N + BLOCK_N - 1) / BLOCK_N) * (long long)splits;
  if (work > 0x7fffffffLL) return set_error(UNIVL_ERR_ARG, "gemm: too many tiles");
  const int pairs = (int)(work < max_pairs ? work : max_pairs);
  e = launch_kernel(kern, dim3(2 * pairs), dim3(GEMM_P_THREADS), (size_t)L::DYN_BYTES, stream, ta, tb, to, tx, p, (int)work);
  if (e != cudaSuccess) return set_error(UNIVL_ERR_CUDA, "gemm launch: %s", cudaGetErrorString(e));
  UNIVL_CHECK_LAUNCH("gemm_tcgen05_2cta");
  return UNIVL_OK;
}

template <int BLOCK_N, int STAGES>
static int dispatch_major_2(bool a_mn, bool b_mn, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to,
                            const CUtensorMap& tx, const GemmParams& p, int splits, cudaStream_t stream) {
  if (!a_mn && !b_mn) return launch_gemm_2cta<BLOCK_N, STAGES, false, false>(ta, tb, to, tx, p, splits, stream);
  if (!a_mn && b_mn) return launch_gemm_2cta<BLOCK_N, STAGES, false, true>(ta, tb, to, tx, p, splits, stream);
  if (a_mn && b_mn) return launch_gemm_2cta<BLOCK_N, STAGES, true, true>(ta, tb, to, tx, p, splits, stream);
  return launch_gemm_2cta<BLOCK_N, STAGES, true, false>(ta, tb, to, tx, p, splits, stream);
}

// 0 = auto (CTA pairs when the problem fills them), 1 = never, 2 = always when N >= 256
static int pair_mode() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("UNIVL_GEMM_PAIR");
    v = e ? atoi(e) : 0;
  }
  return v;
}

static bool use_v1_kernel() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("UNIVL_GEMM_V1");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

template <int BLOCK_N, int STAGES>
static int dispatch_major(bool a_mn, bool b_mn, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p,
                          int splits, cudaStream_t stream) {
  if (!a_mn && !b_mn) return launch_gemm<BLOCK_N, STAGES, false, false>(ta, tb, p, splits, stream);
  if (!a_mn && b_mn) return launch_gemm<BLOCK_N, STAGES, false, true>(ta, tb, p, splits, stream);
  if (a_mn && b_mn) return launch_gemm<BLOCK_N, STAGES, true, true>(ta, tb, p, splits, stream);
  return launch_gemm<BLOCK_N, STAGES, true, false>(ta, tb, p, splits, stream);
}

}  // namespace univl

using namespace univl;

namespace {
struct GemmPlan {
  int bn, splits, kb_per;
  bool pair;
};
// tile width, split-K factor and kernel variant for a problem — a pure function of the arguments (and the bring-up
// environment switches), shared by the launcher and by univl_gemm_plan
int plan_gemm(int M, int N, int Kc, int epilogue, int block_n, int split_k, GemmPlan* plan) {
  const int total_kb = (Kc + BLOCK_K - 1) / BLOCK_K;
  const int m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
  // tile width: widest tile that still yields >= ~1 wave of CTAs on 148 SMs
  // tuning overrides carried in block_n: +512 forces the CTA-pair kernel, +1024 forces single-CTA tiles
  const int force_pair = (block_n & 512) ? 1 : (block_n & 1024) ? -1 : 0;
  int bn = block_n & 511;
  if (bn == 0) {
    bn = 256;
    // without split-K the tile count alone must fill the SMs; with the atomic epilogue split-K supplies the
    // parallelism, so keep the widest (most smem-bandwidth-efficient) MMA shape
    if (epilogue != EPI_ATOMIC_F32)
      // (>= 0.9 wave counts as a wave: 144 tiles of 128x256 move 590 KB each through L2, 288 tiles of 128x128 move
      // 2 x 393 KB per SM — these small-M GEMMs are bound by the per-SM L2 ingress, so fewer bytes per SM wins)
      while (bn > 64 && (long long)m_tiles * ((N + bn - 1) / bn) < 132) bn >>= 1;
    if (N <= 64) bn = 64;
    else if (N <= 128 && bn > 128) bn = 128;
  }
  UNIVL_CHECK_ARG(bn == 64 || bn == 128 || bn == 256, "gemm: block_n must be 64/128/256");
  const int n_tiles = (N + bn - 1) / bn;
  int splits = 1;
  if (epilogue == EPI_ATOMIC_F32) {
    splits = split_k;
    if (splits == 0) {
      const long long tiles = (long long)m_tiles * n_tiles;
      splits = (int)((2 * 148 + tiles - 1) / tiles);
      if (splits > total_kb / 2) splits = total_kb / 2;  // keep >= 2 k-blocks per split
      if (splits < 1) splits = 1;
    }
    if (splits > total_kb) splits = total_kb;
  }
  const int kb_per = (total_kb + splits - 1) / splits;
  splits = (total_kb + kb_per - 1) / kb_per;  // no empty split
  // CTA-pair kernel (256 x 256 tiles, half the B traffic per CTA) when the pairs can be kept busy
  bool pair = false;
  if (!use_v1_kernel() && bn == 256 && pair_mode() != 1) {
    const long long pair_work = (long long)((M + 2 * BLOCK_M - 1) / (2 * BLOCK_M)) * n_tiles * splits;
    pair = pair_mode() == 2 || pair_work >= 60;
  }
  if (force_pair != 0 && !use_v1_kernel() && bn == 256) pair = force_pair > 0;
  plan->bn = bn; plan->splits = splits; plan->kb_per = kb_per; plan->pair = pair;
  return UNIVL_OK;
}
}  // namespace

// Which kernel univl_gemm_bf16 launches for this problem: 2 = CTA-pair (gemm_tcgen05_2cta_kernel), 1 = single-CTA
// persistent, 0 = bring-up kernel; negative = error.  Stateless (a function of its arguments): bench.py uses it to
// attribute per-launch CUDA-event times to the dominant kernel.
extern "C" int univl_gemm_plan(int M, int N, int Kc, int epilogue, int block_n, int split_k) {
  GemmPlan plan;
  if (int rc = plan_gemm(M, N, Kc, epilogue, block_n, split_k, &plan)) return rc;
  if (use_v1_kernel()) return 0;
  return plan.pair ? 2 : 1;
}

extern "C" int univl_gemm_bf16(const void* A, long long lda, int a_mn_major, const void* B, long long ldb,
                               int b_mn_major, int M, int N, int Kc, void* out, long long ldo, int epilogue,
                               const float* bias, const void* aux_in, long long ld_aux_in, void* aux_out,
                               long long ld_aux_out, float alpha, int block_n, int split_k, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  UNIVL_CHECK_ARG(M > 0 && N > 0 && Kc > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, Kc);
  UNIVL_CHECK_ARG(A && B && out, "gemm: null operand");
  UNIVL_CHECK_ARG((lda % 8) == 0 && (ldb % 8) == 0, "gemm: lda/ldb must be multiples of 8 elements (got %lld, %lld)",
                  lda, ldb);
  UNIVL_CHECK_ARG(((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0, "gemm: operands must be 16-byte aligned");
  UNIVL_CHECK_ARG(epilogue >= 0 && epilogue <= 5, "gemm: unknown epilogue %d", epilogue);
  if (epilogue == EPI_GELU_BWD_BF16 || epilogue == EPI_ADD_BF16)
    UNIVL_CHECK_ARG(aux_in != nullptr, "gemm: epilogue %d needs aux_in", epilogue);
  if (epilogue == EPI_BIAS_GELU_BF16) UNIVL_CHECK_ARG(aux_out != nullptr, "gemm: gelu epilogue needs aux_out");
  UNIVL_CHECK_ARG(split_k >= 0, "gemm: bad split_k");
  if (epilogue != EPI_ATOMIC_F32) UNIVL_CHECK_ARG(split_k <= 1, "gemm: split_k>1 needs the atomic epilogue");

  GemmPlan plan;
  if (int prc = plan_gemm(M, N, Kc, epilogue, block_n, split_k, &plan)) return prc;
  const int bn = plan.bn, splits = plan.splits, kb_per = plan.kb_per;
  const bool pair = plan.pair;

  CUtensorMap ta, tb;
  int rc;
  if (!a_mn_major) rc = make_tmap(&ta, A, M, Kc, lda, BLOCK_M);   // [M, Kc], box {64 k, 128 rows}
  else             rc = make_tmap(&ta, A, Kc, M, lda, BLOCK_K);   // [Kc, M], box {64 m, 64 k-rows}
  if (rc) return rc;
  if (!b_mn_major) rc = make_tmap(&tb, B, N, Kc, ldb, pair ? bn / 2 : bn);
  else             rc = make_tmap(&tb, B, Kc, N, ldb, BLOCK_K);
  if (rc) return rc;

  GemmParams p;
  p.M = M; p.N = N; p.Kc = Kc;
  p.k_blocks_per_split = kb_per;
  p.epilogue = epilogue;
  p.alpha = alpha;
  p.out = out; p.ldo = ldo;
  p.bias = bias;
  p.aux_in = reinterpret_cast<const bf16*>(aux_in); p.ld_aux_in = ld_aux_in;
  p.aux_out = reinterpret_cast<bf16*>(aux_out); p.ld_aux_out = ld_aux_out;
  p.tma_epilogue = 0;

  const bool amn = a_mn_major != 0, bmn = b_mn_major != 0;
  if (!use_v1_kernel()) {
    // epilogue operands by bulk tensor copies when their layout allows it (16-byte aligned base and row pitch)
    const bool out_f32 = epilogue == EPI_BIAS_F32 || epilogue == EPI_ATOMIC_F32;
    const void* aux = epilogue == EPI_BIAS_GELU_BF16 ? aux_out
                      : (epilogue == EPI_GELU_BWD_BF16 || epilogue == EPI_ADD_BF16) ? aux_in : nullptr;
    const long long ld_aux = epilogue == EPI_BIAS_GELU_BF16 ? ld_aux_out : ld_aux_in;
    bool tma_ok = ((uintptr_t)out & 15) == 0 && ((ldo * (out_f32 ? 4 : 2)) % 16) == 0;
    if (aux != nullptr) tma_ok = tma_ok && ((uintptr_t)aux & 15) == 0 && ((ld_aux * 2) % 16) == 0;
    static int force_manual = -1;
    if (force_manual < 0) {
      const char* e = getenv("UNIVL_GEMM_MANUAL_EPILOGUE");
      force_manual = (e && e[0] == '1') ? 1 : 0;
    }
    if (force_manual) tma_ok = false;
    CUtensorMap to, tx;
    if (tma_ok) {
      if ((rc = make_tmap_epi(&to, out, out_f32, M, N, ldo))) return rc;
      if (aux != nullptr) {
        if ((rc = make_tmap_epi(&tx, aux, false, M, N, ld_aux))) return rc;
      } else {
        tx = to;
      }
    } else {
      to = ta;  // valid descriptors, never dereferenced
      tx = ta;
    }
    p.tma_epilogue = tma_ok ? 1 : 0;
    if (pair) return dispatch_major_2<256, 5>(amn, bmn, ta, tb, to, tx, p, splits, stream);
    if (bn == 256) return dispatch_major_p<256, 3>(amn, bmn, ta, tb, to, tx, p, splits, stream);
    if (bn == 128) return dispatch_major_p<128, 5>(amn, bmn, ta, tb, to, tx, p, splits, stream);
    return dispatch_major_p<64, 6>(amn, bmn, ta, tb, to, tx, p, splits, stream);
  }
  if (bn == 256) return dispatch_major<256, 4>(amn, bmn, ta, tb, p, splits, stream);
  if (bn == 128) return dispatch_major<128, 3>(amn, bmn, ta, tb, p, splits, stream);
  return dispatch_major<64, 4>(amn, bmn, ta, tb, p, splits, stream);
}
