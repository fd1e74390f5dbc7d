// univl_b200 — persistent, warp-specialised tcgen05 GEMMs (included by gemm_tcgen05.cu).
//
// Same math and operand conventions as gemm_tcgen05_kernel, restructured so the tensor pipe never waits for the
// epilogue:
//   * grid = min(#work items, #SMs); each CTA (or CTA pair) walks work items round-robin in the order decode_work()
//     defines (n-tile fastest without split-K so activation tiles are re-read from L2; split-major with split-K)
//   * TWO accumulator buffers in TMEM (2 x BLOCK_N fp32 columns): the MMA warp fills buffer (t+1)&1 while the eight
//     epilogue warps drain buffer t&1 (tmem_full / tmem_empty mbarriers per buffer)
//   * the TMA producer's smem ring runs continuously across tiles (no pipeline drain between tiles)
//   * epilogue: tcgen05.ld -> registers -> fused math -> swizzled shared-memory staging tile -> ONE bulk tensor store
//     (cp.async.bulk.tensor, or cp.reduce.async.bulk.tensor .add for the split-K / gradient-accumulation epilogue)
//     per 32-row x 128-byte box, issued by one lane and double-buffered, so the epilogue warps spend their issue slots
//     on math instead of address arithmetic and per-row stores; the auxiliary operand of the GELU' / residual-add
//     epilogues comes in through the same staging buffers by TMA load.  Outputs whose leading dimension is not a
//     multiple of 16 bytes fall back to register -> staging -> coalesced st.global / red.global.
#pragma once

namespace univl {

constexpr int EPI_WARPS = 8;                                   // two warps per TMEM lane quarter, alternating chunks
constexpr int GEMM_P_THREADS = 64 + EPI_WARPS * 32;            // TMA warp + MMA warp + epilogue warps
constexpr int STAGING_ROW_BYTES = 128;                         // 128 B of payload per row, 16-B chunks XOR-swizzled
constexpr int STAGING_BUF_BYTES = 32 * STAGING_ROW_BYTES;      // 4096 B: one 32-row box (the TMA 128B-swizzle atom x4)
constexpr int STAGING_WARP_BYTES = 2 * STAGING_BUF_BYTES;      // double-buffered per epilogue warp

template <int STAGE_BYTES_, int STAGES>
struct SmemPlan {
  static constexpr int STAGE_BYTES = STAGE_BYTES_;
  static constexpr int STAGING_OFFSET = STAGES * STAGE_BYTES;  // multiple of 1024
  static constexpr int BAR_OFFSET = STAGING_OFFSET + EPI_WARPS * STAGING_WARP_BYTES;
  // full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], aux[EPI_WARPS], tmem slot
  static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 4 + EPI_WARPS) * 8 + 16;
  static constexpr int DYN_BYTES = TOTAL + 1024;  // slack for manual 1024 B alignment
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
struct EpiState {
  uint8_t* stage;       // this warp's two staging buffers
  uint64_t* aux_bar;    // this warp's mbarrier for auxiliary-operand TMA loads
  uint32_t buf;         // staging buffer to use next
  uint32_t aux_phase;
};

// ---- epilogue math on the 64 (bf16 modes) / 32 (fp32 modes) accumulator columns a thread holds for its row ----
__device__ __forceinline__ void epi_bias(const GemmParams& p, int epi, int CH, float (&v)[64], int col) {
  if (p.bias == nullptr || !(epi == EPI_BIAS_BF16 || epi == EPI_BIAS_GELU_BF16 || epi == EPI_BIAS_F32)) return;
  const float* b = p.bias + col;
  if (col + CH <= p.N && (reinterpret_cast<uintptr_t>(b) & 15) == 0) {
#pragma unroll
    for (int e = 0; e < 64; e += 4)
      if (e < CH) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(b + e));
        v[e] += t.x; v[e + 1] += t.y; v[e + 2] += t.z; v[e + 3] += t.w;
      }
  } else {
#pragma unroll
    for (int e = 0; e < 64; ++e)
      if (e < CH && col + e < p.N) v[e] += __ldg(b + e);
  }
}
__device__ __forceinline__ void epi_apply_aux(int epi, float (&v)[64], const uint4 (&a)[8]) {
#pragma unroll
  for (int e8 = 0; e8 < 8; ++e8) {
    const uint32_t ww[4] = {a[e8].x, a[e8].y, a[e8].z, a[e8].w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16x2(ww[j]);
      if (epi == EPI_GELU_BWD_BF16) {
        v[e8 * 8 + 2 * j] *= gelu_erf_grad(f.x);
        v[e8 * 8 + 2 * j + 1] *= gelu_erf_grad(f.y);
      } else {
        v[e8 * 8 + 2 * j] += f.x;
        v[e8 * 8 + 2 * j + 1] += f.y;
      }
    }
  }
}
// registers -> this lane's row of a swizzled staging buffer (16-byte chunk j of row r at position j ^ (r & 7), which is
// both bank-conflict free and exactly the TMA 128B swizzle of a 1024-byte aligned box)
__device__ __forceinline__ void stage_write(uint8_t* buf, int lane, bool out_f32, const float (&v)[64]) {
  uint8_t* my = buf + lane * STAGING_ROW_BYTES;
  const int sw = lane & 7;
  if (out_f32) {
#pragma unroll
    for (int e = 0; e < 32; e += 4)
      *reinterpret_cast<float4*>(my + (((e >> 2) ^ sw) << 4)) = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
  } else {
#pragma unroll
    for (int e = 0; e < 64; e += 8) {
      uint4 u;
      u.x = pack_bf16x2(v[e], v[e + 1]);     u.y = pack_bf16x2(v[e + 2], v[e + 3]);
      u.z = pack_bf16x2(v[e + 4], v[e + 5]); u.w = pack_bf16x2(v[e + 6], v[e + 7]);
      *reinterpret_cast<uint4*>(my + (((e >> 3) ^ sw) << 4)) = u;
    }
  }
}

// TMA epilogue for one chunk.  `v` holds alpha * accumulator.
__device__ __forceinline__ void epilogue_chunk_tma(const GemmParams& p, int epi, bool out_f32, int CH, float (&v)[64],
                                                   int row0, int col, int lane, EpiState& st,
                                                   const CUtensorMap* tm_out, const CUtensorMap* tm_aux) {
  epi_bias(p, epi, CH, v, col);
  const int passes = (epi == EPI_BIAS_GELU_BF16) ? 2 : 1;
  for (int pass = 0; pass < passes; ++pass) {
    uint8_t* buf = st.stage + st.buf * STAGING_BUF_BYTES;
    st.buf ^= 1;
    if (lane == 0) bulk_wait_read<1>();  // the store that last used this buffer (two groups ago) has drained it
    __syncwarp();
    if (pass == 0 && (epi == EPI_GELU_BWD_BF16 || epi == EPI_ADD_BF16)) {
      if (lane == 0) {
        mbar_arrive_expect_tx(st.aux_bar, STAGING_BUF_BYTES);
        tma_load_2d(buf, tm_aux, st.aux_bar, col, row0);  // rows / columns out of range arrive as zeros
      }
      mbar_wait(st.aux_bar, st.aux_phase);
      st.aux_phase ^= 1;
      uint4 a[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        a[j] = *reinterpret_cast<const uint4*>(buf + lane * STAGING_ROW_BYTES + ((j ^ (lane & 7)) << 4));
      __syncwarp();  // every lane has its row before the buffer is overwritten with results
      epi_apply_aux(epi, v, a);
    }
    if (passes == 2 && pass == 1) {
#pragma unroll
      for (int e = 0; e < 64; ++e) v[e] = gelu_erf(v[e]);
    }
    stage_write(buf, lane, out_f32, v);
    fence_proxy_async_smem();  // generic-proxy writes visible to the bulk-copy engine
    __syncwarp();
    if (lane == 0) {
      const CUtensorMap* tm = (passes == 2 && pass == 0) ? tm_aux : tm_out;
      if (epi == EPI_ATOMIC_F32) tma_reduce_add_2d(tm, buf, col, row0);
      else tma_store_2d(tm, buf, col, row0);
      bulk_commit();
    }
  }
}

// Fallback epilogue (leading dimension not 16-byte aligned): coalesced stores through ONE staging buffer.
__device__ __forceinline__ void epilogue_chunk_manual(const GemmParams& p, const int epi, const bool out_f32,
                                                      const int CH, float (&v)[64], const int row0, const int row,
                                                      const int col, const int lane, uint8_t* stage_w) {
  const int rl = lane >> 3, cl = lane & 7;
  epi_bias(p, epi, CH, v, col);
  if (epi == EPI_GELU_BWD_BF16 || epi == EPI_ADD_BF16) {
    if (row < p.M) {
      const bf16* a = p.aux_in + (long long)row * p.ld_aux_in + col;
#pragma unroll
      for (int e = 0; e < 64; ++e) {
        const float x = (col + e < p.N) ? __bfloat162float(a[e]) : 0.f;
        if (epi == EPI_GELU_BWD_BF16) v[e] *= gelu_erf_grad(x);
        else v[e] += x;
      }
    }
  }
  const int passes = (epi == EPI_BIAS_GELU_BF16) ? 2 : 1;
  for (int pass = 0; pass < passes; ++pass) {
    if (passes == 2 && pass == 1) {
#pragma unroll
      for (int e = 0; e < 64; ++e) v[e] = gelu_erf(v[e]);
    }
    stage_write(stage_w, lane, out_f32, v);
    __syncwarp();
    const int ecol = col + cl * (out_f32 ? 4 : 8);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = i * 4 + rl;
      const int grow = row0 + r;
      if (grow >= p.M || ecol >= p.N) continue;
      const uint4 u = *reinterpret_cast<const uint4*>(stage_w + r * STAGING_ROW_BYTES + ((cl ^ (r & 7)) << 4));
      if (out_f32) {
        float* o = reinterpret_cast<float*>(p.out) + (long long)grow * p.ldo + ecol;
        const float f[4] = {__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w)};
        const bool vec = (ecol + 4 <= p.N) && ((reinterpret_cast<uintptr_t>(o) & 15) == 0);
        if (epi == EPI_ATOMIC_F32) {
          if (vec) red_add_v4(o, f[0], f[1], f[2], f[3]);
          else
            for (int j = 0; j < 4; ++j)
              if (ecol + j < p.N) atomicAdd(o + j, f[j]);
        } else {
          if (vec) *reinterpret_cast<float4*>(o) = make_float4(f[0], f[1], f[2], f[3]);
          else
            for (int j = 0; j < 4; ++j)
              if (ecol + j < p.N) o[j] = f[j];
        }
      } else {
        bf16* base = (passes == 2 && pass == 0) ? p.aux_out : reinterpret_cast<bf16*>(p.out);
        const long long ld = (passes == 2 && pass == 0) ? p.ld_aux_out : p.ldo;
        bf16* o = base + (long long)grow * ld + ecol;
        if ((ecol + 8 <= p.N) && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
          *reinterpret_cast<uint4*>(o) = u;
        } else {
          const bf16* sv = reinterpret_cast<const bf16*>(&u);
          for (int j = 0; j < 8; ++j)
            if (ecol + j < p.N) o[j] = sv[j];
        }
      }
    }
    __syncwarp();
  }
}

// One accumulator tile (this warp's 32 rows x BLOCK_N columns, every other chunk) -> global memory.
template <int BLOCK_N>
__device__ __forceinline__ void epilogue_tile(const GemmParams& p, uint32_t t_acc, int row0, int n0, int lane, int half,
                                              EpiState& st, const CUtensorMap* tm_out, const CUtensorMap* tm_aux) {
  const int epi = p.epilogue;
  const float alpha = p.alpha;
  const bool out_f32 = (epi == EPI_BIAS_F32 || epi == EPI_ATOMIC_F32);
  const int CH = out_f32 ? 32 : 64;  // columns per chunk: 128 bytes of output per row either way
  for (int c = half * CH; c < BLOCK_N; c += 2 * CH) {
    const int col = n0 + c;
    if (col >= p.N) break;  // warp-uniform
    float v[64];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j * 16 < CH) {
        uint32_t r[16];
        tmem_ld_32x32b_x16(t_acc + (uint32_t)(c + j * 16), r);
#pragma unroll
        for (int e = 0; e < 16; ++e) v[j * 16 + e] = __uint_as_float(r[e]);
      }
    }
    tmem_ld_wait();
    if (alpha != 1.0f) {  // warp-uniform; the common case (alpha = 1) spends no issue slot on it
#pragma unroll
      for (int e = 0; e < 64; ++e) v[e] *= alpha;
    }
    if (p.tma_epilogue) epilogue_chunk_tma(p, epi, out_f32, CH, v, row0, col, lane, st, tm_out, tm_aux);
    else epilogue_chunk_manual(p, epi, out_f32, CH, v, row0, row0 + lane, col, lane, st.stage);
  }
}

struct WorkItem {
  int m0, n0, kb_begin, num_kb;
};
// tile_m = rows covered by one work item (128, or 256 for a CTA pair).
// Rasterisation: without split-K the n-tile index runs fastest, so the CTAs in flight at any moment cover a band of a
// few m-tiles across ALL n-tiles — every A tile (activations, streamed from HBM) is fetched from DRAM once and reused
// from L2 by the other n-tiles, while the whole B operand (weights) stays L2-resident.  (m-fastest order re-streamed
// A once per n-tile: measured 1.35 GB of DRAM reads for the 151 MB activation matrix of the QKV projection.)
// With split-K (weight gradients) the tiles of one split run together so they share that split's token slab.
__device__ __forceinline__ WorkItem decode_work(int w, int m_tiles, int n_tiles, int total_kb, int kb_per, int tile_m,
                                                int bn) {
  const int tiles = m_tiles * n_tiles;
  const int split = w / tiles;
  const int rem = w - split * tiles;
  int m_blk, n_blk;
  if (kb_per >= total_kb) {
    m_blk = rem / n_tiles;
    n_blk = rem - m_blk * n_tiles;
  } else {
    n_blk = rem / m_tiles;
    m_blk = rem - n_blk * m_tiles;
  }
  WorkItem it;
  it.m0 = m_blk * tile_m;
  it.n0 = n_blk * bn;
  it.kb_begin = split * kb_per;
  it.num_kb = min(total_kb, it.kb_begin + kb_per) - it.kb_begin;
  return it;
}

// ================================================================================================================
// 1-CTA persistent kernel
// ================================================================================================================
template <int BLOCK_N, int STAGES>
using GemmSmemP = SmemPlan<BLOCK_M * BLOCK_K * 2 + BLOCK_N * BLOCK_K * 2, STAGES>;

template <int BLOCK_N, int STAGES, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(GEMM_P_THREADS, 1)
gemm_tcgen05_persistent_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                               const __grid_constant__ CUtensorMap tmap_out,
                               const __grid_constant__ CUtensorMap tmap_aux, const GemmParams p, const int num_work) {
  using L = GemmSmemP<BLOCK_N, STAGES>;
  constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2]
  uint64_t* aux_bar = tmem_empty_bar + 2;         // [EPI_WARPS]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux_bar + EPI_WARPS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = (p.M + BLOCK_M - 1) / BLOCK_M;
  const int n_tiles = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int total_kb = (p.Kc + BLOCK_K - 1) / BLOCK_K;
  const int kb_per = p.k_blocks_per_split;
  constexpr uint32_t TMEM_COLS = 2 * BLOCK_N;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if (p.tma_epilogue) {
      tma_prefetch_desc(&tmap_out);
      tma_prefetch_desc(&tmap_aux);
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full_bar[b], 1);
      mbar_init(&tmem_empty_bar[b], EPI_WARPS);  // one arrival per epilogue warp
    }
    for (int e = 0; e < EPI_WARPS; ++e) mbar_init(&aux_bar[e], 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // everything above overlapped the previous kernel's tail; global memory is touched only from here on

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      uint32_t it = 0;
      for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
        const WorkItem wi = decode_work(w, m_tiles, n_tiles, total_kb, kb_per, BLOCK_M, BLOCK_N);
        for (int i = 0; i < wi.num_kb; ++i, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = smem + s * L::STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          mbar_arrive_expect_tx(&full_bar[s], L::STAGE_BYTES);
          const int k_elem = (wi.kb_begin + i) * BLOCK_K;
          if (!A_MN) {
            tma_load_2d(sa, &tmap_a, &full_bar[s], k_elem, wi.m0);
          } else {
#pragma unroll
            for (int c = 0; c < BLOCK_M / 64; ++c)
              tma_load_2d(sa + c * (BLOCK_K * 128), &tmap_a, &full_bar[s], wi.m0 + c * 64, k_elem);
          }
          if (!B_MN) {
            tma_load_2d(sb, &tmap_b, &full_bar[s], k_elem, wi.n0);
          } else {
#pragma unroll
            for (int c = 0; c < BLOCK_N / 64; ++c)
              tma_load_2d(sb + c * (BLOCK_K * 128), &tmap_b, &full_bar[s], wi.n0 + c * 64, k_elem);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer --------------------------------
    // All 32 lanes run the warp-uniform loop (addresses and descriptors stay in uniform registers); one elected lane
    // issues the MMAs and the commits.  (An `if (lane == 0)` around the loop costs an ELECT + R2UR re-broadcast per
    // operand of every tcgen05 instruction.)
    constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M, BLOCK_N, A_MN, B_MN);
    uint32_t it = 0, t = 0;
    for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++t) {
      const WorkItem wi = decode_work(w, m_tiles, n_tiles, total_kb, kb_per, BLOCK_M, BLOCK_N);
      const uint32_t acc = t & 1, acc_ph = (t >> 1) & 1;
      mbar_wait(&tmem_empty_bar[acc], acc_ph ^ 1);  // epilogue has drained this accumulator
      tc_fence_after_sync();
      const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
      for (int i = 0; i < wi.num_kb; ++i, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after_sync();
        const uint32_t sa = smem_u32(smem + s * L::STAGE_BYTES);
        const uint32_t sb = sa + A_BYTES;
        // K = 16 step k: K-major operands advance 32 bytes (+2 in the descriptor's 16-byte address field), MN-major
        // operands 16 rows of 128 bytes (+128)
        const uint64_t da0 = A_MN ? make_smem_desc_sw128(sa, BLOCK_K * 128, 1024) : make_smem_desc_sw128(sa, 16, 1024);
        const uint64_t db0 = B_MN ? make_smem_desc_sw128(sb, BLOCK_K * 128, 1024) : make_smem_desc_sw128(sb, 16, 1024);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
            umma_bf16(d_tmem, da0 + (A_MN ? 128 : 2) * k, db0 + (B_MN ? 128 : 2) * k, idesc, (i > 0 || k > 0) ? 1u : 0u);
          umma_commit(&empty_bar[s]);
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(&tmem_full_bar[acc]);
      __syncwarp();
    }
  } else {
    // ------------------------------ epilogue ----------------------------------
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;  // which of the two warps of that quarter: even / odd column chunks
    EpiState st{smem + L::STAGING_OFFSET + (warp - 2) * STAGING_WARP_BYTES, &aux_bar[warp - 2], 0u, 0u};
    uint32_t t = 0;
    for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++t) {
      const WorkItem wi = decode_work(w, m_tiles, n_tiles, total_kb, kb_per, BLOCK_M, BLOCK_N);
      const uint32_t acc = t & 1, acc_ph = (t >> 1) & 1;
      mbar_wait(&tmem_full_bar[acc], acc_ph);
      tc_fence_after_sync();
      epilogue_tile<BLOCK_N>(p, tmem_base + acc * BLOCK_N + ((uint32_t)(q * 32) << 16), wi.m0 + q * 32, wi.n0, lane,
                             half, st, &tmap_out, &tmap_aux);
      // all TMEM reads of this accumulator are complete: hand it back to the MMA warp
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
    }
    if (lane == 0) bulk_wait_read<0>();  // outstanding bulk stores still read this CTA's shared memory (the writes
                                         // themselves complete asynchronously; kernel completion orders them)
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ================================================================================================================
// CTA-pair variant: tcgen05.mma.cta_group::2, one 256 x BLOCK_N output tile per pair of SMs.
//
// On the K = 768 shapes of this model the 1-CTA kernel re-reads a full 256-row B tile per 128 x 256 output tile.  With
// cta_group::2 the two CTAs of a cluster each load their own 128 rows of A and only HALF of the B tile; the tensor
// core reads both halves across the pair, so operand bytes per FLOP drop by a third.  Layout of one UMMA (M = 256,
// N = BLOCK_N, K = 16): accumulator rows 0..127 live in CTA 0's TMEM, rows 128..255 in CTA 1's; CTA r supplies A rows
// [128 r, 128 r + 128) and B rows (N index) [BLOCK_N/2 r, +BLOCK_N/2).  Only the leader CTA (cluster rank 0) issues
// MMAs; its "full" barrier collects the TMA bytes of BOTH CTAs (the peer's copies signal the leader's barrier through
// the shared::cluster window), tcgen05.commit multicasts the "slot free" / "accumulator ready" arrivals to both CTAs,
// and both CTAs' epilogue warps arrive on the leader's "accumulator drained" barrier.
// ================================================================================================================
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address -> leader CTA

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* m, uint64_t* leader_bar, int c0,
                                                int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1),
        "r"(smem_u32(leader_bar) & PEER_BIT_MASK)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all prior MMAs retire) on the barrier at this smem offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .b16 m;\n\tmov.b16 m, 3;\n\t"
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}"
      ::"r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & PEER_BIT_MASK) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

template <int BLOCK_N, int STAGES>
using GemmSmem2 = SmemPlan<BLOCK_M * BLOCK_K * 2 + (BLOCK_N / 2) * BLOCK_K * 2, STAGES>;

template <int BLOCK_N, int STAGES, bool A_MN, bool B_MN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_P_THREADS, 1)
gemm_tcgen05_2cta_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_aux,
                         const GemmParams p, const int num_work) {
  using L = GemmSmem2<BLOCK_N, STAGES>;
  constexpr int PAIR_M = 2 * BLOCK_M;
  constexpr int HALF_N = BLOCK_N / 2;
  constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2] (the leader's copy is the live one)
  uint64_t* aux_bar = tmem_empty_bar + 2;         // [EPI_WARPS]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux_bar + EPI_WARPS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int m_tiles = (p.M + PAIR_M - 1) / PAIR_M;
  const int n_tiles = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int total_kb = (p.Kc + BLOCK_K - 1) / BLOCK_K;
  const int kb_per = p.k_blocks_per_split;
  constexpr uint32_t TMEM_COLS = 2 * BLOCK_N;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if (p.tma_epilogue) {
      tma_prefetch_desc(&tmap_out);
      tma_prefetch_desc(&tmap_aux);
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full_bar[b], 1);
      mbar_init(&tmem_empty_bar[b], 2 * EPI_WARPS);  // epilogue warps of both CTAs
    }
    for (int e = 0; e < EPI_WARPS; ++e) mbar_init(&aux_bar[e], 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2sm(tmem_slot, TMEM_COLS);
  tc_fence_before_sync();
  cluster_sync_all();  // barrier inits of both CTAs visible cluster-wide before any remote arrive / TMA signal
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    // ------------------------------ TMA producer (both CTAs) ------------------------------
    if (lane == 0) {
      uint32_t it = 0;
      for (int w = pair; w < num_work; w += num_pairs) {
        const WorkItem wi = decode_work(w, m_tiles, n_tiles, total_kb, kb_per, PAIR_M, BLOCK_N);
        const int m0 = wi.m0 + (int)rank * BLOCK_M;   // this CTA's rows of A
        const int nb0 = wi.n0 + (int)rank * HALF_N;   // this CTA's half of the B tile
        for (int i = 0; i < wi.num_kb; ++i, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = smem + s * L::STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          if (leader) mbar_arrive_expect_tx(&full_bar[s], 2 * L::STAGE_BYTES);  // bytes of both CTAs land here
          const int k_elem = (wi.kb_begin + i) * BLOCK_K;
          if (!A_MN) {
            tma_load_2d_2sm(sa, &tmap_a, &full_bar[s], k_elem, m0);
          } else {
#pragma unroll
            for (int c = 0; c < BLOCK_M / 64; ++c)
              tma_load_2d_2sm(sa + c * (BLOCK_K * 128), &tmap_a, &full_bar[s], m0 + c * 64, k_elem);
          }
          if (!B_MN) {
            tma_load_2d_2sm(sb, &tmap_b, &full_bar[s], k_elem, nb0);
          } else {
#pragma unroll
            for (int c = 0; c < HALF_N / 64; ++c)
              tma_load_2d_2sm(sb + c * (BLOCK_K * 128), &tmap_b, &full_bar[s], nb0 + c * 64, k_elem);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (leader CTA only) ------------------------------
    if (leader) {   // warp-uniform loop, one elected lane issues (see the 1-CTA kernel)
      constexpr uint32_t idesc = make_idesc_bf16(PAIR_M, BLOCK_N, A_MN, B_MN);
      uint32_t it = 0, t = 0;
      for (int w = pair; w < num_work; w += num_pairs, ++t) {
        const WorkItem wi = decode_work(w, m_tiles, n_tiles, total_kb, kb_per, PAIR_M, BLOCK_N);
        const uint32_t acc = t & 1, acc_ph = (t >> 1) & 1;
        mbar_wait(&tmem_empty_bar[acc], acc_ph ^ 1);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int i = 0; i < wi.num_kb; ++i, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after_sync();
          const uint32_t sa = smem_u32(smem + s * L::STAGE_BYTES);
          const uint32_t sb = sa + A_BYTES;
          const uint64_t da0 = A_MN ? make_smem_desc_sw128(sa, BLOCK_K * 128, 1024) : make_smem_desc_sw128(sa, 16, 1024);
          const uint64_t db0 = B_MN ? make_smem_desc_sw128(sb, BLOCK_K * 128, 1024) : make_smem_desc_sw128(sb, 16, 1024);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
              umma_bf16_2sm(d_tmem, da0 + (A_MN ? 128 : 2) * k, db0 + (B_MN ? 128 : 2) * k, idesc,
                            (i > 0 || k > 0) ? 1u : 0u);
            umma_commit_2sm(&empty_bar[s]);
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit_2sm(&tmem_full_bar[acc]);
        __syncwarp();
      }
    }
  } else {
    // ------------------------------ epilogue (both CTAs, own TMEM rows) ------------------------------
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    EpiState st{smem + L::STAGING_OFFSET + (warp - 2) * STAGING_WARP_BYTES, &aux_bar[warp - 2], 0u, 0u};
    uint32_t t = 0;
    for (int w = pair; w < num_work; w += num_pairs, ++t) {
      const WorkItem wi = decode_work(w, m_tiles, n_tiles, total_kb, kb_per, PAIR_M, BLOCK_N);
      const uint32_t acc = t & 1, acc_ph = (t >> 1) & 1;
      mbar_wait(&tmem_full_bar[acc], acc_ph);
      tc_fence_after_sync();
      epilogue_tile<BLOCK_N>(p, tmem_base + acc * BLOCK_N + ((uint32_t)(q * 32) << 16),
                             wi.m0 + (int)rank * BLOCK_M + q * 32, wi.n0, lane, half, st, &tmap_out, &tmap_aux);
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tmem_empty_bar[acc]);
    }
    if (lane == 0) bulk_wait_read<0>();
  }

  __syncwarp();
  tc_fence_before_sync();
  cluster_sync_all();  // neither CTA may free TMEM / exit while the peer still reads or signals
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc_2sm(tmem_base, TMEM_COLS);
  }
}

}  // namespace univl
