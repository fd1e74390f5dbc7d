// univl_b200 — embedding front-ends fused with their LayerNorm (+ dropout).  HBM-bound, one warp per token row.
//
//   text    : y = dropout(LN(word[id] + pos[s] + type[t]))       reference modules/module_bert.py:132-146,
//             (type table optional: the caption decoder has none)           modules/module_decoder.py:309-320
//   sources : y = dropout(LN(src(row) + pos[s] (+ type[s >= Wa])))
//             visual embeddings (src = Linear(1024->768) output)  reference modules/module_visual.py:118-131
//             cross  embeddings (src = concat(text_i, video_j))   reference modules/module_cross.py:123-138 with
//             modules/modeling.py:315-325; in all-pairs mode sequence p = (i, j) = (p / Nb, p % Nb) reads text i and
//             video j in place — the `repeat`ed [B*B, W+F, H] input of modeling.py:358-367 is never materialised.
// The tables are the fp32 master parameters (no bf16 copy is needed for a gather).  Backward recomputes the pre-LN
// row, runs the LayerNorm backward and scatters: atomics into the fp32 table gradients, direct bf16 writes (summed in
// registers over the pairs that share a source row) into the activation gradients.
#include "common.cuh"

namespace univl {

constexpr int EMB_WARPS = 8;
constexpr int EMB_H = 768;
constexpr int EMB_VEC = EMB_H / 256;  // 3 vectors of 8 per lane

struct EmbDrop {
  uint32_t threshold;
  float scale;
  uint64_t seed, stream;
  int on;
  const unsigned long long* rng;  // device {seed, epoch}, resolved at kernel entry (graph-replayable)
};
__device__ __forceinline__ EmbDrop resolve_emb_drop(EmbDrop d) {
  if (d.on && d.rng != nullptr) {
    d.seed = d.rng[0];
    d.stream += d.rng[1] << 20;
  }
  return d;
}

__device__ __forceinline__ void ld8f(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void ld8h(const bf16* p, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = unpack_bf16x2(w[j]);
    v[2 * j] = f.x;
    v[2 * j + 1] = f.y;
  }
}
__device__ __forceinline__ void st8h(bf16* p, const float (&v)[8]) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
  u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void emb_keep8(const EmbDrop& d, uint64_t idx0, bool (&k)[8]) {
  const uint32_t m = dropout_keep8(d.seed, d.stream, idx0, d.threshold);
#pragma unroll
  for (int j = 0; j < 8; ++j) k[j] = (m >> j) & 1u;
}

// z (registers) -> mean/rstd -> y ; shared by both forward kernels
__device__ __forceinline__ void ln_row_fwd(float (&z)[EMB_VEC][8], const float* gamma, const float* beta, bf16* yrow,
                                           float* mean_out, float* rstd_out, long long row, float eps,
                                           const EmbDrop& drop, int lane) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < EMB_VEC; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) s += z[i][j];
  const float mean = warp_sum(s) * (1.0f / EMB_H);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < EMB_VEC; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float d = z[i][j] - mean;
      q += d * d;
    }
  const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / EMB_H) + eps);
  if (lane == 0) {
    mean_out[row] = mean;
    rstd_out[row] = rstd;
  }
#pragma unroll
  for (int i = 0; i < EMB_VEC; ++i) {
    const int c = (i * 32 + lane) * 8;
    float g[8], b[8], o[8];
    ld8f(gamma + c, g);
    ld8f(beta + c, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = g[j] * ((z[i][j] - mean) * rstd) + b[j];
    if (drop.on) {
      bool k[8];
      emb_keep8(drop, (uint64_t)row * EMB_H + c, k);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = k[j] ? o[j] * drop.scale : 0.f;
    }
    st8h(yrow + c, o);
  }
}

// dy (bf16 row, after-dropout gradient) -> dz in registers; accumulates dgamma/dbeta partials
__device__ __forceinline__ void ln_row_bwd(const float (&z)[EMB_VEC][8], const bf16* dyrow, const float* gamma,
                                           float mean, float rstd, long long row, const EmbDrop& drop, int lane,
                                           float (&dz)[EMB_VEC][8], float (&acc_g)[EMB_VEC][8],
                                           float (&acc_b)[EMB_VEC][8]) {
  float xh[EMB_VEC][8], g[EMB_VEC][8];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < EMB_VEC; ++i) {
    const int c = (i * 32 + lane) * 8;
    float d[8], gm[8];
    ld8h(dyrow + c, d);
    if (drop.on) {
      bool k[8];
      emb_keep8(drop, (uint64_t)row * EMB_H + c, k);
#pragma unroll
      for (int j = 0; j < 8; ++j) d[j] = k[j] ? d[j] * drop.scale : 0.f;
    }
    ld8f(gamma + c, gm);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      xh[i][j] = (z[i][j] - mean) * rstd;
      g[i][j] = d[j] * gm[j];
      s1 += g[i][j];
      s2 += g[i][j] * xh[i][j];
      acc_g[i][j] += d[j] * xh[i][j];
      acc_b[i][j] += d[j];
    }
  }
  s1 = warp_sum(s1) * (1.0f / EMB_H);
  s2 = warp_sum(s2) * (1.0f / EMB_H);
#pragma unroll
  for (int i = 0; i < EMB_VEC; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) dz[i][j] = rstd * (g[i][j] - s1 - xh[i][j] * s2);
}

__device__ __forceinline__ void flush_colsums(float (&acc)[EMB_VEC][8], float* dst, float (*red)[257], int warp,
                                              int lane) {
  for (int i = 0; i < EMB_VEC; ++i) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) red[warp][lane * 8 + j] = acc[i][j];
    __syncthreads();
    for (int e = threadIdx.x; e < 256; e += EMB_WARPS * 32) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < EMB_WARPS; ++w) t += red[w][e];
      atomicAdd(dst + i * 256 + e, t);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// text embeddings
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(EMB_WARPS * 32)
embed_text_fwd_kernel(const long long* __restrict__ ids, const long long* __restrict__ type_ids,
                      const float* __restrict__ word, const float* __restrict__ pos, const float* __restrict__ type,
                      const float* __restrict__ gamma, const float* __restrict__ beta, bf16* __restrict__ y,
                      float* __restrict__ mean_out, float* __restrict__ rstd_out, int n_seq, int S, int vocab,
                      float eps, EmbDrop drop_in) {
  const EmbDrop drop = resolve_emb_drop(drop_in);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long rows = (long long)n_seq * S;
  for (long long row = (long long)blockIdx.x * EMB_WARPS + warp; row < rows; row += (long long)gridDim.x * EMB_WARPS) {
    const int s = (int)(row % S);
    long long id = ids[row];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    const long long t = (type != nullptr && type_ids != nullptr) ? (type_ids[row] != 0 ? 1 : 0) : 0;
    float z[EMB_VEC][8];
#pragma unroll
    for (int i = 0; i < EMB_VEC; ++i) {
      const int c = (i * 32 + lane) * 8;
      float a[8], b[8];
      ld8f(word + id * EMB_H + c, a);
      ld8f(pos + (long long)s * EMB_H + c, b);
#pragma unroll
      for (int j = 0; j < 8; ++j) z[i][j] = a[j] + b[j];
      if (type != nullptr) {
        float tt[8];
        ld8f(type + t * EMB_H + c, tt);
#pragma unroll
        for (int j = 0; j < 8; ++j) z[i][j] += tt[j];
      }
    }
    ln_row_fwd(z, gamma, beta, y + row * EMB_H, mean_out, rstd_out, row, eps, drop, lane);
  }
}

__global__ void __launch_bounds__(EMB_WARPS * 32)
embed_text_bwd_kernel(const bf16* __restrict__ dy, const long long* __restrict__ ids,
                      const long long* __restrict__ type_ids, const float* __restrict__ word,
                      const float* __restrict__ pos, const float* __restrict__ type, const float* __restrict__ gamma,
                      const float* __restrict__ mean_in, const float* __restrict__ rstd_in, float* __restrict__ dword,
                      float* __restrict__ dpos, float* __restrict__ dtype, float* __restrict__ dgamma,
                      float* __restrict__ dbeta, int n_seq, int S, int vocab, EmbDrop drop_in) {
  const EmbDrop drop = resolve_emb_drop(drop_in);
  __shared__ float red[EMB_WARPS][257];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long rows = (long long)n_seq * S;
  float acc_g[EMB_VEC][8], acc_b[EMB_VEC][8];
#pragma unroll
  for (int i = 0; i < EMB_VEC; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc_g[i][j] = acc_b[i][j] = 0.f;
  for (long long row = (long long)blockIdx.x * EMB_WARPS + warp; row < rows; row += (long long)gridDim.x * EMB_WARPS) {
    const int s = (int)(row % S);
    long long id = ids[row];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    const long long t = (type != nullptr && type_ids != nullptr) ? (type_ids[row] != 0 ? 1 : 0) : 0;
    float z[EMB_VEC][8], dz[EMB_VEC][8];
#pragma unroll
    for (int i = 0; i < EMB_VEC; ++i) {
      const int c = (i * 32 + lane) * 8;
      float a[8], b[8];
      ld8f(word + id * EMB_H + c, a);
      ld8f(pos + (long long)s * EMB_H + c, b);
#pragma unroll
      for (int j = 0; j < 8; ++j) z[i][j] = a[j] + b[j];
      if (type != nullptr) {
        float tt[8];
        ld8f(type + t * EMB_H + c, tt);
#pragma unroll
        for (int j = 0; j < 8; ++j) z[i][j] += tt[j];
      }
    }
    ln_row_bwd(z, dy + row * EMB_H, gamma, mean_in[row], rstd_in[row], row, drop, lane, dz, acc_g, acc_b);
#pragma unroll
    for (int i = 0; i < EMB_VEC; ++i) {
      const int c = (i * 32 + lane) * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        atomicAdd(dword + id * EMB_H + c + j, dz[i][j]);
        atomicAdd(dpos + (long long)s * EMB_H + c + j, dz[i][j]);
        if (dtype != nullptr) atomicAdd(dtype + t * EMB_H + c + j, dz[i][j]);
      }
    }
  }
  flush_colsums(acc_g, dgamma, red, warp, lane);
  flush_colsums(acc_b, dbeta, red, warp, lane);
}

// ------------------------------------------------------------------------------------------------------------
// activation-source embeddings (visual / cross)
// ------------------------------------------------------------------------------------------------------------
struct SrcCfg {
  const bf16* a;  // [Na, Wa, H]
  const bf16* b;  // [Nb, Fb, H] or null (Fb = 0)
  int Na, Wa, Nb, Fb;
  int all_pairs;  // 0: sequence p reads (a[p], b[p]); 1: p = i * Nb + j reads (a[i], b[j])
};

__device__ __forceinline__ const bf16* src_row(const SrcCfg& c, long long p, int s) {
  const long long i = c.all_pairs ? p / c.Nb : p;
  const long long j = c.all_pairs ? p % c.Nb : p;
  return s < c.Wa ? c.a + (i * c.Wa + s) * EMB_H : c.b + (j * c.Fb + (s - c.Wa)) * EMB_H;
}

__global__ void __launch_bounds__(EMB_WARPS * 32)
embed_src_fwd_kernel(SrcCfg src, const float* __restrict__ pos, const float* __restrict__ type,
                     const float* __restrict__ gamma, const float* __restrict__ beta, bf16* __restrict__ y,
                     float* __restrict__ mean_out, float* __restrict__ rstd_out, long long n_seq, float eps,
                     EmbDrop drop_in) {
  const EmbDrop drop = resolve_emb_drop(drop_in);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = src.Wa + src.Fb;
  const long long rows = n_seq * S;
  for (long long row = (long long)blockIdx.x * EMB_WARPS + warp; row < rows; row += (long long)gridDim.x * EMB_WARPS) {
    const long long p = row / S;
    const int s = (int)(row % S);
    const bf16* xr = src_row(src, p, s);
    float z[EMB_VEC][8];
#pragma unroll
    for (int i = 0; i < EMB_VEC; ++i) {
      const int c = (i * 32 + lane) * 8;
      float a[8], b[8];
      ld8h(xr + c, a);
      ld8f(pos + (long long)s * EMB_H + c, b);
#pragma unroll
      for (int j = 0; j < 8; ++j) z[i][j] = a[j] + b[j];
      if (type != nullptr) {
        float tt[8];
        ld8f(type + (s < src.Wa ? 0 : EMB_H) + c, tt);
#pragma unroll
        for (int j = 0; j < 8; ++j) z[i][j] += tt[j];
      }
    }
    ln_row_fwd(z, gamma, beta, y + row * EMB_H, mean_out, rstd_out, row, eps, drop, lane);
  }
}

// One warp per SOURCE row: it walks every output sequence that read this row (1 in aligned mode, Nb or Na in
// all-pairs mode), sums dz in registers and writes the source gradient once — deterministic, no atomics on
// activations.  blockIdx.y selects the source (0 = a, 1 = b).
__global__ void __launch_bounds__(EMB_WARPS * 32)
embed_src_bwd_kernel(const bf16* __restrict__ dy, SrcCfg src, const float* __restrict__ pos,
                     const float* __restrict__ type, const float* __restrict__ gamma,
                     const float* __restrict__ mean_in, const float* __restrict__ rstd_in, bf16* __restrict__ da,
                     bf16* __restrict__ db, float* __restrict__ dpos, float* __restrict__ dtype,
                     float* __restrict__ dgamma, float* __restrict__ dbeta, EmbDrop drop_in) {
  const EmbDrop drop = resolve_emb_drop(drop_in);
  __shared__ float red[EMB_WARPS][257];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = src.Wa + src.Fb;
  const int which = blockIdx.y;  // 0: rows of a, 1: rows of b
  const long long n_src_rows = which == 0 ? (long long)src.Na * src.Wa : (long long)src.Nb * src.Fb;
  float acc_g[EMB_VEC][8], acc_b[EMB_VEC][8];
#pragma unroll
  for (int i = 0; i < EMB_VEC; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc_g[i][j] = acc_b[i][j] = 0.f;
  for (long long sr = (long long)blockIdx.x * EMB_WARPS + warp; sr < n_src_rows;
       sr += (long long)gridDim.x * EMB_WARPS) {
    const int len = which == 0 ? src.Wa : src.Fb;
    const long long owner = sr / len;           // i (text) or j (video)
    const int s = (int)(sr % len) + (which == 0 ? 0 : src.Wa);
    const bf16* xr = (which == 0 ? src.a : src.b) + sr * EMB_H;
    float z[EMB_VEC][8], sum[EMB_VEC][8];
#pragma unroll
    for (int i = 0; i < EMB_VEC; ++i) {
      const int c = (i * 32 + lane) * 8;
      float a[8], b[8];
      ld8h(xr + c, a);
      ld8f(pos + (long long)s * EMB_H + c, b);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        z[i][j] = a[j] + b[j];
        sum[i][j] = 0.f;
      }
      if (type != nullptr) {
        float tt[8];
        ld8f(type + (which == 0 ? 0 : EMB_H) + c, tt);
#pragma unroll
        for (int j = 0; j < 8; ++j) z[i][j] += tt[j];
      }
    }
    const int fan = src.all_pairs ? (which == 0 ? src.Nb : src.Na) : 1;
    for (int f = 0; f < fan; ++f) {
      const long long p = src.all_pairs ? (which == 0 ? owner * src.Nb + f : (long long)f * src.Nb + owner) : owner;
      const long long row = p * S + s;
      float dz[EMB_VEC][8];
      ln_row_bwd(z, dy + row * EMB_H, gamma, mean_in[row], rstd_in[row], row, drop, lane, dz, acc_g, acc_b);
#pragma unroll
      for (int i = 0; i < EMB_VEC; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) sum[i][j] += dz[i][j];
    }
    bf16* dst = (which == 0 ? da : db);
#pragma unroll
    for (int i = 0; i < EMB_VEC; ++i) {
      const int c = (i * 32 + lane) * 8;
      if (dst != nullptr) st8h(dst + sr * EMB_H + c, sum[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        atomicAdd(dpos + (long long)s * EMB_H + c + j, sum[i][j]);
        if (dtype != nullptr) atomicAdd(dtype + (which == 0 ? 0 : EMB_H) + c + j, sum[i][j]);
      }
    }
  }
  flush_colsums(acc_g, dgamma, red, warp, lane);
  flush_colsums(acc_b, dbeta, red, warp, lane);
}

static EmbDrop make_emb_drop(float p, const unsigned long long* rng, unsigned long long stream) {
  EmbDrop d;
  d.on = p > 0.f;
  d.threshold = dropout_threshold16(p);
  d.scale = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
  d.seed = 0;
  d.stream = stream;
  d.rng = rng;
  return d;
}
static int emb_grid(long long rows) {
  long long blocks = (rows + EMB_WARPS - 1) / EMB_WARPS;
  const long long cap = 148LL * 8;
  return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

}  // namespace univl

using namespace univl;

extern "C" int univl_embed_text_fwd(const long long* ids, const long long* type_ids, const float* word,
                                    const float* pos, const float* type, const float* gamma, const float* beta,
                                    void* y, float* mean, float* rstd, int n_seq, int S, int H, int vocab, float eps,
                                    float p_drop, const unsigned long long* rng_state, unsigned long long stream_id,
                                    void* stream) {
  UNIVL_CHECK_ARG(H == EMB_H, "embed_text_fwd: hidden size must be %d (got %d)", EMB_H, H);
  UNIVL_CHECK_ARG(ids && word && pos && gamma && beta && y && mean && rstd, "embed_text_fwd: null pointer");
  UNIVL_CHECK_ARG(n_seq >= 0 && S > 0 && vocab > 0, "embed_text_fwd: bad shape");
  if (n_seq == 0) return UNIVL_OK;
  embed_text_fwd_kernel<<<emb_grid((long long)n_seq * S), EMB_WARPS * 32, 0, (cudaStream_t)stream>>>(
      ids, type_ids, word, pos, type, gamma, beta, (bf16*)y, mean, rstd, n_seq, S, vocab, eps,
      make_emb_drop(p_drop, rng_state, stream_id));
  UNIVL_CHECK_LAUNCH("embed_text_fwd");
  return UNIVL_OK;
}

extern "C" int univl_embed_text_bwd(const void* dy, const long long* ids, const long long* type_ids,
                                    const float* word, const float* pos, const float* type, const float* gamma,
                                    const float* mean, const float* rstd, float* dword, float* dpos, float* dtype,
                                    float* dgamma, float* dbeta, int n_seq, int S, int H, int vocab, float p_drop,
                                    const unsigned long long* rng_state, unsigned long long stream_id, void* stream) {
  UNIVL_CHECK_ARG(H == EMB_H, "embed_text_bwd: hidden size must be %d (got %d)", EMB_H, H);
  UNIVL_CHECK_ARG(dy && ids && word && pos && gamma && mean && rstd && dword && dpos && dgamma && dbeta,
                  "embed_text_bwd: null pointer");
  if (n_seq == 0) return UNIVL_OK;
  embed_text_bwd_kernel<<<emb_grid((long long)n_seq * S), EMB_WARPS * 32, 0, (cudaStream_t)stream>>>(
      (const bf16*)dy, ids, type_ids, word, pos, type, gamma, mean, rstd, dword, dpos, dtype, dgamma, dbeta, n_seq, S,
      vocab, make_emb_drop(p_drop, rng_state, stream_id));
  UNIVL_CHECK_LAUNCH("embed_text_bwd");
  return UNIVL_OK;
}

extern "C" int univl_embed_src_fwd(const void* a, const void* b, const float* pos, const float* type,
                                   const float* gamma, const float* beta, void* y, float* mean, float* rstd, int Na,
                                   int Wa, int Nb, int Fb, int all_pairs, int H, float eps, float p_drop,
                                   const unsigned long long* rng_state, unsigned long long stream_id, void* stream) {
  UNIVL_CHECK_ARG(H == EMB_H, "embed_src_fwd: hidden size must be %d (got %d)", EMB_H, H);
  UNIVL_CHECK_ARG(a && pos && gamma && beta && y && mean && rstd, "embed_src_fwd: null pointer");
  UNIVL_CHECK_ARG(Na >= 0 && Wa > 0 && Fb >= 0 && (Fb == 0 || (b != nullptr && Nb > 0)), "embed_src_fwd: bad shape");
  UNIVL_CHECK_ARG(all_pairs || Fb == 0 || Na == Nb, "embed_src_fwd: aligned mode needs Na == Nb");
  SrcCfg src{(const bf16*)a, (const bf16*)b, Na, Wa, Fb == 0 ? 1 : Nb, Fb, all_pairs && Fb > 0};
  const long long n_seq = src.all_pairs ? (long long)Na * Nb : Na;
  if (n_seq == 0) return UNIVL_OK;
  embed_src_fwd_kernel<<<emb_grid(n_seq * (Wa + Fb)), EMB_WARPS * 32, 0, (cudaStream_t)stream>>>(
      src, pos, type, gamma, beta, (bf16*)y, mean, rstd, n_seq, eps, make_emb_drop(p_drop, rng_state, stream_id));
  UNIVL_CHECK_LAUNCH("embed_src_fwd");
  return UNIVL_OK;
}

extern "C" int univl_embed_src_bwd(const void* dy, const void* a, const void* b, const float* pos, const float* type,
                                   const float* gamma, const float* mean, const float* rstd, void* da, void* db,
                                   float* dpos, float* dtype, float* dgamma, float* dbeta, int Na, int Wa, int Nb,
                                   int Fb, int all_pairs, int H, float p_drop, const unsigned long long* rng_state,
                                   unsigned long long stream_id, void* stream) {
  UNIVL_CHECK_ARG(H == EMB_H, "embed_src_bwd: hidden size must be %d (got %d)", EMB_H, H);
  UNIVL_CHECK_ARG(dy && a && pos && gamma && mean && rstd && dpos && dgamma && dbeta, "embed_src_bwd: null pointer");
  UNIVL_CHECK_ARG(Fb == 0 || b != nullptr, "embed_src_bwd: missing second source");
  SrcCfg src{(const bf16*)a, (const bf16*)b, Na, Wa, Fb == 0 ? 1 : Nb, Fb, all_pairs && Fb > 0};
  if (Na == 0) return UNIVL_OK;
  const long long rows_a = (long long)Na * Wa, rows_b = (long long)(Fb == 0 ? 0 : Nb) * Fb;
  dim3 grid(emb_grid(rows_a > rows_b ? rows_a : rows_b), Fb == 0 ? 1 : 2);
  embed_src_bwd_kernel<<<grid, EMB_WARPS * 32, 0, (cudaStream_t)stream>>>(
      (const bf16*)dy, src, pos, type, gamma, mean, rstd, (bf16*)da, (bf16*)db, dpos, dtype, dgamma, dbeta,
      make_emb_drop(p_drop, rng_state, stream_id));
  UNIVL_CHECK_LAUNCH("embed_src_bwd");
  return UNIVL_OK;
}
