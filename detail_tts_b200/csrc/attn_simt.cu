// attn_simt.cu -- exact-fp32 attention on CUDA cores: a 32-query tiled kernel, a KV-cache decode kernel and a
// generic one-query-per-CTA fallback.
// Covers every small attention on the path with one kernel:
//   GPT-2 causal attention, prefill and KV-cache decode  (transformers modeling_gpt2.py:54-72)
//   MelStyleEncoder 2-head attention, temperature sqrt(d_model)  (vqvae/modules/modules.py:565-639)
//   enc_p windowed relative-position attention  (vqvae/modules/attentions.py:198-239; closed form in
//     SURVEY.md Appendix D6)
//   contextual_embedder / latent_conditioner / SIMT check of the diffusion AttentionBlock
//     (vqvae/utils/diff_util.py:145-169, xtransformers.py:177-186)
#include "common.cuh"

namespace {

template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ldf<__half>(const __half* p) { return __half2float(*p); }

template <typename T>
__global__ void __launch_bounds__(256)
attention_simt_kernel(const dtts_attention_params p) {
  extern __shared__ float sm[];
  __shared__ float red[40];
  const int i = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  if (i >= p.q_len[b]) return;
  const int hd = p.head_dim;
  int nk = p.k_len[b];
  if (p.causal) {
    int lim = i + (p.causal_offset ? p.causal_offset[b] : 0) + 1;
    nk = lim < nk ? lim : nk;
  }
  if (nk <= 0) return;
  float* qs = sm;                 // [hd]
  float* sc = sm + hd;            // [nk]
  float* part = sc + nk;          // [JP*hd]
  const T* q = (const T*)p.q + (long)(p.q_off[b] + i) * p.ldq + (long)h * p.head_stride_q;
  const T* kb = (const T*)p.k + (long)p.k_off[b] * p.ldk + (long)h * p.head_stride_k;
  const T* vb = (const T*)p.v + (long)p.k_off[b] * p.ldv + (long)h * p.head_stride_v;
  for (int d = threadIdx.x; d < hd; d += blockDim.x) qs[d] = ldf<T>(q + d) * p.scale;
  __syncthreads();
  // scores
  float lmax = -INFINITY;
  for (int j = threadIdx.x; j < nk; j += blockDim.x) {
    const T* kj = kb + (long)j * p.ldk;
    float s = 0.f;
    for (int d = 0; d < hd; ++d) s = fmaf(qs[d], ldf<T>(kj + d), s);
    if (p.bias_mode == DTTS_ATTN_BIAS_RELPOS_TABLE) {
      int r = j - i;
      r = r < -p.bias_half ? -p.bias_half : (r > p.bias_half ? p.bias_half : r);
      s += p.bias_table[h * (2 * p.bias_half + 1) + r + p.bias_half];
    } else if (p.bias_mode == DTTS_ATTN_BIAS_WINDOW_REL) {
      int r = j - i;
      if (r >= -p.window && r <= p.window) {
        const float* ek = p.rel_k + (r + p.window) * hd;
        float t = 0.f;
        for (int d = 0; d < hd; ++d) t = fmaf(qs[d], ek[d], t);
        s += t;
      }
    }
    sc[j] = s;
    lmax = fmaxf(lmax, s);
  }
  const float mx = block_max(lmax, red);
  float lsum = 0.f;
  for (int j = threadIdx.x; j < nk; j += blockDim.x) {
    float e = expf(sc[j] - mx);
    sc[j] = e;
    lsum += e;
  }
  const float inv = 1.0f / block_sum(lsum, red);  // also orders the sc[] writes before the reads below
  // output: thread (jp, d)
  const int JP = blockDim.x / hd;
  const int d = threadIdx.x % hd, jp = threadIdx.x / hd;
  float o = 0.f;
  if (jp < JP) {
    for (int j = jp; j < nk; j += JP) o = fmaf(sc[j], ldf<T>(vb + (long)j * p.ldv + d), o);
    if (p.bias_mode == DTTS_ATTN_BIAS_WINDOW_REL && jp == 0) {
      for (int r = -p.window; r <= p.window; ++r) {
        int j = i + r;
        if (j >= 0 && j < nk) o = fmaf(sc[j], p.rel_v[(r + p.window) * hd + d], o);
      }
    }
    part[jp * hd + d] = o;
  }
  __syncthreads();
  if (threadIdx.x < hd) {
    float t = 0.f;
    for (int q_ = 0; q_ < JP; ++q_) t += part[q_ * hd + threadIdx.x];
    t *= inv;
    const long orow = (p.o_off ? p.o_off[b] : p.q_off[b]) + i;
    if (p.out_f32) p.out_f32[orow * p.ldo32 + h * hd + threadIdx.x] = t;
    if (p.out_f16) ((__half*)p.out_f16)[orow * p.ldo16 + h * hd + threadIdx.x] = __float2half_rn(t);
  }
}

// ---- Tiled variant for everything with more than one query per utterance (GPT prefill, MelStyleEncoder,
// enc_p, contextual_embedder / latent_conditioner): one CTA = (utterance, head, 32 queries), 8 warps x 4 queries.
// Keys / values stream through shared memory in 64-key chunks (read from global ONCE per 32 queries instead of
// once per query), scores are register-blocked 4 queries x 2 keys per lane, the softmax is online (running max /
// sum per query, fp32, exact expf), P.V has lanes over the head dimension.  Same semantics as the per-query
// kernel below: causal offset, relative-position bias table, windowed relative key/value embeddings.
constexpr int AT_TQ = 32, AT_KT = 64, AT_WARPS = 8, AT_QPW = 4;

template <typename T, int DPL>   // DPL = head-dim elements per lane: head_dim <= 32*DPL
__global__ void __launch_bounds__(AT_WARPS * 32)
attention_tile_kernel(const dtts_attention_params p) {
  extern __shared__ float4 sm4[];
  float* sm = reinterpret_cast<float*>(sm4);
  const int hd = p.head_dim, ldks = hd + 1;
  float* Qs = sm;                                   // [TQ][hd]   (pre-scaled)
  float* Ks = Qs + AT_TQ * hd;                      // [KT][hd+1]
  float* Vs = Ks + AT_KT * ldks;                    // [KT][hd]
  float* Ps = Vs + AT_KT * hd;                      // [TQ][KT]
  float* Rq = Ps + AT_TQ * AT_KT;                   // [TQ][2*window+1]
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * AT_TQ;
  const int qlen = p.q_len[b];
  if (q0 >= qlen) return;
  const int klen = p.k_len[b];
  const int coff = p.causal_offset ? p.causal_offset[b] : 0;
  int nk = klen;
  if (p.causal) nk = min(nk, q0 + AT_TQ + coff);    // the tile's last query sees keys <= q0 + TQ - 1 + coff
  if (nk <= 0) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const T* qg = (const T*)p.q + (long)p.q_off[b] * p.ldq + (long)h * p.head_stride_q;
  const T* kg = (const T*)p.k + (long)p.k_off[b] * p.ldk + (long)h * p.head_stride_k;
  const T* vg = (const T*)p.v + (long)p.k_off[b] * p.ldv + (long)h * p.head_stride_v;
  for (int idx = tid; idx < AT_TQ * hd; idx += AT_WARPS * 32) {
    const int r = idx / hd, d = idx - r * hd;
    Qs[idx] = q0 + r < qlen ? ldf<T>(qg + (long)(q0 + r) * p.ldq + d) * p.scale : 0.f;
  }
  const int nw = 2 * p.window + 1;
  const bool win = p.bias_mode == DTTS_ATTN_BIAS_WINDOW_REL, tab = p.bias_mode == DTTS_ATTN_BIAS_RELPOS_TABLE;
  if (win) {
    __syncthreads();
    for (int idx = tid; idx < AT_TQ * nw; idx += AT_WARPS * 32) {
      const int r = idx / nw, w = idx - r * nw;
      float t = 0.f;
      for (int d = 0; d < hd; ++d) t = fmaf(Qs[r * hd + d], p.rel_k[w * hd + d], t);
      Rq[idx] = t;
    }
  }
  float m_run[AT_QPW], l_run[AT_QPW], o[AT_QPW][DPL];
#pragma unroll
  for (int qi = 0; qi < AT_QPW; ++qi) {
    m_run[qi] = -INFINITY; l_run[qi] = 0.f;
#pragma unroll
    for (int dd = 0; dd < DPL; ++dd) o[qi][dd] = 0.f;
  }
  const int qrow = warp * AT_QPW;            // first query (tile-local) of this warp
  for (int k0 = 0; k0 < nk; k0 += AT_KT) {
    __syncthreads();                         // previous chunk fully consumed (and Qs / Rq visible)
    for (int idx = tid; idx < AT_KT * hd; idx += AT_WARPS * 32) {
      const int j = idx / hd, d = idx - j * hd;
      const bool ok = k0 + j < nk;
      Ks[j * ldks + d] = ok ? ldf<T>(kg + (long)(k0 + j) * p.ldk + d) : 0.f;
      Vs[idx] = ok ? ldf<T>(vg + (long)(k0 + j) * p.ldv + d) : 0.f;
    }
    __syncthreads();
    float s[AT_QPW][2];
#pragma unroll
    for (int qi = 0; qi < AT_QPW; ++qi) s[qi][0] = s[qi][1] = 0.f;
    const float* ka = Ks + lane * ldks;
    const float* kb = Ks + (lane + 32) * ldks;
    for (int d = 0; d < hd; d += 4) {
      float4 qv[AT_QPW];
#pragma unroll
      for (int qi = 0; qi < AT_QPW; ++qi) qv[qi] = *reinterpret_cast<const float4*>(Qs + (qrow + qi) * hd + d);
      const float a0 = ka[d], a1 = ka[d + 1], a2 = ka[d + 2], a3 = ka[d + 3];
      const float b0 = kb[d], b1 = kb[d + 1], b2 = kb[d + 2], b3 = kb[d + 3];
#pragma unroll
      for (int qi = 0; qi < AT_QPW; ++qi) {
        s[qi][0] = fmaf(qv[qi].x, a0, s[qi][0]); s[qi][0] = fmaf(qv[qi].y, a1, s[qi][0]);
        s[qi][0] = fmaf(qv[qi].z, a2, s[qi][0]); s[qi][0] = fmaf(qv[qi].w, a3, s[qi][0]);
        s[qi][1] = fmaf(qv[qi].x, b0, s[qi][1]); s[qi][1] = fmaf(qv[qi].y, b1, s[qi][1]);
        s[qi][1] = fmaf(qv[qi].z, b2, s[qi][1]); s[qi][1] = fmaf(qv[qi].w, b3, s[qi][1]);
      }
    }
#pragma unroll
    for (int qi = 0; qi < AT_QPW; ++qi) {
      const int i = q0 + qrow + qi;
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        const int j = k0 + lane + 32 * jj;
        float v = s[qi][jj];
        if (tab) {
          int r = j - i;
          r = r < -p.bias_half ? -p.bias_half : (r > p.bias_half ? p.bias_half : r);
          v += p.bias_table[h * (2 * p.bias_half + 1) + r + p.bias_half];
        } else if (win) {
          const int r = j - i;
          if (r >= -p.window && r <= p.window) v += Rq[(qrow + qi) * nw + r + p.window];
        }
        const bool ok = j < klen && (!p.causal || j <= i + coff);
        s[qi][jj] = ok ? v : -INFINITY;
      }
      const float mx = warp_max(fmaxf(s[qi][0], s[qi][1]));
      const float mn = fmaxf(m_run[qi], mx);
      float corr = 1.f, p0 = 0.f, p1 = 0.f;
      if (mn != -INFINITY) {
        corr = expf(m_run[qi] - mn);
        p0 = expf(s[qi][0] - mn);
        p1 = expf(s[qi][1] - mn);
      }
      m_run[qi] = mn;
      l_run[qi] = l_run[qi] * corr + warp_sum(p0 + p1);
#pragma unroll
      for (int dd = 0; dd < DPL; ++dd) o[qi][dd] *= corr;
      Ps[(qrow + qi) * AT_KT + lane] = p0;
      Ps[(qrow + qi) * AT_KT + lane + 32] = p1;
    }
    __syncwarp();
    for (int j = 0; j < AT_KT; j += 4) {
      float4 pv[AT_QPW];
#pragma unroll
      for (int qi = 0; qi < AT_QPW; ++qi) pv[qi] = *reinterpret_cast<const float4*>(Ps + (qrow + qi) * AT_KT + j);
#pragma unroll
      for (int dd = 0; dd < DPL; ++dd) {
        const int d = lane + 32 * dd;
        if (d < hd) {
          const float v0 = Vs[j * hd + d], v1 = Vs[(j + 1) * hd + d], v2 = Vs[(j + 2) * hd + d], v3 = Vs[(j + 3) * hd + d];
#pragma unroll
          for (int qi = 0; qi < AT_QPW; ++qi) {
            o[qi][dd] = fmaf(pv[qi].x, v0, o[qi][dd]); o[qi][dd] = fmaf(pv[qi].y, v1, o[qi][dd]);
            o[qi][dd] = fmaf(pv[qi].z, v2, o[qi][dd]); o[qi][dd] = fmaf(pv[qi].w, v3, o[qi][dd]);
          }
        }
      }
    }
    if (win) {
#pragma unroll
      for (int qi = 0; qi < AT_QPW; ++qi) {
        const int i = q0 + qrow + qi;
        for (int w = 0; w < nw; ++w) {
          const int j = i + w - p.window;
          if (j >= k0 && j < k0 + AT_KT) {
            const float pj = Ps[(qrow + qi) * AT_KT + j - k0];
#pragma unroll
            for (int dd = 0; dd < DPL; ++dd) {
              const int d = lane + 32 * dd;
              if (d < hd) o[qi][dd] = fmaf(pj, p.rel_v[w * hd + d], o[qi][dd]);
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int qi = 0; qi < AT_QPW; ++qi) {
    const int i = q0 + qrow + qi;
    if (i >= qlen) continue;
    const float inv = l_run[qi] > 0.f ? 1.0f / l_run[qi] : 0.f;
    const long orow = (long)(p.o_off ? p.o_off[b] : p.q_off[b]) + i;
#pragma unroll
    for (int dd = 0; dd < DPL; ++dd) {
      const int d = lane + 32 * dd;
      if (d < hd) {
        const float t = o[qi][dd] * inv;
        if (p.out_f32) p.out_f32[orow * p.ldo32 + h * hd + d] = t;
        if (p.out_f16) ((__half*)p.out_f16)[orow * p.ldo16 + h * hd + d] = __float2half_rn(t);
      }
    }
  }
}

template <typename T, int DPL>
int launch_tile(const dtts_attention_params* p, cudaStream_t st) {
  const int hd = p->head_dim;
  const size_t smem = (size_t)(AT_TQ * hd + AT_KT * (hd + 1) + AT_KT * hd + AT_TQ * AT_KT + AT_TQ * (2 * p->window + 1) + 4) * sizeof(float);
  static size_t attr = 0;
  if (smem > attr) {
    cudaError_t e = cudaFuncSetAttribute(attention_tile_kernel<T, DPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) DTTS_FAIL(-3, "cudaFuncSetAttribute(attention_tile): %s", cudaGetErrorString(e));
    attr = smem;
  }
  dim3 grid(ceil_div(p->max_q_len, AT_TQ), p->n_heads, p->n_utt);
  attention_tile_kernel<T, DPL><<<grid, AT_WARPS * 32, smem, st>>>(*p);
  DTTS_CHECK_LAUNCH("attention_tile");
  return 0;
}

// ---- KV-cache decode attention: ONE query per (utterance, head) against its cached keys/values.
// The step is latency-bound (ten dependent launches per token), so the kernel is organised for memory-level parallelism:
// one CTA of 4 warps per (utterance, head); the keys are dealt to the warps 8 at a time, so that at 125 cached positions
// every key row (phase 1) and every value row (phase 2) of the head is in flight at once.
// Phase 1: 4 lanes share a key (3 x float4 of the 48-dim row each), 8 keys per warp iteration, scores to shared memory.
// Phase 2: 8 lanes x 6 floats cover a value row, 4 keys per warp iteration; lane groups, then warps, are combined in a
// fixed order.  fp32 throughout; bit-reproducible.  Optionally starts by reducing the split-K partials of the new
// token's q|k|v columns (dtts_attention_params.qkv_ws).
constexpr int DEC_WARPS = 4;
__global__ void __launch_bounds__(DEC_WARPS * 32)
attention_decode_kernel(const dtts_attention_params p) {
  extern __shared__ float sm[];          // [max_k_len] scores
  __shared__ __align__(16) float qs[48];
  __shared__ float red[2 * DEC_WARPS];
  __shared__ __align__(16) float part[DEC_WARPS][48];
  pdl_launch();
  pdl_wait();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = blockIdx.x, b = blockIdx.y;
  if (p.q_len[b] <= 0) return;
  const int nk = p.k_len[b];
  if (nk <= 0) return;
  const long arow = (long)p.q_off[b];
  const float* q = (const float*)p.q + arow * p.ldq + (long)h * p.head_stride_q;
  const float* kb = (const float*)p.k + (long)p.k_off[b] * p.ldk + (long)h * p.head_stride_k;
  const float* vb = (const float*)p.v + (long)p.k_off[b] * p.ldv + (long)h * p.head_stride_v;
  if (p.qkv_ws) {
    // fused split-K reduce of this head's new q|k|v columns (same arithmetic and order as reduce_kernel, gpt_step.cu)
    const int width = p.n_heads * 48;
    const float* wsr = p.qkv_ws + (long)b * p.qkv_ld_ws;
    for (int c = tid; c < 3 * 48; c += DEC_WARPS * 32) {
      const int prt = c / 48, d = c - prt * 48;
      const int col = prt * width + h * 48 + d;
      float v = 0.f;
#pragma unroll 4
      for (int k = 0; k < p.qkv_splits; ++k) v += wsr[(long)k * p.qkv_split_stride + col];   // fixed order
      if (p.qkv_bias) v += __ldg(p.qkv_bias + col);
      float* dst = prt == 0 ? (float*)p.q + arow * p.ldq + (long)h * p.head_stride_q
                 : prt == 1 ? (float*)p.k + arow * p.ldk + (long)h * p.head_stride_k
                            : (float*)p.v + arow * p.ldv + (long)h * p.head_stride_v;
      dst[d] = v;
      if (prt == 0) qs[d] = v * p.scale;
    }
  } else if (tid < 48) {
    qs[tid] = q[tid] * p.scale;
  }
  __syncthreads();   // q staged; the new key/value row (global) is visible to the whole CTA
  const int sub = lane & 3, kl = lane >> 2;
  float4 q4[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) q4[i] = *reinterpret_cast<const float4*>(qs + 16 * i + 4 * sub);
  float lmax = -INFINITY;
#pragma unroll 4
  for (int j0 = warp * 8; j0 < nk; j0 += DEC_WARPS * 8) {
    const int j = j0 + kl;
    float s = 0.f;
    if (j < nk) {
      const float* kj = kb + (long)j * p.ldk + 4 * sub;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const float4 k4 = *reinterpret_cast<const float4*>(kj + 16 * i);
        s = fmaf(q4[i].x, k4.x, s); s = fmaf(q4[i].y, k4.y, s); s = fmaf(q4[i].z, k4.z, s); s = fmaf(q4[i].w, k4.w, s);
      }
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    if (j < nk) {
      if (sub == 0) sm[j] = s;
      lmax = fmaxf(lmax, s);
    }
  }
  lmax = warp_max(lmax);
  if (lane == 0) red[warp] = lmax;
  __syncthreads();
  const float gmax = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  float lsum = 0.f;
  for (int j = tid; j < nk; j += DEC_WARPS * 32) {
    const float e = expf(sm[j] - gmax);
    sm[j] = e;
    lsum += e;
  }
  lsum = warp_sum(lsum);
  if (lane == 0) red[DEC_WARPS + warp] = lsum;
  __syncthreads();
  const float inv = 1.0f / (((red[DEC_WARPS] + red[DEC_WARPS + 1]) + red[DEC_WARPS + 2]) + red[DEC_WARPS + 3]);
  // phase 2: lane group g (8 lanes) takes key j0 + g, lane l of the group the dims [6l, 6l+6): all 32 lanes load, 16 keys
  // per CTA iteration, unrolled so that every value row of the head is in flight at once (measured: 31 us vs 41 us for the
  // 12-lanes-x-float4 mapping at 128 utterances x 125 positions)
  const int g = lane >> 3, l6 = (lane & 7) * 6;
  float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
  for (int j0 = warp * 4; j0 < nk; j0 += DEC_WARPS * 4) {
    const int j = j0 + g;
    if (j < nk) {
      const float pj = sm[j];
      const float2* vj = reinterpret_cast<const float2*>(vb + (long)j * p.ldv + l6);
      const float2 a = vj[0], c = vj[1], e = vj[2];
      acc[0] = fmaf(pj, a.x, acc[0]); acc[1] = fmaf(pj, a.y, acc[1]); acc[2] = fmaf(pj, c.x, acc[2]);
      acc[3] = fmaf(pj, c.y, acc[3]); acc[4] = fmaf(pj, e.x, acc[4]); acc[5] = fmaf(pj, e.y, acc[5]);
    }
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 8);
    acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16);
  }
  if (lane < 8) {
#pragma unroll
    for (int i = 0; i < 6; ++i) part[warp][l6 + i] = acc[i];
  }
  __syncthreads();
  if (tid < 12) {
    const int d = 4 * tid;
    float o4[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) o4[i] = (((part[0][d + i] + part[1][d + i]) + part[2][d + i]) + part[3][d + i]) * inv;
    const float4 o = make_float4(o4[0], o4[1], o4[2], o4[3]);
    const long orow = (p.o_off ? p.o_off[b] : p.q_off[b]);
    if (p.out_lo) {   // tf32 operand split for the projection GEMM (dtts_gemm_tf32x3)
      float4 hi, lo;
      hi.x = __uint_as_float(__float_as_uint(o.x) & 0xFFFFE000u); lo.x = o.x - hi.x;
      hi.y = __uint_as_float(__float_as_uint(o.y) & 0xFFFFE000u); lo.y = o.y - hi.y;
      hi.z = __uint_as_float(__float_as_uint(o.z) & 0xFFFFE000u); lo.z = o.z - hi.z;
      hi.w = __uint_as_float(__float_as_uint(o.w) & 0xFFFFE000u); lo.w = o.w - hi.w;
      *reinterpret_cast<float4*>(p.out_f32 + orow * p.ldo32 + h * 48 + d) = hi;
      *reinterpret_cast<float4*>(p.out_lo + orow * p.ldo_lo + h * 48 + d) = lo;
    } else if (p.out_f32) *reinterpret_cast<float4*>(p.out_f32 + orow * p.ldo32 + h * 48 + d) = o;
    if (p.out_f16) {
      __half* hp = (__half*)p.out_f16 + orow * p.ldo16 + h * 48 + d;
      hp[0] = __float2half_rn(o.x); hp[1] = __float2half_rn(o.y); hp[2] = __float2half_rn(o.z); hp[3] = __float2half_rn(o.w);
    }
  }
}

}  // namespace

extern "C" int dtts_attention_f32(const dtts_attention_params* p, void* stream) {
  DTTS_REQUIRE(p && p->q && p->k && p->v && p->q_off && p->q_len && p->k_off && p->k_len, "attention_f32: null argument");
  DTTS_REQUIRE(p->head_dim > 0 && p->head_dim <= 256, "attention_f32: head_dim out of range");
  DTTS_REQUIRE(p->out_f32 || p->out_f16, "attention_f32: no output");
  DTTS_REQUIRE(p->max_q_len > 0 && p->n_utt > 0 && p->n_heads > 0, "attention_f32: empty problem");
  DTTS_REQUIRE(p->bias_mode != DTTS_ATTN_BIAS_RELPOS_TABLE || p->bias_table, "attention_f32: missing bias table");
  DTTS_REQUIRE(p->bias_mode != DTTS_ATTN_BIAS_WINDOW_REL || (p->rel_k && p->rel_v), "attention_f32: missing rel embeddings");
  static int max_smem = 0;
  if (!max_smem) {
    max_smem = 200 * 1024;
    cudaFuncSetAttribute(attention_simt_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    cudaFuncSetAttribute(attention_simt_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
  }
  DTTS_REQUIRE(p->max_k_len > 0, "attention_f32: max_k_len must be set");
  // KV-cache decode fast path: one fp32 query per utterance, head_dim 48, no bias
  if (p->max_q_len == 1 && !p->is_f16 && p->head_dim == 48 && p->bias_mode == DTTS_ATTN_BIAS_NONE && !p->causal &&
      p->ldq % 4 == 0 && p->ldk % 4 == 0 && p->ldv % 4 == 0 && p->head_stride_q % 4 == 0 && p->head_stride_k % 4 == 0 &&
      p->head_stride_v % 4 == 0 && ((((uintptr_t)p->q) | ((uintptr_t)p->k) | ((uintptr_t)p->v)) & 15) == 0 &&
      (!p->out_f32 || (p->ldo32 % 4 == 0 && (((uintptr_t)p->out_f32) & 15) == 0)) &&
      (size_t)p->max_k_len * sizeof(float) <= 200 * 1024) {
    static bool dec_attr = false;
    if (!dec_attr) {
      cudaFuncSetAttribute(attention_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      dec_attr = true;
    }
    dim3 grid(p->n_heads, p->n_utt);
    launch_maybe_pdl(attention_decode_kernel, grid, dim3(DEC_WARPS * 32), (size_t)p->max_k_len * sizeof(float), (cudaStream_t)stream, *p);
    DTTS_CHECK_LAUNCH("attention_decode");
    return 0;
  }
  DTTS_REQUIRE(!p->out_lo, "attention_f32: out_lo (tf32 split) is only produced by the fp32 KV-cache decode path");
  DTTS_REQUIRE(!p->qkv_ws, "attention_f32: the fused QKV split-K reduce is only available on the fp32 KV-cache decode path");
  if (p->max_q_len > 1 && p->head_dim % 4 == 0 && p->head_dim <= 192 && p->window >= 0 && p->window <= 16) {
    cudaStream_t st = (cudaStream_t)stream;
    const int hd = p->head_dim;
    if (p->is_f16) return hd <= 64 ? launch_tile<__half, 2>(p, st) : hd <= 96 ? launch_tile<__half, 3>(p, st) : launch_tile<__half, 6>(p, st);
    return hd <= 64 ? launch_tile<float, 2>(p, st) : hd <= 96 ? launch_tile<float, 3>(p, st) : launch_tile<float, 6>(p, st);
  }
  const int threads = 256;
  const size_t smem = (size_t)(p->head_dim + p->max_k_len + threads) * sizeof(float);
  DTTS_REQUIRE(smem <= (size_t)max_smem, "attention_f32: too many keys for one CTA (%d)", p->max_k_len);
  dim3 grid(p->max_q_len, p->n_heads, p->n_utt);
  if (p->is_f16)
    attention_simt_kernel<__half><<<grid, threads, smem, (cudaStream_t)stream>>>(*p);
  else
    attention_simt_kernel<float><<<grid, threads, smem, (cudaStream_t)stream>>>(*p);
  DTTS_CHECK_LAUNCH("attention_simt");
  return 0;
}
