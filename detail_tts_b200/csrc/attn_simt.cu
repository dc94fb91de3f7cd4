// attn_simt.cu -- exact-fp32 attention on CUDA cores, one (query, head, utterance) per CTA.
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

// ---- KV-cache decode attention: ONE query per (utterance, head) against its cached keys/values.
// One warp per (utterance, head), 4 heads per CTA.  Phase 1: 4 lanes share a key (each 3 x float4 of the
// 48-dim row, 64 B contiguous per key per load), 8 keys per warp iteration, scores to shared memory.
// Phase 2: 12 lanes x float4 cover a value row (192 B coalesced), two lane groups take even/odd keys and
// are combined in a fixed order.  fp32 throughout; bit-reproducible.
constexpr int DEC_WARPS = 4;
__global__ void __launch_bounds__(DEC_WARPS * 32)
attention_decode_kernel(const dtts_attention_params p) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x * DEC_WARPS + warp, b = blockIdx.y;
  if (h >= p.n_heads || p.q_len[b] <= 0) return;
  const int nk = p.k_len[b];
  if (nk <= 0) return;
  float* sc = sm + warp * p.max_k_len;
  const float* q = (const float*)p.q + (long)p.q_off[b] * p.ldq + (long)h * p.head_stride_q;
  const float* kb = (const float*)p.k + (long)p.k_off[b] * p.ldk + (long)h * p.head_stride_k;
  const float* vb = (const float*)p.v + (long)p.k_off[b] * p.ldv + (long)h * p.head_stride_v;
  const int sub = lane & 3, kl = lane >> 2;
  float4 q4[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    q4[i] = *reinterpret_cast<const float4*>(q + 16 * i + 4 * sub);
    q4[i].x *= p.scale; q4[i].y *= p.scale; q4[i].z *= p.scale; q4[i].w *= p.scale;
  }
  float lmax = -INFINITY;
  for (int j0 = 0; j0 < nk; j0 += 8) {
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
      if (sub == 0) sc[j] = s;
      lmax = fmaxf(lmax, s);
    }
  }
  lmax = warp_max(lmax);
  __syncwarp();
  float lsum = 0.f;
  for (int j = lane; j < nk; j += 32) {
    const float e = expf(sc[j] - lmax);
    sc[j] = e;
    lsum += e;
  }
  const float inv = 1.0f / warp_sum(lsum);
  __syncwarp();
  const int g2 = lane / 12, d4 = lane % 12;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (g2 < 2) {
    for (int j = g2; j < nk; j += 2) {
      const float pj = sc[j];
      const float4 v4 = *reinterpret_cast<const float4*>(vb + (long)j * p.ldv + 4 * d4);
      acc.x = fmaf(pj, v4.x, acc.x); acc.y = fmaf(pj, v4.y, acc.y); acc.z = fmaf(pj, v4.z, acc.z); acc.w = fmaf(pj, v4.w, acc.w);
    }
  }
  const float ox = __shfl_down_sync(0xffffffffu, acc.x, 12), oy = __shfl_down_sync(0xffffffffu, acc.y, 12);
  const float oz = __shfl_down_sync(0xffffffffu, acc.z, 12), ow = __shfl_down_sync(0xffffffffu, acc.w, 12);
  if (lane < 12) {
    const float4 o = make_float4((acc.x + ox) * inv, (acc.y + oy) * inv, (acc.z + oz) * inv, (acc.w + ow) * inv);
    const long orow = (p.o_off ? p.o_off[b] : p.q_off[b]);
    if (p.out_lo) {   // tf32 operand split for the projection GEMM (dtts_gemm_tf32x3)
      float4 hi, lo;
      hi.x = __uint_as_float(__float_as_uint(o.x) & 0xFFFFE000u); lo.x = o.x - hi.x;
      hi.y = __uint_as_float(__float_as_uint(o.y) & 0xFFFFE000u); lo.y = o.y - hi.y;
      hi.z = __uint_as_float(__float_as_uint(o.z) & 0xFFFFE000u); lo.z = o.z - hi.z;
      hi.w = __uint_as_float(__float_as_uint(o.w) & 0xFFFFE000u); lo.w = o.w - hi.w;
      *reinterpret_cast<float4*>(p.out_f32 + orow * p.ldo32 + h * 48 + 4 * d4) = hi;
      *reinterpret_cast<float4*>(p.out_lo + orow * p.ldo_lo + h * 48 + 4 * d4) = lo;
    } else if (p.out_f32) *reinterpret_cast<float4*>(p.out_f32 + orow * p.ldo32 + h * 48 + 4 * d4) = o;
    if (p.out_f16) {
      __half* hp = (__half*)p.out_f16 + orow * p.ldo16 + h * 48 + 4 * d4;
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
      (size_t)DEC_WARPS * p->max_k_len * sizeof(float) <= 200 * 1024) {
    static bool dec_attr = false;
    if (!dec_attr) {
      cudaFuncSetAttribute(attention_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      dec_attr = true;
    }
    dim3 grid(ceil_div(p->n_heads, DEC_WARPS), p->n_utt);
    attention_decode_kernel<<<grid, DEC_WARPS * 32, (size_t)DEC_WARPS * p->max_k_len * sizeof(float), (cudaStream_t)stream>>>(*p);
    DTTS_CHECK_LAUNCH("attention_decode");
    return 0;
  }
  DTTS_REQUIRE(!p->out_lo, "attention_f32: out_lo (tf32 split) is only produced by the fp32 KV-cache decode path");
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
