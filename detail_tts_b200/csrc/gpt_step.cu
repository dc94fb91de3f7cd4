// gpt_step.cu -- glue kernels of the tensor-core GPT decode step (3xTF32 GEMMs, gemm_tc.cu):
//   dtts_split_tf32     x -> (hi, lo), hi tf32-exact: operand preparation
//   dtts_splitk_reduce  fixed-order sum of split-K partial tiles + bias + activation + residual, fused with the
//                       LayerNorm that follows in GPT2Block and the operand split of its output
// Reference ops replaced: HF GPT2Block residual adds / ln_1 / ln_2 / ln_f / gelu_new (transformers
// modeling_gpt2.py:262-310), gpt/model.py:322,173 (final_norm before mel_head).
#include "common.cuh"

namespace {

__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);   // keep sign, exponent, 10 mantissa bits
  lo = x - hi;                                               // exact in fp32
}

__global__ void __launch_bounds__(256)
split_kernel(const dtts_split_params p) {
  const int c4 = p.C >> 2;
  const long total = (long)p.M * c4;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int m = (int)(idx / c4), c = (int)(idx % c4) * 4;
    const float4 v = *reinterpret_cast<const float4*>(p.x + (long)m * p.ldx + c);
    float4 h, l;
    split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
    *reinterpret_cast<float4*>(p.hi + (long)m * p.ld + c) = h;
    *reinterpret_cast<float4*>(p.lo + (long)m * p.ld + c) = l;
  }
}

constexpr int RED_MAX_THREADS = 1024;
constexpr int RED_MAXE = 4;   // LayerNorm path: N <= 4 * blockDim.x

// One CTA per row; the CTA is as wide as the row allows (up to 1024 threads) so that the n_splits dependent-latency
// loads of every column are in flight at once: the kernel is pure latency (a few MB of partials, mostly L2 hits).
template <bool LN>
__global__ void __launch_bounds__(RED_MAX_THREADS)
reduce_kernel(const dtts_reduce_params p) {
  const int RED_THREADS = blockDim.x;
  __shared__ float red[40];
  pdl_launch();
  pdl_wait();
  const int m = blockIdx.x;
  const long orow = p.out_row_map ? (long)p.out_row_map[m] : (long)m;
  float vals[RED_MAXE];
  float s = 0.f;
  for (int i = 0, n = threadIdx.x; n < p.N || (LN && i < RED_MAXE); ++i, n += RED_THREADS) {
    float v = 0.f;
    if (n < p.N) {
      for (int k = 0; k < p.n_splits; ++k) v += p.ws[(long)k * p.split_stride + (long)m * p.ld_ws + n];   // fixed order
      if (p.bias) v += __ldg(p.bias + n);
      v = act_apply(p.act, v, p.act_param);
      if (p.res) v += p.res[(long)m * p.ldr + n];
      if (p.out_f32) p.out_f32[orow * p.ldo32 + n] = v;
      if (!LN) {
        if (p.y_f32) p.y_f32[(long)m * p.ldy + n] = v;
        if (p.y_hi) {
          float h, l;
          split_tf32(v, h, l);
          p.y_hi[(long)m * p.ld_hl + n] = h;
          p.y_lo[(long)m * p.ld_hl + n] = l;
        }
      }
    }
    if (LN) {
      if (i < RED_MAXE) vals[i] = v;
      s += v;
    }
  }
  if (!LN) return;
  const float mean = block_sum(s, red) / p.N;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < RED_MAXE; ++i) {
    const int n = threadIdx.x + i * RED_THREADS;
    if (n < p.N) { const float d = vals[i] - mean; ss += d * d; }
  }
  const float rstd = rsqrtf(block_sum(ss, red) / p.N + p.ln_eps);
#pragma unroll
  for (int i = 0; i < RED_MAXE; ++i) {
    const int n = threadIdx.x + i * RED_THREADS;
    if (n < p.N) {
      const float y = (vals[i] - mean) * rstd * __ldg(p.ln_gamma + n) + __ldg(p.ln_beta + n);
      if (p.y_f32) p.y_f32[(long)m * p.ldy + n] = y;
      if (p.y_hi) {
        float h, l;
        split_tf32(y, h, l);
        p.y_hi[(long)m * p.ld_hl + n] = h;
        p.y_lo[(long)m * p.ld_hl + n] = l;
      }
    }
  }
}

}  // namespace

extern "C" int dtts_split_tf32(const dtts_split_params* p, void* stream) {
  DTTS_REQUIRE(p && p->x && p->hi && p->lo, "split_tf32: null argument");
  DTTS_REQUIRE(p->C % 4 == 0 && p->ldx % 4 == 0 && p->ld % 4 == 0, "split_tf32: C/ld must be multiples of 4");
  DTTS_REQUIRE(((((uintptr_t)p->x) | ((uintptr_t)p->hi) | ((uintptr_t)p->lo)) & 15) == 0, "split_tf32: pointers must be 16-byte aligned");
  if (p->M <= 0) return 0;
  long g = ((long)p->M * p->C / 4 + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  split_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>(*p);
  DTTS_CHECK_LAUNCH("split_tf32");
  return 0;
}

extern "C" int dtts_splitk_reduce(const dtts_reduce_params* p, void* stream) {
  DTTS_REQUIRE(p && (p->n_splits == 0 || p->ws), "splitk_reduce: null partial buffer");
  DTTS_REQUIRE(p->n_splits > 0 || p->res, "splitk_reduce: nothing to reduce");
  DTTS_REQUIRE(!(p->y_hi && !p->y_lo), "splitk_reduce: y_hi needs y_lo");
  DTTS_REQUIRE(p->act < DTTS_ACT_PAIR_TANH_SIGMOID, "splitk_reduce: pair activations are not supported");
  if (p->M <= 0) return 0;
  int threads = (p->N + 31) / 32 * 32;
  if (threads > RED_MAX_THREADS) threads = RED_MAX_THREADS;
  if (threads < 64) threads = 64;
  if (p->ln_gamma) {
    DTTS_REQUIRE(p->ln_beta && p->N <= RED_MAX_THREADS * RED_MAXE, "splitk_reduce: LayerNorm path needs N <= 4096");
    launch_maybe_pdl(reduce_kernel<true>, dim3(p->M), dim3(threads), 0, (cudaStream_t)stream, *p);
  } else {
    launch_maybe_pdl(reduce_kernel<false>, dim3(p->M), dim3(threads), 0, (cudaStream_t)stream, *p);
  }
  DTTS_CHECK_LAUNCH("splitk_reduce");
  return 0;
}
