// norm.cu -- GroupNorm32 (+FiLM +SiLU) per utterance and row LayerNorm.
//   GroupNorm: vqvae/utils/diff_util.py:113-133 (fp32 stats over channels-in-group x frames),
//              vqvae/diff_model.py:107,113-115 (scale-shift norm), :242 (code_norm FiLM).
//   LayerNorm: HF GPT2Block ln_1/ln_2/ln_f, gpt/model.py:322, vqvae/modules/modules.py:36-48.
// GroupNorm is HBM/L2-bound: one CTA owns (utterance, 4 groups); the strip is read once from
// global into shared memory (when it fits), statistics are two-pass exact (mean, then centred
// variance) and the normalised fp16 GEMM operand is written straight from shared memory.
#include <stdlib.h>

#include "common.cuh"
#include "tcgen05.cuh"

namespace {

template <typename T> __device__ __forceinline__ float4 load4(const T* p);
template <> __device__ __forceinline__ float4 load4<float>(const float* p) { return *reinterpret_cast<const float4*>(p); }
template <> __device__ __forceinline__ float4 load4<__half>(const __half* p) {
  uint2 u = *reinterpret_cast<const uint2*>(p);
  __half2 a = *reinterpret_cast<__half2*>(&u.x), b = *reinterpret_cast<__half2*>(&u.y);
  float2 fa = __half22float2(a), fb = __half22float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}

constexpr int GN_GROUPS_PER_CTA = 4;

template <typename T>
__global__ void __launch_bounds__(384)
groupnorm_kernel(const T* __restrict__ x, int ldx, int C, int cpg, const int* __restrict__ utt_off,
                 const int* __restrict__ utt_len, const float* __restrict__ gamma, const float* __restrict__ beta,
                 const float* __restrict__ film_scale, const float* __restrict__ film_shift, int ld_film,
                 const int* __restrict__ film_idx, int act, float eps, float* __restrict__ out32, int ldo32,
                 __half* __restrict__ out16, int ldo16, int cache_rows) {
  extern __shared__ float4 cache[];  // [cache_rows][Q]
  __shared__ float red[40];
  const int b = blockIdx.y;
  const int cw = GN_GROUPS_PER_CTA * cpg;  // channels per CTA
  const int c0 = blockIdx.x * cw;
  const int Q = cw >> 2;                   // float4 columns per row
  const int rs = blockDim.x / Q;           // rows per sweep
  const int col4 = threadIdx.x % Q;
  const int rsub = threadIdx.x / Q;
  const bool active = rsub < rs;
  const int g = (col4 * 4) / cpg;          // group of this thread (static)
  const int T_ = utt_len[b];
  const long row0 = utt_off[b];
  const T* xb = x + row0 * ldx + c0 + col4 * 4;

  // pass 1: sum (and fill the cache)
  float s = 0.f;
  if (active)
    for (int r = rsub; r < T_; r += rs) {
      float4 v = load4<T>(xb + (long)r * ldx);
      if (r < cache_rows) cache[r * Q + col4] = v;
      s += (v.x + v.y) + (v.z + v.w);
    }
  float mean_g[GN_GROUPS_PER_CTA];
#pragma unroll
  for (int i = 0; i < GN_GROUPS_PER_CTA; ++i) mean_g[i] = block_sum((active && g == i) ? s : 0.f, red) / ((float)T_ * cpg);
  const float mean = mean_g[0] * (g == 0) + mean_g[1] * (g == 1) + mean_g[2] * (g == 2) + mean_g[3] * (g == 3);
  // pass 2: centred sum of squares
  float ss = 0.f;
  if (active)
    for (int r = rsub; r < T_; r += rs) {
      float4 v = r < cache_rows ? cache[r * Q + col4] : load4<T>(xb + (long)r * ldx);
      float a0 = v.x - mean, a1 = v.y - mean, a2 = v.z - mean, a3 = v.w - mean;
      ss += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
    }
  float rstd_g[GN_GROUPS_PER_CTA];
#pragma unroll
  for (int i = 0; i < GN_GROUPS_PER_CTA; ++i)
    rstd_g[i] = rsqrtf(block_sum((active && g == i) ? ss : 0.f, red) / ((float)T_ * cpg) + eps);
  const float rstd = rstd_g[0] * (g == 0) + rstd_g[1] * (g == 1) + rstd_g[2] * (g == 2) + rstd_g[3] * (g == 3);
  if (!active) return;
  // pass 3: normalise + affine (+FiLM) (+SiLU)
  const int c = c0 + col4 * 4;
  float ga[4], be[4], fs[4] = {0.f, 0.f, 0.f, 0.f}, fb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int q = 0; q < 4; ++q) { ga[q] = gamma[c + q]; be[q] = beta[c + q]; }
  if (film_scale) {
    const long fr = film_idx ? film_idx[b] : b;
#pragma unroll
    for (int q = 0; q < 4; ++q) { fs[q] = film_scale[fr * ld_film + c + q]; fb[q] = film_shift[fr * ld_film + c + q]; }
  }
  // fold: y = ((x-mean)*rstd*ga + be)*(1+fs) + fb = x*A + B
  float A_[4], B_[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    A_[q] = rstd * ga[q];
    B_[q] = be[q] - mean * A_[q];
  }
  for (int r = rsub; r < T_; r += rs) {
    float4 v = r < cache_rows ? cache[r * Q + col4] : load4<T>(xb + (long)r * ldx);
    float y[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float t = y[q] * A_[q] + B_[q];
      if (film_scale) t = t * (1.0f + fs[q]) + fb[q];
      if (act == DTTS_ACT_SILU) t = t * sigmoidf_(t);
      y[q] = t;
    }
    const long orow = row0 + r;
    if (out32) *reinterpret_cast<float4*>(out32 + orow * ldo32 + c) = make_float4(y[0], y[1], y[2], y[3]);
    if (out16) {
      __half2 h0 = __floats2half2_rn(y[0], y[1]), h1 = __floats2half2_rn(y[2], y[3]);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&h0);
      pk.y = *reinterpret_cast<uint32_t*>(&h1);
      *reinterpret_cast<uint2*>(out16 + orow * ldo16 + c) = pk;
    }
  }
}

// Register-resident variant for the diffusion hot loop (utterances up to GNR_MAXR * rows-per-sweep frames):
// one CTA = (utterance, GPC groups); every thread owns 4-channel vectors (16 B fp32 / 8 B fp16) of MAXR
// rows, all loads in flight at once (HBM-latency bound otherwise); the strip stays in registers in its raw
// storage type, statistics are exact two-pass (mean, then centred sum of squares), and the normalised operand
// is written straight from registers: x is read from HBM exactly once, nothing is staged in shared memory.
// fp32 input: 2 groups (48 ch = 192 B per row) per CTA; fp16 input: 4 groups (96 ch = 192 B per row).
constexpr int GNR_THREADS = 384;
constexpr int GNR_MAXG = 4;

// SiLU with the fast exp/rcp path (2 MUFU + 3 FMA-class ops; |rel err| ~ 2^-21, far below the fp16 operand rounding
// that follows).  The exact expf + IEEE division of sigmoidf_ made this HBM-bound kernel instruction-bound.
__device__ __forceinline__ float silu_fast(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

template <int G>
__device__ __forceinline__ void block_sum_g(float (&v)[G], float (*sh)[GNR_MAXG]) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < G; ++i)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < G; ++i) sh[w][i] = v[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < G; ++i) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < GNR_THREADS / 32; ++k) t += sh[k][i];
    v[i] = t;
  }
}

template <typename T> struct GnVec;
template <> struct GnVec<float> {
  static constexpr int N = 4;
  typedef uint4 Raw;
  static constexpr int MAXR = 11;   // 384 threads / 12 vector columns = 32 rows per sweep -> T <= 352
  __device__ static __forceinline__ Raw zero() { return make_uint4(0u, 0u, 0u, 0u); }
  __device__ static __forceinline__ void unpack(const uint4& u, float (&f)[4]) {
    f[0] = __uint_as_float(u.x); f[1] = __uint_as_float(u.y); f[2] = __uint_as_float(u.z); f[3] = __uint_as_float(u.w);
  }
};
template <> struct GnVec<__half> {
  static constexpr int N = 4;
  typedef uint2 Raw;
  static constexpr int MAXR = 18;   // 384 threads / 24 vector columns = 16 rows per sweep -> T <= 288
  __device__ static __forceinline__ Raw zero() { return make_uint2(0u, 0u); }
  __device__ static __forceinline__ void unpack(const uint2& u, float (&f)[4]) {
    // opaque moves: keep the strip packed in registers (the compiler would otherwise hoist the fp32 copies of all
    // rows across the three passes and spill)
    uint32_t ux, uy;
    asm volatile("mov.b32 %0, %1;" : "=r"(ux) : "r"(u.x));
    asm volatile("mov.b32 %0, %1;" : "=r"(uy) : "r"(u.y));
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&ux));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&uy));
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
  }
};

template <typename T, int GPC>
__global__ void __launch_bounds__(GNR_THREADS, 2)
groupnorm_reg_kernel(const T* __restrict__ x, int ldx, int cpg, const int* __restrict__ utt_off,
                     const int* __restrict__ utt_len, const float* __restrict__ gamma, const float* __restrict__ beta,
                     const float* __restrict__ film_scale, const float* __restrict__ film_shift, int ld_film,
                     const int* __restrict__ film_idx, int act, float eps, float* __restrict__ out32, int ldo32,
                     __half* __restrict__ out16, int ldo16) {
  constexpr int VE = GnVec<T>::N;
  __shared__ float red[GNR_THREADS / 32][GNR_MAXG];
  const int b = blockIdx.y;
  const int cw = GPC * cpg;
  const int c0 = blockIdx.x * cw;
  const int Q = cw / VE;
  const int rs = GNR_THREADS / Q;
  const int col = threadIdx.x % Q, rsub = threadIdx.x / Q;
  const bool active = rsub < rs;
  const int g = (col * VE) / cpg;
  const int T_ = utt_len[b];
  const long row0 = utt_off[b];
  const T* xb = x + row0 * ldx + c0 + col * VE;
  constexpr int GNR_MAXR = GnVec<T>::MAXR;
  typedef typename GnVec<T>::Raw Raw;
  Raw v[GNR_MAXR];
#pragma unroll
  for (int i = 0; i < GNR_MAXR; ++i) {
    const int r = rsub + i * rs;
    v[i] = (active && r < T_) ? __ldg(reinterpret_cast<const Raw*>(xb + (long)r * ldx)) : GnVec<T>::zero();
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < GNR_MAXR; ++i) {
    float f[VE];
    GnVec<T>::unpack(v[i], f);
#pragma unroll
    for (int q = 0; q < VE; q += 2) s += f[q] + f[q + 1];
  }
  const float inv_n = 1.0f / ((float)T_ * cpg);
  float sg[GPC];
#pragma unroll
  for (int i = 0; i < GPC; ++i) sg[i] = g == i ? s : 0.f;
  block_sum_g<GPC>(sg, red);
  float mean = 0.f;
#pragma unroll
  for (int i = 0; i < GPC; ++i) mean = g == i ? sg[i] * inv_n : mean;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < GNR_MAXR; ++i) {
    const int r = rsub + i * rs;
    if (active && r < T_) {
      float f[VE];
      GnVec<T>::unpack(v[i], f);
#pragma unroll
      for (int q = 0; q < VE; q += 2) {
        const float a0 = f[q] - mean, a1 = f[q + 1] - mean;
        ss += a0 * a0 + a1 * a1;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < GPC; ++i) sg[i] = g == i ? ss : 0.f;
  block_sum_g<GPC>(sg, red);
  float var = 0.f;
#pragma unroll
  for (int i = 0; i < GPC; ++i) var = g == i ? sg[i] * inv_n : var;
  const float rstd = rsqrtf(var + eps);
  if (!active) return;
  const int c = c0 + col * VE;
  float A_[VE], B_[VE];
#pragma unroll
  for (int q = 0; q < VE; q += 4) {
    const float4 ga = *reinterpret_cast<const float4*>(gamma + c + q), be = *reinterpret_cast<const float4*>(beta + c + q);
    A_[q] = rstd * ga.x; A_[q + 1] = rstd * ga.y; A_[q + 2] = rstd * ga.z; A_[q + 3] = rstd * ga.w;
    B_[q] = be.x - mean * A_[q]; B_[q + 1] = be.y - mean * A_[q + 1]; B_[q + 2] = be.z - mean * A_[q + 2]; B_[q + 3] = be.w - mean * A_[q + 3];
  }
  if (film_scale) {   // (x*A+B)*(1+fs)+fb = x*A(1+fs) + B(1+fs)+fb
    const long fr = film_idx ? film_idx[b] : b;
#pragma unroll
    for (int q = 0; q < VE; q += 4) {
      const float4 fs = *reinterpret_cast<const float4*>(film_scale + fr * ld_film + c + q);
      const float4 fb = *reinterpret_cast<const float4*>(film_shift + fr * ld_film + c + q);
      const float f1[4] = {1.f + fs.x, 1.f + fs.y, 1.f + fs.z, 1.f + fs.w}, f0[4] = {fb.x, fb.y, fb.z, fb.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) { A_[q + k] *= f1[k]; B_[q + k] = B_[q + k] * f1[k] + f0[k]; }
    }
  }
#pragma unroll
  for (int i = 0; i < GNR_MAXR; ++i) {
    const int r = rsub + i * rs;
    if (r >= T_) break;
    float y[VE];
    GnVec<T>::unpack(v[i], y);
#pragma unroll
    for (int q = 0; q < VE; ++q) y[q] = fmaf(y[q], A_[q], B_[q]);
    if (act == DTTS_ACT_SILU) {
#pragma unroll
      for (int q = 0; q < VE; ++q) y[q] = silu_fast(y[q]);
    }
    const long orow = row0 + r;
    if (out32) {
#pragma unroll
      for (int q = 0; q < VE; q += 4)
        *reinterpret_cast<float4*>(out32 + orow * ldo32 + c + q) = make_float4(y[q], y[q + 1], y[q + 2], y[q + 3]);
    }
    if (out16) {
      uint32_t pk[VE / 2];
#pragma unroll
      for (int q = 0; q < VE; q += 2) {
        __half2 h = __floats2half2_rn(y[q], y[q + 1]);
        pk[q / 2] = *reinterpret_cast<uint32_t*>(&h);
      }
      *reinterpret_cast<uint2*>(out16 + orow * ldo16 + c) = make_uint2(pk[0], pk[1]);
    }
  }
}

// TMA-staged persistent variant (the diffusion hot loop): the kernel is HBM-bound, so what matters is that a
// full strip is always in flight per SM.  One CTA loops over (utterance, channel strip) items; the strip
// [T, 192 bytes] is brought into shared memory by TMA (32-row boxes, mbarrier completion) one item AHEAD of the
// one being normalised, the three passes (sum, centred sum of squares, normalise + FiLM + SiLU) read shared
// memory, and the fp16 GEMM operand is written straight to HBM.  Rows land contiguously in shared memory, so
// thread i reads vector i, i + 384, ...: conflict-free and the channel column of a thread is fixed.
constexpr int GNT_BOX_ROWS = 32, GNT_STAGES = 2, GNT_ROW_BYTES = 192, GNT_MAX_SMEM = 224 * 1024;

template <typename T>
__global__ void __launch_bounds__(GNR_THREADS, 3)
groupnorm_tma_kernel(const __grid_constant__ CUtensorMap tm, int cpg, int gpc, int n_gx, int n_utt, int stage_rows, int n_stages,
                     const int* __restrict__ utt_off, const int* __restrict__ utt_len, const float* __restrict__ gamma,
                     const float* __restrict__ beta, const float* __restrict__ film_scale, const float* __restrict__ film_shift,
                     int ld_film, const int* __restrict__ film_idx, int act, float eps, float* __restrict__ out32, int ldo32,
                     __half* __restrict__ out16, int ldo16) {
  using namespace dtts_tc;
  typedef typename GnVec<T>::Raw Raw;
  extern __shared__ uint8_t gn_smem_raw[];
  uint8_t* smem = gn_smem_raw + ((128u - (smem_u32(gn_smem_raw) & 127u)) & 127u);
  uint64_t* full = (uint64_t*)smem;                        // [GNT_STAGES]
  uint8_t* bufs = smem + 128;
  __shared__ float red[GNR_THREADS / 32][GNR_MAXG];
  const int stage_bytes = stage_rows * GNT_ROW_BYTES;
  const int cw = gpc * cpg;                                // channels per strip: cw * sizeof(T) == 192
  const int Q = cw >> 2;                                   // 4-channel vectors per row (12 or 24)
  const int rs = GNR_THREADS / Q;                          // rows per sweep of the CTA
  const int col = threadIdx.x % Q, rsub = threadIdx.x / Q;
  const int g = (col * 4) / cpg;
  const int n_items = n_utt * n_gx;
  if (threadIdx.x == 0) {
    for (int s = 0; s < n_stages; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int item, int stage) {                  // thread 0 only
    const int b = item / n_gx, gx = item - b * n_gx;
    const int T_ = utt_len[b], row0 = utt_off[b];
    const int nbox = (T_ + GNT_BOX_ROWS - 1) / GNT_BOX_ROWS;
    uint8_t* dst = bufs + stage * stage_bytes;
    mbar_expect_tx(&full[stage], nbox * GNT_BOX_ROWS * GNT_ROW_BYTES);
    for (int i = 0; i < nbox; ++i)
      tma_load_2d(dst + i * GNT_BOX_ROWS * GNT_ROW_BYTES, &tm, &full[stage], gx * cw, row0 + i * GNT_BOX_ROWS);
  };
  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tm) : "memory");
    for (int s = 0; s < n_stages; ++s) {
      const int item = blockIdx.x + s * gridDim.x;
      if (item < n_items) issue(item, s);
    }
  }
  int it = 0;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
    const int stage = it % n_stages;
    const uint32_t phase = (it / n_stages) & 1;
    const int b = item / n_gx, gx = item - b * n_gx;
    const int T_ = utt_len[b];
    const long row0 = utt_off[b];
    const int c = gx * cw + col * 4;
    const int nvec = T_ * Q;
    const Raw* buf = reinterpret_cast<const Raw*>(bufs + stage * stage_bytes);
    mbar_wait(&full[stage], phase);
    float s = 0.f;
    for (int e = threadIdx.x; e < nvec; e += GNR_THREADS) {
      float f[4];
      GnVec<T>::unpack(buf[e], f);
      s += (f[0] + f[1]) + (f[2] + f[3]);
    }
    const float inv_n = 1.0f / ((float)T_ * cpg);
    float sg[GNR_MAXG];
#pragma unroll
    for (int i = 0; i < GNR_MAXG; ++i) sg[i] = g == i ? s : 0.f;
    block_sum_g<GNR_MAXG>(sg, red);
    float mean = 0.f;
#pragma unroll
    for (int i = 0; i < GNR_MAXG; ++i) mean = g == i ? sg[i] * inv_n : mean;
    float ss = 0.f;
    for (int e = threadIdx.x; e < nvec; e += GNR_THREADS) {
      float f[4];
      GnVec<T>::unpack(buf[e], f);
      const float a0 = f[0] - mean, a1 = f[1] - mean, a2 = f[2] - mean, a3 = f[3] - mean;
      ss += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
    }
#pragma unroll
    for (int i = 0; i < GNR_MAXG; ++i) sg[i] = g == i ? ss : 0.f;
    block_sum_g<GNR_MAXG>(sg, red);
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < GNR_MAXG; ++i) var = g == i ? sg[i] * inv_n : var;
    const float rstd = rsqrtf(var + eps);
    const float4 ga = *reinterpret_cast<const float4*>(gamma + c), be = *reinterpret_cast<const float4*>(beta + c);
    float A_[4] = {rstd * ga.x, rstd * ga.y, rstd * ga.z, rstd * ga.w};
    float B_[4] = {be.x - mean * A_[0], be.y - mean * A_[1], be.z - mean * A_[2], be.w - mean * A_[3]};
    if (film_scale) {   // (x*A+B)*(1+fs)+fb = x*A(1+fs) + B(1+fs)+fb
      const long fr = film_idx ? film_idx[b] : b;
      const float4 fs = *reinterpret_cast<const float4*>(film_scale + fr * ld_film + c);
      const float4 fb = *reinterpret_cast<const float4*>(film_shift + fr * ld_film + c);
      const float f1[4] = {1.f + fs.x, 1.f + fs.y, 1.f + fs.z, 1.f + fs.w}, f0[4] = {fb.x, fb.y, fb.z, fb.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) { A_[q] *= f1[q]; B_[q] = B_[q] * f1[q] + f0[q]; }
    }
    {
      float* o32 = out32 ? out32 + (row0 + rsub) * ldo32 + c : nullptr;
      __half* o16 = out16 ? out16 + (row0 + rsub) * ldo16 + c : nullptr;
      const long s32 = (long)rs * ldo32, s16 = (long)rs * ldo16;
      for (int e = threadIdx.x; e < nvec; e += GNR_THREADS) {
        float y[4];
        GnVec<T>::unpack(buf[e], y);
#pragma unroll
        for (int q = 0; q < 4; ++q) y[q] = fmaf(y[q], A_[q], B_[q]);
        if (act == DTTS_ACT_SILU) {
#pragma unroll
          for (int q = 0; q < 4; ++q) y[q] = silu_fast(y[q]);
        }
        if (o32) { *reinterpret_cast<float4*>(o32) = make_float4(y[0], y[1], y[2], y[3]); o32 += s32; }
        if (o16) {
          __half2 h0 = __floats2half2_rn(y[0], y[1]), h1 = __floats2half2_rn(y[2], y[3]);
          *reinterpret_cast<uint2*>(o16) = make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
          o16 += s16;
        }
      }
    }
    __syncthreads();                                         // every thread is done with this stage's buffer
    if (threadIdx.x == 0) {
      const int nxt = item + n_stages * gridDim.x;
      if (nxt < n_items) issue(nxt, stage);
    }
  }
}

// Cluster variant (default for the diffusion hot loop).  The strip kernels above read 192-byte segments at a
// 1.5-3 KB stride, which caps them near 4 TB/s (DRAM page locality), whatever the staging.  Here a thread-block
// cluster owns one utterance and every CTA of it streams a CONTIGUOUS block of whole rows (all channels) with 1-D
// bulk copies (cp.async.bulk, mbarrier completion) into shared memory; the per-group partial sums of the CTAs are
// combined through distributed shared memory (two exchanges: mean, then centred sum of squares = exact two-pass
// statistics, fixed summation order), and the normalised fp16 operand is written back as whole contiguous rows.
constexpr int GNC_THREADS = 384, GNC_MAX_GROUPS = 64, GNC_MAX_CLUSTER = 8, GNC_CHUNK = 32 * 1024;

template <typename T>
__global__ void __launch_bounds__(GNC_THREADS, 2)
groupnorm_cluster_kernel(const T* __restrict__ x, int C, int cpg, int groups, const int* __restrict__ utt_off,
                         const int* __restrict__ utt_len, const float* __restrict__ gamma, const float* __restrict__ beta,
                         const float* __restrict__ film_scale, const float* __restrict__ film_shift, int ld_film,
                         const int* __restrict__ film_idx, int act, float eps, float* __restrict__ out32, int ldo32,
                         __half* __restrict__ out16, int ldo16) {
  using namespace dtts_tc;
  typedef typename GnVec<T>::Raw Raw;
  extern __shared__ uint8_t gnc_smem_raw[];
  uint8_t* smem = gnc_smem_raw + ((128u - (smem_u32(gnc_smem_raw) & 127u)) & 127u);
  __shared__ float part[GNC_THREADS];                 // per-thread partials of the current pass
  __shared__ float gsum[2][GNC_MAX_GROUPS];           // this CTA's per-group partial sums: [0] sum, [1] centred sum of squares
  __shared__ float gstat[2][GNC_MAX_GROUPS];          // cluster-wide mean / rstd
  __shared__ __align__(8) uint64_t bar;
  uint32_t cs, rank;
  asm("mov.u32 %0, %%cluster_nctaid.x;" : "=r"(cs));
  asm("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int b = blockIdx.x / cs;
  const int T_ = utt_len[b];
  const long row0 = utt_off[b];
  const int rows_per = (T_ + (int)cs - 1) / (int)cs;
  const int r0 = min(T_, (int)rank * rows_per), r1 = min(T_, r0 + rows_per), n = r1 - r0;
  const int C4 = C >> 2;                               // 4-channel vectors per row; the launcher guarantees 384 % C4 == 0,
  const int rs = GNC_THREADS / C4;                     // so a thread's channel column (and group) is fixed
  const int col = threadIdx.x % C4;
  const int vpg = cpg >> 2;                            // vectors per group per row
  const int g = col / vpg;
  const long bytes = (long)n * C * sizeof(T);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(&bar, (uint32_t)bytes);
    const uint8_t* src = reinterpret_cast<const uint8_t*>(x + (row0 + r0) * C);
    for (long o = 0; o < bytes; o += GNC_CHUNK) {
      const uint32_t sz = (uint32_t)(bytes - o < GNC_CHUNK ? bytes - o : GNC_CHUNK);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_u32(smem + o)), "l"(src + o), "r"(sz), "r"(smem_u32(&bar)) : "memory");
    }
  }
  __syncthreads();
  const int nvec = n * C4;
  const Raw* buf = reinterpret_cast<const Raw*>(smem);
  const float inv_n = 1.0f / ((float)T_ * cpg);
  mbar_wait(&bar, 0);
  // ---- pass 1 / pass 2: per-group partial sums in a fixed order: thread -> shared -> one thread per group -> cluster (DSMEM)
  for (int pass = 0; pass < 2; ++pass) {
    const float mean = pass ? gstat[0][g] : 0.f;
    float s = 0.f;
    for (int e = threadIdx.x; e < nvec; e += GNC_THREADS) {
      float f[4];
      GnVec<T>::unpack(buf[e], f);
      if (pass == 0) s += (f[0] + f[1]) + (f[2] + f[3]);
      else {
        const float a0 = f[0] - mean, a1 = f[1] - mean, a2 = f[2] - mean, a3 = f[3] - mean;
        s += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
      }
    }
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x < groups) {
      float t = 0.f;
      for (int rr = 0; rr < rs; ++rr)
        for (int j = 0; j < vpg; ++j) t += part[rr * C4 + threadIdx.x * vpg + j];
      gsum[pass][threadIdx.x] = t;
    }
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (threadIdx.x < groups) {
      float t = 0.f;
      const uint32_t local = smem_u32(&gsum[pass][threadIdx.x]);
      for (uint32_t c = 0; c < cs; ++c) {
        uint32_t remote;
        float v;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(c));
        asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote));
        t += v;
      }
      gstat[pass][threadIdx.x] = pass ? rsqrtf(t * inv_n + eps) : t * inv_n;
    }
    __syncthreads();
  }
  // ---- pass 3: normalise + affine (+FiLM) (+SiLU), whole contiguous rows out
  {
    const int c = col * 4;
    const float mean = gstat[0][g], rstd = gstat[1][g];
    const float4 ga = *reinterpret_cast<const float4*>(gamma + c), be = *reinterpret_cast<const float4*>(beta + c);
    float A_[4] = {rstd * ga.x, rstd * ga.y, rstd * ga.z, rstd * ga.w};
    float B_[4] = {be.x - mean * A_[0], be.y - mean * A_[1], be.z - mean * A_[2], be.w - mean * A_[3]};
    if (film_scale) {   // (x*A+B)*(1+fs)+fb = x*A(1+fs) + B(1+fs)+fb
      const long fr = film_idx ? film_idx[b] : b;
      const float4 fs = *reinterpret_cast<const float4*>(film_scale + fr * ld_film + c);
      const float4 fb = *reinterpret_cast<const float4*>(film_shift + fr * ld_film + c);
      const float f1[4] = {1.f + fs.x, 1.f + fs.y, 1.f + fs.z, 1.f + fs.w}, f0[4] = {fb.x, fb.y, fb.z, fb.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) { A_[q] *= f1[q]; B_[q] = B_[q] * f1[q] + f0[q]; }
    }
    float* o32 = out32 ? out32 + (row0 + r0 + threadIdx.x / C4) * ldo32 + c : nullptr;
    __half* o16 = out16 ? out16 + (row0 + r0 + threadIdx.x / C4) * ldo16 + c : nullptr;
    const long s32 = (long)rs * ldo32, s16 = (long)rs * ldo16;
    for (int e = threadIdx.x; e < nvec; e += GNC_THREADS) {
      float y[4];
      GnVec<T>::unpack(buf[e], y);
#pragma unroll
      for (int q = 0; q < 4; ++q) y[q] = fmaf(y[q], A_[q], B_[q]);
      if (act == DTTS_ACT_SILU) {
#pragma unroll
        for (int q = 0; q < 4; ++q) y[q] = silu_fast(y[q]);
      }
      if (o32) { *reinterpret_cast<float4*>(o32) = make_float4(y[0], y[1], y[2], y[3]); o32 += s32; }
      if (o16) {
        __half2 h0 = __floats2half2_rn(y[0], y[1]), h1 = __floats2half2_rn(y[2], y[3]);
        *reinterpret_cast<uint2*>(o16) = make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
        o16 += s16;
      }
    }
  }
  // no CTA may exit while a peer can still read its gsum through DSMEM
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// Apply half of GroupNorm32 when the producing GEMM accumulated the statistics (dtts_gemm_params.gn_stats): one fully
// coalesced elementwise pass.  A CTA owns a 64-row tile (16 / 32 rows for small shards, so that the grid still fills the
// SMs; measured neutral, 14.5 -> 13.7 us per launch at 9 k rows: the pass is then a chain of three dependent loads) and ALL channels: thread -> fixed 8-channel column (its gamma /
// beta / FiLM coefficients stay in registers), rows strided over the row lanes; A, B are refreshed when the utterance
// of the row changes.  16-byte fp16 stores: 8 lanes write one full 128-byte line.
constexpr int GNA_ROWS = 64;

template <typename T>
__global__ void __launch_bounds__(384, sizeof(T) == 4 ? 2 : 3)
groupnorm_apply_kernel(const dtts_gn_apply_params p, const int rows_per_cta) {
  pdl_launch();                                  // (programmatic dependent launch inside the diffusion eval graph; no-ops otherwise)
  pdl_wait();
  const int C8 = p.C >> 3;                       // 8-channel vectors per row
  const int rl = blockDim.x / C8;                // row lanes
  const int col = threadIdx.x % C8, rlane = threadIdx.x / C8;
  if (rlane >= rl) return;
  const int c = col * 8, g = c / p.cpg, G = p.C / p.cpg;
  float A_[8], B_[8];
  int cur = -1;
  const int m_end = min(p.M, (int)(blockIdx.x + 1) * rows_per_cta);
  constexpr int NB = sizeof(T) == 4 ? 3 : 4;     // rows in flight per thread (the pass is pure streaming: latency is hidden by loads in flight)
  for (int base = blockIdx.x * rows_per_cta + rlane; base < m_end; base += NB * rl) {
    int us[NB];
    uint4 raw[NB][sizeof(T) == 4 ? 2 : 1];
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const int m = base + j * rl;
      us[j] = m < m_end ? __ldg(p.row_utt + m) : -1;
    }
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      if (us[j] < 0) continue;
      const size_t m = (size_t)(base + j * rl);
      if (sizeof(T) == 4) {
        raw[j][0] = __ldcs(reinterpret_cast<const uint4*>((const float*)p.x + m * p.ldx + c));
        raw[j][sizeof(T) == 4 ? 1 : 0] = __ldcs(reinterpret_cast<const uint4*>((const float*)p.x + m * p.ldx + c + 4));
      } else {
        raw[j][0] = __ldcs(reinterpret_cast<const uint4*>((const __half*)p.x + m * p.ldx + c));
      }
    }
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const int u = us[j];
      if (u < 0) continue;
      const size_t m = (size_t)(base + j * rl);
      if (u != cur) {
        cur = u;
        const float2 st = *reinterpret_cast<const float2*>(p.stats + ((size_t)u * G + g) * 2);
        const float inv_n = 1.0f / ((float)__ldg(p.utt_len + u) * p.cpg);
        const float mean = st.x * inv_n;
        const float rstd = rsqrtf(fmaxf(st.y * inv_n - mean * mean, 0.f) + p.eps);
#pragma unroll
        for (int q = 0; q < 8; q += 4) {           // (gamma / beta are re-read only when the utterance changes)
          const float4 ga = *reinterpret_cast<const float4*>(p.gamma + c + q), be = *reinterpret_cast<const float4*>(p.beta + c + q);
          A_[q] = rstd * ga.x; A_[q + 1] = rstd * ga.y; A_[q + 2] = rstd * ga.z; A_[q + 3] = rstd * ga.w;
          B_[q] = be.x - mean * A_[q]; B_[q + 1] = be.y - mean * A_[q + 1]; B_[q + 2] = be.z - mean * A_[q + 2]; B_[q + 3] = be.w - mean * A_[q + 3];
        }
        if (p.film_scale) {   // (x*A+B)*(1+fs)+fb
          const long fr = p.film_idx ? p.film_idx[u] : u;
#pragma unroll
          for (int q = 0; q < 8; q += 4) {
            const float4 fs = *reinterpret_cast<const float4*>(p.film_scale + fr * p.ld_film + c + q);
            const float4 fb = *reinterpret_cast<const float4*>(p.film_shift + fr * p.ld_film + c + q);
            const float f1[4] = {1.f + fs.x, 1.f + fs.y, 1.f + fs.z, 1.f + fs.w}, f0[4] = {fb.x, fb.y, fb.z, fb.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) { A_[q + k] *= f1[k]; B_[q + k] = B_[q + k] * f1[k] + f0[k]; }
          }
        }
      }
      float y[8];
      if (sizeof(T) == 4) {
        const uint4 v0 = raw[j][0], v1 = raw[j][sizeof(T) == 4 ? 1 : 0];
        y[0] = __uint_as_float(v0.x); y[1] = __uint_as_float(v0.y); y[2] = __uint_as_float(v0.z); y[3] = __uint_as_float(v0.w);
        y[4] = __uint_as_float(v1.x); y[5] = __uint_as_float(v1.y); y[6] = __uint_as_float(v1.z); y[7] = __uint_as_float(v1.w);
      } else {
        const uint32_t w[4] = {raw[j][0].x, raw[j][0].y, raw[j][0].z, raw[j][0].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[k]));
          y[2 * k] = t.x; y[2 * k + 1] = t.y;
        }
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) y[q] = fmaf(y[q], A_[q], B_[q]);
      if (p.act == DTTS_ACT_SILU) {
#pragma unroll
        for (int q = 0; q < 8; ++q) y[q] = silu_fast(y[q]);
      }
      if (p.out_f32) {
        *reinterpret_cast<float4*>(p.out_f32 + m * p.ldo32 + c) = make_float4(y[0], y[1], y[2], y[3]);
        *reinterpret_cast<float4*>(p.out_f32 + m * p.ldo32 + c + 4) = make_float4(y[4], y[5], y[6], y[7]);
      }
      if (p.out_f16) {
        __half2 h0 = __floats2half2_rn(y[0], y[1]), h1 = __floats2half2_rn(y[2], y[3]);
        __half2 h2 = __floats2half2_rn(y[4], y[5]), h3 = __floats2half2_rn(y[6], y[7]);
        uint4 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
        pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
        *reinterpret_cast<uint4*>((__half*)p.out_f16 + m * p.ldo16 + c) = pk;
      }
    }
  }
}

__global__ void __launch_bounds__(256) zero_f32_kernel(float* p, long n) {
  for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long)gridDim.x * 256) p[i] = 0.f;
}

constexpr int LN_MAXE = 32;  // C <= 1024
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, int ldx, int M, int C, const float* __restrict__ gamma,
                 const float* __restrict__ beta, float eps, const float* __restrict__ res, int ldr,
                 float* __restrict__ out32, int ldo32, __half* __restrict__ out16, int ldo16,
                 const int* __restrict__ row_utt) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  if (row_utt && row_utt[row] < 0) return;
  float v[LN_MAXE];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAXE; ++i) {
    const int c = lane + 32 * i;
    float t = 0.f;
    if (c < C) {
      t = x[(long)row * ldx + c];
      if (res) t += res[(long)row * ldr + c];
    }
    v[i] = t;
    s += t;
  }
  const float mean = warp_sum(s) / C;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAXE; ++i) {
    const int c = lane + 32 * i;
    if (c < C) { float d = v[i] - mean; ss += d * d; }
  }
  const float rstd = rsqrtf(warp_sum(ss) / C + eps);
#pragma unroll
  for (int i = 0; i < LN_MAXE; ++i) {
    const int c = lane + 32 * i;
    if (c < C) {
      float y = (v[i] - mean) * rstd * gamma[c] + beta[c];
      if (out32) out32[(long)row * ldo32 + c] = y;
      if (out16) out16[(long)row * ldo16 + c] = __float2half_rn(y);
    }
  }
}

}  // namespace

extern "C" int dtts_groupnorm(const dtts_groupnorm_params* p, void* stream) {
  DTTS_REQUIRE(p && p->x && p->gamma && p->beta && p->utt_off && p->utt_len, "groupnorm: null argument");
  DTTS_REQUIRE(p->groups > 0 && p->C % p->groups == 0, "groupnorm: C %% groups != 0");
  const int cpg = p->C / p->groups;
  DTTS_REQUIRE(p->groups % GN_GROUPS_PER_CTA == 0 && cpg % 4 == 0, "groupnorm: needs groups %% 4 == 0 and channels/group %% 4 == 0");
  const int Q = GN_GROUPS_PER_CTA * cpg / 4;
  DTTS_REQUIRE(Q <= 384, "groupnorm: channels per group too large");
  DTTS_REQUIRE(p->ldx % 4 == 0 && (!p->out_f32 || p->ldo32 % 4 == 0) && (!p->out_f16 || p->ldo16 % 4 == 0), "groupnorm: leading dims must be multiples of 4");
  DTTS_REQUIRE(p->out_f32 || p->out_f16, "groupnorm: no output");
  DTTS_REQUIRE(p->act == DTTS_ACT_NONE || p->act == DTTS_ACT_SILU, "groupnorm: unsupported activation");
  static int max_smem = 0;
  if (!max_smem) {
    max_smem = 160 * 1024;
    cudaFuncSetAttribute(groupnorm_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    cudaFuncSetAttribute(groupnorm_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
  }
  cudaStream_t st0 = (cudaStream_t)stream;
  {
    // cluster path (see groupnorm_cluster_kernel): contiguous rows, C/4 vector columns dividing the CTA size
    static int cl_on = -1;
    if (cl_on < 0) {
      const char* e = getenv("DTTS_GN_CLUSTER");
      cl_on = e ? atoi(e) : 0;   // measured on B200 (B=128, F=280): 106 / 111 us vs 82 / 78 us for the register path
      cudaError_t e1 = cudaFuncSetAttribute(groupnorm_cluster_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
      cudaError_t e2 = cudaFuncSetAttribute(groupnorm_cluster_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
      if (e1 != cudaSuccess || e2 != cudaSuccess) { cudaGetLastError(); cl_on = 0; }
    }
    const int es = p->x_is_f16 ? 2 : 4;
    const int C4 = p->C / 4;
    const int cs = GNC_MAX_CLUSTER;
    const int rows_per = (p->max_len + cs - 1) / cs;
    const size_t smem = 128 + (size_t)rows_per * p->C * es;
    const bool aligned = ((((uintptr_t)p->gamma) | ((uintptr_t)p->beta) | ((uintptr_t)p->x)) & 15) == 0 &&
                         (!p->film_scale || (((((uintptr_t)p->film_scale) | ((uintptr_t)p->film_shift)) & 15) == 0 && p->ld_film % 4 == 0)) &&
                         (!p->out_f32 || ((((uintptr_t)p->out_f32) & 15) == 0 && p->ldo32 % 4 == 0)) &&
                         (!p->out_f16 || ((((uintptr_t)p->out_f16) & 7) == 0 && p->ldo16 % 4 == 0));
    if (cl_on && p->max_len > 0 && p->ldx == p->C && p->C % 4 == 0 && C4 <= GNC_THREADS && GNC_THREADS % C4 == 0 && cpg % 4 == 0 &&
        p->groups <= GNC_MAX_GROUPS && (p->C * es) % 16 == 0 && aligned && smem <= 112 * 1024) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)(cs * p->n_utt));
      cfg.blockDim = dim3(GNC_THREADS);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = st0;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      cudaError_t le;
      if (p->x_is_f16)
        le = cudaLaunchKernelEx(&cfg, groupnorm_cluster_kernel<__half>, (const __half*)p->x, p->C, cpg, p->groups, p->utt_off, p->utt_len,
                                p->gamma, p->beta, p->film_scale, p->film_shift, p->ld_film, p->film_idx, p->act, p->eps, p->out_f32,
                                p->ldo32, (__half*)p->out_f16, p->ldo16);
      else
        le = cudaLaunchKernelEx(&cfg, groupnorm_cluster_kernel<float>, (const float*)p->x, p->C, cpg, p->groups, p->utt_off, p->utt_len,
                                p->gamma, p->beta, p->film_scale, p->film_shift, p->ld_film, p->film_idx, p->act, p->eps, p->out_f32,
                                p->ldo32, (__half*)p->out_f16, p->ldo16);
      if (le != cudaSuccess) DTTS_FAIL(-3, "groupnorm_cluster launch failed: %s", cudaGetErrorString(le));
      DTTS_CHECK_LAUNCH("groupnorm_cluster");
      return 0;
    }
  }
  {
    // TMA-staged persistent path (see groupnorm_tma_kernel): strips of exactly 192 bytes per row
    const int es = p->x_is_f16 ? 2 : 4;
    const int gpc = cpg * es <= GNT_ROW_BYTES ? GNT_ROW_BYTES / (cpg * es) : 0;
    const int stage_rows = (p->max_len + GNT_BOX_ROWS - 1) / GNT_BOX_ROWS * GNT_BOX_ROWS;
    static int tma_on = -1, sms = 0, n_stages = 1;
    const size_t smem = 128 + 128 + (size_t)n_stages * stage_rows * GNT_ROW_BYTES;
    if (tma_on < 0) {
      const char* e = getenv("DTTS_GN_TMA");
      tma_on = e ? atoi(e) : 0;   // measured on B200 (B=128, F=280): 90.7 / 81.3 us vs 84.7 / 80.2 us for the register path
      const char* e3 = getenv("DTTS_GN_TMA_STAGES");
      n_stages = e3 ? atoi(e3) : 1;
      if (n_stages < 1 || n_stages > GNT_STAGES) n_stages = 1;
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      // (static shared memory counts against the 227 KB limit too)
      cudaError_t e1 = cudaFuncSetAttribute(groupnorm_tma_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, GNT_MAX_SMEM);
      cudaError_t e2 = cudaFuncSetAttribute(groupnorm_tma_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, GNT_MAX_SMEM);
      if (e1 != cudaSuccess || e2 != cudaSuccess) { cudaGetLastError(); tma_on = 0; }
    }
    const bool aligned = ((((uintptr_t)p->gamma) | ((uintptr_t)p->beta) | ((uintptr_t)p->x)) & 15) == 0 && (p->ldx * es) % 16 == 0 &&
                         (!p->film_scale || (((((uintptr_t)p->film_scale) | ((uintptr_t)p->film_shift)) & 15) == 0 && p->ld_film % 4 == 0)) &&
                         (!p->out_f32 || ((((uintptr_t)p->out_f32) & 15) == 0 && p->ldo32 % 4 == 0)) &&
                         (!p->out_f16 || ((((uintptr_t)p->out_f16) & 7) == 0 && p->ldo16 % 4 == 0));
    if (tma_on && sms > 0 && p->n_rows > 0 && p->max_len > 0 && gpc >= 1 && gpc <= GNR_MAXG && gpc * cpg * es == GNT_ROW_BYTES &&
        p->groups % gpc == 0 && cpg % 4 == 0 && aligned && smem <= (size_t)GNT_MAX_SMEM) {
      CUtensorMap tm;
      const int cw = gpc * cpg;
      int rc = dtts_tc::get_map_ex(p->x, p->n_rows, p->C, p->ldx, GNT_BOX_ROWS, cw, &tm, es);
      if (rc) return rc;
      const int n_gx = p->groups / gpc;
      const long items = (long)p->n_utt * n_gx;
      int per_sm = (int)((227 * 1024) / (smem + 1024));
      per_sm = per_sm > 3 ? 3 : per_sm < 1 ? 1 : per_sm;
      const int grid = items < (long)per_sm * sms ? (int)items : per_sm * sms;
      if (p->x_is_f16)
        groupnorm_tma_kernel<__half><<<grid, GNR_THREADS, smem, st0>>>(tm, cpg, gpc, n_gx, p->n_utt, stage_rows, n_stages, p->utt_off, p->utt_len,
            p->gamma, p->beta, p->film_scale, p->film_shift, p->ld_film, p->film_idx, p->act, p->eps, p->out_f32, p->ldo32,
            (__half*)p->out_f16, p->ldo16);
      else
        groupnorm_tma_kernel<float><<<grid, GNR_THREADS, smem, st0>>>(tm, cpg, gpc, n_gx, p->n_utt, stage_rows, n_stages, p->utt_off, p->utt_len,
            p->gamma, p->beta, p->film_scale, p->film_shift, p->ld_film, p->film_idx, p->act, p->eps, p->out_f32, p->ldo32,
            (__half*)p->out_f16, p->ldo16);
      DTTS_CHECK_LAUNCH("groupnorm_tma");
      return 0;
    }
  }
  {
    // register-resident fast path (see groupnorm_reg_kernel)
    const int ve = 4, gpc = p->x_is_f16 ? 4 : 2, maxr = p->x_is_f16 ? GnVec<__half>::MAXR : GnVec<float>::MAXR;
    const int Qr = gpc * cpg / ve;
    const int rs = Qr > 0 ? GNR_THREADS / Qr : 0;
    const bool aligned = ((((uintptr_t)p->gamma) | ((uintptr_t)p->beta) | ((uintptr_t)p->x)) & 15) == 0 && p->ldx % ve == 0 &&
                         (!p->film_scale || (((((uintptr_t)p->film_scale) | ((uintptr_t)p->film_shift)) & 15) == 0 && p->ld_film % 4 == 0)) &&
                         (!p->out_f32 || (((uintptr_t)p->out_f32) & 15) == 0) &&
                         (!p->out_f16 || ((((uintptr_t)p->out_f16) & 7) == 0 && p->ldo16 % 4 == 0));
    if (p->groups % gpc == 0 && rs > 0 && p->max_len > 0 && p->max_len <= maxr * rs && aligned && cpg % ve == 0 && cpg <= 32) {
      dim3 grid(p->groups / gpc, p->n_utt);
      if (p->x_is_f16)
        groupnorm_reg_kernel<__half, 4><<<grid, GNR_THREADS, 0, st0>>>(
            (const __half*)p->x, p->ldx, cpg, p->utt_off, p->utt_len, p->gamma, p->beta, p->film_scale, p->film_shift,
            p->ld_film, p->film_idx, p->act, p->eps, p->out_f32, p->ldo32, (__half*)p->out_f16, p->ldo16);
      else
        groupnorm_reg_kernel<float, 2><<<grid, GNR_THREADS, 0, st0>>>(
            (const float*)p->x, p->ldx, cpg, p->utt_off, p->utt_len, p->gamma, p->beta, p->film_scale, p->film_shift,
            p->ld_film, p->film_idx, p->act, p->eps, p->out_f32, p->ldo32, (__half*)p->out_f16, p->ldo16);
      DTTS_CHECK_LAUNCH("groupnorm_reg");
      return 0;
    }
  }
  // cache as many rows as fit; the caller passes the longest utterance via utt_len on device, so
  // size for the budget (rows beyond cache_rows are re-read from L2).
  const int row_bytes = Q * 16;
  int cache_rows = max_smem / row_bytes;
  if (p->max_len > 0 && p->max_len < cache_rows) cache_rows = p->max_len;
  dim3 grid(p->groups / GN_GROUPS_PER_CTA, p->n_utt);
  const int threads = (384 / Q) * Q;
  cudaStream_t st = (cudaStream_t)stream;
  if (p->x_is_f16)
    groupnorm_kernel<__half><<<grid, threads, cache_rows * row_bytes, st>>>(
        (const __half*)p->x, p->ldx, p->C, cpg, p->utt_off, p->utt_len, p->gamma, p->beta, p->film_scale, p->film_shift,
        p->ld_film, p->film_idx, p->act, p->eps, p->out_f32, p->ldo32, (__half*)p->out_f16, p->ldo16, cache_rows);
  else
    groupnorm_kernel<float><<<grid, threads, cache_rows * row_bytes, st>>>(
        (const float*)p->x, p->ldx, p->C, cpg, p->utt_off, p->utt_len, p->gamma, p->beta, p->film_scale, p->film_shift,
        p->ld_film, p->film_idx, p->act, p->eps, p->out_f32, p->ldo32, (__half*)p->out_f16, p->ldo16, cache_rows);
  DTTS_CHECK_LAUNCH("groupnorm");
  return 0;
}

extern "C" int dtts_groupnorm_apply(const dtts_gn_apply_params* p, void* stream) {
  DTTS_REQUIRE(p && p->x && p->row_utt && p->utt_len && p->stats && p->gamma && p->beta, "groupnorm_apply: null argument");
  DTTS_REQUIRE(p->C % 8 == 0 && p->cpg > 0 && p->cpg % 8 == 0 && p->C % p->cpg == 0 && p->C / 8 <= 384, "groupnorm_apply: needs C %% 8 == 0, cpg %% 8 == 0, C <= 3072");
  DTTS_REQUIRE(p->out_f32 || p->out_f16, "groupnorm_apply: no output");
  DTTS_REQUIRE(p->act == DTTS_ACT_NONE || p->act == DTTS_ACT_SILU, "groupnorm_apply: unsupported activation");
  const int es = p->x_is_f16 ? 2 : 4;
  DTTS_REQUIRE((((uintptr_t)p->x) & 15) == 0 && (p->ldx * es) % 16 == 0 && ((((uintptr_t)p->gamma) | ((uintptr_t)p->beta) | ((uintptr_t)p->stats)) & 15) % 8 == 0 &&
               ((((uintptr_t)p->gamma) | ((uintptr_t)p->beta)) & 15) == 0, "groupnorm_apply: operands must be 16-byte aligned");
  DTTS_REQUIRE(!p->film_scale || (((((uintptr_t)p->film_scale) | ((uintptr_t)p->film_shift)) & 15) == 0 && p->ld_film % 4 == 0), "groupnorm_apply: FiLM rows must be 16-byte aligned");
  DTTS_REQUIRE(!p->out_f16 || ((((uintptr_t)p->out_f16) & 15) == 0 && p->ldo16 % 8 == 0), "groupnorm_apply: fp16 output must be 16-byte aligned");
  DTTS_REQUIRE(!p->out_f32 || ((((uintptr_t)p->out_f32) & 15) == 0 && p->ldo32 % 4 == 0), "groupnorm_apply: fp32 output must be 16-byte aligned");
  if (p->M <= 0) return 0;
  const int C8 = p->C / 8;
  const int threads = (384 / C8) * C8;
  int rows = GNA_ROWS;       // smaller row tiles while the grid would leave SMs (2-3 resident CTAs each) idle
  while (rows > 16 && ceil_div(p->M, rows) < 2 * 148 * (p->x_is_f16 ? 3 : 2)) rows >>= 1;
  const int grid = ceil_div(p->M, rows);
  if (p->x_is_f16) launch_maybe_pdl(groupnorm_apply_kernel<__half>, dim3(grid), dim3(threads), 0, (cudaStream_t)stream, *p, rows);
  else launch_maybe_pdl(groupnorm_apply_kernel<float>, dim3(grid), dim3(threads), 0, (cudaStream_t)stream, *p, rows);
  DTTS_CHECK_LAUNCH("groupnorm_apply");
  return 0;
}

extern "C" int dtts_zero_f32(const dtts_zero_params* p, void* stream) {
  DTTS_REQUIRE(p && p->ptr && p->n >= 0, "zero_f32: bad argument");
  if (p->n == 0) return 0;
  long g = (p->n + 255) / 256;
  if (g > 1184) g = 1184;
  zero_f32_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>(p->ptr, (long)p->n);
  DTTS_CHECK_LAUNCH("zero_f32");
  return 0;
}

extern "C" int dtts_layernorm(const dtts_layernorm_params* p, void* stream) {
  DTTS_REQUIRE(p && p->x && p->gamma && p->beta, "layernorm: null argument");
  DTTS_REQUIRE(p->C > 0 && p->C <= 32 * LN_MAXE, "layernorm: C out of range");
  DTTS_REQUIRE(p->out_f32 || p->out_f16, "layernorm: no output");
  if (p->M <= 0) return 0;
  const int rows_per_cta = 8;
  layernorm_kernel<<<ceil_div(p->M, rows_per_cta), rows_per_cta * 32, 0, (cudaStream_t)stream>>>(
      p->x, p->ldx, p->M, p->C, p->gamma, p->beta, p->eps, p->res, p->ldr, p->out_f32, p->ldo32, (__half*)p->out_f16, p->ldo16, p->row_utt);
  DTTS_CHECK_LAUNCH("layernorm");
  return 0;
}
