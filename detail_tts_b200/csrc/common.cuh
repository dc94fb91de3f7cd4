// common.cuh -- shared device/host helpers for the dtts kernels (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <utility>

#include "../../include/dtts.h"

// ---------------------------------------------------------------------------------------------
// host: error reporting + launch accounting
// ---------------------------------------------------------------------------------------------
extern thread_local char g_dtts_err[512];
extern int g_dtts_launches;

#define DTTS_FAIL(code, ...)                                   \
  do {                                                         \
    snprintf(g_dtts_err, sizeof(g_dtts_err), __VA_ARGS__);     \
    return (code);                                             \
  } while (0)

#define DTTS_REQUIRE(cond, ...)                                \
  do {                                                         \
    if (!(cond)) DTTS_FAIL(-2, __VA_ARGS__);                   \
  } while (0)

#define DTTS_CHECK_LAUNCH(name)                                                          \
  do {                                                                                   \
    cudaError_t e__ = cudaGetLastError();                                                \
    if (e__ != cudaSuccess) DTTS_FAIL(-3, "%s launch failed: %s", name, cudaGetErrorString(e__)); \
    ++g_dtts_launches;                                                                   \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------------------------
// programmatic dependent launch (PDL) for the latency-bound GPT decode step: the ~84 small kernels of a
// step are launched with cudaLaunchAttributeProgrammaticStreamSerialization while dtts_set_pdl(1) is in
// effect, so kernel N+1 is scheduled (and runs its prologue) while kernel N drains.  Every such kernel
// calls pdl_launch() first and pdl_wait() before it touches memory written by its predecessor; both are
// no-ops for a normal launch.
// ---------------------------------------------------------------------------------------------
extern int g_dtts_pdl;
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_maybe_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_dtts_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
#endif

// ---------------------------------------------------------------------------------------------
// device: activations
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float act_apply(int act, float x, float p) {
  switch (act) {
    case DTTS_ACT_GELU_NEW: {
      float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
      return 0.5f * x * (1.0f + tanhf(u));
    }
    case DTTS_ACT_RELU: return fmaxf(x, 0.0f);
    case DTTS_ACT_SILU: return x * sigmoidf_(x);
    case DTTS_ACT_MISH: {
      float sp = x > 20.0f ? x : log1pf(expf(x));
      return x * tanhf(sp);
    }
    case DTTS_ACT_LRELU: return x > 0.0f ? x : x * p;
    case DTTS_ACT_TANH: return tanhf(x);
    case DTTS_ACT_LOG_CLAMP: return logf(fmaxf(x, p));
    default: return x;
  }
}

__device__ __forceinline__ float act_pair(int act, float a, float b) {
  if (act == DTTS_ACT_PAIR_TANH_SIGMOID) return tanhf(a) * sigmoidf_(b);
  return a * sigmoidf_(b);  // DTTS_ACT_PAIR_GLU
}

// ---------------------------------------------------------------------------------------------
// device: the GEMM epilogue shared by the tcgen05 and the fp32 CUDA-core GEMMs
// ---------------------------------------------------------------------------------------------
struct EpiParams {
  int M, N;
  const float* bias;
  const float* bias_utt;
  const int* row_utt;
  const int* out_row_map;
  const float* res;
  float* out_f32;
  __half* out_f16;
  int ldr, ldo32, ldo16;
  int act, act16;
  float act_param, act16_param, alpha;
  int accumulate;
  float* gn_stats;   // [n_utt, N/gn_cpg, 2] or nullptr
  int gn_cpg;
};

static inline EpiParams make_epi(const dtts_gemm_params* p) {
  EpiParams e;
  e.M = p->M; e.N = p->N;
  e.bias = p->bias; e.bias_utt = p->bias_utt; e.row_utt = p->row_utt; e.out_row_map = p->out_row_map;
  e.res = p->res; e.out_f32 = p->out_f32; e.out_f16 = (__half*)p->out_f16;
  e.ldr = p->ldr; e.ldo32 = p->ldo32; e.ldo16 = p->ldo16;
  e.act = p->act; e.act16 = p->act16; e.act_param = p->act_param; e.act16_param = p->act16_param;
  e.alpha = p->alpha; e.accumulate = p->accumulate;
  e.gn_stats = p->gn_stats; e.gn_cpg = p->gn_cpg;
  return e;
}

// v[0..NV) are accumulator values of row m, columns n0..n0+NV (NV multiple of 4, n0 multiple of 4).
template <int NV>
__device__ __forceinline__ void epilogue_chunk(const EpiParams& e, int m, int n0, float (&v)[NV]) {
  if (m >= e.M || n0 >= e.N) return;
  int u = 0;
  if (e.row_utt) {
    u = e.row_utt[m];
    if (u < 0) return;  // separator row: stays zero
  }
  const long orow = e.out_row_map ? (long)e.out_row_map[m] : (long)m;
  const bool pair = e.act >= DTTS_ACT_PAIR_TANH_SIGMOID;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    int n = n0 + j;
    if (n < e.N) {
      float b = e.bias ? __ldg(e.bias + n) : 0.0f;
      if (e.bias_utt) b += __ldg(e.bias_utt + (long)u * e.N + n);
      v[j] += b;
    }
  }
  const int nout_total = pair ? e.N / 2 : e.N;
  const int no0 = pair ? n0 / 2 : n0;
  constexpr int NO = NV;  // upper bound on outputs
  float o[NO];
  const int nvalid_out = pair ? NV / 2 : NV;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    if (pair) {
      if (j < NV / 2) o[j] = act_pair(e.act, v[2 * j], v[2 * j + 1]);
    } else {
      o[j] = act_apply(e.act, v[j], e.act_param);
    }
  }
#pragma unroll
  for (int j0 = 0; j0 < NV; j0 += 4) {
    if (j0 >= nvalid_out) break;
    const int no = no0 + j0;
    if (no >= nout_total) break;
    const bool full4 = (no + 4 <= nout_total) && (j0 + 4 <= nvalid_out);
    float r[4] = {0.f, 0.f, 0.f, 0.f};
    if (e.res) {
      const float* rp = e.res + orow * e.ldr + no;
      if (full4 && ((e.ldr & 3) == 0) && ((((uintptr_t)e.res) & 15) == 0)) {
        float4 t = *reinterpret_cast<const float4*>(rp);
        r[0] = t.x; r[1] = t.y; r[2] = t.z; r[3] = t.w;
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (no + q < nout_total && j0 + q < nvalid_out) r[q] = rp[q];
      }
    }
    float w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) w[q] = e.alpha * (o[j0 + q] + r[q]);
    if (e.out_f32) {
      float* op = e.out_f32 + orow * e.ldo32 + no;
      const bool vec = full4 && ((e.ldo32 & 3) == 0) && ((((uintptr_t)e.out_f32) & 15) == 0);
      if (e.accumulate) {
        if (vec) {
          float4 t = *reinterpret_cast<const float4*>(op);
          w[0] += t.x; w[1] += t.y; w[2] += t.z; w[3] += t.w;
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (no + q < nout_total && j0 + q < nvalid_out) w[q] += op[q];
        }
      }
      if (vec) {
        *reinterpret_cast<float4*>(op) = make_float4(w[0], w[1], w[2], w[3]);
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (no + q < nout_total && j0 + q < nvalid_out) op[q] = w[q];
      }
    }
    if (e.out_f16) {
      __half* hp = e.out_f16 + orow * e.ldo16 + no;
      float h[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) h[q] = act_apply(e.act16, w[q], e.act16_param);
      if (full4 && ((e.ldo16 & 3) == 0) && ((((uintptr_t)e.out_f16) & 7) == 0)) {
        __half2 a = __floats2half2_rn(h[0], h[1]);
        __half2 b = __floats2half2_rn(h[2], h[3]);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&a);
        pk.y = *reinterpret_cast<uint32_t*>(&b);
        *reinterpret_cast<uint2*>(hp) = pk;
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (no + q < nout_total && j0 + q < nvalid_out) hp[q] = __float2half_rn(h[q]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// device: reductions
// ---------------------------------------------------------------------------------------------
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
// block-wide sum; `sh` must hold >= 33 floats; all threads get the result.
__device__ __forceinline__ float block_sum(float v, float* sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  if (w == 0) {
    float t = lane < nw ? sh[lane] : 0.0f;
    t = warp_sum(t);
    if (lane == 0) sh[32] = t;
  }
  __syncthreads();
  return sh[32];
}
__device__ __forceinline__ float block_max(float v, float* sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  if (w == 0) {
    float t = lane < nw ? sh[lane] : -INFINITY;
    t = warp_max(t);
    if (lane == 0) sh[32] = t;
  }
  __syncthreads();
  return sh[32];
}
