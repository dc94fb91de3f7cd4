// gemm_simt.cu -- exact-fp32 CUDA-core GEMM / multi-tap conv-GEMM with the shared epilogue.
// Same contract as dtts_gemm_f16_tc (include/dtts.h) in fp32 FMA arithmetic: used where the
// reference's fp32 numerics must be reproduced closely enough for token-exact sampling (GPT
// trunk, gpt/model.py:149-173) and for the small one-off encoders.
#include "common.cuh"

namespace {

constexpr int TM = 64, TN = 64, TK = 16;

// 256 threads, each computes a 4x4 micro-tile. As[k][m], Ws[k][n] (transposed in smem).
__global__ void __launch_bounds__(256)
gemm_f32_kernel(const float* __restrict__ A, const float* __restrict__ W, int K, int lda, int ldw,
                int taps, int shift0, int tap_stride, const EpiParams epi) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Ws[TK][TN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int tx = tid & 15, ty = tid >> 4;  // tx -> n, ty -> m
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  // loaders: 64 rows x 16 k = 1024 elements / 256 threads = 4 each (one float4 along k when aligned)
  const int lr = tid >> 2;        // 0..63 row within tile
  const int lk = (tid & 3) * 4;   // 0,4,8,12
  const bool a_vec = ((lda & 3) == 0) && ((((uintptr_t)A) & 15) == 0);
  const bool w_vec = ((ldw & 3) == 0) && ((((uintptr_t)W) & 15) == 0);

  for (int tap = 0; tap < taps; ++tap) {
    const int arow = m0 + lr + shift0 + tap * tap_stride;
    const bool arow_ok = arow >= 0 && arow < epi.M;
    const int wrow = n0 + lr;
    const bool wrow_ok = wrow < epi.N;
    const float* ap = A + (long)arow * lda;
    const float* wp = W + ((long)tap * epi.N + wrow) * ldw;
    for (int k0 = 0; k0 < K; k0 += TK) {
      float av[4] = {0.f, 0.f, 0.f, 0.f}, wv[4] = {0.f, 0.f, 0.f, 0.f};
      const int k = k0 + lk;
      if (arow_ok) {
        if (a_vec && k + 4 <= K) {
          float4 t = *reinterpret_cast<const float4*>(ap + k);
          av[0] = t.x; av[1] = t.y; av[2] = t.z; av[3] = t.w;
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) if (k + q < K) av[q] = ap[k + q];
        }
      }
      if (wrow_ok) {
        if (w_vec && k + 4 <= K) {
          float4 t = *reinterpret_cast<const float4*>(wp + k);
          wv[0] = t.x; wv[1] = t.y; wv[2] = t.z; wv[3] = t.w;
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) if (k + q < K) wv[q] = wp[k + q];
        }
      }
      __syncthreads();
#pragma unroll
      for (int q = 0; q < 4; ++q) { As[lk + q][lr] = av[q]; Ws[lk + q][lr] = wv[q]; }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < TK; ++kk) {
        float a[4], b[4];
        *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
        *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float v[4] = {acc[i][0], acc[i][1], acc[i][2], acc[i][3]};
    epilogue_chunk<4>(epi, m0 + ty * 4 + i, n0 + tx * 4, v);
  }
}

// Small-M variant (GPT decode step, M = utterances): one CTA per 8 output columns x all rows
// would starve the SMs on K; instead each CTA owns 16 columns and 32 rows and the grid covers
// N/16 x M/32 -> 4x more CTAs streaming the weight matrix.
constexpr int SM_ = 32, SN_ = 16, SK_ = 32;
__global__ void __launch_bounds__(128)
gemm_f32_smallm_kernel(const float* __restrict__ A, const float* __restrict__ W, int K, int lda, int ldw,
                       const EpiParams epi) {
  __shared__ float As[SK_][SM_ + 1];
  __shared__ float Ws[SK_][SN_ + 1];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * SM_, n0 = blockIdx.x * SN_;
  const int tx = tid & 3, ty = tid >> 2;  // tx -> 4 columns each (16), ty -> 1 row each (32)
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k0 = 0; k0 < K; k0 += SK_) {
    __syncthreads();
    // A tile 32x32: 1024 elements / 128 threads = 8
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int idx = tid + i * 128;
      int r = idx >> 5, k = idx & 31;
      float v = 0.f;
      if (m0 + r < epi.M && k0 + k < K) v = A[(long)(m0 + r) * lda + k0 + k];
      As[k][r] = v;
    }
    // W tile 16x32: 512 / 128 = 4
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int idx = tid + i * 128;
      int r = idx >> 5, k = idx & 31;
      float v = 0.f;
      if (n0 + r < epi.N && k0 + k < K) v = W[(long)(n0 + r) * ldw + k0 + k];
      Ws[k][r] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SK_; ++kk) {
      float a = As[kk][ty];
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] = fmaf(a, Ws[kk][tx * 4 + j], acc[j]);
    }
  }
  epilogue_chunk<4>(epi, m0 + ty, n0 + tx * 4, acc);
}

}  // namespace

extern "C" int dtts_gemm_f32(const dtts_gemm_params* p, void* stream) {
  DTTS_REQUIRE(p && p->A && p->W, "gemm_f32: null operand");
  DTTS_REQUIRE(p->M > 0 && p->N > 0 && p->K > 0 && p->taps >= 1, "gemm_f32: bad shape");
  DTTS_REQUIRE(!(p->bias_utt && !p->row_utt), "gemm_f32: bias_utt requires row_utt");
  DTTS_REQUIRE(p->out_f32 || p->out_f16, "gemm_f32: no output");
  DTTS_REQUIRE(!(p->act >= DTTS_ACT_PAIR_TANH_SIGMOID && (p->N & 1)), "gemm_f32: pair activation needs even N");
  EpiParams e = make_epi(p);
  cudaStream_t st = (cudaStream_t)stream;
  if (p->taps == 1 && p->tap_shift0 == 0 && p->M <= 256) {
    dim3 grid(ceil_div(p->N, SN_), ceil_div(p->M, SM_));
    gemm_f32_smallm_kernel<<<grid, 128, 0, st>>>((const float*)p->A, (const float*)p->W, p->K, p->lda, p->ldw, e);
    DTTS_CHECK_LAUNCH("gemm_f32_smallm");
    return 0;
  }
  dim3 grid(ceil_div(p->N, TN), ceil_div(p->M, TM));
  gemm_f32_kernel<<<grid, 256, 0, st>>>((const float*)p->A, (const float*)p->W, p->K, p->lda, p->ldw, p->taps,
                                        p->tap_shift0, p->tap_stride, e);
  DTTS_CHECK_LAUNCH("gemm_f32");
  return 0;
}
