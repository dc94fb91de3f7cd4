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

// Small-M variant (GPT decode step, M = utterances <= a few hundred): the step is a weight-streaming
// GEMM that has to keep every SM pulling weights.  CTA = 64 rows x 8 output columns (N/8 x M/64 CTAs:
// 96..384 per GEMM of the trunk), 128 threads = (row, K-parity): a thread accumulates its row against the
// 8 weight columns over the even or odd k of each 32-wide K chunk (8 independent FMA chains), chunks are
// double-buffered through registers, and the two K-parities are combined with one shuffle in a FIXED
// order, so results are bit-reproducible run to run (no atomics / split-K races).
constexpr int RP_ROWS = 64, RP_COLS = 8, RP_K = 32;
__global__ void __launch_bounds__(128)
gemm_f32_rowpar_kernel(const float* __restrict__ A, const float* __restrict__ W, int K, int lda, int ldw,
                       const EpiParams epi) {
  __shared__ float As[RP_ROWS][RP_K + 2];   // stride 34: (row, K-parity) lanes hit 32 distinct banks
  __shared__ __align__(16) float Ws[RP_K][RP_COLS];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * RP_ROWS, n0 = blockIdx.x * RP_COLS;
  const int r = tid >> 1, kh = tid & 1;
  // loaders: A chunk 64x32 floats = 512 float4 -> 4 per thread (row = i*16 + tid/8, k4 = tid%8);
  //          W chunk 8x32 floats = 64 float4 -> threads 0..63 (col = tid/8, k4 = tid%8)
  const int lrow = tid >> 3, lk = (tid & 7) * 4;
  const bool vecA = ((lda & 3) == 0) && ((((uintptr_t)A) & 15) == 0);
  const bool vecW = ((ldw & 3) == 0) && ((((uintptr_t)W) & 15) == 0);
  float4 ra[4], rw;
  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = m0 + i * 16 + lrow, k = k0 + lk;
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < epi.M) {
        const float* ap = A + (long)row * lda + k;
        if (vecA && k + 4 <= K) t = *reinterpret_cast<const float4*>(ap);
        else {
          if (k < K) t.x = ap[0];
          if (k + 1 < K) t.y = ap[1];
          if (k + 2 < K) t.z = ap[2];
          if (k + 3 < K) t.w = ap[3];
        }
      }
      ra[i] = t;
    }
    rw = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid < 64) {
      const int col = n0 + lrow, k = k0 + lk;
      if (col < epi.N) {
        const float* wp = W + (long)col * ldw + k;
        if (vecW && k + 4 <= K) rw = __ldg(reinterpret_cast<const float4*>(wp));
        else {
          if (k < K) rw.x = wp[0];
          if (k + 1 < K) rw.y = wp[1];
          if (k + 2 < K) rw.z = wp[2];
          if (k + 3 < K) rw.w = wp[3];
        }
      }
    }
  };
  float acc[RP_COLS];
#pragma unroll
  for (int j = 0; j < RP_COLS; ++j) acc[j] = 0.f;
  gload(0);
  for (int k0 = 0; k0 < K; k0 += RP_K) {
    __syncthreads();   // previous chunk fully consumed
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float* d = &As[i * 16 + lrow][lk];
      d[0] = ra[i].x; d[1] = ra[i].y; d[2] = ra[i].z; d[3] = ra[i].w;
    }
    if (tid < 64) { Ws[lk][lrow] = rw.x; Ws[lk + 1][lrow] = rw.y; Ws[lk + 2][lrow] = rw.z; Ws[lk + 3][lrow] = rw.w; }
    __syncthreads();
    if (k0 + RP_K < K) gload(k0 + RP_K);   // next chunk's loads in flight during the FMAs
#pragma unroll
    for (int kk = 0; kk < RP_K; kk += 2) {
      const float a = As[r][kk + kh];
      const float4 w0 = *reinterpret_cast<const float4*>(&Ws[kk + kh][0]);
      const float4 w1 = *reinterpret_cast<const float4*>(&Ws[kk + kh][4]);
      acc[0] = fmaf(a, w0.x, acc[0]); acc[1] = fmaf(a, w0.y, acc[1]);
      acc[2] = fmaf(a, w0.z, acc[2]); acc[3] = fmaf(a, w0.w, acc[3]);
      acc[4] = fmaf(a, w1.x, acc[4]); acc[5] = fmaf(a, w1.y, acc[5]);
      acc[6] = fmaf(a, w1.z, acc[6]); acc[7] = fmaf(a, w1.w, acc[7]);
    }
  }
#pragma unroll
  for (int j = 0; j < RP_COLS; ++j) {
    const float o = __shfl_xor_sync(0xffffffffu, acc[j], 1);
    acc[j] = kh == 0 ? acc[j] + o : o + acc[j];   // even-k partial + odd-k partial, same order in both lanes
  }
  if (kh == 0) {
    float v0[4] = {acc[0], acc[1], acc[2], acc[3]}, v1[4] = {acc[4], acc[5], acc[6], acc[7]};
    epilogue_chunk<4>(epi, m0 + r, n0, v0);
    epilogue_chunk<4>(epi, m0 + r, n0 + 4, v1);
  }
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
    dim3 grid(ceil_div(p->N, RP_COLS), ceil_div(p->M, RP_ROWS));
    gemm_f32_rowpar_kernel<<<grid, 128, 0, st>>>((const float*)p->A, (const float*)p->W, p->K, p->lda, p->ldw, e);
    DTTS_CHECK_LAUNCH("gemm_f32_rowpar");
    return 0;
  }
  dim3 grid(ceil_div(p->N, TN), ceil_div(p->M, TM));
  gemm_f32_kernel<<<grid, 256, 0, st>>>((const float*)p->A, (const float*)p->W, p->K, p->lda, p->ldw, p->taps,
                                        p->tap_shift0, p->tap_stride, e);
  DTTS_CHECK_LAUNCH("gemm_f32");
  return 0;
}
