// voc_fused.cu -- fused multi-receptive-field stage of the HiFi-GAN vocoder for the narrow stages
// (vqvae/model_24k.py:269-288 Generator.forward, vqvae/modules/modules.py:240-328 ResBlock1):
//
//   xs = ( RB_3(x) + RB_7(x) + RB_11(x) ) / 3,   RB_k: for d in (1,3,5): x = x + conv_{k,1}(lrelu(conv_{k,d}(lrelu(x))))
//   out16 = lrelu_{slope_out}(xs)                 (the operand of the next ConvTranspose1d / conv_post)
//
// At 25 and 12 channels (stages 4 and 5: 128x and 256x the frame rate, up to 9.3 M positions) a conv is far too
// thin for a tcgen05 tile; the 18 convs of a stage ran as 18 GEMM launches at 10-70 TFLOP/s, each a full pass over
// HBM.  Here ONE CTA owns a tile of 512 positions (392 outputs + the 60-position halo of the k=11 chain on both
// sides) and runs the whole chain with the activations in shared memory: fp32 residual stream and the running
// sum of the three ResBlocks in registers (mma.sync accumulator fragments), fp16 operand copies in two
// shared-memory buffers (channels-last, 16-byte padded rows: conflict-free ldmatrix, tap shifts are row offsets),
// weights as pre-packed B fragments straight from L1/L2.  HBM traffic per stage: x read, out16 written.
// Zero padding at utterance boundaries = the rows layout's separator rows, re-imposed after every conv through
// the row->utterance map.
#include "common.cuh"

namespace {

constexpr int VF_THREADS = 256, VF_WARPS = 8;
constexpr int VF_MT = 4;                                  // 16-position m-tiles per warp
constexpr int VF_TL_IN = VF_WARPS * VF_MT * 16;           // 512 positions per tile (incl. halo)
constexpr int VF_HALO = 60;                               // receptive field of RB_11 per side: 5 * (1+1 + 3+1 + 5+1)
constexpr int VF_TL_OUT = VF_TL_IN - 2 * VF_HALO;         // 392
constexpr int VF_MARGIN = 32;                             // >= 25 = largest tap shift (k=11, d=5)

__device__ __forceinline__ uint32_t sm_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float lrelu(float x, float s) { return x > 0.f ? x : x * s; }

// acc[mg][nt] (+)= conv over `k` taps with dilation `dil` of the fp16 activations at `src` (shared-memory byte address of
// tile position 0) for the MG m-tiles starting at position pos0.  wf: B fragments [tap][nt][ks][lane] (uint2).
template <int CP, int MG>
__device__ __forceinline__ void conv_taps(float (&acc)[MG][CP / 8][4], uint32_t src, int pos0, const uint2* __restrict__ wf,
                                          int k, int dil, int lane) {
  constexpr int NT = CP / 8, KS = CP / 16, PITCH = CP * 2 + 16;
  const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8, lk = ((lane >> 4) & 1) * 16;
  const uint32_t a_base = src + (uint32_t)((pos0 + lrow) * PITCH + lk);
  uint2 b[NT][KS], bn[NT][KS];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) bn[nt][ks] = __ldg(wf + (nt * KS + ks) * 32 + lane);
  for (int tap = 0; tap < k; ++tap) {
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) b[nt][ks] = bn[nt][ks];
    if (tap + 1 < k) {
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) bn[nt][ks] = __ldg(wf + (((tap + 1) * NT + nt) * KS + ks) * 32 + lane);
    }
    const int shift = (tap - (k >> 1)) * dil;
    const uint32_t a_tap = a_base + (uint32_t)(shift * PITCH);
#pragma unroll
    for (int mg = 0; mg < MG; ++mg) {
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        uint32_t a[4];
        ldsm_x4(a, a_tap + (uint32_t)(mg * 16 * PITCH + ks * 32));
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) mma16816(acc[mg][nt], a, b[nt][ks].x, b[nt][ks].y);
      }
    }
  }
}

template <int CP>
__global__ void __launch_bounds__(VF_THREADS, CP == 16 ? 2 : 1)
voc_mrf_kernel(const dtts_voc_mrf_params p) {
  constexpr int NT = CP / 8, KS = CP / 16, PITCH = CP * 2 + 16;
  constexpr int MG1 = CP == 16 ? VF_MT : 2;                // m-tiles per conv1 accumulator group (register budget)
  extern __shared__ __align__(16) uint8_t vf_smem[];
  uint8_t* bufA = vf_smem;                                                  // lrelu(x) fp16, positions [-MARGIN, TL_IN+MARGIN)
  uint8_t* bufT = bufA + (VF_TL_IN + 2 * VF_MARGIN) * PITCH;                // lrelu(conv1) fp16
  uint8_t* valid = bufT + (VF_TL_IN + 2 * VF_MARGIN) * PITCH;               // [TL_IN] 1 = inside an utterance
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tig = lane & 3;
  const long tile0 = (long)blockIdx.x * VF_TL_OUT - VF_HALO;                // global row of tile position 0
  // zero both operand buffers (margins stay zero; interior is overwritten), build the validity map
  for (int i = tid; i < 2 * (VF_TL_IN + 2 * VF_MARGIN) * PITCH / 16; i += VF_THREADS)
    reinterpret_cast<uint4*>(vf_smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = tid; i < VF_TL_IN; i += VF_THREADS) {
    const long r = tile0 + i;
    valid[i] = (r >= 0 && r < p.M && p.row_utt[r] >= 0) ? 1 : 0;
  }
  __syncthreads();
  const uint32_t sA = sm_addr(bufA) + VF_MARGIN * PITCH, sT = sm_addr(bufT) + VF_MARGIN * PITCH;
  uint8_t* const pA = bufA + VF_MARGIN * PITCH;
  uint8_t* const pT = bufT + VF_MARGIN * PITCH;
  const int wpos = warp * VF_MT * 16;                                       // first tile position of this warp
  // validity of the rows this thread owns in the accumulator fragments: bit (2*mt + hi)
  uint32_t vbits = 0;
#pragma unroll
  for (int mt = 0; mt < VF_MT; ++mt) {
    vbits |= (uint32_t)valid[wpos + mt * 16 + g] << (2 * mt);
    vbits |= (uint32_t)valid[wpos + mt * 16 + g + 8] << (2 * mt + 1);
  }
  const uint2* wf = reinterpret_cast<const uint2*>(p.w_frag);
  float accum[VF_MT][NT][4];
  float x[VF_MT][NT][4];
  int conv = 0;
  for (int rb = 0; rb < 3; ++rb) {
    const int k = 3 + 4 * rb;
    // residual stream of this ResBlock: x from HBM (L2 after the first ResBlock), operand copy lrelu(x) -> bufA
    if (rb > 0) __syncthreads();                                            // previous ResBlock done with bufA
#pragma unroll
    for (int mt = 0; mt < VF_MT; ++mt) {
#pragma unroll
      for (int hi = 0; hi < 2; ++hi) {
        const int pos = wpos + mt * 16 + g + hi * 8;
        const bool ok = (vbits >> (2 * mt + hi)) & 1u;
        const float* xr = p.x + (tile0 + pos) * p.ldx + 2 * tig;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          float2 v = make_float2(0.f, 0.f);
          if (ok) v = *reinterpret_cast<const float2*>(xr + nt * 8);
          x[mt][nt][2 * hi] = v.x; x[mt][nt][2 * hi + 1] = v.y;
          *reinterpret_cast<uint32_t*>(pA + pos * PITCH + (nt * 8 + 2 * tig) * 2) = pack_h2(lrelu(v.x, p.slope), lrelu(v.y, p.slope));
        }
      }
    }
    __syncthreads();
    for (int pair = 0; pair < 3; ++pair) {
      const int dil = 1 + 2 * pair;
      // conv1 (dilated): bufA -> t; bufT = lrelu(t + bias)
      const float* b1 = p.bias + conv * CP;
#pragma unroll
      for (int m0 = 0; m0 < VF_MT; m0 += MG1) {
        float t[MG1][NT][4];
#pragma unroll
        for (int mg = 0; mg < MG1; ++mg)
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            const float2 bb = *reinterpret_cast<const float2*>(b1 + nt * 8 + 2 * tig);
            t[mg][nt][0] = bb.x; t[mg][nt][1] = bb.y; t[mg][nt][2] = bb.x; t[mg][nt][3] = bb.y;
          }
        conv_taps<CP, MG1>(t, sA, wpos + m0 * 16, wf, k, dil, lane);
#pragma unroll
        for (int mg = 0; mg < MG1; ++mg)
#pragma unroll
          for (int hi = 0; hi < 2; ++hi) {
            const int mt = m0 + mg, pos = wpos + mt * 16 + g + hi * 8;
            const bool ok = (vbits >> (2 * mt + hi)) & 1u;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
              const float a = ok ? lrelu(t[mg][nt][2 * hi], p.slope) : 0.f, b = ok ? lrelu(t[mg][nt][2 * hi + 1], p.slope) : 0.f;
              *reinterpret_cast<uint32_t*>(pT + pos * PITCH + (nt * 8 + 2 * tig) * 2) = pack_h2(a, b);
            }
          }
      }
      wf += k * NT * KS * 32;
      ++conv;
      __syncthreads();
      // conv2 (dilation 1): bufT -> accumulated onto the residual stream; bufA = lrelu(x) for the next pair
      const float* b2 = p.bias + conv * CP;
#pragma unroll
      for (int mt = 0; mt < VF_MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const float2 bb = *reinterpret_cast<const float2*>(b2 + nt * 8 + 2 * tig);
          x[mt][nt][0] += bb.x; x[mt][nt][1] += bb.y; x[mt][nt][2] += bb.x; x[mt][nt][3] += bb.y;
        }
      conv_taps<CP, VF_MT>(x, sT, wpos, wf, k, 1, lane);
      wf += k * NT * KS * 32;
      ++conv;
      if (pair < 2) {
#pragma unroll
        for (int mt = 0; mt < VF_MT; ++mt)
#pragma unroll
          for (int hi = 0; hi < 2; ++hi) {
            const int pos = wpos + mt * 16 + g + hi * 8;
            const bool ok = (vbits >> (2 * mt + hi)) & 1u;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
              const float a = ok ? lrelu(x[mt][nt][2 * hi], p.slope) : 0.f, b = ok ? lrelu(x[mt][nt][2 * hi + 1], p.slope) : 0.f;
              *reinterpret_cast<uint32_t*>(pA + pos * PITCH + (nt * 8 + 2 * tig) * 2) = pack_h2(a, b);
            }
          }
        __syncthreads();
      }
    }
#pragma unroll
    for (int mt = 0; mt < VF_MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) accum[mt][nt][e] = rb == 0 ? x[mt][nt][e] : accum[mt][nt][e] + x[mt][nt][e];
  }
  // mean of the three ResBlocks -> outputs for the tile's interior positions
#pragma unroll
  for (int mt = 0; mt < VF_MT; ++mt)
#pragma unroll
    for (int hi = 0; hi < 2; ++hi) {
      const int pos = wpos + mt * 16 + g + hi * 8;
      const long r = tile0 + pos;
      if (pos < VF_HALO || pos >= VF_TL_IN - VF_HALO || !((vbits >> (2 * mt + hi)) & 1u)) continue;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const float a = accum[mt][nt][2 * hi] * (1.0f / 3.0f), b = accum[mt][nt][2 * hi + 1] * (1.0f / 3.0f);
        const int c = nt * 8 + 2 * tig;
        if (p.out_f32) *reinterpret_cast<float2*>(p.out_f32 + r * p.ldo32 + c) = make_float2(a, b);
        if (p.out_f16) *reinterpret_cast<uint32_t*>((__half*)p.out_f16 + r * p.ldo16 + c) = pack_h2(lrelu(a, p.slope_out), lrelu(b, p.slope_out));
      }
    }
}

// conv_post: wav[m] = tanh( sum_{tap<7} sum_{c<C} w[tap][c] * x16[m + tap - 3][c] )   (no bias; model_24k.py:284-286)
// x16 already holds lrelu_{0.01}(xs).  One thread per output sample; rows are 32 bytes, neighbours share them through L1.
__global__ void __launch_bounds__(256)
conv_post_kernel(const dtts_conv_post_params p) {
  __shared__ float w[7 * 16];
  for (int i = threadIdx.x; i < 7 * 16; i += 256) w[i] = (i % 16) < p.C ? p.w[(i / 16) * p.C + (i % 16)] : 0.f;
  __syncthreads();
  const long m = (long)blockIdx.x * 256 + threadIdx.x;
  if (m >= p.M) return;
  if (p.row_utt[m] < 0) return;
  float acc = 0.f;
#pragma unroll
  for (int tap = 0; tap < 7; ++tap) {
    const long r = m + tap - 3;
    if (r < 0 || r >= p.M) continue;
    const uint4* xr = reinterpret_cast<const uint4*>((const __half*)p.x + r * p.ldx);
    const uint4 v0 = __ldg(xr), v1 = __ldg(xr + 1);
    const uint32_t u[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&u[q]));
      acc = fmaf(f.x, w[tap * 16 + 2 * q], acc);
      acc = fmaf(f.y, w[tap * 16 + 2 * q + 1], acc);
    }
  }
  p.out[m * p.ldo] = tanhf(acc);
}

}  // namespace

extern "C" int dtts_voc_mrf(const dtts_voc_mrf_params* p, void* stream) {
  DTTS_REQUIRE(p && p->x && p->row_utt && p->w_frag && p->bias, "voc_mrf: null argument");
  DTTS_REQUIRE(p->Cp == 16 || p->Cp == 32, "voc_mrf: padded channel count must be 16 or 32 (wider stages run on the tcgen05 GEMM)");
  DTTS_REQUIRE(p->out_f16 || p->out_f32, "voc_mrf: no output");
  DTTS_REQUIRE(p->ldx % 2 == 0 && (((uintptr_t)p->x) & 7) == 0 && (((uintptr_t)p->w_frag) & 7) == 0 && (((uintptr_t)p->bias) & 7) == 0,
               "voc_mrf: operands must be 8-byte aligned");
  DTTS_REQUIRE(!p->out_f16 || (p->ldo16 % 2 == 0 && (((uintptr_t)p->out_f16) & 3) == 0), "voc_mrf: fp16 output alignment");
  DTTS_REQUIRE(!p->out_f32 || (p->ldo32 % 2 == 0 && (((uintptr_t)p->out_f32) & 7) == 0), "voc_mrf: fp32 output alignment");
  if (p->M <= 0) return 0;
  const int tiles = ceil_div(p->M, VF_TL_OUT);
  const int pitch = p->Cp * 2 + 16;
  const size_t smem = (size_t)2 * (VF_TL_IN + 2 * VF_MARGIN) * pitch + VF_TL_IN;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(voc_mrf_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(voc_mrf_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    attr = true;
  }
  if (p->Cp == 16) voc_mrf_kernel<16><<<tiles, VF_THREADS, smem, (cudaStream_t)stream>>>(*p);
  else voc_mrf_kernel<32><<<tiles, VF_THREADS, smem, (cudaStream_t)stream>>>(*p);
  DTTS_CHECK_LAUNCH("voc_mrf");
  return 0;
}

extern "C" int dtts_conv_post(const dtts_conv_post_params* p, void* stream) {
  DTTS_REQUIRE(p && p->x && p->w && p->out && p->row_utt, "conv_post: null argument");
  DTTS_REQUIRE(p->C >= 1 && p->C <= 16 && p->ldx == 16 && (((uintptr_t)p->x) & 15) == 0, "conv_post: needs fp16 rows of exactly 16 (padded) channels");
  if (p->M <= 0) return 0;
  conv_post_kernel<<<ceil_div(p->M, 256), 256, 0, (cudaStream_t)stream>>>(*p);
  DTTS_CHECK_LAUNCH("conv_post");
  return 0;
}
