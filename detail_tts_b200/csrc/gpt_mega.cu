// gpt_mega.cu -- the KV-cached GPT decode step for SMALL batches (B <= 32 utterances) as ONE persistent kernel.
//
// Why: the step (gpt/model.py:107-185 + HF GPT-2 block, modeling_gpt2.py:229-310) is a chain of ~84 dependent launches of
// 5-13 us each when run kernel by kernel; it does not get faster when the batch shrinks, which is what bounds strong
// scaling (16 utterances per GPU at 8 GPUs) and single-utterance latency.  At B <= 32 the arithmetic is tiny (B x 154
// MFLOP) and the step is a weight stream (308 MB of fp32 weights), so here it is: one CTA per SM, all CTAs co-resident
// (cooperative launch), phases separated by a grid barrier (~1-2 us instead of a launch boundary):
//   per layer:  QKV  (stage x = residual + split-K partials of the previous MLP, LayerNorm ln_1 in shared memory; GEMV
//                     over the 2304 output columns, written straight into the KV arena row of the new token)
//               ATT  (one CTA per (utterance, head): q.K^T, softmax, P.V over the cached positions)
//               PROJ (GEMV, K split in two -> partials)
//               FC   (stage x = residual + partials + bias, LayerNorm ln_2; GEMV 3072 columns, bias + gelu_new)
//               OUT  (GEMV, K split in two halves of the CTAs -> partials)
//   then HEAD (stage, ln_f, final_norm -> latent row; GEMV over the 8194 mel-head columns -> logits).
// Every CTA stages the (small) [B, K] activation block itself, so no phase needs a separate reduce / LayerNorm launch.
// GEMV: a warp owns CPW consecutive output columns, its lanes split K (512-byte coalesced weight-row reads, CPW loads in
// flight per lane), accumulates all B rows in registers against the staged activations, and reduces the CPW x B partial
// sums with a halving butterfly (62 shuffles for 64 values) after which each lane owns 1-2 finished outputs.
// Exact fp32 FMA arithmetic (the kernel-by-kernel path uses 3xTF32 on the tensor cores; both are fp32-class).
// Measured (B200, step as one CUDA graph): B=1 0.45 ms vs 0.64 ms kernel by kernel, B=4 0.47 vs 0.61, B=16 0.65 vs 0.63,
// B=32 0.98 vs 0.66: each of the 52 grid barriers costs ~3.5 us on the 148-SM / two-die part (runtime barrier and a
// hand-written one alike) and a phase ~4-5 us of dependent L2 round trips, so the host selects this kernel for B <= 4 only.
#include "common.cuh"
#include <cooperative_groups.h>

namespace {

constexpr int MG_D = 768, MG_H = 16, MG_HD = 48, MG_FF = 3072;
constexpr int MG_THREADS = 256, MG_WARPS = MG_THREADS / 32;
enum { LP_LN1_G, LP_LN1_B, LP_W_QKV, LP_B_QKV, LP_W_PROJ, LP_B_PROJ, LP_LN2_G, LP_LN2_B, LP_W_FC, LP_B_FC, LP_W_OUT, LP_B_OUT,
       LP_ARENA, LP_COUNT };
static_assert((int)LP_COUNT == (int)DTTS_GPT_LAYER_PTRS, "layer pointer table layout");

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Grid barrier over co-resident CTAs: arrival counter bar[0] (reset by the last arriver), generation bar[1].
struct GridBar {
  uint32_t* bar;
  unsigned gen;
};
__device__ __forceinline__ void grid_sync(GridBar& gb) {
#ifndef DTTS_MEGA_OWN_BARRIER
  // the runtime's grid barrier (cooperative launch): polls without invalidating L1 on every read -- the hand-written
  // variant below spent its time in CCTL.IVALL (one per ld.acquire poll; ncu: 66 % of all samples waiting here)
  cooperative_groups::this_grid().sync();
#else
  __syncthreads();                       // the CTA's writes happen-before thread 0's release below (cumulativity)
  if (threadIdx.x == 0) {
    unsigned t;
    asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(t) : "l"(gb.bar) : "memory");
    if (t == gridDim.x - 1) {
      asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(gb.bar), "r"(0u) : "memory");
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(gb.bar + 1) : "memory");
    } else {
      const long long t0 = clock64();
      while (*(volatile unsigned*)&gb.bar[1] == gb.gen) {
        if (clock64() - t0 > 4000000000LL) __trap();   // ~2 s: a lost CTA must not hang the device
      }
      asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
  }
  gb.gen++;
  __syncthreads();
#endif
}

// L2 prefetch of the weight rows this warp will stream in an upcoming phase (weights do not depend on the activations, so
// the HBM latency of a phase is paid while the grid is still waiting at the barriers before it).
template <int CPW>
__device__ __forceinline__ void prefetch_cols(const float* __restrict__ W, int ldw, int k0, int klen, int N, int wg, int wtot) {
  const int lane = threadIdx.x & 31;
  const int n_groups = (N + CPW - 1) / CPW;
  for (int gi = wg + wtot; gi < n_groups; gi += wtot) {   // (the warp's first group is preloaded into registers)
#pragma unroll
    for (int c = 0; c < CPW; ++c) {
      const float* row = W + (size_t)min(gi * CPW + c, N - 1) * ldw + k0;
      for (int k = lane * 32; k < klen; k += 32 * 32)      // one 128-byte line per lane
        asm volatile("prefetch.global.L2 [%0];" ::"l"(row + k));
    }
  }
}

__device__ __forceinline__ float4 f4add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

// LayerNorm of one 768-row held as 6 float4 per lane (element index lane*4 + 128*i).
__device__ __forceinline__ void ln_row(float4 (&v)[6], const float* __restrict__ g, const float* __restrict__ be, float eps, int lane) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 6; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = warp_sum(s) * (1.0f / MG_D);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    ss += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(ss) * (1.0f / MG_D) + eps);
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const int e = lane * 4 + 128 * i;
    const float4 gg = __ldg(reinterpret_cast<const float4*>(g + e)), bb = __ldg(reinterpret_cast<const float4*>(be + e));
    v[i].x = (v[i].x - mean) * rstd * gg.x + bb.x;
    v[i].y = (v[i].y - mean) * rstd * gg.y + bb.y;
    v[i].z = (v[i].z - mean) * rstd * gg.z + bb.z;
    v[i].w = (v[i].w - mean) * rstd * gg.w + bb.w;
  }
}

// Stage the [B, 768] input block of a phase into shared memory: x = base (+ part[0] + part[1]) (+ bias); CTA 0 writes x
// back (the new residual stream); optional LayerNorm(s); CTA 0 optionally writes the normalised rows (latent).
__device__ void stage_rows(float* xs, int B, const float* base, const float* part, const float* bias, float* writeback,
                           const float* g1, const float* b1, const float* g2, const float* b2, float* norm_out, float eps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // two rows per warp iteration: their loads and the LayerNorm shuffle chains overlap
  for (int b0 = warp; b0 < B; b0 += 2 * MG_WARPS) {
    float4 v[2][6];
    const int nr = b0 + MG_WARPS < B ? 2 : 1;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      if (r < nr) {
        const int b = b0 + r * MG_WARPS;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          const int e = lane * 4 + 128 * i;
          v[r][i] = *reinterpret_cast<const float4*>(base + (size_t)b * MG_D + e);
          if (part) {
            v[r][i] = f4add(v[r][i], *reinterpret_cast<const float4*>(part + (size_t)b * MG_D + e));
            v[r][i] = f4add(v[r][i], *reinterpret_cast<const float4*>(part + (size_t)(B + b) * MG_D + e));
          }
          if (bias) v[r][i] = f4add(v[r][i], __ldg(reinterpret_cast<const float4*>(bias + e)));
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      if (r < nr) {
        const int b = b0 + r * MG_WARPS;
        if (writeback && blockIdx.x == 0) {
#pragma unroll
          for (int i = 0; i < 6; ++i) *reinterpret_cast<float4*>(writeback + (size_t)b * MG_D + lane * 4 + 128 * i) = v[r][i];
        }
        if (g1) ln_row(v[r], g1, b1, eps, lane);
        if (g2) ln_row(v[r], g2, b2, eps, lane);
        if (norm_out && blockIdx.x == 0) {
#pragma unroll
          for (int i = 0; i < 6; ++i) *reinterpret_cast<float4*>(norm_out + (size_t)b * MG_D + lane * 4 + 128 * i) = v[r][i];
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) *reinterpret_cast<float4*>(xs + (size_t)b * MG_D + lane * 4 + 128 * i) = v[r][i];
      }
    }
  }
  __syncthreads();
}

// Plain copy of a [B, klen] block (row pitch ld) into shared memory [B][klen].
__device__ void stage_copy(float* xs, int B, const float* src, int ld, int klen) {
  const int per_row = klen >> 2;
  for (int i = threadIdx.x; i < B * per_row; i += MG_THREADS) {
    const int b = i / per_row, e = (i - b * per_row) * 4;
    *reinterpret_cast<float4*>(xs + (size_t)b * klen + e) = *reinterpret_cast<const float4*>(src + (size_t)b * ld + e);
  }
  __syncthreads();
}

enum { EP_QKV = 0, EP_PART = 1, EP_FC = 2, EP_HEAD = 3 };
struct EpiArgs {
  int mode;
  const float* bias;
  float* out;          // EP_QKV: arena base; EP_PART: partial buffer of this K half [B][768]; EP_FC: u; EP_HEAD: logits
  int ld;              // row pitch of out
  const int* kv_row;   // EP_QKV: arena row of utterance b
};

__device__ __forceinline__ void epi_store(const EpiArgs& ea, int b, int n, float v) {
  switch (ea.mode) {
    case EP_QKV: ea.out[(size_t)ea.kv_row[b] * ea.ld + n] = v + __ldg(ea.bias + n); break;
    case EP_PART: ea.out[(size_t)b * ea.ld + n] = v; break;
    case EP_FC: ea.out[(size_t)b * ea.ld + n] = act_apply(DTTS_ACT_GELU_NEW, v + __ldg(ea.bias + n), 0.f); break;
    default: ea.out[(size_t)b * ea.ld + n] = v + __ldg(ea.bias + n); break;
  }
}

// The first column group of a warp is held in registers ACROSS the grid barrier in front of its phase: weights do not depend
// on activations, so their HBM round trip overlaps the barrier wait and the staging instead of sitting on the phase's
// critical path (ncu: the phases were a chain of 2-3 dependent global round trips, 66 % of the samples waiting at the barrier).
constexpr int MG_PRE = 6;                         // k-steps of 128 floats kept per column (a whole 768-long row)
template <int CPW>
struct WPre {
  float4 w[MG_PRE][CPW];
};

template <int CPW>
__device__ __forceinline__ void preload_cols(WPre<CPW>& wp, const float* __restrict__ W, int ldw, int k0, int klen, int N, int wg) {
  const int lane = threadIdx.x & 31;
  if (wg * CPW >= N) return;
#pragma unroll
  for (int c = 0; c < CPW; ++c) {
    const float* row = W + (size_t)min(wg * CPW + c, N - 1) * ldw + k0 + lane * 4;
#pragma unroll
    for (int i = 0; i < MG_PRE; ++i)
      if (128 * i < klen) wp.w[i][c] = __ldcs(reinterpret_cast<const float4*>(row + 128 * i));
  }
}

template <int BMAX, int CPW>
__device__ __forceinline__ void fma_step(float (&acc)[BMAX * CPW], const float4 (&w)[CPW], const float* xs, int ldx, int k, int B) {
#pragma unroll
  for (int b = 0; b < BMAX; ++b) {
    if (b < B) {
      const float4 x = *reinterpret_cast<const float4*>(xs + (size_t)b * ldx + k);
#pragma unroll
      for (int c = 0; c < CPW; ++c) {
        float a = acc[c * BMAX + b];
        a = fmaf(w[c].x, x.x, a); a = fmaf(w[c].y, x.y, a); a = fmaf(w[c].z, x.z, a); a = fmaf(w[c].w, x.w, a);
        acc[c * BMAX + b] = a;
      }
    }
  }
}

// Halving butterfly over the warp (after the step with offset o a lane keeps the half of its values selected by its bit o),
// then lane L owns the finished sums j = L * (NV/32) + r, j = c * BMAX + b.
template <int BMAX, int CPW>
__device__ __forceinline__ void reduce_store(float (&acc)[BMAX * CPW], int n0, int N, int B, const EpiArgs& ea) {
  constexpr int NV = BMAX * CPW;
  const int lane = threadIdx.x & 31;
  int n = NV;
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool up = (lane & o) != 0;
    const int h = n >> 1;
#pragma unroll
    for (int i = 0; i < NV / 2; ++i) {
      if (i < h) {
        const float send = up ? acc[i] : acc[i + h];
        const float keep = up ? acc[i + h] : acc[i];
        acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
      }
    }
    n = h;
  }
#pragma unroll
  for (int r = 0; r < NV / 32; ++r) {
    const int j = lane * (NV / 32) + r;
    const int c = j / BMAX, b = j - c * BMAX;
    if (b < B && n0 + c < N) epi_store(ea, b, n0 + c, acc[r]);
  }
}

// y[b][n] = sum_k W[n][k0 + k] * xs[b * ldx + k], k < klen, for the column groups dealt to this warp (group wg first, its
// first MG_PRE k-steps from the registers preloaded before the barrier).
template <int BMAX, int CPW>
__device__ void gemv_phase(const WPre<CPW>& wp, const float* __restrict__ W, int ldw, int k0, int klen, int N, const float* xs, int ldx,
                           int B, int wg, int wtot, const EpiArgs ea) {
  constexpr int NV = BMAX * CPW;
  static_assert(NV == 32 || NV == 64, "butterfly reduction needs 32 or 64 partial sums per lane");
  const int lane = threadIdx.x & 31;
  const int n_groups = (N + CPW - 1) / CPW;
  int gi = wg;
  if (gi < n_groups) {
    const int n0 = gi * CPW;
    float acc[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) acc[j] = 0.f;
#pragma unroll
    for (int i = 0; i < MG_PRE; ++i)
      if (128 * i < klen) fma_step<BMAX, CPW>(acc, wp.w[i], xs, ldx, lane * 4 + 128 * i, B);
    if (klen > 128 * MG_PRE) {
      const float* wrow[CPW];
#pragma unroll
      for (int c = 0; c < CPW; ++c) wrow[c] = W + (size_t)min(n0 + c, N - 1) * ldw + k0;
#pragma unroll 6
      for (int k = lane * 4 + 128 * MG_PRE; k < klen; k += 128) {
        float4 w[CPW];
#pragma unroll
        for (int c = 0; c < CPW; ++c) w[c] = __ldcs(reinterpret_cast<const float4*>(wrow[c] + k));
        fma_step<BMAX, CPW>(acc, w, xs, ldx, k, B);
      }
    }
    reduce_store<BMAX, CPW>(acc, n0, N, B, ea);
    gi += wtot;
  }
  for (; gi < n_groups; gi += wtot) {
    const int n0 = gi * CPW;
    float acc[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) acc[j] = 0.f;
    const float* wrow[CPW];
#pragma unroll
    for (int c = 0; c < CPW; ++c) wrow[c] = W + (size_t)min(n0 + c, N - 1) * ldw + k0;
#pragma unroll 6                      // every weight load of a 768-long row is in flight at once
    for (int k = lane * 4; k < klen; k += 128) {
      float4 w[CPW];
#pragma unroll
      for (int c = 0; c < CPW; ++c) w[c] = __ldcs(reinterpret_cast<const float4*>(wrow[c] + k));   // streamed once per step
      fma_step<BMAX, CPW>(acc, w, xs, ldx, k, B);
    }
    reduce_store<BMAX, CPW>(acc, n0, N, B, ea);
  }
}

// One (utterance, head): q of the new token against the cached keys/values of the head (fp32, 8 warps split the keys).
__device__ void attend_item(const dtts_gpt_step_params& p, const float* arena, int b, int h, float* sc /*smem [max_k_len]*/,
                            float* qs, float* red, float (*part)[MG_HD]) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nk = p.kv_len[b];
  const int ld = p.arena_ld;
  const float* q = arena + (size_t)p.kv_row[b] * ld + h * MG_HD;
  const float* kb = arena + (size_t)p.k_off[b] * ld + MG_D + h * MG_HD;
  const float* vb = arena + (size_t)p.k_off[b] * ld + 2 * MG_D + h * MG_HD;
  const float scale = rsqrtf((float)MG_HD);
  if (tid < MG_HD) qs[tid] = q[tid] * scale;
  __syncthreads();
  const int sub = lane & 3, kl = lane >> 2;
  float4 q4[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) q4[i] = *reinterpret_cast<const float4*>(qs + 16 * i + 4 * sub);
  float lmax = -INFINITY;
#pragma unroll 2
  for (int j0 = warp * 8; j0 < nk; j0 += MG_WARPS * 8) {
    const int j = j0 + kl;
    float s = 0.f;
    if (j < nk) {
      const float* kj = kb + (size_t)j * ld + 4 * sub;
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
  if (lane == 0) red[warp] = lmax;
  __syncthreads();
  float gmax = red[0];
#pragma unroll
  for (int w = 1; w < MG_WARPS; ++w) gmax = fmaxf(gmax, red[w]);
  float lsum = 0.f;
  for (int j = tid; j < nk; j += MG_THREADS) {
    const float e = expf(sc[j] - gmax);
    sc[j] = e;
    lsum += e;
  }
  lsum = warp_sum(lsum);
  if (lane == 0) red[MG_WARPS + warp] = lsum;
  __syncthreads();
  float tot = red[MG_WARPS];
#pragma unroll
  for (int w = 1; w < MG_WARPS; ++w) tot += red[MG_WARPS + w];
  const float inv = 1.0f / tot;
  const int g = lane >> 3, l6 = (lane & 7) * 6;
  float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
  for (int j0 = warp * 4; j0 < nk; j0 += MG_WARPS * 4) {
    const int j = j0 + g;
    if (j < nk) {
      const float pj = sc[j];
      const float2* vj = reinterpret_cast<const float2*>(vb + (size_t)j * ld + l6);
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
  if (tid < MG_HD) {
    float o = part[0][tid];
#pragma unroll
    for (int w = 1; w < MG_WARPS; ++w) o += part[w][tid];
    p.att[(size_t)b * MG_D + h * MG_HD + tid] = o * inv;
  }
  __syncthreads();   // sc / qs / part are reused by the CTA's next item
}

template <int BMAX, int CPW>
__global__ void __launch_bounds__(MG_THREADS, 1)
gpt_step_kernel(const dtts_gpt_step_params p) {
  extern __shared__ __align__(16) float xs[];          // [B][<=1536] staged activations (attention scores in the ATT phase)
  __shared__ __align__(16) float qs[MG_HD];
  __shared__ float red[2 * MG_WARPS];
  __shared__ __align__(16) float apart[MG_WARPS][MG_HD];
  const int warp = threadIdx.x >> 5;
  const int B = p.B;
  GridBar gb;
  gb.bar = p.barrier;
  gb.gen = threadIdx.x == 0 ? ld_acquire_u32(&p.barrier[1]) : 0u;
  const int wg = blockIdx.x * MG_WARPS + warp, wtot = gridDim.x * MG_WARPS;
  const int half = blockIdx.x & 1;                     // OUT phase: K half of this CTA
  const int wg_h = (blockIdx.x >> 1) * MG_WARPS + warp, wtot_h = (gridDim.x >> 1) * MG_WARPS;
  const float* base = p.x_in;
  const float* prev_bias = nullptr;
  const float* prev_part = nullptr;
  float* X0 = p.xa;
  float* X1 = p.xb;
  WPre<CPW> wp;
  preload_cols<CPW>(wp, (const float*)p.layer_ptrs[LP_W_QKV], MG_D, 0, MG_D, 3 * MG_D, wg);
  for (int l = 0; l < p.n_layers; ++l) {
    const uint64_t* lp = p.layer_ptrs + (size_t)l * LP_COUNT;
    float* arena = (float*)lp[LP_ARENA];
    const int hh = wg & 1;                             // PROJ phase: K half of this warp
    // ---- QKV
    stage_rows(xs, B, base, prev_part, prev_bias, X0, (const float*)lp[LP_LN1_G], (const float*)lp[LP_LN1_B], nullptr, nullptr,
               nullptr, p.ln_eps);
    {
      EpiArgs ea{EP_QKV, (const float*)lp[LP_B_QKV], arena, p.arena_ld, p.kv_row};
      gemv_phase<BMAX, CPW>(wp, (const float*)lp[LP_W_QKV], MG_D, 0, MG_D, 3 * MG_D, xs, MG_D, B, wg, wtot, ea);
    }
    preload_cols<CPW>(wp, (const float*)lp[LP_W_PROJ], MG_D, hh * (MG_D / 2), MG_D / 2, MG_D, wg >> 1);
    prefetch_cols<CPW>((const float*)lp[LP_W_FC], MG_D, 0, MG_D, MG_FF, wg, wtot);
    grid_sync(gb);
    // ---- ATT
    for (int it = blockIdx.x; it < B * MG_H; it += gridDim.x) attend_item(p, arena, it / MG_H, it % MG_H, xs, qs, red, apart);
    grid_sync(gb);
    // ---- PROJ: work item = (column group, K half); even warps take half 0, odd warps half 1 of the same staged block
    stage_copy(xs, B, p.att, MG_D, MG_D);
    {
      EpiArgs ea{EP_PART, nullptr, p.part + (size_t)hh * B * MG_D, MG_D, nullptr};
      gemv_phase<BMAX, CPW>(wp, (const float*)lp[LP_W_PROJ], MG_D, hh * (MG_D / 2), MG_D / 2, MG_D, xs + hh * (MG_D / 2), MG_D, B,
                            wg >> 1, wtot >> 1, ea);
    }
    preload_cols<CPW>(wp, (const float*)lp[LP_W_FC], MG_D, 0, MG_D, MG_FF, wg);
    grid_sync(gb);
    // ---- FC
    stage_rows(xs, B, X0, p.part, (const float*)lp[LP_B_PROJ], X1, (const float*)lp[LP_LN2_G], (const float*)lp[LP_LN2_B], nullptr,
               nullptr, nullptr, p.ln_eps);
    {
      EpiArgs ea{EP_FC, (const float*)lp[LP_B_FC], p.u, MG_FF, nullptr};
      gemv_phase<BMAX, CPW>(wp, (const float*)lp[LP_W_FC], MG_D, 0, MG_D, MG_FF, xs, MG_D, B, wg, wtot, ea);
    }
    preload_cols<CPW>(wp, (const float*)lp[LP_W_OUT], MG_FF, half * (MG_FF / 2), MG_FF / 2, MG_D, wg_h);
    grid_sync(gb);
    // ---- OUT: the CTA stages its K half of u
    stage_copy(xs, B, p.u + half * (MG_FF / 2), MG_FF, MG_FF / 2);
    {
      EpiArgs ea{EP_PART, nullptr, p.part + (size_t)half * B * MG_D, MG_D, nullptr};
      gemv_phase<BMAX, CPW>(wp, (const float*)lp[LP_W_OUT], MG_FF, half * (MG_FF / 2), MG_FF / 2, MG_D, xs, MG_FF / 2, B, wg_h, wtot_h, ea);
    }
    if (l + 1 < p.n_layers) {
      preload_cols<CPW>(wp, (const float*)(lp + LP_COUNT)[LP_W_QKV], MG_D, 0, MG_D, 3 * MG_D, wg);
    } else {
      preload_cols<CPW>(wp, p.w_head, MG_D, 0, MG_D, p.vocab, wg);
      prefetch_cols<CPW>(p.w_head, MG_D, 0, MG_D, p.vocab, wg, wtot);
    }
    grid_sync(gb);
    base = X1;
    prev_part = p.part;
    prev_bias = (const float*)lp[LP_B_OUT];
  }
  // ---- HEAD: ln_f, final_norm (gpt/model.py:41,403: both are applied), latent row, mel_head logits
  stage_rows(xs, B, base, prev_part, prev_bias, nullptr, p.lnf_g, p.lnf_b, p.fn_g, p.fn_b, p.hn, p.ln_eps);
  {
    EpiArgs ea{EP_HEAD, p.b_head, p.logits, p.ld_logits, nullptr};
    gemv_phase<BMAX, CPW>(wp, p.w_head, MG_D, 0, MG_D, p.vocab, xs, MG_D, B, wg, wtot, ea);
  }
}

template <int BMAX, int CPW>
int launch_step(const dtts_gpt_step_params* p, int grid, size_t smem, cudaStream_t st) {
  static size_t attr = 0;
  if (smem > attr) {
    cudaError_t e = cudaFuncSetAttribute(gpt_step_kernel<BMAX, CPW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) DTTS_FAIL(-3, "cudaFuncSetAttribute(gpt_step): %s", cudaGetErrorString(e));
    attr = smem;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(MG_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;     // all CTAs co-resident, or the launch fails (the grid barrier needs it)
  at[0].val.cooperative = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gpt_step_kernel<BMAX, CPW>, *p);
  if (e != cudaSuccess) DTTS_FAIL(-3, "gpt_decode_step launch failed: %s", cudaGetErrorString(e));
  ++g_dtts_launches;
  return 0;
}

}  // namespace

extern "C" int dtts_gpt_decode_step(const dtts_gpt_step_params* p, void* stream) {
  DTTS_REQUIRE(p && p->layer_ptrs && p->x_in && p->xa && p->xb && p->att && p->u && p->part && p->logits && p->barrier &&
               p->kv_row && p->k_off && p->kv_len && p->w_head && p->b_head && p->lnf_g && p->lnf_b && p->fn_g && p->fn_b,
               "gpt_decode_step: null argument");
  DTTS_REQUIRE(p->B >= 1 && p->B <= 32, "gpt_decode_step: 1 <= B <= 32 (larger batches use the kernel-by-kernel step)");
  DTTS_REQUIRE(p->n_layers >= 1 && p->arena_ld >= 3 * MG_D && p->arena_ld % 4 == 0, "gpt_decode_step: bad arena layout");
  DTTS_REQUIRE(p->d_model == MG_D && p->n_heads == MG_H && p->d_ff == MG_FF, "gpt_decode_step: built for d_model 768, 16 heads, d_ff 3072");
  DTTS_REQUIRE(p->vocab > 0 && p->ld_logits >= p->vocab && p->max_k_len > 0, "gpt_decode_step: bad vocab / max_k_len");
  static int n_sm = 0;
  if (!n_sm) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  }
  const int grid = n_sm & ~1;                      // even: the OUT phase pairs CTAs over the two K halves
  const int BM = p->B <= 16 ? 16 : 32;
  size_t smem = (size_t)BM * (MG_FF / 2) * sizeof(float);
  if ((size_t)p->max_k_len * sizeof(float) > smem) smem = (size_t)p->max_k_len * sizeof(float);
  DTTS_REQUIRE(smem <= 200 * 1024, "gpt_decode_step: staging block does not fit in shared memory");
  cudaStream_t st = (cudaStream_t)stream;
  // CPW columns per warp: 2304 / 2 = 1152 column groups ~ one per warp of the grid (148 x 8), so a phase is ONE latency chain
  if (BM <= 16) return launch_step<16, 2>(p, grid, smem, st);
  return launch_step<32, 1>(p, grid, smem, st);
}
