// gpt_dgemm.cu -- the GEMMs of the KV-cached GPT decode step, fused with everything around them (round 2).
//
//   out[b, n] = act( sum_k LN(x)[b, k] * W[n, k] + bias[n] ) + res[b, n]          b < B <= 128 utterance rows
//
// Replaces, per GPT2Block of the decode step (transformers modeling_gpt2.py:262-310 as driven by gpt/model.py:107-185):
// ln_1 + c_attn, c_proj + residual, ln_2 + c_fc + gelu_new, mlp.c_proj + residual -- and gpt/model.py:324 mel_head -- i.e.
// the former {split-K GEMM, reduce + LayerNorm + operand split} launch pairs (gemm_tc.cu + gpt_step.cu).
//
// Design ("swap-AB" decode GEMM, sm_100a):
//  * A 128-row slab of W is the UMMA M operand; the activation rows are the N operand (padded to BP = 16/32/64/128; more
//    than 64 utterances may be split over two CTAs), so a 16-utterance shard costs 1/8 of the tensor time and shared-memory
//    operand traffic of a 128-row activation tile.
//  * 3xTF32 with fp32 accumulators in TMEM (fp32-class logits: token-exact sampling) in TWO
//    instructions per K=8 step instead of three: the x_hi and x_lo tiles lie back to back in shared memory, so ONE
//    tcgen05.mma with N = 2*BP computes W_hi*x_hi (columns [0, BP)) and W_hi*x_lo (columns [BP, 2BP)); a second one adds
//    W_lo*x_hi into columns [BP, 2BP); the epilogue adds the two column groups.  (The step is bound by the issue rate of the
//    single MMA thread at small B: measured ~55 clocks per tcgen05.mma.)
//  * W tiles (tf32-exact high part and low part, pre-split at load) arrive by TMA (128B swizzle) through an mbarrier
//    ring; they do not depend on the previous kernel, so the producer warp issues them BEFORE griddepcontrol.wait: under
//    programmatic dependent launch the weight stream of kernel N+1 overlaps the tail of kernel N.
//  * The activation tile is written by 8 warps straight into the swizzled UMMA layout: load fp32 x (L2, software-pipelined
//    one K block ahead), apply the LayerNorm that precedes the GEMM (row statistics come from the producer's epilogue as
//    per-128-column partial (sum, centred sum of squares) pairs, merged with Chan's formula -- deterministic and as accurate
//    as two-pass), split into tf32 hi + lo.  No LayerNorm / split / reduce launches.
//  * K is split over the CTAs of one thread-block cluster.  Every CTA PUSHES row b of its partial accumulator tile into the
//    landing buffer of the CTA that owns row b (st.shared::cluster), one cluster barrier later every CTA sums its rows of
//    all CL partials from its OWN shared memory in a fixed order (deterministic), applies bias / gelu_new / residual,
//    stores 512 contiguous bytes per row and emits the LayerNorm statistics of its output rows for the next GEMM.  Nobody
//    reads remote memory after the barrier, so CTAs exit independently.
#include <stdlib.h>

#include "common.cuh"
#include "tcgen05.cuh"

using namespace dtts_tc;

namespace {

constexpr int PREP_WARPS = 8;
constexpr int PREP_THREADS = PREP_WARPS * 32;
constexpr int PREP_WARP0 = 3;
constexpr int DG_THREADS = PREP_WARP0 * 32 + PREP_THREADS;   // warp 0: TMA producer, warps 1-2: MMA issuers (warp 1 also allocates TMEM), warps 3..10: activation tiles + epilogue
constexpr int W_TILE_BYTES = 128 * 128;         // 128 weight rows x 32 fp32 (one 128-byte swizzle span per row)
constexpr int MAX_STAGES = 6;
constexpr int LN_MAX_K = 768;                   // largest K slice of one CTA when the LayerNorm is fused (weight / bias staged in shared memory)

template <int BP>
struct DCfg {
  static constexpr int X_TILE_BYTES = BP * 128;
  static constexpr int STAGE_BYTES = 2 * W_TILE_BYTES + 2 * X_TILE_BYTES;      // {W_hi, W_lo, x_hi, x_lo}
  static constexpr int LAND_BYTES = BP * 128 * 4;                              // [CL][BP/CL][128] fp32 partial rows pushed by the cluster
  static constexpr int AUX_BYTES = 256 + 44 * BP + 2 * 4 * LN_MAX_K;           // barriers | mean, rstd [BP] | reduction scratch [2][4][BP] | out rows [BP] | gamma, beta
  static constexpr int STAGES_FIT = (226 * 1024 - LAND_BYTES - AUX_BYTES - 1024) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_FIT > MAX_STAGES ? MAX_STAGES : STAGES_FIT;
  static constexpr int NACC = BP <= 64 ? 2 : 1;                                // MMA issuer threads = independent accumulators (K blocks alternate)
  static constexpr int TMEM_COLS = NACC * 2 * BP < 32 ? 32 : NACC * 2 * BP;
  static constexpr int NI = BP >= 32 ? BP / 32 : 1;                            // activation rows per prep thread and K block
  static constexpr int PFD = BP <= 32 ? 4 : 2;                                 // K blocks of activation rows in flight (registers) per prep thread
};

__device__ __forceinline__ void bar_sync_prep() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t map_cluster(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void split1(float x, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);   // tf32-exact: sign, exponent, 10 mantissa bits
  lo = x - hi;                                               // exact in fp32
}
template <int NC>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, float (&v)[NC]);
template <>
__device__ __forceinline__ void tmem_ld<32>(uint32_t taddr, float (&v)[32]) { tmem_ld32(taddr, v); }
template <>
__device__ __forceinline__ void tmem_ld<16>(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
template <>
__device__ __forceinline__ void tmem_ld<8>(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// BP: activation rows of this CTA (padded); blockIdx.y selects the row block when B > BP.
template <int BP>
__global__ void __launch_bounds__(DG_THREADS, 1)
dgemm_kernel(const __grid_constant__ CUtensorMap tmHi, const __grid_constant__ CUtensorMap tmLo, const dtts_dgemm_params p,
             const int nkb, const int stages, long long* trace) {
  using C = DCfg<BP>;
#define TR(slot) do { if (trace && blockIdx.x == 0 && blockIdx.y == 0) trace[slot] = clock64(); } while (0)
  pdl_launch();            // dependents may be scheduled right away: they only prefetch weights until their own griddepcontrol.wait
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // swizzled tiles need 1024-byte alignment
  float* land = reinterpret_cast<float*>(smem + stages * C::STAGE_BYTES);
  uint8_t* aux = smem + stages * C::STAGE_BYTES + C::LAND_BYTES;
  uint64_t* full_w = (uint64_t*)aux;          // [MAX_STAGES]
  uint64_t* full_x = full_w + MAX_STAGES;     // [MAX_STAGES]
  uint64_t* empty = full_x + MAX_STAGES;      // [MAX_STAGES]
  uint64_t* acc_full = empty + MAX_STAGES;    // [1]
  uint32_t* tmem_slot = (uint32_t*)(acc_full + 1);
  float* s_mean = (float*)(aux + 256);
  float* s_rstd = s_mean + BP;
  float* s_red = s_rstd + BP;                 // [2][4][BP]
  int* s_orow = (int*)(s_red + 8 * BP);       // [BP] output row of local utterance row
  float* s_gamma = (float*)(s_orow + BP);     // [nkb * 32] LayerNorm weight / bias of this CTA's K slice
  float* s_beta = s_gamma + LN_MAX_K;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t csize = (uint32_t)p.k_splits;
  uint32_t crank = 0;
  if (csize > 1) asm("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  const int slab = blockIdx.x / csize;
  const int n0 = slab * 128;
  const int kb0 = (int)crank * nkb;
  const int b_base = blockIdx.y * BP;                      // first utterance row of this CTA
  const int Bl = min(BP, p.B - b_base);                    // valid local rows

  if (threadIdx.x == PREP_WARP0 * 32) TR(0);
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full_w[s], 1); mbar_init(&full_x[s], PREP_WARPS); mbar_init(&empty[s], 1); }
    mbar_init(acc_full, C::NACC);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)C::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // distributed shared memory may only be written once every CTA of the cluster runs: arrive now, wait right before the
  // first remote store (long complete by then)
  if (csize > 1) cluster_arrive_relaxed();

  if (warp == 0) {
    // ================================ weight producer (never waits for the previous kernel) ================================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmHi) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmLo) : "memory");
      int s = 0; uint32_t ph = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&empty[s], ph ^ 1);
        uint8_t* st = smem + s * C::STAGE_BYTES;
        mbar_expect_tx(&full_w[s], 2 * W_TILE_BYTES);
        tma_load_2d(st, &tmHi, &full_w[s], (kb0 + kb) * 32, n0);
        tma_load_2d(st + W_TILE_BYTES, &tmLo, &full_w[s], (kb0 + kb) * 32, n0);
        TR(16 + kb * 8 + 0);
        if (++s == stages) { s = 0; ph ^= 1; }
      }
    }
    __syncwarp();
    if (csize > 1) { cluster_wait(); cluster_sync_all(); } else __syncthreads();     // (see the epilogue warps)
  } else if (warp < PREP_WARP0) {
    // ================================ MMA issuers ================================
    // One thread issues ~55 clocks per tcgen05.mma (descriptor arithmetic + election), far more than a narrow-N MMA takes to
    // execute: NACC issuer threads take alternate K blocks into their own TMEM accumulators (summed by the epilogue).
    const int me = warp - 1;
    if (me < C::NACC) {          // the whole warp runs the issue loop with warp-uniform operands; one elected lane issues (umma_tf32_e)
      const uint32_t idesc2 = make_idesc(128, 2 * BP) | (2u << 7) | (2u << 10);   // a/b format TF32, N = 2*BP: [x_hi ; x_lo]
      const uint32_t idesc1 = make_idesc(128, BP) | (2u << 7) | (2u << 10);       // N = BP: x_hi only
      const uint32_t acc = tmem_base + (uint32_t)(me * 2 * BP);
      int s = 0; uint32_t ph = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        if (kb % C::NACC == me) {
          mbar_wait(&full_w[s], ph);
          if (me == 0 && lane == 0) TR(16 + kb * 8 + 1);
          mbar_wait(&full_x[s], ph);
          if (me == 0 && lane == 0) TR(16 + kb * 8 + 2);
          tc_fence_after();
          const uint32_t w_hi = smem_u32(smem + s * C::STAGE_BYTES);
          const uint64_t d_whi = make_desc(w_hi), d_wlo = make_desc(w_hi + W_TILE_BYTES), d_x = make_desc(w_hi + 2 * W_TILE_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) {       // 4 x (K = 8 tf32 = 32 bytes: +2 in the descriptor's 16-byte address units)
            umma_tf32_e(acc, d_whi + 2 * k, d_x + 2 * k, idesc2, (kb >= C::NACC || k > 0) ? 1u : 0u);   // W_hi*x_hi | W_hi*x_lo
            umma_tf32_e(acc + BP, d_wlo + 2 * k, d_x + 2 * k, idesc1, 1u);                              // += W_lo*x_hi
          }
          tc_commit_e(&empty[s]);
          if (me == 0 && lane == 0) TR(16 + kb * 8 + 3);
        }
        if (++s == stages) { s = 0; ph ^= 1; }
      }
      tc_commit_e(acc_full);
    }
    __syncwarp();
    if (csize > 1) { cluster_wait(); cluster_sync_all(); } else __syncthreads();     // (see the epilogue warps)
  } else {
    // ================================ activation tiles, then the epilogue ================================
    const int et = threadIdx.x - PREP_WARP0 * 32;       // 0..255
    const bool ln = p.ln_stats != nullptr;
    const int c4 = et & 7;                 // 16-byte chunk of the 128-byte K row this thread fills (fixed: 256 % 8 == 0)
    const int brow = et >> 3;              // local row, + 32 * i
    // everything that does not depend on the previous kernel first: LayerNorm weight / bias of the K slice
    if (ln)
      for (int i = et; i < nkb * 32; i += PREP_THREADS) { s_gamma[i] = __ldg(p.ln_gamma + kb0 * 32 + i); s_beta[i] = __ldg(p.ln_beta + kb0 * 32 + i); }
    if (et == 0) TR(1);
    pdl_wait();                            // x, the LayerNorm statistics and the residual come from the previous kernels
    if (et == 0) TR(2);
    // all dependent loads are issued before anything consumes them (a warp stalls at its first use in program order):
    // the activation rows of the first PFD K blocks (xq[d] = K block kb + d)
    float4 xq[C::PFD][C::NI];
#pragma unroll
    for (int d = 0; d < C::PFD; ++d)
#pragma unroll
      for (int i = 0; i < C::NI; ++i) {
        const int bl = brow + 32 * i;
        xq[d][i] = (bl < Bl && d < nkb) ? *reinterpret_cast<const float4*>(p.x + (size_t)(b_base + bl) * p.ldx + (kb0 + d) * 32 + c4 * 4)
                                        : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    if (et < BP) s_orow[et] = et < Bl ? (p.out_row_map ? p.out_row_map[b_base + et] : b_base + et) : 0;
    if (ln && et < BP) {
      float mean = 0.f, rstd = 1.f;
      if (et < Bl) {
        float s1[8], q1[8];
#pragma unroll
        for (int pp = 0; pp < 8; ++pp) {
          const float2 t = pp < p.ln_parts ? *reinterpret_cast<const float2*>(p.ln_stats + ((size_t)pp * p.B + b_base + et) * 2) : make_float2(0.f, 0.f);
          s1[pp] = t.x; q1[pp] = t.y;
        }
        const float npp = (float)(p.K / p.ln_parts), inv_npp = 1.0f / npp, inv_k = 1.0f / (float)p.K;
        float tot = 0.f;
#pragma unroll
        for (int pp = 0; pp < 8; ++pp) tot += s1[pp];
        mean = tot * inv_k;
        float m2 = 0.f;
#pragma unroll
        for (int pp = 0; pp < 8; ++pp)
          if (pp < p.ln_parts) { const float d = s1[pp] * inv_npp - mean; m2 += q1[pp] + npp * d * d; }
        rstd = 1.0f / sqrtf(m2 * inv_k + p.ln_eps);
      }
      s_mean[et] = mean; s_rstd[et] = rstd;
    }
    bar_sync_prep();
    if (et == 0) TR(3);
    {
      float rm[C::NI], rr[C::NI];          // this thread's rows are the same in every K block
#pragma unroll
      for (int i = 0; i < C::NI; ++i) {
        const int bl = brow + 32 * i;
        rm[i] = (ln && bl < BP) ? s_mean[bl] : 0.f;
        rr[i] = (ln && bl < BP) ? s_rstd[bl] : 1.f;
      }
      int s = 0; uint32_t ph = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        float4 xv[C::NI];
#pragma unroll
        for (int i = 0; i < C::NI; ++i) xv[i] = xq[0][i];
#pragma unroll
        for (int d = 0; d + 1 < C::PFD; ++d)
#pragma unroll
          for (int i = 0; i < C::NI; ++i) xq[d][i] = xq[d + 1][i];
        if (kb + C::PFD < nkb) {           // software pipeline: K block kb + PFD goes in flight while this one is written
#pragma unroll
          for (int i = 0; i < C::NI; ++i) {
            const int bl = brow + 32 * i;
            xq[C::PFD - 1][i] = bl < Bl ? *reinterpret_cast<const float4*>(p.x + (size_t)(b_base + bl) * p.ldx + (kb0 + kb + C::PFD) * 32 + c4 * 4)
                                        : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        float4 g4 = make_float4(1.f, 1.f, 1.f, 1.f), be4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ln) { g4 = *reinterpret_cast<const float4*>(s_gamma + kb * 32 + c4 * 4); be4 = *reinterpret_cast<const float4*>(s_beta + kb * 32 + c4 * 4); }
        mbar_wait(&empty[s], ph ^ 1);
        if (et == 0) TR(16 + kb * 8 + 4);
        uint8_t* x_hi = smem + s * C::STAGE_BYTES + 2 * W_TILE_BYTES;
        uint8_t* x_lo = x_hi + C::X_TILE_BYTES;
#pragma unroll
        for (int i = 0; i < C::NI; ++i) {
          const int bl = brow + 32 * i;
          if (bl < BP) {
            float4 v = xv[i];
            if (ln && bl < Bl) {
              v.x = (v.x - rm[i]) * rr[i] * g4.x + be4.x; v.y = (v.y - rm[i]) * rr[i] * g4.y + be4.y;
              v.z = (v.z - rm[i]) * rr[i] * g4.z + be4.z; v.w = (v.w - rm[i]) * rr[i] * g4.w + be4.w;
            }
            float4 h, l;
            split1(v.x, h.x, l.x); split1(v.y, h.y, l.y); split1(v.z, h.z, l.z); split1(v.w, h.w, l.w);
            const uint32_t off = (uint32_t)bl * 128u + (uint32_t)((c4 ^ (bl & 7)) << 4);   // SWIZZLE_128B: chunk ^= row % 8
            *reinterpret_cast<float4*>(x_hi + off) = h;
            *reinterpret_cast<float4*>(x_lo + off) = l;
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to tcgen05.mma
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_x[s]);
        if (et == 0) TR(16 + kb * 8 + 5);
        if (++s == stages) { s = 0; ph ^= 1; }
      }
    }
    // ---- accumulator (TMEM lane = weight row n; columns [0, BP) + [BP, 2BP) = utterance b) -> pushed, row by row, into the
    //      landing buffer of the cluster CTA that finishes row b: land[src rank][b % Bc][n]
    if (et == 0) TR(4);
    mbar_wait(acc_full, 0);
    if (et == 0) TR(5);
    tc_fence_after();
    const int q = warp & 3;                // TMEM lane quarter this warp may read
    const int hsel = (warp - PREP_WARP0) >> 2;      // which half of the utterance columns this warp handles
    const int nl = q * 32 + lane;
    constexpr int HB = BP / 2;             // columns per half
    constexpr int CH = HB >= 32 ? 32 : HB; // columns per tcgen05.ld
    const int Bc = BP / (int)csize;        // rows each CTA finishes
    const uint32_t land_addr = smem_u32(land);
    if (csize > 1) cluster_wait();         // (phase 1: every peer has started)
#pragma unroll
    for (int c = 0; c < HB / CH; ++c) {
      float va[CH], vb[CH];
      const int col0 = hsel * HB + c * CH;
      tmem_ld<CH>(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)col0, va);
      tmem_ld<CH>(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(BP + col0), vb);
      if (C::NACC > 1 && nkb > 1) {        // second issuer's accumulator (odd K blocks)
        float vc[CH], vd[CH];
        tmem_ld<CH>(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(2 * BP + col0), vc);
        tmem_ld<CH>(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(3 * BP + col0), vd);
#pragma unroll
        for (int j = 0; j < CH; ++j) { va[j] += vc[j]; vb[j] += vd[j]; }
      }
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        const int bl = col0 + j;
        const int owner = bl / Bc, slot = bl - owner * Bc;
        const uint32_t off = (uint32_t)((((int)crank * Bc + slot) * 128 + nl) * 4);
        const float v = va[j] + vb[j];
        if (csize > 1) st_cluster_f32(map_cluster(land_addr, (uint32_t)owner) + off, v);
        else land[(slot) * 128 + nl] = v;
      }
    }
    tc_fence_before();
    if (et == 0) TR(6);

    // ---- finish rows [bl_first, bl_first + Bc) of this CTA: thread = (weight row n, half), the halves alternate over
    //      batches of U rows.  The first batch's residual values do not depend on the peers: fetch them before the barrier.
    const int nr = et & 127, hr = et >> 7, wq = (et >> 5) & 3;
    const int n = n0 + nr;
    const bool nvalid = n < p.N;
    const int bl_first = (int)crank * Bc;
    constexpr int U = 4;
    float r0[U];
    size_t orow0[U];
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const int bb = hr * U + j, bl = bl_first + bb;
      const bool bv = bb < Bc && bl < Bl;
      orow0[j] = bv ? (size_t)s_orow[bl] : 0;
      r0[j] = (bv && p.res && nvalid) ? p.res[orow0[j] * p.ldr + n] : 0.f;
    }
    // every partial row has landed at its owner (release / acquire at cluster scope)
    if (csize > 1) cluster_sync_all(); else __syncthreads();
    if (et == 0) TR(7);
    const float bias_n = (p.bias && nvalid) ? __ldg(p.bias + n) : 0.f;
    const bool stats = p.out_stats != nullptr;
    for (int bb0 = hr * U; bb0 < Bc; bb0 += 2 * U) {      // the two thread halves alternate over batches of U rows
      if (bl_first + bb0 >= Bl) break;
      float r[U], t[U][8];
      size_t orow[U];
      const bool first = bb0 == hr * U;
#pragma unroll
      for (int j = 0; j < U; ++j) {
        const int bb = bb0 + j, bl = bl_first + bb;
        const bool bv = bb < Bc && bl < Bl;
        orow[j] = first ? orow0[j] : (bv ? (size_t)s_orow[bl] : 0);
        r[j] = first ? r0[j] : ((bv && p.res && nvalid) ? p.res[orow[j] * p.ldr + n] : 0.f);
#pragma unroll
        for (int c = 0; c < 8; ++c) t[j][c] = (c < (int)csize && bv) ? land[(c * Bc + bb) * 128 + nr] : 0.f;
      }
#pragma unroll
      for (int j = 0; j < U; ++j) {
        const int bb = bb0 + j, bl = bl_first + bb;
        if (bb < Bc && bl < Bl) {          // uniform over each warp
          float a = t[j][0];
#pragma unroll
          for (int c = 1; c < 8; ++c) a += t[j][c];      // fixed order: deterministic (absent partials are exact zeros)
          a = act_apply(p.act, a + bias_n, 0.f) + r[j];
          if (nvalid) p.out[orow[j] * p.ldo + n] = a;
          if (stats) {
            land[bb * 128 + nr] = a;       // slot (rank 0, bb): this thread is its only reader from here on
            const float s1 = warp_sum(a);
            if (lane == 0) s_red[wq * BP + bb] = s1;
          }
        }
      }
    }
    if (stats) {
      bar_sync_prep();
      for (int bb0 = hr * U; bb0 < Bc; bb0 += 2 * U) {
        if (bl_first + bb0 >= Bl) break;
#pragma unroll
        for (int j = 0; j < U; ++j) {
          const int bb = bb0 + j;
          if (bb < Bc && bl_first + bb < Bl) {
            const float mean = (((s_red[bb] + s_red[BP + bb]) + s_red[2 * BP + bb]) + s_red[3 * BP + bb]) * (1.0f / 128.0f);
            const float d = land[bb * 128 + nr] - mean;
            const float s2 = warp_sum(d * d);
            if (lane == 0) s_red[4 * BP + wq * BP + bb] = s2;
          }
        }
      }
      bar_sync_prep();
      if (et < Bc && bl_first + et < Bl) {
        const int bb = et;
        const float s1 = ((s_red[bb] + s_red[BP + bb]) + s_red[2 * BP + bb]) + s_red[3 * BP + bb];
        const float m2 = ((s_red[4 * BP + bb] + s_red[5 * BP + bb]) + s_red[6 * BP + bb]) + s_red[7 * BP + bb];
        float* o = p.out_stats + ((size_t)slab * p.B + b_base + bl_first + bb) * 2;
        o[0] = s1; o[1] = m2;
      }
    }
    if (et == 0) TR(8);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS) : "memory");
  }
  if (threadIdx.x == PREP_WARP0 * 32) TR(9);
#undef TR
}

long long* g_dgemm_trace = nullptr;   // debug: device buffer receiving CTA 0's clock64() timeline (dtts_dgemm_set_trace)

template <int BP>
int launch_dgemm(const dtts_dgemm_params* p, cudaStream_t st, int b_blocks) {
  using C = DCfg<BP>;
  static_assert(C::STAGES >= 2, "decode GEMM: shared memory budget leaves no pipeline");
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(dgemm_kernel<BP>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::STAGES * C::STAGE_BYTES + C::LAND_BYTES + C::AUX_BYTES + 1024);
    if (e != cudaSuccess) DTTS_FAIL(-3, "cudaFuncSetAttribute(dgemm<%d>): %s", BP, cudaGetErrorString(e));
    attr_set = true;
  }
  CUtensorMap mh, ml;
  int rc = get_map(p->W_hi, p->w_rows, p->K, p->ldw, 128, &mh, 4);
  if (rc) return rc;
  rc = get_map(p->W_lo, p->w_rows, p->K, p->ldw, 128, &ml, 4);
  if (rc) return rc;
  const int nkb = p->K / 32 / p->k_splits;
  static int max_stages = -1;   // DTTS_DGEMM_STAGES: cap on the ring depth (a CTA below ~110 KB lets the next kernel's CTAs co-reside under PDL)
  if (max_stages < 0) { const char* e = getenv("DTTS_DGEMM_STAGES"); max_stages = e ? atoi(e) : MAX_STAGES; if (max_stages < 1) max_stages = 1; }
  int stages = nkb < C::STAGES ? nkb : C::STAGES;
  if (stages > max_stages) stages = max_stages;
  const int slabs = ceil_div(p->N, 128);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(slabs * p->k_splits), (unsigned)b_blocks);
  cfg.blockDim = dim3(DG_THREADS);
  cfg.dynamicSmemBytes = (size_t)stages * C::STAGE_BYTES + C::LAND_BYTES + C::AUX_BYTES + 1024;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (p->k_splits > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = (unsigned)p->k_splits; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (g_dtts_pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  cudaError_t le = cudaLaunchKernelEx(&cfg, dgemm_kernel<BP>, mh, ml, *p, nkb, stages, g_dgemm_trace);
  if (le != cudaSuccess) DTTS_FAIL(-3, "decode_gemm launch failed: %s", cudaGetErrorString(le));
  DTTS_CHECK_LAUNCH("decode_gemm");
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// ln_f -> final_norm of the B new rows (+ latent capture).  One CTA per row, exact two-pass statistics in registers.
constexpr int FL_THREADS = 256;
constexpr int FL_MAXE = 4;   // C <= 1024

__global__ void __launch_bounds__(FL_THREADS)
final_ln_kernel(const dtts_final_ln_params p) {
  __shared__ float red[40];
  pdl_launch();
  pdl_wait();
  const int b = blockIdx.x, tid = threadIdx.x;
  float v[FL_MAXE];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < FL_MAXE; ++i) {
    const int c = tid + i * FL_THREADS;
    v[i] = c < p.C ? p.x[(size_t)b * p.ldx + c] : 0.f;
    s += v[i];
  }
  for (int pass = 0; pass < 2; ++pass) {
    const float* g = pass == 0 ? p.g1 : p.g2;
    const float* be = pass == 0 ? p.b1 : p.b2;
    if (!g) break;
    if (pass == 1) {
      s = 0.f;
#pragma unroll
      for (int i = 0; i < FL_MAXE; ++i) s += v[i];
    }
    const float mean = block_sum(s, red) / (float)p.C;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < FL_MAXE; ++i) {
      const int c = tid + i * FL_THREADS;
      if (c < p.C) { const float d = v[i] - mean; ss += d * d; }
    }
    const float rstd = 1.0f / sqrtf(block_sum(ss, red) / (float)p.C + p.eps);
#pragma unroll
    for (int i = 0; i < FL_MAXE; ++i) {
      const int c = tid + i * FL_THREADS;
      v[i] = c < p.C ? (v[i] - mean) * rstd * __ldg(g + c) + __ldg(be + c) : 0.f;
    }
  }
  float* lat = nullptr;
  if (p.lat) {
    const int pos = p.lat_pos0 + (p.step_dev ? *p.step_dev : 0) - (p.row_step0 ? p.row_step0[b] : 0);
    if (p.lat_T <= 0 || (pos >= 0 && pos < p.lat_T)) lat = p.lat + (size_t)b * p.lat_stride_b + (size_t)pos * p.C;
  }
#pragma unroll
  for (int i = 0; i < FL_MAXE; ++i) {
    const int c = tid + i * FL_THREADS;
    if (c < p.C) {
      p.y[(size_t)b * p.ldy + c] = v[i];
      if (lat) lat[c] = v[i];
    }
  }
}

}  // namespace

extern "C" int dtts_decode_gemm(const dtts_dgemm_params* p, void* stream) {
  DTTS_REQUIRE(p && p->x && p->W_hi && p->W_lo && p->out, "decode_gemm: null argument");
  DTTS_REQUIRE(p->B >= 1 && p->B <= 128, "decode_gemm: B=%d out of range (1..128)", p->B);
  DTTS_REQUIRE(p->N > 0 && p->K > 0 && p->K % 32 == 0, "decode_gemm: bad shape N=%d K=%d", p->N, p->K);
  DTTS_REQUIRE(p->k_splits == 1 || p->k_splits == 2 || p->k_splits == 4 || p->k_splits == 8, "decode_gemm: k_splits must be 1, 2, 4 or 8");
  DTTS_REQUIRE((p->K / 32) % p->k_splits == 0, "decode_gemm: K/32 must be divisible by k_splits");
  DTTS_REQUIRE(p->ldx % 4 == 0 && p->ldw % 4 == 0 && p->ldw >= p->K && p->ldx >= p->K, "decode_gemm: bad leading dimensions");
  DTTS_REQUIRE(p->w_rows >= p->N, "decode_gemm: w_rows < N");
  DTTS_REQUIRE(((((uintptr_t)p->x) | ((uintptr_t)p->W_hi) | ((uintptr_t)p->W_lo)) & 15) == 0, "decode_gemm: operands must be 16-byte aligned");
  DTTS_REQUIRE(!p->ln_stats || (p->ln_gamma && p->ln_beta && p->ln_parts > 0 && p->K % p->ln_parts == 0 &&
                                ((((uintptr_t)p->ln_gamma) | ((uintptr_t)p->ln_beta)) & 15) == 0),
               "decode_gemm: LayerNorm needs gamma, beta and ln_parts dividing K");
  DTTS_REQUIRE(!p->ln_stats || (p->ln_parts <= 8 && p->K / p->k_splits <= LN_MAX_K), "decode_gemm: LayerNorm needs ln_parts <= 8 and K/k_splits <= %d", LN_MAX_K);
  DTTS_REQUIRE(!p->out_stats || p->N % 128 == 0, "decode_gemm: out_stats needs N %% 128 == 0");
  DTTS_REQUIRE(p->act == DTTS_ACT_NONE || p->act == DTTS_ACT_GELU_NEW, "decode_gemm: unsupported activation");
  cudaStream_t st = (cudaStream_t)stream;
  if (p->B <= 16) return launch_dgemm<16>(p, st, 1);
  if (p->B <= 32) return launch_dgemm<32>(p, st, 1);
  if (p->B <= 64) return launch_dgemm<64>(p, st, 1);
  // more than 64 utterances: two CTAs per (slab, K split) with 64 rows each when that still fits one wave of SMs -- the
  // tensor time per CTA halves (the step is MMA-bound at this size) for twice the L2 -> SM weight traffic
  static int sm_count = 0;
  if (!sm_count) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev); }
  static int bsplit = -1;
  if (bsplit < 0) { const char* e = getenv("DTTS_DGEMM_BSPLIT"); bsplit = e ? atoi(e) : 1; }
  if (bsplit && 2 * ceil_div(p->N, 128) * p->k_splits <= sm_count) return launch_dgemm<64>(p, st, 2);
  return launch_dgemm<128>(p, st, 1);
}

// debug hook (tools/dgemm_probe.py): CTA 0 of every following dtts_decode_gemm launch writes clock64() stamps to `buf` (>= 256 int64)
extern "C" int dtts_dgemm_set_trace(void* buf) { g_dgemm_trace = (long long*)buf; return 0; }

extern "C" int dtts_final_ln(const dtts_final_ln_params* p, void* stream) {
  DTTS_REQUIRE(p && p->x && p->y && p->g1 && p->b1, "final_ln: null argument");
  DTTS_REQUIRE(p->C > 0 && p->C <= FL_THREADS * FL_MAXE, "final_ln: C out of range");
  DTTS_REQUIRE(!p->g2 || p->b2, "final_ln: g2 without b2");
  if (p->B <= 0) return 0;
  cudaError_t le = launch_maybe_pdl(final_ln_kernel, dim3(p->B), dim3(FL_THREADS), 0, (cudaStream_t)stream, *p);
  if (le != cudaSuccess) DTTS_FAIL(-3, "final_ln launch failed: %s", cudaGetErrorString(le));
  DTTS_CHECK_LAUNCH("final_ln");
  return 0;
}
