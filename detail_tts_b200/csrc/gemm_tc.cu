// gemm_tc.cu -- persistent, warp-specialised tcgen05 GEMM / multi-tap conv-GEMM for sm_100a.
//
//   D[m, n] = epilogue( sum_{tap} sum_k A[m + shift(tap), k] * W[tap*N + n, k] )
//
// A: fp16 activations in rows layout (K-major), W: fp16 weights [taps*N, K] (K-major).
// Operands are staged by TMA (cp.async.bulk.tensor, 128B swizzle) into a multi-stage shared-memory
// ring; one elected thread issues tcgen05.mma (M=128, N=BN, K=16) with fp32 accumulators in TMEM;
// two accumulator stages let the 4 epilogue warps (tcgen05.ld -> bias/act/residual -> global)
// overlap the next tile's main loop.  Conv taps are extra K-iterations whose A tile is the same
// tensor map read at a shifted row coordinate: TMA's out-of-bounds zero fill and the zero
// separator rows of the rows layout implement the reference's per-utterance zero padding.
//
// Replaces nn.Conv1d/nn.Linear/ConvTranspose1d + their eager epilogues (see include/dtts.h).
#include <stdlib.h>
#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "tcgen05.cuh"

using namespace dtts_tc;

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // fp16 elements = 128 bytes = one swizzle span
constexpr int A_BYTES = BM * BK * 2;
constexpr int EPI_WARPS = 8;      // two warps per TMEM lane quarter, alternating 32-column chunks
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;  // warp0 TMA, warp1 MMA(+TMEM alloc), warps2.. epilogue
constexpr int EPI_LD = 36;        // padded fp32 row of the per-warp 32x32 transpose staging tile
constexpr int EPI_LD_W = 68;      // ... of the 32x64 tile of the fp16-only epilogue (full 128-byte output lines per row)
constexpr int epi_stage_bytes(int bn, bool tf32) { return EPI_WARPS * 32 * ((bn >= 256 && !tf32) ? EPI_LD_W : EPI_LD) * 4; }

// TF32 = true: 3xTF32 fp32-class GEMM.  Operands are fp32 split on the host side of the ABI into a
// tf32-exact high part and a low part (x = hi + lo); a stage holds {A_hi, A_lo, W_hi, W_lo} tiles (32 fp32 =
// 128 B rows, same swizzle span) and the MMA warp issues hi*hi + lo*hi + hi*lo (kind::tf32, K=8 per MMA).
// CTAS = 2: cta_group::2 -- a pair of CTAs (one cluster, two SMs of a TPC) computes a 256 x BN tile; each CTA stages its own
// 128 rows of A and HALF of the B tile, the leader's tcgen05.mma reads both halves, each CTA's TMEM receives its 128 rows
// of the accumulator.  Shared-memory operand traffic per SM per MMA drops from 4 + BN/32 KB to 4 + BN/64 KB.
// EPI: 0 = register/LSU epilogue (every activation / row map / accumulate variant), 1 = fp16-only output through per-warp
// TMA stores, 2 = fp32 output (+ fp32 residual) through per-warp TMA loads and stores (see the epilogue below).
constexpr int EPI_BUF_BYTES = 4096;      // one per-warp TMA box: 32 rows x 128 bytes (32 fp32 or 64 fp16 columns), 128B swizzle
// boxes per epilogue warp: EPI 2 keeps two residual boxes in flight; EPI 1 double-buffers its stores where the 32 KB stages of
// the CTA-pair tiles leave room (a single-CTA 256-wide tile would drop to three operand stages)
constexpr int epi_bufs(int epi, int ctas) { return epi == 2 ? 3 : (ctas == 2 ? 2 : 1); }
constexpr int BAR_BYTES = 512;
template <int BN, bool TF32 = false, int CTAS = 1, int EPI = 0>
struct Cfg {
  static constexpr int B_BYTES = (BN / CTAS) * BK * 2;   // rows of B staged by one CTA x 128 B, for fp16 (64 el) and tf32 (32 el) alike
  static constexpr int STAGE_BYTES = (A_BYTES + B_BYTES) * (TF32 ? 2 : 1);
  static constexpr int EPI_STAGE_BYTES = EPI == 0 ? epi_stage_bytes(BN, TF32) : EPI_WARPS * epi_bufs(EPI, CTAS) * EPI_BUF_BYTES;
  static constexpr int STAGES_RAW = (227 * 1024 - 1024 - BAR_BYTES - EPI_STAGE_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int ACC_COLS = 2 * BN;
  static constexpr int TMEM_COLS = ACC_COLS <= 32 ? 32 : ACC_COLS <= 64 ? 64 : ACC_COLS <= 128 ? 128 : ACC_COLS <= 256 ? 256 : 512;
  // layout: [operand stages | epilogue staging (1024-byte aligned: the stages are multiples of 1 KB) | barriers]
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + BAR_BYTES + EPI_STAGE_BYTES;
  static_assert(STAGE_BYTES % 1024 == 0, "swizzled tiles and TMA boxes need 1024-byte aligned bases");
  static_assert(STAGES >= 2, "operand ring too shallow");
};

// GroupNorm statistics of the output tile, accumulated by the epilogue (EpiParams.gn_stats): this lane's per-row partial
// sums ps / pss (its 4 or 8 columns all lie in group g) for rows mrow + 4*it of utterances uid[it] (-1 = not a row).
// Utterance ids increase with the row index, so the 32 rows of a warp touch the first utterance, the last one, and
// (only for utterances shorter than 32 frames) some in between: two shuffle-reduced segments + per-row atomics for the rest.
__device__ __forceinline__ void gn_accumulate(float* stats, int groups, int g, const int (&uid)[8], const float (&ps)[8],
                                              const float (&pss)[8], int lane) {
  int uf = 0x7fffffff, ul = -1;
#pragma unroll
  for (int it = 0; it < 8; ++it)
    if (uid[it] >= 0) { uf = min(uf, uid[it]); ul = max(ul, uid[it]); }
  uf = min(uf, __shfl_xor_sync(0xffffffffu, uf, 8)); uf = min(uf, __shfl_xor_sync(0xffffffffu, uf, 16));
  ul = max(ul, __shfl_xor_sync(0xffffffffu, ul, 8)); ul = max(ul, __shfl_xor_sync(0xffffffffu, ul, 16));
  if (ul < 0) return;
  float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int u = uid[it];
    if (u == uf) { a0 += ps[it]; a1 += pss[it]; }
    else if (u == ul) { b0 += ps[it]; b1 += pss[it]; }
    else if (u >= 0) {
      atomicAdd(stats + ((size_t)u * groups + g) * 2, ps[it]);
      atomicAdd(stats + ((size_t)u * groups + g) * 2 + 1, pss[it]);
    }
  }
  a0 += __shfl_xor_sync(0xffffffffu, a0, 8); a0 += __shfl_xor_sync(0xffffffffu, a0, 16);
  a1 += __shfl_xor_sync(0xffffffffu, a1, 8); a1 += __shfl_xor_sync(0xffffffffu, a1, 16);
  if (lane < 8) {
    atomicAdd(stats + ((size_t)uf * groups + g) * 2, a0);
    atomicAdd(stats + ((size_t)uf * groups + g) * 2 + 1, a1);
  }
  if (ul != uf) {   // warp-uniform
    b0 += __shfl_xor_sync(0xffffffffu, b0, 8); b0 += __shfl_xor_sync(0xffffffffu, b0, 16);
    b1 += __shfl_xor_sync(0xffffffffu, b1, 8); b1 += __shfl_xor_sync(0xffffffffu, b1, 16);
    if (lane < 8) {
      atomicAdd(stats + ((size_t)ul * groups + g) * 2, b0);
      atomicAdd(stats + ((size_t)ul * groups + g) * 2 + 1, b1);
    }
  }
}

// ------------------------------------------------------------------------------------ the kernel
template <int BN, bool TF32, int CTAS = 1, bool GN = false, int EPI = 0>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
               const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmW2,
               const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmR,
               EpiParams epi, const int K, const int taps, const int tap_shift0, const int tap_stride,
               const int kb_per_split, const long split_stride, const int debug, const int slab_bytes, const int n_slab, const int n_wst) {
  using C = Cfg<BN, TF32, CTAS, EPI>;
  static_assert(EPI == 0 || !TF32, "the TMA epilogues are fp16-GEMM only");
  static_assert(CTAS == 1 || !TF32, "the CTA-pair variant is fp16 only");
  uint32_t cta_rank = 0;
  if (CTAS == 2) asm("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
  const bool leader = cta_rank == 0;
  const int work_id = CTAS == 2 ? blockIdx.x >> 1 : blockIdx.x;        // tile stream of this CTA (pair)
  const int work_stride = CTAS == 2 ? gridDim.x >> 1 : gridDim.x;
  constexpr int BKE = TF32 ? 32 : 64;   // elements per 128-byte K block
  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzled tiles must start on 1024-byte boundaries of the SHARED address space
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* epi_region = smem + C::STAGES * C::STAGE_BYTES;      // 1024-byte aligned
  uint64_t* bars = (uint64_t*)(epi_region + C::EPI_STAGE_BYTES);
  uint64_t* full = bars;                       // [8]
  uint64_t* empty = bars + 8;                  // [8]
  uint64_t* tfull = bars + 16;                 // [2]
  uint64_t* tempty = bars + 18;                // [2]
  uint32_t* tmem_slot = (uint32_t*)(bars + 20);
  uint64_t* ebars = bars + 22;                 // [EPI_WARPS][3]: residual boxes landed (EPI == 2)
  // Slab mode (multi-tap convs, slab_bytes > 0; see the producer): the operand space is re-cut into a ring of n_slab A slabs and
  // a ring of n_wst weight tiles; full / empty above serve the weight ring (n_wst <= 8), these the slab ring (n_slab <= 4).
  uint64_t* sfull = ebars + EPI_WARPS * 3;     // [4]
  uint64_t* sempty = sfull + 4;                // [4]
  static_assert((2 * 8 + 6 + EPI_WARPS * 3 + 8) * 8 <= BAR_BYTES, "barrier block too small");
  const bool slab = !TF32 && slab_bytes > 0;
  uint8_t* wring = smem + n_slab * slab_bytes; // weight ring behind the slabs (slab mode)
  float* epi_stage = (float*)epi_region;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = (epi.M + CTAS * BM - 1) / (CTAS * BM);
  const int n_tiles = (epi.N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  // split-K: blockIdx.y owns K blocks [kb0, kb0 + k_blocks) and writes a raw partial tile set at
  // out_f32 + blockIdx.y * split_stride (the host strips bias/act/residual from `epi` in that mode)
  const int k_blocks_all = (K + BKE - 1) / BKE;
  const int kb0 = blockIdx.y * kb_per_split;
  const int k_blocks = min(kb_per_split, k_blocks_all - kb0);
  const int k_iters = taps * k_blocks;
  if (epi.out_f32) epi.out_f32 += (long)blockIdx.y * split_stride;

  if (threadIdx.x == 0) {
    for (int s = 0; s < 8; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 4; ++s) { mbar_init(&sfull[s], 1); mbar_init(&sempty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], CTAS * EPI_WARPS); }
    if (EPI == 2) for (int i = 0; i < EPI_WARPS * 3; ++i) mbar_init(&ebars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    if (CTAS == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)C::TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)C::TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CTAS == 2) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");   // peer barriers initialised
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // programmatic dependent launch (no-ops for a normal launch): the prologue above overlapped the predecessor's tail; nothing
  // below reads memory the predecessor wrote before this wait (the decode step of the GPT and, round 2, the diffusion eval)
  pdl_launch();
  pdl_wait();

  if (warp == 0) {
    // ================================ TMA producer =================================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmW) : "memory");
      if (TF32) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA2) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmW2) : "memory");
      }
      if (EPI != 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmO) : "memory");
        if (EPI == 2) asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmR) : "memory");
      }
      int stage = 0; uint32_t phase = 0;
      if (slab) {
        // Multi-tap convs: per K block ONE slab of A rows [m0 + shift0, m0 + shift0 + 128 + (taps-1)*stride) (a single TMA box) serves
        // every tap -- tap t is the same slab read from row t*stride on (a shifted UMMA descriptor) -- instead of one shifted
        // 128-row tile per tap: the A traffic of a k3 conv drops 3x, of the vocoder's k11 convs 11x.  Weight tiles stream
        // through their own ring, tap by tap.
        int ss = 0; uint32_t sphase = 0;
        for (int tile = work_id; tile < num_tiles; tile += work_stride) {
          const int m0 = (tile / n_tiles) * (CTAS * BM) + (int)cta_rank * BM;
          const int n0 = (tile % n_tiles) * BN + (int)cta_rank * (BN / CTAS);
          for (int kb = 0; kb < k_blocks; ++kb) {
            mbar_wait(&sempty[ss], sphase ^ 1);
            if (CTAS == 2) {
              if (leader) mbar_expect_tx(&sfull[ss], 2 * slab_bytes);
              tma_load_2d_pair(smem + ss * slab_bytes, &tmA, &sfull[ss], (kb0 + kb) * BKE, m0 + tap_shift0);
            } else {
              mbar_expect_tx(&sfull[ss], slab_bytes);
              tma_load_2d(smem + ss * slab_bytes, &tmA, &sfull[ss], (kb0 + kb) * BKE, m0 + tap_shift0);
            }
            if (++ss == n_slab) { ss = 0; sphase ^= 1; }
            for (int tap = 0; tap < taps; ++tap) {
              mbar_wait(&empty[stage], phase ^ 1);
              if (CTAS == 2) {
                if (leader) mbar_expect_tx(&full[stage], 2 * C::B_BYTES);
                tma_load_2d_pair(wring + stage * C::B_BYTES, &tmW, &full[stage], (kb0 + kb) * BKE, tap * epi.N + n0);
              } else {
                mbar_expect_tx(&full[stage], C::B_BYTES);
                tma_load_2d(wring + stage * C::B_BYTES, &tmW, &full[stage], (kb0 + kb) * BKE, tap * epi.N + n0);
              }
              if (++stage == n_wst) { stage = 0; phase ^= 1; }
            }
          }
        }
      } else
      for (int tile = work_id; tile < num_tiles; tile += work_stride) {
        const int m0 = (tile / n_tiles) * (CTAS * BM) + (int)cta_rank * BM;
        const int n0 = (tile % n_tiles) * BN + (int)cta_rank * (BN / CTAS);
        for (int it = 0; it < k_iters; ++it) {
          const int tap = it / k_blocks;
          const int kb = it - tap * k_blocks;
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * C::STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          if (CTAS == 2) {
            // both CTAs' loads complete on the LEADER's barrier (its own arrive.expect_tx covers the pair's bytes)
            if (leader) mbar_expect_tx(&full[stage], 2 * C::STAGE_BYTES);
            tma_load_2d_pair(sa, &tmA, &full[stage], (kb0 + kb) * BKE, m0 + tap_shift0 + tap * tap_stride);
            tma_load_2d_pair(sb, &tmW, &full[stage], (kb0 + kb) * BKE, tap * epi.N + n0);
            if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
          mbar_expect_tx(&full[stage], C::STAGE_BYTES);
          tma_load_2d(sa, &tmA, &full[stage], (kb0 + kb) * BKE, m0 + tap_shift0 + tap * tap_stride);
          tma_load_2d(sb, &tmW, &full[stage], (kb0 + kb) * BKE, tap * epi.N + n0);
          if (TF32) {
            tma_load_2d(sb + C::B_BYTES, &tmA2, &full[stage], (kb0 + kb) * BKE, m0 + tap_shift0 + tap * tap_stride);
            tma_load_2d(sb + C::B_BYTES + A_BYTES, &tmW2, &full[stage], (kb0 + kb) * BKE, tap * epi.N + n0);
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ===================================
    if (leader) {   // CTA pair: the leader issues for both CTAs (cta_group::2).  The WHOLE warp runs the loop (warp-uniform
                    // operands, one elected lane issues: see umma_f16_e)
      const uint32_t idesc = make_idesc(CTAS * BM, BN) | (TF32 ? ((2u << 7) | (2u << 10)) : 0u);   // a/b format: F16 = 0, TF32 = 2
      int stage = 0; uint32_t phase = 0;
      int sstage = 0; uint32_t sphase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = work_id; tile < num_tiles; tile += work_stride) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        if (slab) {
          if constexpr (!TF32) {
            for (int kb = 0; kb < k_blocks; ++kb) {
              mbar_wait(&sfull[sstage], sphase);
              const uint32_t sa0 = smem_u32(smem + sstage * slab_bytes);
              for (int tap = 0; tap < taps; ++tap) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                // rows of a 128B-swizzled tile are 128 bytes apart and the swizzle is a function of the shared-memory address,
                // so the tile that starts tap*stride rows into the slab is the slab's descriptor advanced by that many rows
                const uint32_t sa = sa0 + (uint32_t)(tap * tap_stride) * 128u;
                const uint32_t sb = smem_u32(wring + stage * C::B_BYTES);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const uint64_t da = make_desc(sa + k * 32);
                  const uint64_t db = make_desc(sb + k * 32);
                  if (CTAS == 2) umma_f16_pair_e(d_tmem, da, db, idesc, (kb > 0 || tap > 0 || k > 0) ? 1u : 0u);
                  else umma_f16_e(d_tmem, da, db, idesc, (kb > 0 || tap > 0 || k > 0) ? 1u : 0u);
                }
                if (CTAS == 2) tc_commit_pair_e(&empty[stage]);
                else tc_commit_e(&empty[stage]);
                if (++stage == n_wst) { stage = 0; phase ^= 1; }
              }
              if (CTAS == 2) tc_commit_pair_e(&sempty[sstage]);
              else tc_commit_e(&sempty[sstage]);
              if (++sstage == n_slab) { sstage = 0; sphase ^= 1; }
            }
          }
        } else
        for (int it = 0; it < k_iters; ++it) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint32_t sb = sa + A_BYTES;
          if (!TF32) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {   // 4 x (K=16 fp16 = 32 B)
              const uint64_t da = make_desc(sa + k * 32);
              const uint64_t db = make_desc(sb + k * 32);
              if (CTAS == 2) umma_f16_pair_e(d_tmem, da, db, idesc, (it > 0 || k > 0) ? 1u : 0u);
              else umma_f16_e(d_tmem, da, db, idesc, (it > 0 || k > 0) ? 1u : 0u);
            }
          } else {
            const uint32_t sa2 = sb + C::B_BYTES, sb2 = sa2 + A_BYTES;
#pragma unroll
            for (int seg = 0; seg < 3; ++seg) {   // hi*hi, lo*hi, hi*lo -- small terms last is not needed: fp32 accumulate
              const uint32_t xa = seg == 1 ? sa2 : sa, xb = seg == 2 ? sb2 : sb;
#pragma unroll
              for (int k = 0; k < 4; ++k) {   // 4 x (K=8 tf32 = 32 B)
                umma_tf32_e(d_tmem, make_desc(xa + k * 32), make_desc(xb + k * 32), idesc, (it > 0 || seg > 0 || k > 0) ? 1u : 0u);
              }
            }
          }
          if (CTAS == 2) tc_commit_pair_e(&empty[stage]);   // multicast: frees the slot in both CTAs
          else tc_commit_e(&empty[stage]);  // frees the smem slot when these MMAs retire
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        if (CTAS == 2) tc_commit_pair_e(&tfull[acc]);
        else tc_commit_e(&tfull[acc]);      // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // ================================ epilogue warps ===============================
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int cpar = (warp - 2) >> 2;  // which of the two warps of the quarter: takes chunks c = cpar, cpar+2, ...
    if constexpr (EPI != 0) {
      // ---------------------------------------------------------------- TMA epilogues (round 2, second session)
      // The register/LSU epilogue below moves a 128 x 256 fp32 tile (+ its residual) at ~19 GB/s per SM: ncu shows the epilogue
      // warps parked on the first instruction that reuses an address register of their LDG.128s (the LSU queue is backed up),
      // whatever the number of loads in flight (two-deep register prefetch and an L2 prefetch of the residual tile both
      // measured neutral).  Here every epilogue warp owns 32 accumulator rows (its TMEM lane quarter) x 128-byte chunks and
      // talks to HBM only through the TMA unit: one thread per row, the row's 128 bytes live in a 128B-swizzled per-warp box
      // (chunk k of row r at k ^ (r & 7): conflict-free for row-per-lane AND for the transposed statistics pass), so there is
      // no transpose, no per-lane address arithmetic and no LSU global access at all.
      //   EPI == 2: fp32 output (+ fp32 residual).  Three boxes per warp: the residual boxes of the next two chunks are in
      //             flight (across tile boundaries, i.e. during the next tile's main loop) while the current one is updated
      //             in place and stored.
      //   EPI == 1: fp16-only output, 64 columns per box, one box per warp.
      // Separator rows (row_utt < 0) are written as zeros: they are zero in every rows-layout buffer.
      constexpr int NB = EPI == 2 ? 3 : (CTAS == 2 ? 2 : 1);
      constexpr int CW = EPI == 2 ? 32 : 64;     // columns per chunk = one 128-byte box row
      constexpr int NCH = BN / CW;               // chunks per tile; this warp takes c = cpar, cpar + 2, ...
      const int ew = warp - 2;
      uint8_t* ebuf = epi_region + ew * NB * EPI_BUF_BYTES;
      uint64_t* ebar = ebars + ew * 3;
      const int r7 = lane & 7;
      const bool gn = GN && epi.gn_stats != nullptr;
      const int gn_groups = gn ? epi.N / epi.gn_cpg : 0;
      const bool has_res = EPI == 2 && epi.res != nullptr;
      const float alpha = epi.alpha;
      // load cursor: walks the same (tile, chunk) sequence as the compute loop, NB - 1 chunks ahead
      int ld_tile = work_id, ld_c = cpar;
      uint32_t ld_seq = 0, use_seq = 0;
      auto ld_normalize = [&]() {
        while (ld_tile < num_tiles && (ld_c >= NCH || (ld_tile % n_tiles) * BN + ld_c * CW >= epi.N)) { ld_tile += work_stride; ld_c = cpar; if (cpar >= NCH) { ld_tile = num_tiles; } }
      };
      auto issue_load = [&]() {
        if (ld_tile >= num_tiles) return;
        if (lane == 0) {
          const int b = (int)(ld_seq % NB);
          mbar_expect_tx(&ebar[b], EPI_BUF_BYTES);
          tma_load_2d(ebuf + b * EPI_BUF_BYTES, &tmR, &ebar[b], (ld_tile % n_tiles) * BN + ld_c * CW,
                      (ld_tile / n_tiles) * (CTAS * BM) + (int)cta_rank * BM + q * 32);
        }
        ++ld_seq;
        ld_c += 2;
        ld_normalize();
      };
      ld_normalize();
      if (has_res) {
#pragma unroll
        for (int i = 0; i < NB - 1; ++i) issue_load();
      }
      int acc = 0; uint32_t acc_phase = 0;
      // utterance id of this lane's row, fetched one tile ahead (its L2 latency was 7 % of the epilogue warps' time)
      auto own_utt = [&](int t) -> int {
        if (t >= num_tiles) return -1;
        const int m = (t / n_tiles) * (CTAS * BM) + (int)cta_rank * BM + q * 32 + lane;
        return m < epi.M ? (epi.row_utt ? __ldg(epi.row_utt + m) : 0) : -1;
      };
      int utt_next = own_utt(work_id);
      // ... and the utterance ids of the statistics pass' rows (lane >> 3) + 4 it
      int uid_next[8];
      auto stat_utts = [&](int t, int (&u)[8]) {
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          u[it] = -1;
          if (GN && epi.row_utt && t < num_tiles) {
            const int m = (t / n_tiles) * (CTAS * BM) + (int)cta_rank * BM + q * 32 + (lane >> 3) + it * 4;
            if (m < epi.M) u[it] = __ldg(epi.row_utt + m);
          }
        }
      };
      stat_utts(work_id, uid_next);
      // The accumulator stage goes back to the MMA issuer as soon as this warp's LAST tcgen05.ld of the tile has landed.  The arrive
      // is fence-free on the CTA-pair tiles (mbar_arrive_leader_relaxed) and comes from lane 1, which has no TMA stores in flight:
      // with the cluster-scope RELEASE arrive from lane 0 (MEMBAR + ERRBAR) the lane waited for its own stores -- 25 % of the
      // epilogue warps' samples, with the tensor pipe idle behind it.
      auto release_acc = [&](int a) {
        tc_fence_before();
        __syncwarp();
        if (lane == 1) {
          if (CTAS == 2) mbar_arrive_leader_relaxed(&tempty[a]);
          else mbar_arrive(&tempty[a]);
        }
      };
      for (int tile = work_id; tile < num_tiles; tile += work_stride) {
        const int m0 = (tile / n_tiles) * (CTAS * BM) + (int)cta_rank * BM;
        const int n0 = (tile % n_tiles) * BN;
        const bool valid = utt_next >= 0;
        utt_next = own_utt(tile + work_stride);
        int uid[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) uid[it] = uid_next[it];
        stat_utts(tile + work_stride, uid_next);
        int c_last = -1;                             // this warp's last chunk of the tile
        for (int c = cpar; c < NCH && n0 + c * CW < epi.N && debug != 1; c += 2) c_last = c;
        mbar_wait(&tfull[acc], acc_phase);
        tc_fence_after();
        if (c_last < 0) release_acc(acc);
#pragma unroll 1
        for (int c = cpar; c <= c_last; c += 2) {
          const int nc = n0 + c * CW;
          const int b = (int)(use_seq % NB);
          uint8_t* box = ebuf + b * EPI_BUF_BYTES;
          uint8_t* row = box + lane * 128;
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c * CW);
          if constexpr (EPI == 2) {
            float4 bb[8];                              // bias of the chunk's 32 columns: requested before the TMEM load
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              bb[k] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (epi.bias && nc + 4 * k < epi.N) bb[k] = __ldg(reinterpret_cast<const float4*>(epi.bias + nc) + k);
            }
            float v[32];
            tmem_ld32(taddr, v);
            if (c == c_last) release_acc(acc);
            if (has_res) {
              mbar_wait(&ebar[b], (use_seq / NB) & 1u);
            } else {
              if (lane == 0) bulk_wait_read<NB - 1>();       // the store that last used this box has read it
              __syncwarp();
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              float4* p4 = reinterpret_cast<float4*>(row + ((k ^ r7) << 4));
              float4 w = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
              w.x += bb[k].x; w.y += bb[k].y; w.z += bb[k].z; w.w += bb[k].w;
              if (has_res) { const float4 r = *p4; w.x += r.x; w.y += r.y; w.z += r.z; w.w += r.w; }
              if (alpha != 1.0f) { w.x *= alpha; w.y *= alpha; w.z *= alpha; w.w *= alpha; }
              if (!valid) w = make_float4(0.f, 0.f, 0.f, 0.f);
              *p4 = w;
            }
          } else {
            float4 bb[16];                             // bias of the chunk's 64 columns: requested before the TMEM loads
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              bb[k] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (epi.bias && nc + 4 * k < epi.N) bb[k] = __ldg(reinterpret_cast<const float4*>(epi.bias + nc) + k);
            }
            float v[32], v2[32];
            tmem_ld32(taddr, v);
            tmem_ld32(taddr + 32, v2);
            if (c == c_last) release_acc(acc);
            uint4 pk[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {        // 16-byte chunk k = columns 8k .. 8k+7
              float w[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) w[e] = k < 4 ? v[8 * k + e] : v2[8 * (k - 4) + e];
              {
                const float4 b0 = bb[2 * k], b1 = bb[2 * k + 1];
                w[0] += b0.x; w[1] += b0.y; w[2] += b0.z; w[3] += b0.w; w[4] += b1.x; w[5] += b1.y; w[6] += b1.z; w[7] += b1.w;
              }
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                if (alpha != 1.0f) w[e] *= alpha;
                if (!valid) w[e] = 0.f;
              }
              __half2 h0 = __floats2half2_rn(w[0], w[1]), h1 = __floats2half2_rn(w[2], w[3]);
              __half2 h2 = __floats2half2_rn(w[4], w[5]), h3 = __floats2half2_rn(w[6], w[7]);
              pk[k].x = *reinterpret_cast<uint32_t*>(&h0); pk[k].y = *reinterpret_cast<uint32_t*>(&h1);
              pk[k].z = *reinterpret_cast<uint32_t*>(&h2); pk[k].w = *reinterpret_cast<uint32_t*>(&h3);
            }
            if (lane == 0) bulk_wait_read<NB - 1>();       // the store that last used this box has read it
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 8; ++k) *reinterpret_cast<uint4*>(row + ((k ^ r7) << 4)) = pk[k];
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0 && debug != 2) {            // DTTS_GEMM_DEBUG=2: everything but the stores (timing experiments)
            tma_store_2d(&tmO, box, nc, m0 + q * 32);
            bulk_commit();
          }
          if (has_res) {
            // refill the box of the PREVIOUS chunk (its store is the older of the two pending groups) with the residual of the
            // chunk after the next, before the statistics pass below delays it; every lane's reads of that box (the previous
            // chunk's statistics pass) precede the async write
            if (lane == 0) bulk_wait_read<1>();
            __syncwarp();
            issue_load();
          }
          if (GN && gn) {
            // statistics of the stored values (fp32 / the fp16-rounded outputs): transposed read of the box, lane = rows
            // (lane >> 3) + 4 it, 16-byte chunk lane & 7 (4 fp32 / 8 fp16 columns, all inside one group: gn_cpg % 8 == 0)
            float ps[8], pss[8];
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              const int rr = (lane >> 3) + it * 4;
              const uint4 t = *reinterpret_cast<const uint4*>(box + rr * 128 + ((r7 ^ (rr & 7)) << 4));
              if constexpr (EPI == 2) {
                const float a0 = __uint_as_float(t.x), a1 = __uint_as_float(t.y), a2 = __uint_as_float(t.z), a3 = __uint_as_float(t.w);
                ps[it] = (a0 + a1) + (a2 + a3);
                pss[it] = fmaf(a0, a0, fmaf(a1, a1, fmaf(a2, a2, a3 * a3)));
              } else {
                const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&t.x)), f1 = __half22float2(*reinterpret_cast<const __half2*>(&t.y));
                const float2 f2 = __half22float2(*reinterpret_cast<const __half2*>(&t.z)), f3 = __half22float2(*reinterpret_cast<const __half2*>(&t.w));
                ps[it] = ((f0.x + f0.y) + (f1.x + f1.y)) + ((f2.x + f2.y) + (f3.x + f3.y));
                pss[it] = fmaf(f0.x, f0.x, fmaf(f0.y, f0.y, fmaf(f1.x, f1.x, fmaf(f1.y, f1.y, fmaf(f2.x, f2.x, fmaf(f2.y, f2.y, fmaf(f3.x, f3.x, f3.y * f3.y)))))));
              }
            }
            gn_accumulate(epi.gn_stats, gn_groups, (nc + r7 * (EPI == 2 ? 4 : 8)) / epi.gn_cpg, uid, ps, pss, lane);
          }
          ++use_seq;
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      if (lane == 0) bulk_wait_all();          // shared memory must outlive the last stores
    } else {
    constexpr bool WIDE_OK = BN >= 256 && !TF32;
    float* st = epi_stage + (warp - 2) * 32 * (WIDE_OK ? EPI_LD_W : EPI_LD);
    const int cc = (lane & 7) * 4;
    // fast path: plain/act epilogue on 4 consecutive columns per lane, everything 16-byte aligned
    const bool fast = epi.act < DTTS_ACT_PAIR_TANH_SIGMOID && !epi.out_row_map && !epi.accumulate && (epi.N & 3) == 0 &&
                      (!epi.res || ((epi.ldr & 3) == 0 && (((uintptr_t)epi.res) & 15) == 0)) &&
                      (!epi.out_f32 || ((epi.ldo32 & 3) == 0 && (((uintptr_t)epi.out_f32) & 15) == 0)) &&
                      (!epi.out_f16 || ((epi.ldo16 & 3) == 0 && (((uintptr_t)epi.out_f16) & 7) == 0)) &&
                      (!epi.bias || (((uintptr_t)epi.bias) & 15) == 0) && (!epi.bias_utt || (((uintptr_t)epi.bias_utt) & 15) == 0);
    // lean path (every diffusion / vocoder conv): bias + residual + alpha, fp32 and/or fp16 stores (optional
    // leaky-ReLU on the fp16 copy).  ~25 instructions per 4 outputs: pointers advance by constant strides, row
    // validity comes from a per-tile bitmask loaded while the MMAs still run, all residual loads are issued first.
    const bool lean = fast && epi.act == DTTS_ACT_NONE && (epi.act16 == DTTS_ACT_NONE || epi.act16 == DTTS_ACT_LRELU) && !epi.bias_utt;
    if (GN && !lean) __trap();     // the statistics are only accumulated by the lean epilogue: refuse instead of silently skipping them
    const bool wide16 = lean && !epi.out_f32 && !epi.res && epi.out_f16 && (epi.ldo16 & 7) == 0 && (((uintptr_t)epi.out_f16) & 15) == 0 &&
                        (epi.N & 63) == 0 && debug != 4;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = work_id; tile < num_tiles; tile += work_stride) {
      const int m0 = (tile / n_tiles) * (CTAS * BM) + (int)cta_rank * BM;
      const int n0 = (tile % n_tiles) * BN;
      const int mrow = m0 + q * 32 + (lane >> 3);   // + it*4
      uint32_t vmask = 0;
      int uid[8];          // utterance of row mrow + 4*it (only needed for the GroupNorm statistics)
      if (lean) {
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int m = mrow + it * 4;
          const int u = m < epi.M ? (epi.row_utt ? __ldg(epi.row_utt + m) : 0) : -1;
          uid[it] = u;
          vmask |= u >= 0 ? (1u << it) : 0u;
        }
      }
      const bool gn = GN && lean && epi.gn_stats != nullptr;     // GN = false instantiations carry none of the statistics code
      const int gn_groups = gn ? epi.N / epi.gn_cpg : 0;
      // Residual rows of this warp's first 32-column chunk: they do not depend on the accumulator, so they are requested
      // BEFORE waiting for the MMAs (their HBM latency hides behind the main loop); inside the chunk loop the next chunk's
      // residual is requested before the current chunk is processed.  (The 1x1 + residual convs are HBM-bound: with the loads
      // issued chunk by chunk after the accumulator was ready the epilogue ran at ~18 GB/s per SM, 162 us vs 86 us at HBM peak.)
      const bool res_pf = lean && epi.res != nullptr && !(WIDE_OK && wide16);
      float4 rs_cur[8];
      auto load_res = [&](int c, float4 (&rs)[8]) {
        const int nn = n0 + c * 32 + cc;
        const bool ok = c < BN / 32 && nn < epi.N;
        const float* rp = epi.res + (size_t)mrow * epi.ldr + nn;
        const size_t rstep = (size_t)4 * epi.ldr;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          rs[it] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ok && ((vmask >> it) & 1u)) rs[it] = *reinterpret_cast<const float4*>(rp);
          rp += rstep;
        }
      };
      if (res_pf) load_res(cpar, rs_cur);
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      if (WIDE_OK && wide16 && debug != 1) {
        // fp16-only outputs (c1 / qkv convs): 64 columns per pass, a lane owns 8 consecutive columns of a row, so every store
        // is 16 bytes and 8 lanes write one full 128-byte line (the 32-column path writes half lines from two warps)
        const int c8 = (lane & 7) * 8;
        const bool lrelu16 = epi.act16 == DTTS_ACT_LRELU;
        const float slope = epi.act16_param, alpha = epi.alpha;
#pragma unroll 1
        for (int gi = cpar; gi < BN / 64; gi += 2) {
          const int ng = n0 + gi * 64;
          if (ng >= epi.N) break;
#pragma unroll
          for (int hc = 0; hc < 2; ++hc) {
            float v[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + gi * 64 + hc * 32);
            tmem_ld32(taddr, v);
            float4* srow = reinterpret_cast<float4*>(st + lane * EPI_LD_W + hc * 32);
#pragma unroll
            for (int k = 0; k < 8; ++k) srow[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
          }
          __syncwarp();
          if (debug != 2) {
            float4 ba = make_float4(0.f, 0.f, 0.f, 0.f), bb = ba;
            if (epi.bias) { ba = __ldg(reinterpret_cast<const float4*>(epi.bias + ng + c8)); bb = __ldg(reinterpret_cast<const float4*>(epi.bias + ng + c8 + 4)); }
            __half* o16 = epi.out_f16 + (size_t)mrow * epi.ldo16 + ng + c8;
            const size_t s16 = (size_t)4 * epi.ldo16;
            const float* sp = st + (lane >> 3) * EPI_LD_W + c8;
            float ps[8], pss[8];
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              ps[it] = 0.f; pss[it] = 0.f;
              if ((vmask >> it) & 1u) {
                const float4 t0 = *reinterpret_cast<const float4*>(sp + it * 4 * EPI_LD_W);
                const float4 t1 = *reinterpret_cast<const float4*>(sp + it * 4 * EPI_LD_W + 4);
                float w[8] = {t0.x + ba.x, t0.y + ba.y, t0.z + ba.z, t0.w + ba.w, t1.x + bb.x, t1.y + bb.y, t1.z + bb.z, t1.w + bb.w};
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  if (alpha != 1.0f) w[e] *= alpha;
                  if (GN) { ps[it] += w[e]; pss[it] = fmaf(w[e], w[e], pss[it]); }
                  if (lrelu16) w[e] = w[e] > 0.f ? w[e] : w[e] * slope;
                }
                __half2 h0 = __floats2half2_rn(w[0], w[1]), h1 = __floats2half2_rn(w[2], w[3]);
                __half2 h2 = __floats2half2_rn(w[4], w[5]), h3 = __floats2half2_rn(w[6], w[7]);
                uint4 pk;
                pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
                pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
                *reinterpret_cast<uint4*>(o16) = pk;
              }
              o16 += s16;
            }
            if (gn) gn_accumulate(epi.gn_stats, gn_groups, (ng + c8) / epi.gn_cpg, uid, ps, pss, lane);
          }
          __syncwarp();
        }
      } else
      // TMEM lane = output row.  Each 32x32 chunk is transposed through a per-warp shared-memory tile so that
      // a lane owns 4 CONSECUTIVE columns of a row: residual loads and fp32/fp16 stores are 128 B / 64 B
      // contiguous per row (8 lanes).
#pragma unroll 1
      for (int c = cpar; c < (debug == 1 ? 0 : BN / 32); c += 2) {
        const int n = n0 + c * 32 + cc;
        if (n0 + c * 32 >= epi.N) break;
        if (debug != 3) {
          float v[32];
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c * 32);
          tmem_ld32(taddr, v);
          float4* srow = reinterpret_cast<float4*>(st + lane * EPI_LD);
#pragma unroll
          for (int k = 0; k < 8; ++k) srow[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
          __syncwarp();
        }
        if (debug == 2) continue;
        if (lean) {
          if (n < epi.N) {
            float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (epi.bias) b4 = __ldg(reinterpret_cast<const float4*>(epi.bias + n));
            float4 (&rs)[8] = rs_cur;            // requested before the accumulator wait (first chunk) / right after the previous chunk
            float* o32 = epi.out_f32 ? epi.out_f32 + (size_t)mrow * epi.ldo32 + n : nullptr;
            __half* o16 = epi.out_f16 ? epi.out_f16 + (size_t)mrow * epi.ldo16 + n : nullptr;
            const size_t s32 = (size_t)4 * epi.ldo32, s16 = (size_t)4 * epi.ldo16;
            const float* sp = st + (lane >> 3) * EPI_LD + cc;
            const bool lrelu16 = epi.act16 == DTTS_ACT_LRELU;
            const float slope = epi.act16_param, alpha = epi.alpha;
            float ps[8], pss[8];
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              ps[it] = 0.f; pss[it] = 0.f;
              if ((vmask >> it) & 1u) {
                const float4 t = *reinterpret_cast<const float4*>(sp + it * 4 * EPI_LD);
                float w0 = t.x + b4.x, w1 = t.y + b4.y, w2 = t.z + b4.z, w3 = t.w + b4.w;
                if (epi.res) { w0 += rs[it].x; w1 += rs[it].y; w2 += rs[it].z; w3 += rs[it].w; }
                if (alpha != 1.0f) { w0 *= alpha; w1 *= alpha; w2 *= alpha; w3 *= alpha; }
                if (GN) {
                  ps[it] = (w0 + w1) + (w2 + w3);
                  pss[it] = fmaf(w0, w0, fmaf(w1, w1, fmaf(w2, w2, w3 * w3)));
                }
                if (o32) *reinterpret_cast<float4*>(o32) = make_float4(w0, w1, w2, w3);
                if (o16) {
                  if (lrelu16) {
                    w0 = w0 > 0.f ? w0 : w0 * slope; w1 = w1 > 0.f ? w1 : w1 * slope;
                    w2 = w2 > 0.f ? w2 : w2 * slope; w3 = w3 > 0.f ? w3 : w3 * slope;
                  }
                  __half2 h0 = __floats2half2_rn(w0, w1), h1 = __floats2half2_rn(w2, w3);
                  uint2 pk;
                  pk.x = *reinterpret_cast<uint32_t*>(&h0);
                  pk.y = *reinterpret_cast<uint32_t*>(&h1);
                  *reinterpret_cast<uint2*>(o16) = pk;
                }
              }
              if (o32) o32 += s32;
              if (o16) o16 += s16;
            }
            if (epi.res) load_res(c + 2, rs_cur);   // next chunk of this warp: in flight during its tcgen05.ld + transpose
            if (gn) gn_accumulate(epi.gn_stats, gn_groups, n / epi.gn_cpg, uid, ps, pss, lane);
          }
        } else {
#pragma unroll 1
          for (int it = 0; it < 8; ++it) {
            const int r = it * 4 + (lane >> 3);
            const float4 t = *reinterpret_cast<const float4*>(st + r * EPI_LD + cc);
            float w[4] = {t.x, t.y, t.z, t.w};
            epilogue_chunk<4>(epi, m0 + q * 32 + r, n, w);
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CTAS == 2) mbar_arrive_leader_relaxed(&tempty[acc]);   // the leader's issuer waits for both CTAs' epilogues
        else mbar_arrive(&tempty[acc]);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    }   // EPI == 0
  }

  tc_fence_before();
  __syncthreads();
  if (CTAS == 2) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    if (CTAS == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

struct MapKey {
  const void* ptr; int rows, cols, ld, box_rows, esize, box_cols;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows && esize == o.esize && box_cols == o.box_cols;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = (size_t)k.ptr;
    h = h * 1000003u ^ (size_t)k.rows; h = h * 1000003u ^ (size_t)k.cols;
    h = h * 1000003u ^ (size_t)k.ld; h = h * 1000003u ^ (size_t)k.box_rows; h = h * 1000003u ^ (size_t)k.esize; h = h * 1000003u ^ (size_t)k.box_cols;
    return h;
  }
};
std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;
std::mutex g_maps_mu;

int g_sm_count = 0;
int g_debug = -1;   // DTTS_GEMM_DEBUG=1: skip the epilogue (main-loop timing only; results are garbage)

// CTA-pair (cta_group::2) launch: fp16 only, 256 x BN tiles, one 2-CTA cluster per tile stream
// Slab mode of a multi-tap conv (see the producer): rows of the A slab (multiple of 8, <= 256 = the TMA box limit) and the split
// of the operand space into n_slab slabs + n_wst weight tiles.  Returns false when the conv keeps the tile-per-tap scheme.
struct SlabGeo { int rows = 0, bytes = 0, n_slab = 0, n_wst = 0; };
inline bool slab_geometry(const dtts_gemm_params* p, int space, int b_bytes, SlabGeo* g) {
  static int on = -1;
  if (on < 0) { const char* e = getenv("DTTS_GEMM_SLAB"); on = e ? atoi(e) : 1; }
  if (!on || p->taps < 2 || p->tap_stride < 1) return false;
  const int rows = (BM + (p->taps - 1) * p->tap_stride + 7) & ~7;
  if (rows > 256) return false;
  g->rows = rows; g->bytes = rows * 128; g->n_slab = 2;
  int n = (space - g->n_slab * g->bytes) / b_bytes;
  g->n_wst = n > 8 ? 8 : n;
  return g->n_wst >= 3;
}

// tensor maps of the TMA epilogues: per-warp boxes of 32 rows x 128 bytes over the output (EPI 1: fp16, EPI 2: fp32) and the
// fp32 residual.  Unused maps alias the A map (never dereferenced).
template <int EPI>
int epi_maps(const dtts_gemm_params* p, const CUtensorMap& dflt, CUtensorMap* mo, CUtensorMap* mr) {
  *mo = dflt; *mr = dflt;
  if (EPI == 1) return get_map(p->out_f16, p->M, p->N, p->ldo16, 32, mo, 2);
  if (EPI == 2) {
    int rc = get_map(p->out_f32, p->M, p->N, p->ldo32, 32, mo, 4);
    if (rc) return rc;
    if (p->res) return get_map(p->res, p->M, p->N, p->ldr, 32, mr, 4);
  }
  return 0;
}

template <int BN, bool GN = false, int EPI = 0>
int launch_pair(const dtts_gemm_params* p, cudaStream_t st) {
  using C = Cfg<BN, false, 2, EPI>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN, false, 2, GN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) DTTS_FAIL(-3, "cudaFuncSetAttribute(gemm_tc pair<%d>): %s", BN, cudaGetErrorString(e));
    attr_set = true;
  }
  CUtensorMap ma, mw;
  SlabGeo sg;
  const bool use_slab = slab_geometry(p, C::STAGES * C::STAGE_BYTES, C::B_BYTES, &sg);
  int rc = get_map(p->A, p->M, p->K, p->lda, use_slab ? sg.rows : BM, &ma, 2);
  if (rc) return rc;
  rc = get_map(p->W, p->taps * p->N, p->K, p->ldw, BN / 2, &mw, 2);
  if (rc) return rc;
  CUtensorMap mo, mr;
  rc = epi_maps<EPI>(p, ma, &mo, &mr);
  if (rc) return rc;
  const int pairs = ceil_div(p->M, 2 * BM) * ceil_div(p->N, BN);
  const int kb_all = ceil_div(p->K, 64);
  EpiParams e = make_epi(p);
  int grid = 2 * pairs < g_sm_count ? 2 * pairs : (g_sm_count & ~1);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_dtts_pdl ? 2 : 1;
  cudaError_t le = cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, false, 2, GN, EPI>, ma, mw, ma, mw, mo, mr, e, p->K, p->taps, p->tap_shift0, p->tap_stride,
                                      kb_all, (long)0, g_debug, use_slab ? sg.bytes : 0, sg.n_slab, sg.n_wst);
  if (le != cudaSuccess) DTTS_FAIL(-3, "gemm_tc pair launch failed: %s", cudaGetErrorString(le));
  DTTS_CHECK_LAUNCH("gemm_tc_pair");
  return 0;
}

template <int BN, bool TF32, bool GN = false, int EPI = 0>
int launch(const dtts_gemm_params* p, cudaStream_t st) {
  using C = Cfg<BN, TF32, 1, EPI>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN, TF32, 1, GN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) DTTS_FAIL(-3, "cudaFuncSetAttribute(gemm_tc<%d>): %s", BN, cudaGetErrorString(e));
    attr_set = true;
  }
  const int es = TF32 ? 4 : 2;
  CUtensorMap ma, mw, ma2, mw2;
  SlabGeo sg;
  const bool use_slab = !TF32 && slab_geometry(p, C::STAGES * C::STAGE_BYTES, C::B_BYTES, &sg);
  int rc = get_map(p->A, p->M, p->K, p->lda, use_slab ? sg.rows : BM, &ma, es);
  if (rc) return rc;
  rc = get_map(p->W, p->taps * p->N, p->K, p->ldw, BN, &mw, es);
  if (rc) return rc;
  ma2 = ma; mw2 = mw;
  if (TF32) {
    rc = get_map(p->A_lo, p->M, p->K, p->lda, BM, &ma2, es);
    if (rc) return rc;
    rc = get_map(p->W_lo, p->taps * p->N, p->K, p->ldw, BN, &mw2, es);
    if (rc) return rc;
  }
  CUtensorMap mo, mr;
  rc = epi_maps<EPI>(p, ma, &mo, &mr);
  if (rc) return rc;
  const int tiles = ceil_div(p->M, BM) * ceil_div(p->N, BN);
  const int kb_all = ceil_div(p->K, TF32 ? 32 : 64);
  int splits = TF32 && p->split_k > 1 ? p->split_k : 1;
  if (splits > kb_all) splits = kb_all;
  const int kb_per = ceil_div(kb_all, splits);
  splits = ceil_div(kb_all, kb_per);
  EpiParams e = make_epi(p);
  if (TF32 && p->split_k > 1) {   // raw partials: the reduce kernel applies bias / activation / residual
    e.bias = nullptr; e.bias_utt = nullptr; e.res = nullptr; e.out_f16 = nullptr; e.out_row_map = nullptr;
    e.act = DTTS_ACT_NONE; e.act16 = DTTS_ACT_NONE; e.alpha = 1.0f; e.accumulate = 0; e.row_utt = nullptr;
  }
  dim3 grid(tiles < g_sm_count ? tiles : g_sm_count, splits);
  {
    cudaError_t le = launch_maybe_pdl(gemm_tc_kernel<BN, TF32, 1, GN, EPI>, grid, dim3(NUM_THREADS), (size_t)C::SMEM_BYTES, st, ma, mw, ma2, mw2, mo, mr, e,
                                      p->K, p->taps, p->tap_shift0, p->tap_stride, kb_per, (long)p->split_stride, g_debug,
                                      use_slab ? sg.bytes : 0, sg.n_slab, sg.n_wst);
    if (le != cudaSuccess) DTTS_FAIL(-3, "gemm_tc launch failed: %s", cudaGetErrorString(le));
  }
  DTTS_CHECK_LAUNCH("gemm_tc");
  return 0;
}

int common_checks(const dtts_gemm_params* p, const char* who, int ld_mult) {
  DTTS_REQUIRE(p && p->A && p->W, "%s: null operand", who);
  DTTS_REQUIRE(p->M > 0 && p->N > 0 && p->K > 0 && p->taps >= 1, "%s: bad shape M=%d N=%d K=%d taps=%d", who, p->M, p->N, p->K, p->taps);
  DTTS_REQUIRE((p->lda % ld_mult) == 0 && (p->ldw % ld_mult) == 0, "%s: lda/ldw must keep 16-byte TMA strides", who);
  DTTS_REQUIRE((((uintptr_t)p->A) & 15) == 0 && (((uintptr_t)p->W) & 15) == 0, "%s: operands must be 16-byte aligned", who);
  DTTS_REQUIRE(p->lda >= p->K && p->ldw >= p->K, "%s: leading dimension smaller than K", who);
  DTTS_REQUIRE(!(p->bias_utt && !p->row_utt), "%s: bias_utt requires row_utt", who);
  DTTS_REQUIRE(p->out_f32 || p->out_f16, "%s: no output", who);
  DTTS_REQUIRE(!(p->act >= DTTS_ACT_PAIR_TANH_SIGMOID && (p->N & 1)), "%s: pair activation needs even N", who);
  DTTS_REQUIRE(!p->gn_stats || (p->row_utt && p->gn_cpg > 0 && p->gn_cpg % 8 == 0 && p->N % p->gn_cpg == 0 && p->act == DTTS_ACT_NONE &&
                                (p->act16 == DTTS_ACT_NONE || p->act16 == DTTS_ACT_LRELU) && !p->bias_utt && !p->out_row_map && !p->accumulate &&
                                (p->N & 3) == 0),
               "%s: gn_stats needs row_utt, gn_cpg %% 8 == 0 and the plain bias/residual epilogue", who);
  if (g_debug < 0) {
    const char* d = getenv("DTTS_GEMM_DEBUG");   // 1: no epilogue, 2: no global stores, 4: no 64-column fp16 epilogue
    g_debug = d ? atoi(d) : 0;
  }
  if (!g_sm_count) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (g_sm_count <= 0) DTTS_FAIL(-6, "%s: no CUDA device", who);
  }
  return 0;
}

}  // namespace

// 2D tensor map (fp16: esize 2, fp32: esize 4) over a row-major [rows, cols] matrix (ld elements),
// box [box_rows, 128 bytes], 128B swizzle.
int dtts_tc::get_map(const void* ptr, int rows, int cols, int ld, int box_rows, CUtensorMap* out, int esize) {
  return get_map_ex(ptr, rows, cols, ld, box_rows, 0, out, esize);
}

// box_cols == 0: box [box_rows, 128 bytes] with 128B swizzle (tcgen05 operand tiles); otherwise a plain
// [box_rows, box_cols] box without swizzle (rows land contiguously in shared memory).
int dtts_tc::get_map_ex(const void* ptr, int rows, int cols, int ld, int box_rows, int box_cols, CUtensorMap* out, int esize) {
  MapKey key{ptr, rows, cols, ld, box_rows, esize, box_cols};
  {
    std::lock_guard<std::mutex> g(g_maps_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) { *out = it->second; return 0; }
  }
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) DTTS_FAIL(-4, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * esize};
  cuuint32_t box[2] = {(cuuint32_t)(box_cols ? box_cols : 128 / esize), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, esize == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    DTTS_FAIL(-5, "cuTensorMapEncodeTiled failed (%d) rows=%d cols=%d ld=%d box_rows=%d ptr=%p", (int)r, rows, cols, ld, box_rows, ptr);
  std::lock_guard<std::mutex> g(g_maps_mu);
  if (g_maps.size() > 65536) g_maps.clear();
  g_maps[key] = *out;
  return 0;
}



extern "C" int dtts_gemm_f16_tc(const dtts_gemm_params* p, void* stream) {
  int rc = common_checks(p, "gemm_f16_tc", 8);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int N = p->N;
  static int bn256 = -1;   // 256-wide tiles (less shared-memory read traffic per MMA): +5 % on the diffusion shapes
  if (bn256 < 0) { const char* e = getenv("DTTS_GEMM_BN256"); bn256 = e ? atoi(e) : 1; }
  // cta_group::2 tiles (256 x 256 per CTA pair) for the long-K GEMMs (k3 convs, K = 1536): +8..12 % measured; the K = 768
  // 1x1 convs are epilogue / HBM-bound and lose 3-5 % to the pair synchronisation, so they stay on single-CTA tiles
  static int pair = -1;
  if (pair < 0) { const char* e = getenv("DTTS_GEMM_PAIR"); pair = e ? atoi(e) : 1; }
  // Tile shape by wave quantisation: cost = waves over the SMs x tile width / relative efficiency of the shape.  At the
  // bench shape (M = 72k rows) the 256-wide shapes win; on a 1/8 shard (M = 9k) 256-wide tiles leave the second wave
  // 42 % full and the 192-wide shape is ~20 % cheaper.
  const bool gs = p->gn_stats != nullptr;
  // TMA epilogues (EPI 1 / 2, see the kernel): the plain bias / residual / alpha epilogue of the diffusion convs with 16-byte
  // aligned rows; everything else (activations, row maps, accumulate, per-utterance bias, both outputs) stays on EPI 0.
  static int tma_epi = -1;
  if (tma_epi < 0) { const char* e = getenv("DTTS_GEMM_TMAEPI"); tma_epi = e ? atoi(e) : 3; }
  int epi = 0;
  if (tma_epi && p->act == DTTS_ACT_NONE && !p->bias_utt && !p->out_row_map && !p->accumulate && (N & 3) == 0 &&
      (!p->bias || (((uintptr_t)p->bias) & 15) == 0) && (g_debug == 0 || g_debug == 1 || g_debug == 2)) {
    // (fp16 output WITH GroupNorm statistics stays on EPI 0 by default: its statistics pass has to convert the staged fp16 box
    //  back to fp32 and measured 87 vs 85 us in situ on the c1 convs; DTTS_GEMM_TMAEPI bit 2 forces it for tests)
    if (p->out_f16 && !p->out_f32 && !p->res && p->act16 == DTTS_ACT_NONE && (p->ldo16 & 7) == 0 && (((uintptr_t)p->out_f16) & 15) == 0 &&
        (!gs || ((N & 63) == 0 && (tma_epi & 4))))
      epi = 1;
    else if (p->out_f32 && !p->out_f16 && (p->ldo32 & 3) == 0 && (((uintptr_t)p->out_f32) & 15) == 0 &&
             (!p->res || ((p->ldr & 3) == 0 && (((uintptr_t)p->res) & 15) == 0)) && (!gs || (N & 31) == 0))
      epi = 2;
    if (epi & tma_epi & 3) {} else epi = 0;          // DTTS_GEMM_TMAEPI = 0 / 1 / 2 / 3: off / fp16 stores only / fp32 only / both (default)
  }
  // the fp32 + residual boxes need 96 KB: a single-CTA 256-wide tile would be left with two operand stages, so those GEMMs
  // take the CTA-pair tile (32 KB stages) whatever their K
  static int pair_mink = -1;   // smallest K * taps that takes the CTA-pair tile
  if (pair_mink < 0) { const char* e = getenv("DTTS_GEMM_PAIR_MINK"); pair_mink = e ? atoi(e) : 768; }
  const bool can_pair = pair && N % 256 == 0 && p->M >= 4096 && ((long)p->K * p->taps >= pair_mink || epi == 2);
  const bool can256 = bn256 && N % 256 == 0 && epi != 2, can192 = N % 192 == 0;
#define DTTS_LAUNCH(BN_, PAIR_)                                                                                      \
  do {                                                                                                               \
    if (PAIR_) {                                                                                                     \
      if (epi == 2) return gs ? launch_pair<BN_, true, 2>(p, st) : launch_pair<BN_, false, 2>(p, st);                \
      if (epi == 1) return gs ? launch_pair<BN_, true, 1>(p, st) : launch_pair<BN_, false, 1>(p, st);                \
      return gs ? launch_pair<BN_, true>(p, st) : launch_pair<BN_>(p, st);                                           \
    }                                                                                                                \
    if (epi == 2) return gs ? launch<BN_, false, true, 2>(p, st) : launch<BN_, false, false, 2>(p, st);              \
    if (epi == 1) return gs ? launch<BN_, false, true, 1>(p, st) : launch<BN_, false, false, 1>(p, st);              \
    return gs ? launch<BN_, false, true>(p, st) : launch<BN_, false>(p, st);                                         \
  } while (0)
  if ((can_pair || can256) && can192) {
    const long mt = ceil_div(p->M, BM), mt2 = ceil_div(p->M, 2 * BM);
    const double c256 = (double)((mt * (N / 256) + g_sm_count - 1) / g_sm_count) * 256.0;
    const double c192 = (double)((mt * (N / 192) + g_sm_count - 1) / g_sm_count) * 192.0 / 0.95;
    const double cpair = (double)((mt2 * (N / 256) + g_sm_count / 2 - 1) / (g_sm_count / 2)) * 256.0 / 1.08;
    if (c192 < (can256 ? c256 : 1e30) && c192 < (can_pair ? cpair : 1e30)) DTTS_LAUNCH(192, false);
    if (can_pair && (!can256 || cpair <= c256)) DTTS_LAUNCH(256, true);
  }
  if (can_pair) DTTS_LAUNCH(256, true);
  if (can256) DTTS_LAUNCH(256, false);
  if (can192) DTTS_LAUNCH(192, false);
#undef DTTS_LAUNCH
  epi = 0;
  if (bn256 && N % 256 == 0) return gs ? launch<256, false, true>(p, st) : launch<256, false>(p, st);   // EPI 2 without a CTA pair: legacy epilogue
  DTTS_REQUIRE(!gs, "gemm_f16_tc: gn_stats is implemented for N %% 192 == 0 or N %% 256 == 0");
  if (N > 64) return launch<128, false>(p, st);
  if (N > 32) return launch<64, false>(p, st);
  return launch<32, false>(p, st);
}

extern "C" int dtts_gemm_tf32x3(const dtts_gemm_params* p, void* stream) {
  int rc = common_checks(p, "gemm_tf32x3", 4);
  if (rc) return rc;
  DTTS_REQUIRE(p->A_lo && p->W_lo, "gemm_tf32x3: missing low-part operands");
  DTTS_REQUIRE((((uintptr_t)p->A_lo) & 15) == 0 && (((uintptr_t)p->W_lo) & 15) == 0, "gemm_tf32x3: low parts must be 16-byte aligned");
  DTTS_REQUIRE(p->split_k <= 1 || (p->out_f32 && p->split_stride >= (int64_t)p->M * p->ldo32), "gemm_tf32x3: split-K needs an fp32 partial buffer of split_k x M x ldo32");
  cudaStream_t st = (cudaStream_t)stream;
  static int bn128 = -1;
  if (bn128 < 0) { const char* e = getenv("DTTS_TF32_BN128"); bn128 = e ? atoi(e) : 0; }
  if ((p->M > 256 || bn128) && p->N >= 128) return launch<128, true>(p, st);
  return launch<64, true>(p, st);
}
