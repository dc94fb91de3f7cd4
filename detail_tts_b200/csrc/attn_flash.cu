// attn_flash.cu -- flash-style attention for the diffusion AttentionBlock hot loop
// (vqvae/utils/diff_util.py:145-169: QKVAttentionLegacy; xtransformers.py:177-186: T5-style
// relative-position bias added to the scores before the fp32 softmax).
//
// fp16 operands, fp32 scores / online softmax / output accumulation.  One CTA = (utterance, head,
// 64-query tile), 4 warps x 16 queries; keys/values stream through a double-buffered cp.async
// ring in 64-key tiles; the [F, F] score matrix the reference materialises in HBM never leaves
// registers.  head_dim 48 = 3 k-steps of the m16n8k16 MMA (the score/PV term is ~7 % of the
// diffusion FLOPs at F=280; the 93 % in the 1x1/k3 conv GEMMs run on tcgen05 in gemm_tc.cu).
// The bias is a per-head table over clamp(key - query, +-bias_half) staged in shared memory.
#include "common.cuh"

namespace {

constexpr int HD = 48;
constexpr int LDS = 56;     // padded smem row (halfs): 112 B rows -> conflict-free ldmatrix
constexpr int BQ = 96, BKV = 48;   // 6 warps x 16 queries; 48-key tiles: F=280 pads to 288 (2.9 %) on both axes
constexpr int NWARP = BQ / 16;
constexpr int NTHR = NWARP * 32;
constexpr int NT = BKV / 8;        // score n-tiles per key tile
constexpr int MAX_BIAS = 513;
constexpr int BIAS_PAD = 80;       // table padded with its end values: one clamp per (thread, key tile) instead of per score
static_assert(BIAS_PAD >= BKV + 16 + 9, "padding must cover one key tile + one warp's query rows");

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_addr(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_addr(p)));
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
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Scores are kept in the log2 domain: s2 = (q.k)*scale*log2e + bias*log2e, p = 2^(s2 - m2).  The scale is
// applied to the fp32 accumulator (one FMUL fused with the bias add as an FFMA); the bias table is staged
// pre-multiplied by log2e.  Key tiles entirely beyond the clamp distance of the table (|key - query| >=
// bias_half for every pair) take a tile-uniform bias with no table lookups; the key mask is only evaluated
// on the last key tile.
__global__ void __launch_bounds__(NTHR)
flash48_kernel(const dtts_attention_params p) {
  __shared__ __align__(16) __half sQ[BQ][LDS];
  __shared__ __align__(16) __half sK[2][BKV][LDS];
  __shared__ __align__(16) __half sV[2][BKV][LDS];
  __shared__ float sBiasPad[MAX_BIAS + 2 * BIAS_PAD];
  float* const sBias = sBiasPad + BIAS_PAD;   // sBias[-BIAS_PAD .. 2*half + BIAS_PAD]
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int qlen = p.q_len[b], klen = p.k_len[b];
  const int q0 = qt * BQ;
  if (q0 >= qlen || klen <= 0) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, tig = lane & 3;
  const __half* qg = (const __half*)p.q + (long)p.q_off[b] * p.ldq + (long)h * p.head_stride_q;
  const __half* kg = (const __half*)p.k + (long)p.k_off[b] * p.ldk + (long)h * p.head_stride_k;
  const __half* vg = (const __half*)p.v + (long)p.k_off[b] * p.ldv + (long)h * p.head_stride_v;
  const float LOG2E = 1.4426950408889634f;

  for (int c = tid; c < BQ * 6; c += NTHR) {
    const int r = c / 6, ch = (c % 6) * 8;
    if (q0 + r < qlen) cp_async16(&sQ[r][ch], qg + (long)(q0 + r) * p.ldq + ch);
    else *reinterpret_cast<uint4*>(&sQ[r][ch]) = make_uint4(0, 0, 0, 0);
  }
  const int half = p.bias_half;
  const int nb = p.bias_mode == DTTS_ATTN_BIAS_RELPOS_TABLE ? 2 * half + 1 : 0;
  if (nb)
    for (int i = tid - BIAS_PAD; i < nb + BIAS_PAD; i += NTHR) sBias[i] = p.bias_table[h * nb + min(max(i, 0), nb - 1)] * LOG2E;

  auto load_kv = [&](int t, int buf) {
    const int k0 = t * BKV;
    for (int c = tid; c < BKV * 6; c += NTHR) {
      const int r = c / 6, ch = (c % 6) * 8;
      if (k0 + r < klen) {
        cp_async16(&sK[buf][r][ch], kg + (long)(k0 + r) * p.ldk + ch);
        cp_async16(&sV[buf][r][ch], vg + (long)(k0 + r) * p.ldv + ch);
      } else {
        *reinterpret_cast<uint4*>(&sK[buf][r][ch]) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(&sV[buf][r][ch]) = make_uint4(0, 0, 0, 0);
      }
    }
  };
  const int ntiles = (klen + BKV - 1) / BKV;
  load_kv(0, 0);
  cp_commit();

  uint32_t qa[3][4];
  float o[6][4];
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) o[i][e] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  const float sc2 = p.scale * LOG2E;
  const int qi0 = q0 + warp * 16 + g;  // query index of c0/c1 rows; +8 for c2/c3

  for (int t = 0; t < ntiles; ++t) {
    const int buf = t & 1;
    if (t + 1 < ntiles) {
      load_kv(t + 1, buf ^ 1);
      cp_commit();
      cp_wait<1>();
    } else {
      cp_wait<0>();
    }
    __syncthreads();
    if (t == 0) {
#pragma unroll
      for (int ks = 0; ks < 3; ++ks)
        ldsm_x4(qa[ks], &sQ[warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8][ks * 16 + ((lane >> 4) & 1) * 8]);
    }
    float s[NT][4];
#pragma unroll
    for (int i = 0; i < NT; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) s[i][e] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 3; ++ks) {
#pragma unroll
      for (int np = 0; np < NT / 2; ++np) {
        uint32_t kb[4];
        const int mat = lane >> 3;
        ldsm_x4(kb, &sK[buf][np * 16 + (lane & 7) + (mat >> 1) * 8][ks * 16 + (mat & 1) * 8]);
        mma16816(s[2 * np], qa[ks], kb[0], kb[1]);
        mma16816(s[2 * np + 1], qa[ks], kb[2], kb[3]);
      }
    }
    // log2-domain scores: scale + bias (+ key mask on the last tile), running max
    const int k0 = t * BKV;
    float mx[2] = {-INFINITY, -INFINITY};
    // signed distance range of this (warp query rows, key tile): d = key - query
    const int dmin = k0 - (q0 + warp * 16 + 15), dmax = k0 + BKV - 1 - (q0 + warp * 16);
    if (nb == 0 || dmin >= half || dmax <= -half) {
      const float ub = nb == 0 ? 0.f : (dmin >= half ? sBias[2 * half] : sBias[0]);
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) s[nt][e] = fmaf(s[nt][e], sc2, ub);
    } else {
      // table index of (nt, e): base + nt*8 + (e&1) - (e>>1)*8, i.e. rows g+8 at n-tile nt reuse the pair of rows g at
      // n-tile nt-1: NT+1 pair loads per thread.  The padded table makes one clamp of `base` per tile exact.
      int base = k0 + tig * 2 - qi0 + half;
      base = min(max(base, -BIAS_PAD + 8), 2 * half + BIAS_PAD - 8 * NT);
      const float* bp = sBias + base;
      float b0 = bp[-8], b1 = bp[-7];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const float c0 = bp[nt * 8], c1 = bp[nt * 8 + 1];
        s[nt][0] = fmaf(s[nt][0], sc2, c0);
        s[nt][1] = fmaf(s[nt][1], sc2, c1);
        s[nt][2] = fmaf(s[nt][2], sc2, b0);
        s[nt][3] = fmaf(s[nt][3], sc2, b1);
        b0 = c0; b1 = c1;
      }
    }
    if (k0 + BKV > klen) {
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (k0 + nt * 8 + tig * 2 + (e & 1) >= klen) s[nt][e] = -INFINITY;
    }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
    }
    float corr[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float mn = fmaxf(m_run[r], mx[r]);
      corr[r] = ex2(m_run[r] - mn);
      m_run[r] = mn;
    }
    float rs[2] = {0.f, 0.f};
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float pv = ex2(s[nt][e] - m_run[e >> 1]);
        s[nt][e] = pv;
        rs[e >> 1] += pv;
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 1);
      rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 2);
      l_run[r] = l_run[r] * corr[r] + rs[r];
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      o[i][0] *= corr[0]; o[i][1] *= corr[0]; o[i][2] *= corr[1]; o[i][3] *= corr[1];
    }
    // O += P V
#pragma unroll
    for (int kk = 0; kk < BKV / 16; ++kk) {
      uint32_t pa[4];
      pa[0] = pack_h2(s[2 * kk][0], s[2 * kk][1]);
      pa[1] = pack_h2(s[2 * kk][2], s[2 * kk][3]);
      pa[2] = pack_h2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pa[3] = pack_h2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int dp = 0; dp < 3; ++dp) {
        uint32_t vb[4];
        const int mat = lane >> 3;
        ldsm_x4_t(vb, &sV[buf][kk * 16 + (lane & 7) + (mat & 1) * 8][dp * 16 + (mat >> 1) * 8]);
        mma16816(o[2 * dp], pa, vb[0], vb[1]);
        mma16816(o[2 * dp + 1], pa, vb[2], vb[3]);
      }
    }
    __syncthreads();
  }
  // normalise + store (head-major [rows, n_heads*48])
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int i = qi0 + r * 8;
    if (i >= qlen) continue;
    const float inv = 1.0f / l_run[r];
    const long orow = (long)(p.o_off ? p.o_off[b] : p.q_off[b]) + i;
#pragma unroll
    for (int dt = 0; dt < 6; ++dt) {
      const int d = h * HD + dt * 8 + tig * 2;
      const float a = o[dt][2 * r] * inv, c = o[dt][2 * r + 1] * inv;
      if (p.out_f16) *reinterpret_cast<__half2*>((__half*)p.out_f16 + orow * p.ldo16 + d) = __floats2half2_rn(a, c);
      if (p.out_f32) *reinterpret_cast<float2*>(p.out_f32 + orow * p.ldo32 + d) = make_float2(a, c);
    }
  }
}

}  // namespace

extern "C" int dtts_attention_f16_flash(const dtts_attention_params* p, void* stream) {
  DTTS_REQUIRE(p && p->q && p->k && p->v && p->q_off && p->q_len && p->k_off && p->k_len, "attention_f16_flash: null argument");
  DTTS_REQUIRE(p->is_f16 && p->head_dim == HD, "attention_f16_flash: needs fp16 operands and head_dim 48");
  DTTS_REQUIRE(!p->causal, "attention_f16_flash: causal masks are not supported (use dtts_attention_f32)");
  DTTS_REQUIRE(p->bias_mode == DTTS_ATTN_BIAS_NONE || (p->bias_mode == DTTS_ATTN_BIAS_RELPOS_TABLE && p->bias_table && 2 * p->bias_half + 1 <= MAX_BIAS),
               "attention_f16_flash: unsupported bias mode");
  DTTS_REQUIRE(p->ldq % 8 == 0 && p->ldk % 8 == 0 && p->ldv % 8 == 0 && p->head_stride_q % 8 == 0 && p->head_stride_k % 8 == 0 && p->head_stride_v % 8 == 0,
               "attention_f16_flash: strides must keep 16-byte alignment");
  DTTS_REQUIRE((((uintptr_t)p->q | (uintptr_t)p->k | (uintptr_t)p->v) & 15) == 0, "attention_f16_flash: operands must be 16-byte aligned");
  DTTS_REQUIRE(p->out_f32 || p->out_f16, "attention_f16_flash: no output");
  DTTS_REQUIRE(!p->out_f16 || p->ldo16 % 2 == 0, "attention_f16_flash: ldo16 must be even");
  DTTS_REQUIRE(!p->out_f32 || p->ldo32 % 2 == 0, "attention_f16_flash: ldo32 must be even");
  if (p->n_utt <= 0 || p->max_q_len <= 0) return 0;
  dim3 grid(ceil_div(p->max_q_len, BQ), p->n_heads, p->n_utt);
  flash48_kernel<<<grid, NTHR, 0, (cudaStream_t)stream>>>(*p);
  DTTS_CHECK_LAUNCH("attention_f16_flash");
  return 0;
}
