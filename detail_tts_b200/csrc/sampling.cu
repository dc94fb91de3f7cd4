// sampling.cu -- the per-token host-free tail of the GPT decode step.
//   dtts_process_logits: HF processor chain RepetitionPenalty -> [Typical] -> Temperature -> TopK -> TopP ->
//     softmax (or argmax when greedy): transformers generation/logits_process.py:298,407-410,
//     522-532,582-585, as configured by vqvae/model_24k.py:782-792 / gpt/model.py:540-544.
//   dtts_append_token: HF _sample bookkeeping (generation/utils.py:2797-2805) fused with the next
//     step's input embedding mel_embedding[id] + mel_pos_embedding[pos] (gpt/model.py:145-148).
// One CTA per utterance row; the 8194-entry logits row lives in shared memory for the whole chain
// (read once from HBM, dense probabilities written once).  Top-k uses an exact 4-pass radix select
// on order-preserving integer keys, so ties at the k-th value are kept exactly like torch.topk +
// `scores < kth` does.
#include "common.cuh"

namespace {

constexpr int PL_THREADS = 512;
constexpr int PL_MAX_CAND = 1024;

__device__ __forceinline__ uint32_t f2key(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // ascending float order == ascending uint order
}

// LayerNorm statistics of one row held in shared memory, in the format dtts_decode_gemm consumes: per 128-column group g,
// stats[(g * n_rows + row) * 2 + {0, 1}] = (sum, sum of squared deviations from the group mean).  One warp per group.
__device__ __forceinline__ void row_group_stats(const float* xs, int dim, float* stats, int row, int n_rows) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int g = warp; g * 128 < dim; g += nw) {
    float v[4];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[i] = xs[g * 128 + lane + 32 * i]; s += v[i]; }
    s = warp_sum(s);
    const float mean = s * (1.0f / 128.0f);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float d = v[i] - mean; q += d * d; }
    q = warp_sum(q);
    if (lane == 0) { stats[((long)g * n_rows + row) * 2] = s; stats[((long)g * n_rows + row) * 2 + 1] = q; }
  }
}

// TAIL = false: dtts_process_logits (dense probabilities / argmax out).  TAIL = true: dtts_decode_tail -- the token is
// chosen on the device (argmax, or inverse-CDF sampling from a pre-drawn uniform) and appended (HF _sample bookkeeping +
// next-step embedding + its LayerNorm statistics) by the same CTA; the last CTA to finish bumps the step counter.
template <bool TAIL>
__global__ void __launch_bounds__(PL_THREADS)
process_logits_kernel(const dtts_decode_tail_params p, int64_t* argmax_out) {
  pdl_launch();
  pdl_wait();
  extern __shared__ float sh[];          // [vocab] scores
  __shared__ int sel_tok;
  __shared__ uint32_t bitmap[512];       // vocab <= 16384
  __shared__ uint32_t hist[256];
  __shared__ float red[40];
  __shared__ int redi[40];
  __shared__ uint32_t sel_prefix, sel_remaining;
  __shared__ float cand_v[PL_MAX_CAND];
  __shared__ int cand_i[PL_MAX_CAND];
  __shared__ int n_cand;
  __shared__ int n_keep;
  __shared__ float z_keep;
  const int row = blockIdx.x, tid = threadIdx.x, V = p.vocab;
  const float* lg = p.logits + (long)row * p.ldl;
  for (int i = tid; i < V; i += PL_THREADS) sh[i] = lg[i];
  for (int i = tid; i < 512; i += PL_THREADS) bitmap[i] = 0;
  __syncthreads();
  // repetition penalty: every distinct id of the history is penalised once (gather/scatter semantics)
  const int gstep = p.step_dev ? *p.step_dev : 0;                        // global step (uniform row)
  const int step = gstep - (TAIL && p.row_step0 ? p.row_step0[row] : 0);  // this row's own step (continuous batching)
  const bool frozen = TAIL && p.row_step0 && !p.unfinished[row];         // finished slot waiting to be harvested / rebound
  if (frozen) {
    // nothing of this row may change any more; still take part in the step-counter handshake
    if (tid == 0 && p.done_counter) {
      __threadfence();
      const unsigned prev = atomicAdd(p.done_counter, 1u);
      if (prev == (unsigned)p.n_rows - 1u) { *p.done_counter = 0u; *p.step_dev = gstep + 1; }
    }
    return;
  }
  const int n_ids = p.n_ids + step;
  const int64_t* ids = p.ids + (long)row * p.ld_ids;
  for (int i = tid; i < n_ids; i += PL_THREADS) {
    const int id = (int)ids[i];
    if (id < 0 || id >= V) continue;
    const uint32_t bit = 1u << (id & 31);
    const uint32_t old = atomicOr(&bitmap[id >> 5], bit);
    if (!(old & bit)) {
      const float s = sh[id];
      sh[id] = s < 0.f ? s * p.penalty : s / p.penalty;
    }
  }
  __syncthreads();
  if (p.suppress_token >= 0 && p.suppress_token < V && tid == 0) sh[p.suppress_token] = -INFINITY;
  __syncthreads();

  if (p.typical_mass > 0.f) {
    // TypicalLogitsWarper (gpt/modules/typical_sampling.py:14-33): keep the tokens whose surprise -log p is closest to the
    // entropy, up to cumulative probability `mass`.  The reference sorts |(-log p) - H| ascending, takes the value at the
    // first sorted index whose cumulative probability reaches the mass and removes everything above it; that value is the
    // smallest t with sum{p_i : shifted_i <= t} >= mass, found here by a 32-step bisection over the (order-preserving) bit
    // pattern of the non-negative shifted scores, every step one deterministic block reduction.
    float m = -INFINITY;
    for (int i = tid; i < V; i += PL_THREADS) m = fmaxf(m, sh[i]);
    const float mx = block_max(m, red);
    float z = 0.f;
    for (int i = tid; i < V; i += PL_THREADS) z += expf(sh[i] - mx);
    const float Z = block_sum(z, red);
    const float logZ = mx + logf(Z);
    float e = 0.f;
    for (int i = tid; i < V; i += PL_THREADS) {
      const float lp = sh[i] - logZ;
      const float pr = expf(lp);
      if (pr > 0.f) e -= lp * pr;            // nansum: -inf * 0 terms are skipped
    }
    const float ent = block_sum(e, red);
    uint32_t lo = 0;
    for (int bit = 31; bit >= 0; --bit) {
      const uint32_t cand = lo | ((1u << bit) - 1u);
      float acc = 0.f;
      for (int i = tid; i < V; i += PL_THREADS) {
        const float lp = sh[i] - logZ;
        const float shifted = fabsf(-lp - ent);
        if (__float_as_uint(shifted) <= cand) acc += expf(lp);
      }
      if (block_sum(acc, red) < p.typical_mass) lo |= 1u << bit;
    }
    for (int i = tid; i < V; i += PL_THREADS) {
      const float shifted = fabsf(-(sh[i] - logZ) - ent);
      if (__float_as_uint(shifted) > lo) sh[i] = -INFINITY;
    }
    __syncthreads();
  }

  if (!p.do_sample) {
    // argmax, lowest index on ties
    float bv = -INFINITY; int bi = 0x7fffffff;
    for (int i = tid; i < V; i += PL_THREADS) {
      const float s = sh[i];
      if (s > bv || (s == bv && i < bi)) { bv = s; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if ((tid & 31) == 0) { red[tid >> 5] = bv; redi[tid >> 5] = bi; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < PL_THREADS / 32; ++w)
        if (red[w] > bv || (red[w] == bv && redi[w] < bi)) { bv = red[w]; bi = redi[w]; }
      if (argmax_out) argmax_out[row] = bi;
      sel_tok = bi;
    }
    if (!TAIL) return;
  } else {

  // temperature
  for (int i = tid; i < V; i += PL_THREADS) sh[i] = sh[i] / p.temperature;
  __syncthreads();
  // exact k-th largest via radix select over 4 x 8 bits
  const int k = p.top_k < V ? p.top_k : V;
  if (tid == 0) { sel_prefix = 0; sel_remaining = (uint32_t)k; n_cand = 0; }
  __syncthreads();
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    for (int i = tid; i < 256; i += PL_THREADS) hist[i] = 0;
    __syncthreads();
    const uint32_t prefix = sel_prefix;
    const uint32_t mask_hi = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
    for (int i = tid; i < V; i += PL_THREADS) {
      const uint32_t key = f2key(sh[i]);
      if ((key & mask_hi) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (tid < 32) {
      // bucket holding the rem-th largest key: scan the 256 bins from the top, one warp (lane l owns bins 8l..8l+7;
      // suffix sums over lanes by shuffles) instead of a 256-step dependent loop in one thread
      const uint32_t rem = sel_remaining;
      uint32_t c[8], tot = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) { c[i] = hist[tid * 8 + i]; tot += c[i]; }
      uint32_t suf = tot;                       // -> sum over lanes >= tid
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_down_sync(0xffffffffu, suf, o);
        if (tid + o < 32) suf += t;
      }
      const uint32_t above = suf - tot;         // keys in bins owned by higher lanes
      const bool mine = above < rem && suf >= rem;
      const uint32_t who = __ballot_sync(0xffffffffu, mine);
      if (mine) {
        uint32_t r = rem - above;
        int i = 7;
        for (; i > 0; --i) {
          if (c[i] >= r) break;
          r -= c[i];
        }
        sel_prefix = prefix | ((uint32_t)(tid * 8 + i) << shift);
        sel_remaining = r;
      } else if (who == 0 && tid == 0) {        // fewer than rem keys under this prefix (cannot happen: kept for safety)
        sel_prefix = prefix;
        sel_remaining = rem - (suf - c[0]);
      }
    }
    __syncthreads();
  }
  const uint32_t kth_key = sel_prefix;
  // candidates: everything >= kth value (ties kept)
  for (int i = tid; i < V; i += PL_THREADS) {
    const float s = sh[i];
    if (f2key(s) >= kth_key && s > -INFINITY) {
      const int slot = atomicAdd(&n_cand, 1);
      if (slot < PL_MAX_CAND) { cand_v[slot] = s; cand_i[slot] = i; }
    }
  }
  __syncthreads();
  const int nc = n_cand < PL_MAX_CAND ? n_cand : PL_MAX_CAND;
  // sort candidates descending by (value, then index ascending) -- bitonic over next pow2
  int np2 = 1;
  while (np2 < nc) np2 <<= 1;
  for (int i = nc + tid; i < np2; i += PL_THREADS) { cand_v[i] = -INFINITY; cand_i[i] = 0x7fffffff; }
  __syncthreads();
  for (int size = 2; size <= np2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = tid; t < np2 / 2; t += PL_THREADS) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const float a = cand_v[lo], b = cand_v[hi];
        const int ai = cand_i[lo], bi = cand_i[hi];
        const bool a_first = (a > b) || (a == b && ai < bi);  // a should precede b in descending order
        if (desc ? !a_first : a_first) {
          cand_v[lo] = b; cand_v[hi] = a; cand_i[lo] = bi; cand_i[hi] = ai;
        }
      }
      __syncthreads();
    }
  }
  // top-p on the sorted candidates (one thread; nc ~ 50): HF sorts ascending, softmax, cumsum,
  // removes cum <= 1 - top_p, always keeps the largest.
  if (tid == 0) {
    const float mx = cand_v[0];
    float Z = 0.f;
    for (int i = nc - 1; i >= 0; --i) Z += expf(cand_v[i] - mx);   // ascending accumulation
    float cum = 0.f;
    int keep = nc;   // keep candidates [0, keep)
    for (int i = nc - 1; i >= 1; --i) {
      cum += expf(cand_v[i] - mx) / Z;
      if (cum <= 1.0f - p.top_p) keep = i; else break;
    }
    float Zk = 0.f;
    for (int i = keep - 1; i >= 0; --i) Zk += expf(cand_v[i] - mx);
    n_keep = keep;
    z_keep = Zk;
  }
  __syncthreads();
  const float mx = cand_v[0];
  if (p.probs) {
    float* pr = p.probs + (long)row * p.ldp;
    for (int i = tid; i < V; i += PL_THREADS) pr[i] = 0.f;
    __syncthreads();
    for (int i = tid; i < n_keep; i += PL_THREADS) pr[cand_i[i]] = expf(cand_v[i] - mx) / z_keep;
  }
  if (TAIL) {
    // inverse-CDF sampling over the kept tokens in ascending id order (what a sequential fp32 cumulative sum over the
    // dense probability row does): the first id whose cumulative probability exceeds u; the last kept id if none does.
    float* sp = sh;                          // the score row is dead: reuse it ([0, nk) probabilities, [nk, 2nk) ids)
    int* si = reinterpret_cast<int*>(sh + PL_MAX_CAND);
    __syncthreads();
    const int nk = n_keep;
    for (int i = tid; i < nk; i += PL_THREADS) {
      const int my = cand_i[i];
      int r = 0;
      for (int j = 0; j < nk; ++j) r += cand_i[j] < my;
      sp[r] = expf(cand_v[i] - mx) / z_keep;
      si[r] = my;
    }
    __syncthreads();
    if (tid == 0) {
      const float u = p.uniforms[(long)gstep * p.ld_u + row];
      float c = 0.f;
      int tok = si[nk - 1];
      for (int i = 0; i < nk; ++i) {
        c += sp[i];
        if (c > u) { tok = si[i]; break; }
      }
      sel_tok = tok;
    }
  }
  }   // do_sample
  if (!TAIL) return;

  // ---- append (dtts_append_token's arithmetic) ----
  __syncthreads();
  __shared__ int64_t tok_s;
  if (tid == 0) {
    int64_t tok = sel_tok;
    const int unf = p.unfinished[row];
    if (!unf) tok = p.stop_token;
    p.ids[(long)row * p.ld_ids + p.n_ids + step] = tok;
    p.unfinished[row] = unf && (tok != p.stop_token) && !(p.max_new > 0 && step + 1 >= p.max_new);
    const int pos0 = p.kv_pos_rows ? p.kv_pos_rows[row] : 0;
    if (p.kv_row) p.kv_row[row] = row * p.kv_stride + pos0 + step;
    if (p.kv_len) p.kv_len[row] = pos0 + step + 1;
    tok_s = tok;
  }
  __syncthreads();
  if (p.x_out) {
    const float* te = p.tok_emb + tok_s * p.dim;
    const float* pe = p.pos_emb + (long)(p.pos + step) * p.dim;
    float* xo = p.x_out + (long)row * p.ldx;
    for (int d = tid; d < p.dim; d += PL_THREADS) { const float v = te[d] + pe[d]; xo[d] = v; sh[d] = v; }
    __syncthreads();
    if (p.x_stats) row_group_stats(sh, p.dim, p.x_stats, row, p.n_rows);
  }
  // the last row to arrive advances the step counter (every row has read it before arriving)
  __syncthreads();
  if (tid == 0 && p.done_counter) {
    __threadfence();
    const unsigned prev = atomicAdd(p.done_counter, 1u);
    if (prev == (unsigned)p.n_rows - 1u) {
      *p.done_counter = 0u;
      if (p.step_dev) *p.step_dev = gstep + 1;
    }
  }
}

__global__ void __launch_bounds__(256)
append_token_kernel(const dtts_append_params p) {
  const int row = blockIdx.x;
  const int step = p.step_dev ? *p.step_dev : 0;
  __shared__ int64_t tok_s;
  if (threadIdx.x == 0) {
    int64_t tok = p.next[row];
    const int unf = p.unfinished[row];
    if (!unf) tok = p.stop_token;
    p.ids[(long)row * p.ld_ids + p.n_ids + step] = tok;
    const int still = unf && (tok != p.stop_token);
    p.unfinished[row] = still;
    if (still && p.n_unfinished) atomicAdd(p.n_unfinished, 1);
    const int pos0 = p.kv_pos_rows ? p.kv_pos_rows[row] : p.kv_pos0;
    if (p.kv_row) p.kv_row[row] = row * p.kv_stride + pos0 + step;
    if (p.kv_len) p.kv_len[row] = pos0 + step + 1;
    tok_s = tok;
  }
  __syncthreads();
  const int64_t tok = tok_s;
  if (p.x_out) {
    __shared__ float xrow[1024];
    const float* te = p.tok_emb + tok * p.dim;
    const float* pe = p.pos_emb + (long)(p.pos + step) * p.dim;
    float* xo = p.x_out + (long)row * p.ldx;
    for (int d = threadIdx.x; d < p.dim; d += blockDim.x) {
      const float v = te[d] + pe[d];
      xo[d] = v;
      if (p.x_stats) xrow[d] = v;
    }
    if (p.x_stats) {
      __syncthreads();
      row_group_stats(xrow, p.dim, p.x_stats, row, p.n_rows);
    }
  }
}

__global__ void bump_step_kernel(int* step) { *step += 1; }

}  // namespace

extern "C" int dtts_process_logits(const dtts_logits_params* p, void* stream) {
  DTTS_REQUIRE(p && p->logits && p->ids, "process_logits: null argument");
  DTTS_REQUIRE(p->vocab > 0 && p->vocab <= 16384, "process_logits: vocab out of range");
  DTTS_REQUIRE(p->do_sample ? (p->probs != nullptr) : (p->argmax != nullptr), "process_logits: missing output");
  DTTS_REQUIRE(!p->do_sample || (p->temperature > 0.f && p->top_k > 0 && p->top_k <= 512), "process_logits: bad sampling params");
  if (p->n_rows <= 0) return 0;
  dtts_decode_tail_params t;
  memset(&t, 0, sizeof(t));
  t.logits = p->logits; t.ldl = p->ldl; t.n_rows = p->n_rows; t.vocab = p->vocab;
  t.ids = const_cast<int64_t*>(p->ids); t.ld_ids = p->ld_ids; t.n_ids = p->n_ids; t.step_dev = const_cast<int*>(p->step_dev);
  t.penalty = p->penalty; t.temperature = p->temperature; t.top_p = p->top_p; t.top_k = p->top_k;
  t.do_sample = p->do_sample; t.suppress_token = p->suppress_token; t.typical_mass = p->typical_mass;
  t.probs = p->probs; t.ldp = p->ldp;
  launch_maybe_pdl(process_logits_kernel<false>, dim3(p->n_rows), dim3(PL_THREADS), p->vocab * sizeof(float), (cudaStream_t)stream, t, p->argmax);
  DTTS_CHECK_LAUNCH("process_logits");
  return 0;
}

extern "C" int dtts_decode_tail(const dtts_decode_tail_params* p, void* stream) {
  DTTS_REQUIRE(p && p->logits && p->ids && p->unfinished && p->step_dev && p->done_counter, "decode_tail: null argument");
  DTTS_REQUIRE(p->vocab >= 2 * PL_MAX_CAND && p->vocab <= 16384, "decode_tail: vocab out of range");
  DTTS_REQUIRE(!p->do_sample || (p->uniforms && p->temperature > 0.f && p->top_k > 0 && p->top_k <= 512), "decode_tail: bad sampling params");
  DTTS_REQUIRE(!p->x_out || (p->tok_emb && p->pos_emb && p->dim <= p->vocab), "decode_tail: missing embedding tables");
  DTTS_REQUIRE(!p->x_stats || (p->x_out && p->dim % 128 == 0), "decode_tail: x_stats needs x_out and dim %% 128 == 0");
  if (p->n_rows <= 0) return 0;
  launch_maybe_pdl(process_logits_kernel<true>, dim3(p->n_rows), dim3(PL_THREADS), p->vocab * sizeof(float), (cudaStream_t)stream, *p, (int64_t*)nullptr);
  DTTS_CHECK_LAUNCH("decode_tail");
  return 0;
}

extern "C" int dtts_append_token(const dtts_append_params* p, void* stream) {
  DTTS_REQUIRE(p && p->next && p->ids && p->unfinished, "append_token: null argument");
  DTTS_REQUIRE(!p->x_out || (p->tok_emb && p->pos_emb), "append_token: missing embedding tables");
  DTTS_REQUIRE(!p->x_stats || (p->x_out && p->dim % 128 == 0 && p->dim <= 1024), "append_token: x_stats needs x_out and dim %% 128 == 0, dim <= 1024");
  if (p->n_rows <= 0) return 0;
  append_token_kernel<<<p->n_rows, 256, 0, (cudaStream_t)stream>>>(*p);
  DTTS_CHECK_LAUNCH("append_token");
  if (p->step_dev) {
    bump_step_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(p->step_dev);
    DTTS_CHECK_LAUNCH("bump_step");
  }
  return 0;
}
