// eltwise.cu -- HBM-bound elementwise / layout kernels of the synthesis path (rows layout).
// All are single-pass, float4-vectorised where the leading dimensions allow; each cites the
// reference op it replaces.  Also hosts the library-info entry points.
#include "common.cuh"

thread_local char g_dtts_err[512] = {0};
int g_dtts_launches = 0;
int g_dtts_pdl = 0;

namespace {

// ---- ancestral DDPM update with CFG and learned-range variance (vqvae/utils/diffusion.py:317-386,472-485)
__global__ void __launch_bounds__(256)
pstep_kernel(const dtts_pstep_params p) {
  const int c4 = p.C >> 2;
  const long total = (long)p.M * c4;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int m = (int)(idx / c4), c = (int)(idx % c4) * 4;
    float4 x = *reinterpret_cast<const float4*>(p.x + (long)m * p.ldx + c);
    const float4 ec = *reinterpret_cast<const float4*>(p.out_c + (long)m * p.ldo + c);
    const float4 vc = *reinterpret_cast<const float4*>(p.out_c + (long)m * p.ldo + p.C + c);
    const float4 eu = *reinterpret_cast<const float4*>(p.out_u + (long)m * p.ldo + c);
    const float4 nz = *reinterpret_cast<const float4*>(p.noise + (long)m * p.ldn + c);
    float xs[4] = {x.x, x.y, x.z, x.w}, e1[4] = {ec.x, ec.y, ec.z, ec.w}, v1[4] = {vc.x, vc.y, vc.z, vc.w};
    float e2[4] = {eu.x, eu.y, eu.z, eu.w}, nn[4] = {nz.x, nz.y, nz.z, nz.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float frac = (v1[q] + 1.0f) / 2.0f;
      const float logvar = frac * p.max_log + (1.0f - frac) * p.min_log;
      const float eps = (1.0f + p.cfk) * e1[q] - p.cfk * e2[q];
      float x0 = p.sqrt_recip * xs[q] - p.sqrt_recipm1 * eps;
      x0 = fminf(fmaxf(x0, -1.0f), 1.0f);
      const float mean = p.coef1 * x0 + p.coef2 * xs[q];
      xs[q] = mean + p.nonzero * expf(0.5f * logvar) * nn[q];
    }
    *reinterpret_cast<float4*>(p.x + (long)m * p.ldx + c) = make_float4(xs[0], xs[1], xs[2], xs[3]);
    if (p.x_f16) {
      __half2 h0 = __floats2half2_rn(xs[0], xs[1]), h1 = __floats2half2_rn(xs[2], xs[3]);
      uint2 pk; pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
      *reinterpret_cast<uint2*>((__half*)p.x_f16 + (long)m * p.ldx16 + c) = pk;
    }
  }
}

// ---- [B, C, T] (reference layout) <-> rows, 32x32 smem-tiled transpose
__global__ void __launch_bounds__(256)
bct2rows_kernel(const dtts_bct2rows_params p) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, len = p.utt_len[b];
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  if (t0 >= len) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 8 rows per sweep
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, t = t0 + tx;
    tile[i][tx] = (c < p.C && t < len) ? p.src[((long)b * p.C + c) * p.T + t] * p.scale + p.shift : 0.f;
  }
  __syncthreads();
  const long row0 = p.utt_off[b];
  for (int i = ty; i < 32; i += 8) {
    const int t = t0 + i, c = c0 + tx;
    if (t < len && c < p.C) {
      const float v = tile[tx][i];
      if (p.dst_f32) p.dst_f32[(row0 + t) * p.ld32 + c] = v;
      if (p.dst_f16) ((__half*)p.dst_f16)[(row0 + t) * p.ld16 + c] = __float2half_rn(v);
    }
  }
}

__global__ void __launch_bounds__(256)
rows2bct_kernel(const dtts_rows2bct_params p) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, len = p.utt_len[b];
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const long row0 = p.utt_off[b];
  for (int i = ty; i < 32; i += 8) {
    const int t = t0 + i, c = c0 + tx;
    tile[i][tx] = (t < len && c < p.C) ? p.src[(row0 + t) * p.ld + c] * p.scale + p.shift : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, t = t0 + tx;
    if (c < p.C && t < p.T) p.dst[((long)b * p.C + c) * p.T + t] = t < len ? tile[tx][i] : 0.f;
  }
}

__global__ void __launch_bounds__(256)
eltwise_kernel(const dtts_eltwise_params p) {
  const long total = (long)p.M * p.C;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int m = (int)(idx / p.C), c = (int)(idx % p.C);
    float v = 0.f;
    if (!p.row_utt || p.row_utt[m] >= 0) v = act_apply(p.act, p.x[(long)m * p.ldx + c], p.act_param) * p.scale;
    if (p.out_f32) p.out_f32[(long)m * p.ldo32 + c] = v;
    if (p.out_f16) ((__half*)p.out_f16)[(long)m * p.ldo16 + c] = __float2half_rn(v);
  }
}

// ---- embedding gather + learned positions (gpt/model.py:134-136,203-215,517-519)
__global__ void __launch_bounds__(256)
embed_kernel(const dtts_embed_params p) {
  const int i = blockIdx.x;
  const long id = p.ids[i];
  const float* te = p.table + id * p.dim;
  const float* pe = p.pos_table ? p.pos_table + (long)(p.pos ? p.pos[i] : i) * p.dim : nullptr;
  float* o = p.out + (long)(p.dst_row ? p.dst_row[i] : i) * p.ldo;
  for (int d = threadIdx.x; d < p.dim; d += blockDim.x) o[d] = te[d] + (pe ? pe[d] : 0.f);
}

// ---- F.interpolate(mode='nearest') by an integer factor (vqvae/diff_model.py:252)
__global__ void __launch_bounds__(256)
repeat_rows_kernel(const dtts_repeat_rows_params p) {
  const int b = blockIdx.y;
  const int len = p.utt_len[b] * p.repeat;
  const int c4 = p.C >> 2;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < (long)len * c4; idx += (long)gridDim.x * blockDim.x) {
    const int r = (int)(idx / c4), c = (int)(idx % c4) * 4;
    const float4 v = *reinterpret_cast<const float4*>(p.x + (long)(p.utt_off[b] + r / p.repeat) * p.ldx + c);
    *reinterpret_cast<float4*>(p.out + (long)(p.out_off[b] + r) * p.ldo + c) = v;
  }
}

// ---- masked temporal mean (modules.py:686-694; diff_model.py:228)
__global__ void __launch_bounds__(256)
mean_rows_kernel(const dtts_mean_rows_params p) {
  const int b = blockIdx.y;
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int ty = threadIdx.x >> 5;  // 8 row lanes
  __shared__ float part[8][33];
  const int len = p.utt_len[b];
  float s = 0.f;
  if (c < p.C)
    for (int r = ty; r < len; r += 8) s += p.x[(long)(p.utt_off[b] + r) * p.ldx + c];
  part[ty][threadIdx.x & 31] = s;
  __syncthreads();
  if (ty == 0 && c < p.C) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += part[i][threadIdx.x];
    p.out[(long)b * p.ldo + c] = t / (float)len;
  }
}

// ---- sinusoidal timestep embedding, cos half first (vqvae/diff_model.py:20-38)
__global__ void __launch_bounds__(256)
tsemb_kernel(const dtts_tsemb_params p) {
  const int i = blockIdx.x, half = p.dim / 2;
  const float t = p.t[i];
  for (int k = threadIdx.x; k < half; k += blockDim.x) {
    const float f = (float)exp(-9.210340371976184 * (double)k / (double)half);
    const float a = t * f;
    const float c = cosf(a), s = sinf(a);
    if (p.out) { p.out[(long)i * p.ldo + k] = c; p.out[(long)i * p.ldo + half + k] = s; }
    if (p.out_f16) {
      ((__half*)p.out_f16)[(long)i * p.ldo16 + k] = __float2half_rn(c);
      ((__half*)p.out_f16)[(long)i * p.ldo16 + half + k] = __float2half_rn(s);
    }
  }
}

// ---- coupling update + Flip (vqvae/modules/modules.py:393-400,472-474)
__global__ void __launch_bounds__(256)
couple_kernel(const dtts_couple_params p) {
  const int H = p.half;
  const long total = (long)p.M * H;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int m = (int)(idx / H), c = (int)(idx % H);
    float* xr = p.x + (long)m * p.ldx;
    const bool valid = !p.row_utt || p.row_utt[m] >= 0;
    float lo = xr[c];                 // x0[c]
    const int hc = 2 * H - 1 - c;     // mirrored channel in x1
    float hi = xr[hc];
    if (p.m) hi = valid ? hi - p.m[(long)m * p.ldm + (hc - H)] : 0.f;
    if (p.flip_after) { const float t = lo; lo = hi; hi = t; }
    xr[c] = lo;
    xr[hc] = hi;
    if (p.x0_f16) ((__half*)p.x0_f16)[(long)m * p.ld16 + c] = __float2half_rn(lo);
  }
}

// ---- z_p = m + eps * exp(logs) * noise_scale (vqvae/model_24k.py:860)
__global__ void __launch_bounds__(256)
zp_kernel(const dtts_zp_params p) {
  const long total = (long)p.M * p.C;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int m = (int)(idx / p.C), c = (int)(idx % p.C);
    float v = 0.f;
    if (!p.row_utt || p.row_utt[m] >= 0)
      v = p.m[(long)m * p.ld + c] + p.noise[(long)m * p.ldn + c] * expf(p.logs[(long)m * p.ld + c]) * p.noise_scale;
    p.out[(long)m * p.ldo + c] = v;
  }
}

__global__ void __launch_bounds__(256)
copy_utts_kernel(const dtts_copy_utts_params p) {
  const int b = blockIdx.y;
  const int c8 = p.C >> 3;   // 16-byte chunks per row
  const long total = (long)p.utt_len[b] * c8;
  const __half* s = (const __half*)p.src + (long)p.src_off[b] * p.ld_src;
  __half* d = (__half*)p.dst + (long)p.dst_off[b] * p.ld_dst;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int r = (int)(idx / c8), c = (int)(idx % c8) * 8;
    *reinterpret_cast<uint4*>(d + (long)r * p.ld_dst + c) = *reinterpret_cast<const uint4*>(s + (long)r * p.ld_src + c);
  }
}

__global__ void __launch_bounds__(256)
rowutt_kernel(const dtts_rowutt_params p) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= p.M) return;
  int u = -1;
  for (int b = 0; b < p.n_utt; ++b)
    if (m >= p.utt_off[b] && m < p.utt_off[b] + p.utt_len[b]) { u = b; break; }
  p.row_utt[m] = u;
}

inline int grid_for(long total, int threads = 256) {
  long g = (total + threads - 1) / threads;
  const long cap = 148L * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

// ---- prompt front-end: mel_spectrogram_torch (vqvae/utils/data_utils.py:105-155) ---------------------------------
// STFT as a DFT GEMM on the 3xTF32 tensor-core path (fp32-class): this kernel builds the windowed frames
// (reflect padding of (n_fft-hop)/2 samples at both ends, center=False) already split into tf32 hi/lo operands.
__global__ void __launch_bounds__(256)
stft_frames_kernel(const dtts_stft_frames_params p) {
  const int b = blockIdx.y;
  const int n = p.wav_len[b], nfr = p.utt_len[b];
  const long row0 = p.utt_off[b];
  const float* w = p.wav + (long)b * p.ldw;
  const long total = (long)nfr * p.n_fft;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int r = (int)(idx / p.n_fft), k = (int)(idx - (long)r * p.n_fft);
    int i = r * p.hop + k - p.pad;
    if (!p.zero_pad) {
      if (i < 0) i = -i;                     // reflect (no edge repeat), torch F.pad(mode="reflect")
      if (i >= n) i = 2 * (n - 1) - i;
    }
    const float v = (i >= 0 && i < n) ? (p.window ? w[i] * __ldg(p.window + k) : w[i]) : 0.f;
    const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    p.out_hi[(row0 + r) * p.ld + k] = hi;
    p.out_lo[(row0 + r) * p.ld + k] = v - hi;
  }
}

// |X| = sqrt(re^2 + im^2 + 1e-6) from the DFT GEMM output [M, >= 2*n_bins] (re | im), split into tf32 hi/lo for the mel GEMM
__global__ void __launch_bounds__(256)
spec_mag_kernel(const dtts_spec_mag_params p) {
  const long total = (long)p.M * p.ld;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int m = (int)(idx / p.ld), f = (int)(idx - (long)m * p.ld);
    float v = 0.f;
    if (f < p.n_bins) {
      const float re = p.spec[(long)m * p.lds + f], im = p.spec[(long)m * p.lds + p.n_bins + f];
      v = sqrtf(re * re + im * im + p.eps);
    }
    const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    p.out_hi[idx] = hi;
    p.out_lo[idx] = v - hi;
  }
}

}  // namespace

extern "C" int dtts_p_sample_step(const dtts_pstep_params* p, void* stream) {
  DTTS_REQUIRE(p && p->x && p->out_c && p->out_u && p->noise, "p_sample_step: null argument");
  DTTS_REQUIRE(p->C % 4 == 0 && p->ldx % 4 == 0 && p->ldo % 4 == 0 && p->ldn % 4 == 0, "p_sample_step: C/ld must be multiples of 4");
  if (p->M <= 0) return 0;
  pstep_kernel<<<grid_for((long)p->M * p->C / 4), 256, 0, (cudaStream_t)stream>>>(*p);
  DTTS_CHECK_LAUNCH("p_sample_step");
  return 0;
}

extern "C" int dtts_bct_to_rows(const dtts_bct2rows_params* p, void* stream) {
  DTTS_REQUIRE(p && p->src && p->utt_off && p->utt_len && (p->dst_f32 || p->dst_f16), "bct_to_rows: null argument");
  if (p->B <= 0 || p->T <= 0) return 0;
  dim3 grid(ceil_div(p->T, 32), ceil_div(p->C, 32), p->B);
  bct2rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*p);
  DTTS_CHECK_LAUNCH("bct_to_rows");
  return 0;
}

extern "C" int dtts_rows_to_bct(const dtts_rows2bct_params* p, void* stream) {
  DTTS_REQUIRE(p && p->src && p->utt_off && p->utt_len && p->dst, "rows_to_bct: null argument");
  if (p->B <= 0 || p->T <= 0) return 0;
  dim3 grid(ceil_div(p->T, 32), ceil_div(p->C, 32), p->B);
  rows2bct_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*p);
  DTTS_CHECK_LAUNCH("rows_to_bct");
  return 0;
}

extern "C" int dtts_eltwise(const dtts_eltwise_params* p, void* stream) {
  DTTS_REQUIRE(p && p->x && (p->out_f32 || p->out_f16), "eltwise: null argument");
  if (p->M <= 0) return 0;
  eltwise_kernel<<<grid_for((long)p->M * p->C), 256, 0, (cudaStream_t)stream>>>(*p);
  DTTS_CHECK_LAUNCH("eltwise");
  return 0;
}

extern "C" int dtts_embed(const dtts_embed_params* p, void* stream) {
  DTTS_REQUIRE(p && p->ids && p->table && p->out, "embed: null argument");
  if (p->n <= 0) return 0;
  embed_kernel<<<p->n, 256, 0, (cudaStream_t)stream>>>(*p);
  DTTS_CHECK_LAUNCH("embed");
  return 0;
}

extern "C" int dtts_repeat_rows(const dtts_repeat_rows_params* p, void* stream) {
  DTTS_REQUIRE(p && p->x && p->out && p->utt_off && p->utt_len && p->out_off, "repeat_rows: null argument");
  DTTS_REQUIRE(p->C % 4 == 0 && p->ldx % 4 == 0 && p->ldo % 4 == 0 && p->repeat >= 1, "repeat_rows: C/ld must be multiples of 4");
  if (p->n_utt <= 0) return 0;
  dim3 grid(64, p->n_utt);
  repeat_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*p);
  DTTS_CHECK_LAUNCH("repeat_rows");
  return 0;
}

extern "C" int dtts_mean_rows(const dtts_mean_rows_params* p, void* stream) {
  DTTS_REQUIRE(p && p->x && p->out && p->utt_off && p->utt_len, "mean_rows: null argument");
  if (p->n_utt <= 0) return 0;
  dim3 grid(ceil_div(p->C, 32), p->n_utt);
  mean_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*p);
  DTTS_CHECK_LAUNCH("mean_rows");
  return 0;
}

extern "C" int dtts_timestep_embedding(const dtts_tsemb_params* p, void* stream) {
  DTTS_REQUIRE(p && p->t && (p->out || p->out_f16) && p->dim % 2 == 0, "timestep_embedding: bad argument");
  if (p->n <= 0) return 0;
  tsemb_kernel<<<p->n, 256, 0, (cudaStream_t)stream>>>(*p);
  DTTS_CHECK_LAUNCH("timestep_embedding");
  return 0;
}

extern "C" int dtts_flow_couple(const dtts_couple_params* p, void* stream) {
  DTTS_REQUIRE(p && p->x && p->half > 0, "flow_couple: null argument");
  if (p->M <= 0) return 0;
  couple_kernel<<<grid_for((long)p->M * p->half), 256, 0, (cudaStream_t)stream>>>(*p);
  DTTS_CHECK_LAUNCH("flow_couple");
  return 0;
}

extern "C" int dtts_sample_zp(const dtts_zp_params* p, void* stream) {
  DTTS_REQUIRE(p && p->m && p->logs && p->noise && p->out, "sample_zp: null argument");
  if (p->M <= 0) return 0;
  zp_kernel<<<grid_for((long)p->M * p->C), 256, 0, (cudaStream_t)stream>>>(*p);
  DTTS_CHECK_LAUNCH("sample_zp");
  return 0;
}

extern "C" int dtts_copy_utt_rows(const dtts_copy_utts_params* p, void* stream) {
  DTTS_REQUIRE(p && p->src && p->dst && p->src_off && p->dst_off && p->utt_len, "copy_utt_rows: null argument");
  DTTS_REQUIRE(p->C % 8 == 0 && p->ld_src % 8 == 0 && p->ld_dst % 8 == 0, "copy_utt_rows: C/ld must be multiples of 8");
  DTTS_REQUIRE(((((uintptr_t)p->src) | ((uintptr_t)p->dst)) & 15) == 0, "copy_utt_rows: pointers must be 16-byte aligned");
  if (p->n_utt <= 0) return 0;
  dim3 grid(32, p->n_utt);
  copy_utts_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*p);
  DTTS_CHECK_LAUNCH("copy_utt_rows");
  return 0;
}

extern "C" int dtts_fill_row_utt(const dtts_rowutt_params* p, void* stream) {
  DTTS_REQUIRE(p && p->row_utt && p->utt_off && p->utt_len, "fill_row_utt: null argument");
  if (p->M <= 0) return 0;
  rowutt_kernel<<<ceil_div(p->M, 256), 256, 0, (cudaStream_t)stream>>>(*p);
  DTTS_CHECK_LAUNCH("fill_row_utt");
  return 0;
}

// ---------------------------------------------------------------------------------------------
extern "C" int dtts_abi_version(void) { return DTTS_ABI_VERSION; }
extern "C" const char* dtts_last_error(void) { return g_dtts_err; }
extern "C" int dtts_kernel_launches(void) { return g_dtts_launches; }
extern "C" int dtts_set_pdl(int on) { const int old = g_dtts_pdl; g_dtts_pdl = on ? 1 : 0; return old; }
extern "C" int dtts_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) DTTS_FAIL(-6, "device_info: no CUDA device");
  cudaDeviceProp pr;
  if (cudaGetDeviceProperties(&pr, dev) != cudaSuccess) DTTS_FAIL(-6, "device_info: cudaGetDeviceProperties failed");
  if (sm_count) *sm_count = pr.multiProcessorCount;
  if (cc_major) *cc_major = pr.major;
  if (cc_minor) *cc_minor = pr.minor;
  return 0;
}
#define SZ(name) if (!strcmp(struct_name, #name)) return (int)sizeof(name)
extern "C" int dtts_stft_frames(const dtts_stft_frames_params* p, void* stream) {
  DTTS_REQUIRE(p && p->wav && p->wav_len && p->utt_off && p->utt_len && (p->window || p->zero_pad) && p->out_hi && p->out_lo, "stft_frames: null argument");
  DTTS_REQUIRE(p->n_fft > 0 && p->hop > 0 && p->pad >= 0 && p->ld >= p->n_fft, "stft_frames: bad shape");
  if (p->n_utt <= 0 || p->max_frames <= 0) return 0;
  long g = ((long)p->max_frames * p->n_fft + 255) / 256;
  if (g > 2048) g = 2048;
  stft_frames_kernel<<<dim3((unsigned)g, p->n_utt), 256, 0, (cudaStream_t)stream>>>(*p);
  DTTS_CHECK_LAUNCH("stft_frames");
  return 0;
}

extern "C" int dtts_spec_mag(const dtts_spec_mag_params* p, void* stream) {
  DTTS_REQUIRE(p && p->spec && p->out_hi && p->out_lo, "spec_mag: null argument");
  DTTS_REQUIRE(p->n_bins > 0 && p->ld >= p->n_bins && p->lds >= 2 * p->n_bins, "spec_mag: bad shape");
  if (p->M <= 0) return 0;
  long g = ((long)p->M * p->ld + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  spec_mag_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>(*p);
  DTTS_CHECK_LAUNCH("spec_mag");
  return 0;
}

extern "C" int dtts_sizeof(const char* struct_name) {
  if (!struct_name) return -1;
  SZ(dtts_gemm_params); SZ(dtts_groupnorm_params); SZ(dtts_layernorm_params); SZ(dtts_attention_params);
  SZ(dtts_logits_params); SZ(dtts_append_params); SZ(dtts_pstep_params); SZ(dtts_bct2rows_params);
  SZ(dtts_rows2bct_params); SZ(dtts_eltwise_params); SZ(dtts_embed_params); SZ(dtts_repeat_rows_params);
  SZ(dtts_mean_rows_params); SZ(dtts_tsemb_params); SZ(dtts_couple_params); SZ(dtts_zp_params);
  SZ(dtts_rowutt_params); SZ(dtts_copy_utts_params); SZ(dtts_split_params); SZ(dtts_reduce_params);
  SZ(dtts_voc_mrf_params); SZ(dtts_conv_post_params); SZ(dtts_stft_frames_params); SZ(dtts_spec_mag_params);
  SZ(dtts_gn_apply_params); SZ(dtts_zero_params); SZ(dtts_gpt_step_params);
  SZ(dtts_dgemm_params); SZ(dtts_final_ln_params); SZ(dtts_decode_tail_params);
  return -1;
}
