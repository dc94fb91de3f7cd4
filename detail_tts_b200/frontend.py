"""Prompt front-end of the synthesis path on the dtts kernels: wav -> log-mel spectrogram.

Mirrors `mel_spectrogram_torch(y, n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax, center=False)`
(vqvae/utils/data_utils.py:105-155, called from api.py:39-45): reflect padding of (n_fft - hop)/2 samples,
STFT with a periodic Hann window (center=False), magnitude sqrt(re^2 + im^2 + 1e-6), slaney mel filterbank
(librosa.filters.mel defaults), log(clamp(., 1e-5)).
Design: the STFT is a DFT GEMM and the filterbank a second GEMM, both on the 3xTF32 tcgen05 path (fp32-class:
the prompt mel feeds the GPT conditioning encoder, whose logits must stay token-exact); framing/windowing,
the magnitude and the log are fused into the operand-producing kernels / the GEMM epilogue.  Varlen batches
use the rows layout (one frame per row).
"""
import math

import numpy as np
import torch

from . import ops, pack
from .ops import RowsLayout


def slaney_mel_filterbank(sr, n_fft, n_mels, fmin=0.0, fmax=None):
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax) with its defaults (htk=False, norm='slaney'), restated
    from its published definition in float64 -> [n_mels, 1 + n_fft//2] float32 (the reference calls it at
    vqvae/utils/data_utils.py:113-118)."""
    fmax = sr / 2.0 if fmax is None else float(fmax)
    f_sp, min_log_hz = 200.0 / 3.0, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, math.log(6.4) / 27.0

    def hz_to_mel(f):
        f = np.asarray(f, dtype=np.float64)
        return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, f / f_sp)

    def mel_to_hz(m):
        m = np.asarray(m, dtype=np.float64)
        return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)

    fftfreqs = np.linspace(0.0, sr / 2.0, 1 + n_fft // 2)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    w = np.maximum(0.0, np.minimum(lower, upper))
    w *= (2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels]))[:, None]
    return w.astype(np.float32)


class MelFrontEnd:
    """Packed constants (DFT basis, mel filterbank, window) for one STFT configuration."""

    def __init__(self, n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax, device):
        assert win_size == n_fft, "the reference configuration uses win_length == filter_length"
        self.n_fft, self.hop, self.n_mels, self.device = n_fft, hop_size, num_mels, torch.device(device)
        self.pad = int((n_fft - hop_size) / 2)
        self.n_bins = n_fft // 2 + 1
        k = np.arange(n_fft, dtype=np.float64)
        f = np.arange(self.n_bins, dtype=np.float64)
        ang = 2.0 * np.pi * np.outer(f, k) / n_fft
        basis = np.concatenate([np.cos(ang), -np.sin(ang)], 0)                  # [2*n_bins, n_fft]
        self.dft = pack.pack_linear_tf32x3(torch.from_numpy(basis).float(), None, self.device, n_pad=8)
        mel = slaney_mel_filterbank(sampling_rate, n_fft, num_mels, fmin, fmax)   # [n_mels, n_bins]
        self.kp = (self.n_bins + 7) // 8 * 8
        melp = torch.zeros(num_mels, self.kp)
        melp[:, :self.n_bins] = torch.from_numpy(mel)
        self.mel = pack.pack_linear_tf32x3(melp, None, self.device, n_pad=4)
        self.window = torch.hann_window(win_size, dtype=torch.float32).to(self.device)

    @torch.no_grad()
    def __call__(self, y, lengths=None):
        """y [B, N] fp32 waveform (values in [-1, 1]); lengths: valid samples per row -> (mel [B, n_mels, Rmax], frames)."""
        dev = self.device
        y = y.to(dev, torch.float32).contiguous()
        B, N = y.shape
        lens = [N] * B if lengths is None else [int(v) for v in lengths]
        assert min(lens) > self.pad, "reflect padding needs more samples than the pad width"
        frames = [(n + 2 * self.pad - self.n_fft) // self.hop + 1 for n in lens]
        lay = RowsLayout(frames, 0, dev)
        M = lay.M
        L = ops._lib.lib()
        fh = torch.empty(M, self.n_fft, device=dev)
        fl = torch.empty(M, self.n_fft, device=dev)
        L.call("dtts_stft_frames", wav=y, ldw=N, wav_len=torch.tensor(lens, dtype=torch.int32, device=dev), n_utt=B,
               utt_off=lay.off, utt_len=lay.len, max_frames=max(frames), n_fft=self.n_fft, hop=self.hop, pad=self.pad,
               window=self.window, out_hi=fh, out_lo=fl, ld=self.n_fft)
        spec = torch.empty(M, self.dft.N, device=dev)
        ops.gemm_tf32x3(fh, fl, self.dft, spec, bias=False)
        mh = torch.empty(M, self.kp, device=dev)
        ml = torch.empty(M, self.kp, device=dev)
        L.call("dtts_spec_mag", spec=spec, lds=self.dft.N, M=M, n_bins=self.n_bins, eps=1e-6, out_hi=mh, out_lo=ml, ld=self.kp)
        mel = torch.empty(M, self.n_mels, device=dev)
        ops.gemm_tf32x3(mh, ml, self.mel, mel, bias=False, act=ops.ACT_LOG_CLAMP, act_param=1e-5)
        out = torch.empty(B, self.n_mels, max(frames), device=dev)
        ops.rows_to_bct(mel, lay, out)
        return out, frames


def sinc_resample_kernel(orig_freq, new_freq, lowpass_filter_width=6, rolloff=0.99):
    """The windowed-sinc polyphase filter bank of torchaudio.functional.resample (`sinc_interp_hann`, the defaults
    torchaudio.transforms.Resample uses at api.py:37), restated from its definition in float64:
    kernel[j, k] = sinc(t) * cos^2(t*pi/(2*lpw)) * base/orig with t = clamp((-j/new + (k - width)/orig) * base, +-lpw).
    Returns (kernel [new, 2*width+orig] fp32, width, orig, new) for the reduced ratio orig/new."""
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    base = min(orig, new) * rolloff
    width = int(math.ceil(lowpass_filter_width * orig / base))
    idx = np.arange(-width, width + orig, dtype=np.float64)[None, :] / orig
    t = (np.arange(0, -new, -1, dtype=np.float64)[:, None] / new + idx) * base
    t = np.clip(t, -lowpass_filter_width, lowpass_filter_width)
    window = np.cos(t * np.pi / lowpass_filter_width / 2) ** 2
    t = t * np.pi
    with np.errstate(invalid="ignore", divide="ignore"):
        k = np.where(t == 0, 1.0, np.sin(t) / t) * window * (base / orig)
    return k.astype(np.float32), width, orig, new


class Resample:
    """torchaudio.transforms.Resample(orig_freq, new_freq) (api.py:37) on the GPU: zero-padded framing (dtts_stft_frames,
    frame f = samples [f*orig - width, f*orig + width + orig)) x the [new, 2*width+orig] sinc kernel as one 3xTF32
    tensor-core GEMM (fp32-class); row f of the result holds output samples f*new .. f*new+new-1."""

    def __init__(self, orig_freq, new_freq, device="cuda", lowpass_filter_width=6, rolloff=0.99):
        self.device = torch.device(device)
        k, self.width, self.orig, self.new = sinc_resample_kernel(orig_freq, new_freq, lowpass_filter_width, rolloff)
        self.klen = k.shape[1]
        self.kp = (self.klen + 3) // 4 * 4
        kpad = torch.zeros(self.new, self.kp)
        kpad[:, :self.klen] = torch.from_numpy(k)
        self.kernel = pack.pack_linear_tf32x3(kpad, None, self.device, n_pad=4)

    @torch.no_grad()
    def __call__(self, waveform, lengths=None):
        """waveform [B, L] (or [L]) fp32 -> [B, ceil(new*L/orig)]; `lengths`: valid samples per row (ragged batch: every
        row is resampled as if alone; the tail beyond its own target length is zero)."""
        if self.orig == self.new:
            return waveform
        dev = self.device
        squeeze = waveform.dim() == 1
        y = waveform.reshape(-1, waveform.shape[-1]).to(dev, torch.float32).contiguous()
        B, L = y.shape
        lens = [L] * B if lengths is None else [int(v) for v in lengths]
        frames = [n // self.orig + 1 for n in lens]                   # conv1d(stride=orig) over the padded signal
        target = [-(-self.new * n // self.orig) for n in lens]        # ceil(new * L / orig)
        lay = RowsLayout(frames, 0, dev)
        fh = torch.zeros(lay.M, self.kp, device=dev)
        fl = torch.zeros(lay.M, self.kp, device=dev)
        ops._lib.lib().call("dtts_stft_frames", wav=y, ldw=L, wav_len=torch.tensor(lens, dtype=torch.int32, device=dev), n_utt=B,
                            utt_off=lay.off, utt_len=lay.len, max_frames=max(frames), n_fft=self.klen, hop=self.orig,
                            pad=self.width, window=None, out_hi=fh, out_lo=fl, ld=self.kp, zero_pad=1)
        res = torch.empty(lay.M, self.kernel.N, device=dev)
        ops.gemm_tf32x3(fh, fl, self.kernel, res, bias=False)
        out = torch.zeros(B, max(target), device=dev)
        for b in range(B):
            o, f = lay.offs[b], frames[b]
            out[b, :target[b]] = res[o:o + f, :self.new].reshape(-1)[:target[b]]
        return out[0] if squeeze else out


_CACHE = {}


def mel_spectrogram_torch(y, n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax, center=False, device="cuda"):
    """Reference signature (vqvae/utils/data_utils.py:105).  y [B, N] -> log-mel [B, num_mels, N // hop_size]."""
    assert not center, "the synthesis path calls it with center=False (api.py:39-45)"
    key = (n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax, str(device))
    if key not in _CACHE:
        _CACHE[key] = MelFrontEnd(n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax, device)
    return _CACHE[key](y)[0]
