"""ORACLE (test infrastructure, not a product path) -- CPU restatement of the prompt front-end
`mel_spectrogram_torch` (vqvae/utils/data_utils.py:105-155) as api.py:39-45 calls it (center=False).
Pinned: tests/golden/make_melspec.py runs the UNMODIFIED reference function (librosa's mel basis shimmed with
torchaudio's slaney filterbank, SURVEY.md section 8c shim 4) and refuses to write tests/golden/melspec.pt unless
this restatement agrees."""
import math

import numpy as np
import torch


def mel_basis(sr, n_fft, n_mels, fmin, fmax):
    """librosa.filters.mel defaults (slaney scale, slaney area normalisation), from its published definition;
    called at vqvae/utils/data_utils.py:113-118."""
    fmax = sr / 2.0 if fmax is None else float(fmax)
    f_sp, brk = 200.0 / 3.0, 1000.0
    brk_mel, step = brk / f_sp, math.log(6.4) / 27.0

    def to_mel(f):
        return brk_mel + math.log(f / brk) / step if f >= brk else f / f_sp

    def to_hz(m):
        return brk * math.exp(step * (m - brk_mel)) if m >= brk_mel else f_sp * m

    lo, hi = to_mel(fmin), to_mel(fmax)
    pts = [to_hz(lo + (hi - lo) * i / (n_mels + 1)) for i in range(n_mels + 2)]
    freqs = [sr / 2.0 * i / (n_fft // 2) for i in range(n_fft // 2 + 1)]
    W = np.zeros((n_mels, len(freqs)), dtype=np.float64)
    for m in range(n_mels):
        l, c, r = pts[m], pts[m + 1], pts[m + 2]
        for j, f in enumerate(freqs):
            W[m, j] = max(0.0, min((f - l) / (c - l), (r - f) / (r - c))) * 2.0 / (r - l)
    return torch.from_numpy(W.astype(np.float32))


def mel_spectrogram(y, n_fft=1024, num_mels=128, sampling_rate=24000, hop_size=256, win_size=1024, fmin=0.0, fmax=None):
    """data_utils.py:120-153: reflect pad -> torch.stft(center=False, periodic Hann) -> sqrt(|X|^2 + 1e-6) -> mel -> log(clamp 1e-5)."""
    pad = int((n_fft - hop_size) / 2)
    y = torch.nn.functional.pad(y.unsqueeze(1), (pad, pad), mode="reflect").squeeze(1)
    spec = torch.stft(y, n_fft, hop_length=hop_size, win_length=win_size, window=torch.hann_window(win_size), center=False,
                      pad_mode="reflect", normalized=False, onesided=True, return_complex=True)
    spec = torch.sqrt(torch.view_as_real(spec).pow(2).sum(-1) + 1e-6)
    mel = torch.matmul(mel_basis(sampling_rate, n_fft, num_mels, fmin, fmax), spec)
    return torch.log(torch.clamp(mel, min=1e-5))


def resample(y, orig_freq, new_freq, lowpass_filter_width=6, rolloff=0.99):
    """torchaudio.transforms.Resample(orig_freq, new_freq)(y) as api.py:37 applies it to the prompt (torchaudio
    functional.resample, method sinc_interp_hann), restated: float64 windowed-sinc kernel [new, 1, 2*width+orig],
    zero padding (width, width+orig), conv1d with stride orig, interleave, trim to ceil(new*L/orig).  y [B, L] fp32."""
    import math
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    if orig == new:
        return y
    base = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base)
    idx = torch.arange(-width, width + orig, dtype=torch.float64)[None, None] / orig
    t = torch.arange(0, -new, -1, dtype=torch.float64)[:, None, None] / new + idx
    t = (t * base).clamp_(-lowpass_filter_width, lowpass_filter_width)
    window = torch.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t = t * math.pi
    kernel = (torch.where(t == 0, torch.tensor(1.0, dtype=torch.float64), t.sin() / t) * window * (base / orig)).float()
    L = y.shape[-1]
    x = torch.nn.functional.pad(y.float(), (width, width + orig))
    out = torch.nn.functional.conv1d(x[:, None], kernel, stride=orig)          # [B, new, frames]
    out = out.transpose(1, 2).reshape(y.shape[0], -1)
    return out[:, :math.ceil(new * L / orig)]
