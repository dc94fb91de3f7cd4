"""Golden vectors for the prompt front-end: the UNMODIFIED reference mel_spectrogram_torch
(/root/reference/vqvae/utils/data_utils.py:105-155, imported under the shims of refshim.py) on seeded synthetic
waveforms and on the reference's own prompt 1.wav (first 1.5 s, resampled to 24 kHz as api.py:36-38 does with
torchaudio.transforms.Resample).  Refuses to write the fixture unless oracle/frontend.py agrees.
Run in the build container:  python tests/golden/make_melspec.py"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import refshim  # noqa: E402

refshim.install()
from vqvae.utils.data_utils import mel_spectrogram_torch  # noqa: E402
import oracle.frontend as ofe  # noqa: E402

g = torch.Generator().manual_seed(1234)
N = 24000 // 2
t = torch.arange(N) / 24000.0
wav = torch.stack([0.3 * torch.sin(2 * torch.pi * 220.0 * t) + 0.05 * torch.randn(N, generator=g),
                   0.5 * torch.sin(2 * torch.pi * (300.0 + 2000.0 * t) * t) * torch.hann_window(N),
                   0.8 * (torch.rand(N, generator=g) * 2 - 1) * (t < 0.3)])
items = {"synthetic": wav}
try:
    import scipy.io.wavfile
    import torchaudio
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sr, data = scipy.io.wavfile.read("/root/reference/1.wav")
    a = torch.from_numpy(data.astype("float32") / 32768.0)[None, :int(1.5 * sr)]
    items["prompt_1wav"] = torchaudio.transforms.Resample(sr, 24000)(a)
except Exception as e:  # pragma: no cover
    print("1.wav not used:", e)
out = {}
for name, y in items.items():
    ref = mel_spectrogram_torch(y, 1024, 128, 24000, 256, 1024, 0.0, None)
    mine = ofe.mel_spectrogram(y)
    err = float((ref - mine).abs().max())
    print(name, tuple(y.shape), "->", tuple(ref.shape), "oracle max abs err", err)
    assert err < 2e-4, err
    out[name] = {"wav": y.clone(), "mel": ref.clone()}
torch.save(out, os.path.join(HERE, "melspec.pt"))
print("wrote", os.path.join(HERE, "melspec.pt"), os.path.getsize(os.path.join(HERE, "melspec.pt")), "bytes")
