"""Golden vectors for the prompt resampler: torchaudio.transforms.Resample(sr, 24000) exactly as /root/reference/api.py:37
applies it (torchaudio is the third-party dependency the reference calls; version as installed here), on seeded signals and
on the first 0.5 s of the reference's own prompt 1.wav (44.1 kHz).  Refuses to write unless oracle/frontend.py agrees.
Run in the build container:  python tests/golden/make_resample.py"""
import os
import sys
import warnings

import torch
import torchaudio

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle.frontend as ofe  # noqa: E402

g = torch.Generator().manual_seed(77)
items = {}
t = torch.arange(22050) / 44100.0
items["synthetic_44100"] = (44100, torch.stack([0.4 * torch.sin(2 * torch.pi * 440.0 * t) + 0.1 * torch.randn(22050, generator=g),
                                                0.6 * torch.sin(2 * torch.pi * (100.0 + 9000.0 * t) * t)]))
items["synthetic_16000"] = (16000, 0.5 * torch.randn(1, 8001, generator=g))
try:
    import scipy.io.wavfile
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sr, data = scipy.io.wavfile.read("/root/reference/1.wav")
    items["prompt_1wav"] = (sr, torch.from_numpy(data.astype("float32") / 32768.0)[None, :sr // 2])
except Exception as e:  # pragma: no cover
    print("1.wav not used:", e)
out = {}
for name, (sr, y) in items.items():
    ref = torchaudio.transforms.Resample(sr, 24000)(y)
    mine = ofe.resample(y, sr, 24000)
    err = float((ref - mine).abs().max())
    print(name, sr, tuple(y.shape), "->", tuple(ref.shape), "oracle max abs err", err)
    assert ref.shape == mine.shape and err < 1e-5, err     # fp32 summation order of the 171-tap filter
    out[name] = {"sr": sr, "wav": y.clone(), "out": ref.clone()}
torch.save(out, os.path.join(HERE, "resample.pt"))
print("wrote resample.pt", os.path.getsize(os.path.join(HERE, "resample.pt")), "bytes")
