"""The reference's `api.py` (lines 34-49) as two calls: prompt waveform -> log-mel on the GPU, then synthesis.

    model = load_model('vqvae', 'vqvae.pth', 'vqvae/configs/config_24k.json', 'cuda')
    wav = synthesize(model, text_tokens, audio, sr)            # audio [1, N] at sr Hz, text_tokens [1, L] int

Text normalisation (pypinyin) and BPE tokenisation stay on the host exactly as in the reference (api.py:19-26)."""
import torch

from .frontend import Resample, mel_spectrogram_torch

# vqvae/configs/config_24k.json `data`: filter_length, n_mel_channels, sampling_rate, hop_length, win_length, mel_fmin, mel_fmax
MEL_ARGS = (1024, 128, 24000, 256, 1024, 0.0, None)
_RESAMPLERS = {}


@torch.no_grad()
def prompt_from_wav(audio, sr, device="cuda"):
    """api.py:34-45: first channel -> torchaudio.transforms.Resample(sr, 24000) -> mel_spectrogram_torch.
    audio [C, N] or [N] float in [-1, 1]; returns (spec [1, 128, R] on `device`, spec_lengths [1])."""
    audio = audio.reshape(-1, audio.shape[-1])[:1]
    key = (int(sr), str(device))
    if key not in _RESAMPLERS:
        _RESAMPLERS[key] = Resample(int(sr), MEL_ARGS[2], device=device)
    y = _RESAMPLERS[key](audio)
    spec = mel_spectrogram_torch(y, *MEL_ARGS, device=device)
    return spec, torch.tensor([spec.shape[-1]], device=spec.device)


@torch.no_grad()
def synthesize(model, text_tokens, audio, sr, noise_scale=0.667, **kw):
    """api.py:24-49: `text_tokens` [1, L] as produced by VoiceBpeTokenizer.encode (the trailing pad of api.py:25 is added
    here), prompt `audio` at `sr` Hz.  Returns wav [1, 1, 1024*T] at 24 kHz (what api.py:49 saves)."""
    text_tokens = torch.nn.functional.pad(text_tokens.reshape(1, -1).to(torch.int32), (0, 1))     # api.py:25
    spec, spec_lengths = prompt_from_wav(audio, sr, device=model.device)
    text_lengths = torch.tensor([text_tokens.shape[-1]])
    return model.infer(text_tokens, text_lengths, spec, spec_lengths, noise_scale=noise_scale, **kw)
