"""detail_tts_b200 -- B200-native implementation of detail_tts's end-to-end synthesis hot path.

Public surface (mirrors the reference):
    load_model(model_name, model_path, config_path, device) -> SynthesizerTrn   prepare/load_infer.py:8
    SynthesizerTrn.infer / infer_batch / infer_flowvae                            vqvae/model_24k.py:774,848
    SynthesizerTrn.gpt.inference_speech_tortoise (alias inference_speech)         gpt/model.py:514
    SynthesizerTrn.diffusion / infer_diffuser.p_sample_loop / do_spectrogram_diffusion
    SynthesizerTrn.dec (Generator.forward)                                        vqvae/model_24k.py:269
The arithmetic runs in the C-ABI CUDA library declared in include/dtts.h (libdtts.so, sm_100a).
"""
__all__ = ["load_model", "SynthesizerTrn", "do_spectrogram_diffusion"]


def __getattr__(name):
    if name in ("load_model", "SynthesizerTrn"):
        from . import model
        return getattr(model, name)
    if name == "do_spectrogram_diffusion":
        from .diffusion import do_spectrogram_diffusion
        return do_spectrogram_diffusion
    raise AttributeError(name)
