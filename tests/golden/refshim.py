"""Import shims that let the UNMODIFIED reference (/root/reference) run in this container.

Only used by tests/golden/make_golden.py (fixture generation, run in the build container where
/root/reference exists).  Nothing on the GPU box imports this file's targets: /root/reference does
not travel.  Shim list follows SURVEY.md section 8c.
"""
import sys
import types
import json

import os

# The unmodified reference tree: /root/reference in the build container; on the GPU box the driver's snapshot carries the
# git-ignored install baseline/_ref (copied there by __graft_entry__.build(), never committed).
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("DTTS_REFERENCE") or ("/root/reference" if os.path.isdir("/root/reference/vqvae") else os.path.join(_ROOT, "baseline", "_ref"))


def available():
    return os.path.isfile(os.path.join(REF, "vqvae", "model_24k.py"))


def install():
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import torch  # noqa: F401

    # 1) k_diffusion (vqvae/utils/diffusion.py:8,13) - never called on the infer path
    kd = types.ModuleType("k_diffusion")
    kds = types.ModuleType("k_diffusion.sampling")
    kds.sample_dpmpp_2m = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("stub"))
    kds.sample_euler_ancestral = kds.sample_dpmpp_2m
    kd.sampling = kds
    sys.modules.setdefault("k_diffusion", kd)
    sys.modules.setdefault("k_diffusion.sampling", kds)

    # 2) transformers.utils.model_parallel_utils (gpt/model.py:8), removed in transformers 5
    import transformers
    from transformers import GPT2Model, GPT2Config, GPT2PreTrainedModel, LogitsProcessorList  # noqa: F401
    mpu = types.ModuleType("transformers.utils.model_parallel_utils")
    mpu.get_device_map = lambda *a, **k: None
    mpu.assert_device_map = lambda *a, **k: None
    sys.modules["transformers.utils.model_parallel_utils"] = mpu

    # 3) transformers.LogitsWarper (gpt/modules/typical_sampling.py:2)
    from transformers import LogitsProcessor
    sys.modules["transformers"].LogitsWarper = LogitsProcessor
    try:
        transformers.LogitsWarper = LogitsProcessor
    except Exception:
        pass

    # 4) librosa (vqvae/utils/data_utils.py:10-14); mel basis restated with torchaudio
    lib = types.ModuleType("librosa")
    libu = types.ModuleType("librosa.util")
    libf = types.ModuleType("librosa.filters")
    libu.normalize = libu.pad_center = libu.tiny = lambda *a, **k: None

    def _mel(sr, n_fft, n_mels, fmin, fmax):
        import torchaudio
        fmax = sr / 2 if fmax is None else fmax
        fb = torchaudio.functional.melscale_fbanks(n_fft // 2 + 1, fmin, fmax, n_mels, sr,
                                                   norm="slaney", mel_scale="slaney")
        return fb.T.contiguous().numpy()
    libf.mel = _mel
    lib.util, lib.filters = libu, libf
    sys.modules.setdefault("librosa", lib)
    sys.modules.setdefault("librosa.util", libu)
    sys.modules.setdefault("librosa.filters", libf)

    # 5) GenerationMixin for GPT2InferenceModel (transformers >= 4.50)
    import gpt.model as gm
    from transformers import GenerationMixin
    if GenerationMixin not in gm.GPT2InferenceModel.__bases__:
        gm.GPT2InferenceModel.__bases__ = gm.GPT2InferenceModel.__bases__ + (GenerationMixin,)
    return gm


def load_config():
    cfg = json.load(open(f"{REF}/vqvae/configs/config_24k.json"))
    cfg["diffusion"].pop("g_channels", None)  # 6) stale key, DiffusionTts.__init__ rejects it
    return cfg


def build_reference_model():
    """SynthesizerTrn as prepare/load_infer.py:8-34 builds it (random init)."""
    install()
    from vqvae.model_24k import SynthesizerTrn
    from vqvae.utils.data_utils import HParams
    cfg = load_config()
    hps = HParams(**cfg)
    model = SynthesizerTrn(hps.data.filter_length // 2 + 1,
                           hps.train.segment_size // hps.data.hop_length,
                           **hps.vaegan, cfg=hps)
    return model.eval(), cfg
