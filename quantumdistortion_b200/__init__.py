"""quantumdistortion_b200 -- B200-native (sm_100a) implementation of the STFT processing path of
TGALLOWAY1/QuantumDistortion behind the reference's ``process_audio`` / ``PipelineConfig`` API.

    from quantumdistortion_b200 import process_audio, process_batch, PipelineConfig

The heavy lifting is in ``csrc/`` (hand-written CUDA, C ABI in ``include/qd_b200.h``).
"""
from .config import PipelineConfig, ensure_mono_float32  # noqa: F401
from .presets import SPECTRAL_FX_PRESETS, get_preset, list_presets, spectral_fx_preset_kwargs  # noqa: F401


def __getattr__(name):  # lazy: importing the package must not need torch or a GPU
    if name in ("process_audio", "process_batch", "make_renderer", "Renderer", "RenderTiming"):
        from . import pipeline
        return getattr(pipeline, name)
    if name in ("process_file_to_file", "process_files"):
        from . import harness
        return getattr(harness, name)
    if name in ("load_audio", "save_audio"):
        from . import audio_io
        return getattr(audio_io, name)
    if name in ("peak_limiter", "linkwitz_riley_split", "apply_distortion"):
        from . import stages
        return getattr(stages, name)
    raise AttributeError(name)
