"""WAV file I/O for the file harness (reference: quantum_distortion/io/audio_io.py:10-26).

The reference uses ``soundfile`` (libsndfile), which is not part of this image; ``scipy.io.wavfile`` is used
instead with libsndfile's conversions restated: integer PCM is read as ``value / 2**(bits-1)`` and float data
is written as 16-bit PCM ``lrint(x * 32767)`` (libsndfile's default normalisation, WAV default subtype PCM_16).
"""
from __future__ import annotations

from pathlib import Path
from typing import Tuple, Union

import numpy as np
from scipy.io import wavfile


def load_audio(path: Union[str, Path]) -> Tuple[np.ndarray, int]:
    """-> (float32 audio [n] or [n, channels], sample rate)."""
    sr, data = wavfile.read(str(path))
    if data.dtype == np.int16:
        audio = data.astype(np.float32) / 32768.0
    elif data.dtype == np.int32:
        audio = (data.astype(np.float64) / 2147483648.0).astype(np.float32)
    elif data.dtype == np.uint8:
        audio = (data.astype(np.float32) - 128.0) / 128.0
    else:
        audio = data.astype(np.float32)
    return audio, int(sr)


def load_pcm16(path: Union[str, Path]):
    """-> (int16 samples [n] or [n, channels], sample rate) when the file holds 16-bit PCM, else (None, sample rate).
    The batch harness ships such files to the GPU as they are (int16 on PCIe, ``sample / 32768`` on the device)."""
    sr, data = wavfile.read(str(path))
    return (data if data.dtype == np.int16 else None), int(sr)


def save_pcm16(path: Union[str, Path], pcm: np.ndarray, sr: int) -> None:
    """Write int16 samples that are already in the file format (the device applied ``float_to_pcm16``'s rule)."""
    path = Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    wavfile.write(str(path), int(sr), np.ascontiguousarray(pcm, dtype=np.int16))


def float_to_pcm16(audio: np.ndarray) -> np.ndarray:
    """libsndfile float -> PCM_16: round-half-even of x * 0x7FFF (clipped to the int16 range)."""
    return np.clip(np.rint(np.asarray(audio, dtype=np.float64) * 32767.0), -32768, 32767).astype(np.int16)


def save_audio(path: Union[str, Path], audio: np.ndarray, sr: int) -> None:
    path = Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    wavfile.write(str(path), int(sr), float_to_pcm16(audio))
