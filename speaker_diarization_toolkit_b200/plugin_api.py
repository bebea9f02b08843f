"""Host-side mirror of the reference's backend plugin interface.

Mirrors (same names, arguments and error behaviour; re-implemented, not copied):
  * `EmbeddingBackend`            speaker_detection_backends/base.py:22-200
  * `get_backend`, `list_backends`, `reload_backends_config`   base.py:203-304
  * `AudioProfile`, `PROFILES`, `get_profile`, `format_ffmpeg_args`, `register_profile`
                                  speaker_detection_backends/audio_profiles.py:12-111

When the reference package itself is importable (`speaker_detection_backends` on sys.path) its own
`EmbeddingBackend` is used as the base class, so a `Backend()` built here passes the reference CLI's
isinstance / duck-typing unchanged.  On a box without the reference (the GPU box) the mirror stands alone.
"""
from __future__ import annotations

import importlib
import os
import sys
from abc import ABC, abstractmethod
from dataclasses import dataclass
from pathlib import Path
from typing import Any, Dict, List, Optional, Tuple, Union

from . import transcript as _transcript


# ---- audio_profiles.py mirror (boundary only: no arithmetic on the hot path) ---------------------
@dataclass
class AudioProfile:
    sample_rate: int = 16000
    channels: int = 1
    format: str = "wav"
    bit_depth: int = 16
    max_duration_sec: Optional[float] = None


PROFILES: Dict[str, AudioProfile] = {
    "speechmatics": AudioProfile(16000, 1, "wav", 16),
    "pyannote": AudioProfile(16000, 1, "wav", 16),
    "b200": AudioProfile(16000, 1, "wav", 16),
    "default": AudioProfile(),
}

_WAV_CODECS = {8: "pcm_u8", 16: "pcm_s16le", 24: "pcm_s24le", 32: "pcm_s32le"}


def get_profile(backend_name: str) -> AudioProfile:
    return PROFILES.get(backend_name, PROFILES["default"])


def register_profile(name: str, profile: AudioProfile) -> None:
    PROFILES[name] = profile


def format_ffmpeg_args(profile: AudioProfile) -> List[str]:
    args = ["-ar", str(profile.sample_rate), "-ac", str(profile.channels), "-f", profile.format]
    if profile.format == "wav" and profile.bit_depth in _WAV_CODECS:
        args += ["-acodec", _WAV_CODECS[profile.bit_depth]]
    return args


# ---- EmbeddingBackend mirror --------------------------------------------------------------------
class _MirrorEmbeddingBackend(ABC):
    """Same contract as base.py:22-200."""

    @property
    @abstractmethod
    def name(self) -> str: ...

    @property
    @abstractmethod
    def requires_api_key(self) -> bool: ...

    @property
    def embedding_dim(self) -> Optional[int]:
        return None

    @property
    def model_version(self) -> str:
        return f"{self.name}-unknown"

    @property
    def audio_profile(self) -> Union[str, AudioProfile]:
        return "default"

    def get_audio_profile(self) -> AudioProfile:
        prof = self.audio_profile
        return get_profile(prof) if isinstance(prof, str) else prof

    def check_embedding_compatibility(self, embedding: Dict[str, Any]) -> Dict[str, Any]:
        version = embedding.get("model_version", "unknown")
        ok = version.startswith(f"{self.name}-")
        return {
            "compatible": ok,
            "version": version,
            "current": self.model_version,
            "warning": None if ok else (f"Embedding created with {version} may not work with "
                                        f"backend {self.name}. Consider re-enrolling."),
        }

    @abstractmethod
    def enroll_speaker(self, audio_path: Path, segments: Optional[List[Tuple[float, float]]] = None) -> Dict[str, Any]: ...

    @abstractmethod
    def identify_speaker(self, audio_path: Path, candidates: List[Dict[str, Any]], threshold: float = 0.354) -> List[Dict[str, Any]]: ...

    def verify_speaker(self, audio_path: Path, speaker_profile: Dict[str, Any], threshold: float = 0.354) -> Dict[str, Any]:
        hits = self.identify_speaker(audio_path, [speaker_profile], threshold)
        if not hits:
            return {"match": False, "similarity": 0.0, "embedding_id": None}
        return {"match": True, "similarity": hits[0]["similarity"], "embedding_id": hits[0].get("embedding_id")}

    def extract_segments_from_transcript(self, transcript_path: Path, speaker_label: str) -> List[Tuple[float, float]]:
        return _transcript.extract_segments_as_tuples(_transcript.load_transcript(transcript_path), speaker_label)


try:  # prefer the reference's own ABC when it is on sys.path (drop-in under the reference CLIs)
    from speaker_detection_backends.base import EmbeddingBackend as EmbeddingBackend  # type: ignore
    USING_REFERENCE_ABC = True
except Exception:  # pragma: no cover - exercised on boxes without the reference
    EmbeddingBackend = _MirrorEmbeddingBackend  # type: ignore
    USING_REFERENCE_ABC = False


# ---- registry mirror (base.py:203-304) ----------------------------------------------------------
_DEFAULT_BACKENDS = {"b200": "speaker_diarization_toolkit_b200.backend"}
_loaded: Optional[Dict[str, str]] = None


def _config_file() -> Optional[Path]:
    env = os.environ.get("SPEAKER_BACKENDS_CONFIG")
    if env:
        p = Path(env)
        if p.exists():
            return p
        print(f"Warning: SPEAKER_BACKENDS_CONFIG not found: {p}", file=sys.stderr)
    sibling = Path(__file__).parent / "backends.yaml"
    return sibling if sibling.exists() else None


def _load_backends_config() -> Dict[str, str]:
    global _loaded
    if _loaded is not None:
        return _loaded
    table = None
    cfg = _config_file()
    if cfg is not None:
        try:
            import yaml
            doc = yaml.safe_load(cfg.read_text()) or {}
            table = {}
            for name, info in (doc.get("backends") or {}).items():
                if isinstance(info, dict):
                    table[name] = info.get("module", "")
                elif isinstance(info, str):
                    table[name] = info
        except ImportError:
            table = None
        except Exception as exc:
            print(f"Warning: Failed to load backends config: {exc}", file=sys.stderr)
            table = None
    _loaded = table if table is not None else dict(_DEFAULT_BACKENDS)
    return _loaded


def get_backend(name: str):
    table = _load_backends_config()
    if name not in table:
        raise ValueError(f"Unknown backend: {name}. Available: {', '.join(table.keys())}")
    return importlib.import_module(table[name]).Backend()


def list_backends() -> List[str]:
    return list(_load_backends_config().keys())


def reload_backends_config() -> None:
    global _loaded
    _loaded = None
