"""B200-native embedding-matching hot path of the speaker-diarization toolkit.

Drop-in scope: the arithmetic behind
`EmbeddingBackend.identify_speaker` (speaker_detection_backends/base.py:130-151) and the
`combine_signals` assignment (speaker-assign:418-492), as hand-written sm_100a CUDA behind a C-ABI
(`include/sdk_b200.h`), plus the host-side mirror of the plugin / CLI interface that calls it.
"""
__version__ = "0.1.0"
BACKEND_NAME = "b200"
