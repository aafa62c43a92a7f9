"""B200-native spectrogram-domain audio hot path (STFT / iSTFT / mel / Griffin-Lim) behind the reference's
``utils/audio.py::AudioProcessor`` API.  Import as ``your_voice_tts_b200`` (alias module at the repo root)."""
from .audio import AsyncAudioLogger, AudioProcessor, BatchLayout, HostPipeline  # noqa: F401
from . import _lib  # noqa: F401

__version__ = "0.1.0"
