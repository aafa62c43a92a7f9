"""ctypes binding of libttsa_b200.so (include/ttsa.h).  Fails loudly: there is no fallback path."""
import ctypes
import os
from ctypes import (POINTER, Structure, c_char_p, c_double, c_int, c_int32, c_int64, c_size_t, c_uint32, c_uint64,
                    c_void_p)

HERE = os.path.dirname(os.path.abspath(__file__))
# TTSA_LIB: A/B experiments against a second build (build.py with TTSA_BUILD_TAG); the shipped library otherwise
LIB_PATH = os.environ.get("TTSA_LIB") or os.path.join(HERE, "libttsa_b200.so")

TTSA_OK = 0
TTSA_ERR_BAD_ARG = -1
TTSA_ERR_BAD_CONFIG = -2
TTSA_ERR_UNSUPPORTED = -3
TTSA_ERR_WORKSPACE = -4
TTSA_ERR_CUDA = -5
TTSA_ERR_NO_DEVICE = -6

FEAT_PREEMPHASIS = 1
SPEC_MAGNITUDE, SPEC_NORM_DB = 0, 1
GL_DEEMPHASIS = 1
MEL_IN_AMPLITUDE, MEL_IN_NORM_DB = 0, 1
MEL_OUT_PLAIN, MEL_OUT_POWER, MEL_OUT_NORM_DB = 0, 1, 2
PCM_JOINT_PEAK, PCM_F32_ARITH = 1, 2
PW_NORMALIZE, PW_DENORMALIZE, PW_AMP_TO_DB, PW_DB_TO_AMP = 0, 1, 2, 3


class TtsaConfig(Structure):
    _fields_ = [("sample_rate", c_int32), ("num_mels", c_int32), ("num_freq", c_int32), ("n_fft", c_int32),
                ("hop_length", c_int32), ("win_length", c_int32), ("signal_norm", c_int32),
                ("symmetric_norm", c_int32), ("clip_norm", c_int32), ("griffin_lim_iters", c_int32),
                ("min_level_db", c_double), ("ref_level_db", c_double), ("power", c_double),
                ("preemphasis", c_double), ("max_norm", c_double), ("mel_fmin", c_double), ("mel_fmax", c_double)]


# name -> (restype, argtypes); every symbol include/ttsa.h declares
PROTOTYPES = {
    "ttsa_version": (c_int, []),
    "ttsa_last_error": (c_char_p, []),
    "ttsa_launch_count": (c_uint64, []),
    "ttsa_plan_create": (c_int, [POINTER(TtsaConfig), c_int, POINTER(c_void_p)]),
    "ttsa_plan_destroy": (c_int, [c_void_p]),
    "ttsa_plan_mel_basis": (c_int, [c_void_p, POINTER(c_double)]),
    "ttsa_plan_inv_mel_basis": (c_int, [c_void_p, POINTER(c_double)]),
    "ttsa_plan_mel_schedule": (c_int64, [c_void_p, POINTER(c_int32), POINTER(c_uint32), c_int64]),
    "ttsa_batch_from_frames": (c_int, [c_void_p, POINTER(c_int32), c_int32, POINTER(c_void_p)]),
    "ttsa_batch_from_wav_lengths": (c_int, [c_void_p, POINTER(c_int32), c_int32, POINTER(c_void_p)]),
    "ttsa_batch_from_frames_strided": (c_int, [c_void_p, POINTER(c_int32), c_int32, c_int64, POINTER(c_void_p)]),
    "ttsa_batch_from_wav_lengths_strided": (c_int, [c_void_p, POINTER(c_int32), c_int32, c_int64, POINTER(c_void_p)]),
    "ttsa_batch_destroy": (c_int, [c_void_p]),
    "ttsa_batch_total_frames": (c_int64, [c_void_p]),
    "ttsa_batch_total_samples": (c_int64, [c_void_p]),
    "ttsa_batch_offsets": (c_int, [c_void_p, POINTER(c_int64), POINTER(c_int64), POINTER(c_int32)]),
    "ttsa_stft_features": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_uint32, c_void_p]),
    "ttsa_stft": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ttsa_istft": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ttsa_griffin_lim_workspace_bytes": (c_size_t, [c_void_p, c_void_p]),
    "ttsa_griffin_lim": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_uint64, c_uint32, c_void_p,
                                 c_void_p, c_void_p, c_size_t, c_void_p]),
    "ttsa_griffin_lim_fast": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_uint64, c_uint32, c_double,
                                      c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "ttsa_mel_to_linear": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p]),
    "ttsa_linear_to_mel": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p]),
    "ttsa_preemphasis": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ttsa_deemphasis_workspace_bytes": (c_size_t, [c_void_p, c_void_p]),
    "ttsa_deemphasis": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "ttsa_wav_peaks": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ttsa_find_endpoint": (c_int, [c_void_p, c_void_p, c_void_p, c_double, c_double, c_void_p, c_void_p]),
    "ttsa_pcm16_workspace_bytes": (c_size_t, [c_void_p, c_void_p]),
    "ttsa_wav_to_pcm16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_uint32, c_int64, c_void_p, c_void_p, c_int64,
                                  c_void_p, c_size_t, c_void_p]),
    "ttsa_pointwise": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int64, c_void_p]),
    "ttsa_transpose": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p]),
}

_lib = None


class TtsaError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("ttsa error %d: %s" % (code, message))
        self.code = code
        self.message = message


def load():
    """Load the CUDA library.  Raises if it is missing -- the product path never falls back to CPU code."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        # a clean checkout: compile the CUDA library in-tree (nvcc, sm_100a); still no non-CUDA path
        try:
            import importlib.util
            spec = importlib.util.spec_from_file_location("ttsa_build", os.path.join(HERE, "build.py"))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            mod.build()
        except Exception as exc:          # noqa: BLE001 -- reported below
            raise RuntimeError("%s is missing and could not be built (%s). There is no CPU or library fallback for "
                               "the audio hot path." % (LIB_PATH, exc))
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "%s is missing: build it with `python your-voice-tts_b200/build.py` (or __graft_entry__.build()). "
            "There is no CPU or library fallback for the audio hot path." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc):
    if rc != TTSA_OK:
        msg = load().ttsa_last_error()
        raise TtsaError(rc, msg.decode() if msg else "")
    return rc
