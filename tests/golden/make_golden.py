#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE's own utils/audio.py.

Run in the dev container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference module imports ``librosa`` and ``soundfile`` at top level; both
are absent here (no network).  We satisfy the imports with a shim whose three
entry points used by the hot path are backed by independent implementations of
the same published algorithms:

    librosa.stft         -> torch.stft   (float64, hann periodic, center, reflect)
    librosa.istft        -> torch.istft  (float64)
    librosa.filters.mel  -> torchaudio.functional.melscale_fbanks(norm='slaney',
                            mel_scale='slaney')  (float32 internally)

Everything else executed is the reference's unmodified code (utils/audio.py:
_normalize, _denormalize, _amp_to_db, _db_to_amp, apply_preemphasis (scipy
lfilter), spectrogram, melspectrogram, inv_spectrogram, inv_mel_spectrogram,
out_linear_to_mel, _griffin_lim incl. its np.random.rand phases).  ``np.complex``
(removed from numpy >= 1.24, used at utils/audio.py:184) is aliased to
``complex``.  The fixtures pin the oracle (oracle/audio_oracle.py) and the CUDA
path; nothing here is shipped in the product path.
"""
import json
import os
import re
import sys
import types

import numpy as np
import torch
import torchaudio

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def _install_librosa_shim():
    lib = types.ModuleType("librosa")
    filt = types.ModuleType("librosa.filters")
    eff = types.ModuleType("librosa.effects")

    def stft(y, n_fft=2048, hop_length=None, win_length=None):
        y = torch.as_tensor(np.asarray(y, dtype=np.float64))
        win = torch.hann_window(win_length, periodic=True, dtype=torch.float64)
        D = torch.stft(y, n_fft, hop_length=hop_length, win_length=win_length, window=win,
                       center=True, pad_mode="reflect", return_complex=True)
        return D.numpy()

    def istft(Y, hop_length=None, win_length=None):
        Y = torch.as_tensor(np.asarray(Y, dtype=np.complex128))
        n_fft = 2 * (Y.shape[0] - 1)
        win = torch.hann_window(win_length, periodic=True, dtype=torch.float64)
        y = torch.istft(Y, n_fft, hop_length=hop_length, win_length=win_length, window=win, center=True)
        return y.numpy()

    def mel(sr, n_fft, n_mels=128, fmin=0.0, fmax=None):
        if fmax is None:
            fmax = float(sr) / 2
        fb = torchaudio.functional.melscale_fbanks(
            n_freqs=1 + n_fft // 2, f_min=float(fmin), f_max=float(fmax), n_mels=int(n_mels),
            sample_rate=int(sr), norm="slaney", mel_scale="slaney")
        return fb.T.double().numpy()

    lib.stft, lib.istft = stft, istft
    filt.mel = mel
    lib.filters, lib.effects = filt, eff
    sys.modules["librosa"] = lib
    sys.modules["librosa.filters"] = filt
    sys.modules["librosa.effects"] = eff
    sys.modules["soundfile"] = types.ModuleType("soundfile")
    if not hasattr(np, "complex"):
        np.complex = complex  # utils/audio.py:184


def _load_ref_audio_processor():
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_audio", os.path.join(REF, "utils", "audio.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.AudioProcessor


def load_json_with_comments(path):
    # same comment stripping idea as utils/generic_utils.py:18-32
    txt = open(path).read()
    txt = re.sub(r"\\\n", "", txt)
    txt = re.sub(r"//.*\n", "\n", txt)
    return json.loads(txt)


def main():
    _install_librosa_shim()
    AP = _load_ref_audio_processor()
    from scipy.io import wavfile

    sr, wav_i16 = wavfile.read(os.path.join(REF, "tests", "inputs", "example_1.wav"))
    assert sr == 22050 and wav_i16.dtype == np.int16
    wav = wav_i16.astype(np.float64) / 32768.0  # soundfile.read convention

    cfgs = {
        "test": load_json_with_comments(os.path.join(REF, "tests", "test_config.json"))["audio"],
        "main": load_json_with_comments(os.path.join(REF, "config.json"))["audio"],
    }
    out = {"wav_i16": wav_i16, "sample_rate": np.int64(sr)}
    import contextlib, io
    for name, audio in cfgs.items():
        with contextlib.redirect_stdout(io.StringIO()):
            ap = AP(**audio)
        out[f"{name}_audio_json"] = np.frombuffer(json.dumps(audio).encode(), dtype=np.uint8)
        lin = ap.spectrogram(wav)            # [1025, 153]
        mel = ap.melspectrogram(wav)         # [80, 153]
        out[f"{name}_lin_sub4"] = lin[:, ::4].astype(np.float32)
        out[f"{name}_mel"] = mel.astype(np.float32)
        out[f"{name}_lin2mel"] = ap.out_linear_to_mel(lin.astype(np.float32)).astype(np.float32)
        # Griffin-Lim on a 40-frame excerpt with the reference's own RNG call
        T0, T1 = 60, 100
        lin_x = lin[:, T0:T1].astype(np.float32)
        mel_x = mel[:, T0:T1].astype(np.float32)
        out[f"{name}_gl_lin_in"] = lin_x
        out[f"{name}_gl_mel_in"] = mel_x
        # phases: np.random.seed(1234); 2*pi*np.random.rand(1025, 40) -- regenerated by the tests
        np.random.seed(1234)
        out[f"{name}_inv_spectrogram"] = ap.inv_spectrogram(lin_x).astype(np.float64)
        np.random.seed(1234)
        out[f"{name}_inv_mel_spectrogram"] = ap.inv_mel_spectrogram(mel_x).astype(np.float64)
        # raw stft / istft pair on a short excerpt
        y_x = wav[5000:5000 + 3000]
        D = ap._stft(y_x)
        out[f"{name}_stft_in"] = y_x
        out[f"{name}_stft_re"] = D.real.astype(np.float64)
        out[f"{name}_stft_im"] = D.imag.astype(np.float64)
        out[f"{name}_istft"] = ap._istft(D)
        out[f"{name}_mel_basis"] = ap._build_mel_basis().astype(np.float64)
    # mel basis tables for the other shipped geometries (config_tacotron_de.json, config_libritts.json)
    for (sr_, fmin, fmax) in [(16000, 0.0, 8000.0), (24000, 0.0, 8000.0), (22050, 50.0, None)]:
        import librosa
        out[f"melbasis_{sr_}_{int(fmin)}_{'none' if fmax is None else int(fmax)}"] = librosa.filters.mel(
            sr_, 2048, n_mels=80, fmin=fmin, fmax=fmax)
    path = os.path.join(HERE, "ref_shim_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
