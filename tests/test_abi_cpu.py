"""CPU-only checks of the C-ABI library: it loads, exports every symbol include/ttsa.h declares, and its host
logic (mel basis, pseudo-inverse, batch layout, argument/ config errors) matches the oracle.  No compute calls."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import MAIN_AUDIO, ROOT, TEST_AUDIO
from oracle.audio_oracle import OracleAudioProcessor, lr_mel

import your_voice_tts_b200 as pkg
from your_voice_tts_b200 import AudioProcessor, _lib as L


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "ttsa.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(ttsa_[a-z_0-9]+)\s*\(", hdr))
    assert len(declared) >= 24
    assert declared == set(L.PROTOTYPES), declared ^ set(L.PROTOTYPES)
    lib = L.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.ttsa_version() == 100


def _host_plan(audio):
    ap = AudioProcessor(verbose=False, **audio)
    return ap, ap._plan(host_only=True)


@pytest.mark.parametrize("audio", [MAIN_AUDIO, TEST_AUDIO, dict(MAIN_AUDIO, sample_rate=16000),
                                   dict(MAIN_AUDIO, sample_rate=24000, mel_fmax=None, mel_fmin=50.0)])
def test_mel_basis_and_pinv_match_oracle(audio):
    ap = AudioProcessor(verbose=False, **audio)
    orc = OracleAudioProcessor(**audio)
    M = ap._build_mel_basis()
    np.testing.assert_allclose(M, orc._build_mel_basis(), atol=1e-14)
    P = ap._inv_mel_basis()
    Pref = np.linalg.pinv(orc._build_mel_basis())
    assert P.shape == Pref.shape == (1025, audio["num_mels"])
    np.testing.assert_allclose(P, Pref, atol=1e-10 * np.abs(Pref).max())


def test_other_transform_sizes_host_tables():
    """num_freq 257 / 513 / 2049 (n_fft 512 / 1024 / 4096) are served by the any-size path: plan creation succeeds
    and the mel basis / pseudo-inverse follow the oracle for that n_fft."""
    for nf, flm, fsm in ((257, 20.0, 5.0), (513, 40.0, 10.0), (2049, 50.0, 12.5)):
        audio = dict(MAIN_AUDIO, num_freq=nf, frame_length_ms=flm, frame_shift_ms=fsm)
        ap = AudioProcessor(verbose=False, **audio)
        orc = OracleAudioProcessor(**audio)
        assert ap.n_fft == (nf - 1) * 2 and ap.win_length <= ap.n_fft
        M = ap._build_mel_basis()
        assert M.shape == (80, nf)
        np.testing.assert_allclose(M, orc._build_mel_basis(), atol=1e-14)
        Pref = np.linalg.pinv(orc._build_mel_basis())
        np.testing.assert_allclose(ap._inv_mel_basis(), Pref, atol=1e-9 * np.abs(Pref).max())


def test_mel_basis_matches_reference_shim_golden(golden):
    ap = AudioProcessor(verbose=False, **MAIN_AUDIO)
    np.testing.assert_allclose(ap._build_mel_basis(), golden["main_mel_basis"], atol=2e-7)


def test_pinv_rank_deficient_basis():
    # 128 mels over 0-2 kHz at n_fft 2048: the lowest filters fall between bins -> all-zero rows, rank < num_mels
    audio = dict(MAIN_AUDIO, num_mels=128, mel_fmax=1200.0)
    M = lr_mel(22050, 2048, 128, 0.0, 1200.0)
    assert np.linalg.matrix_rank(M) < 128
    ap = AudioProcessor(verbose=False, **audio)
    P = ap._inv_mel_basis()
    Pref = np.linalg.pinv(M)
    np.testing.assert_allclose(P, Pref, atol=1e-7 * np.abs(Pref).max())


def test_batch_layout_from_frames_and_lengths():
    ap, plan = _host_plan(MAIN_AUDIO)
    lay = pkg.BatchLayout(plan, n_frames=[482, 1, 0, 5, 153])
    assert lay.total_frames == 641
    assert list(lay.frame_off) == [0, 482, 483, 483, 488, 641]
    assert list(lay.wav_len) == [275 * 481, 0, 0, 275 * 4, 275 * 152]
    assert all(o % 4 == 0 for o in lay.wav_off)
    assert all(lay.wav_off[u + 1] - lay.wav_off[u] >= lay.wav_len[u] for u in range(5))
    lay2 = pkg.BatchLayout(plan, wav_lengths=[132300, 41885, 1, 274, 275])
    assert list(lay2.n_frames) == [482, 153, 1, 1, 2]          # T = 1 + L // hop
    assert lay2.total_samples >= 132300 + 41885 + 1 + 274 + 275


def test_error_codes_and_messages():
    lib = L.load()
    ap, plan = _host_plan(MAIN_AUDIO)
    # a host-only plan has no compute path: loud failure, not a CPU fallback
    lay = pkg.BatchLayout(plan, n_frames=[4])
    rc = lib.ttsa_stft(plan.handle, lay.handle, None, None, None)
    assert rc == L.TTSA_ERR_NO_DEVICE and b"no CPU path" in lib.ttsa_last_error()
    rc = lib.ttsa_pointwise(plan.handle, 0, None, None, 4, None)
    assert rc == L.TTSA_ERR_NO_DEVICE
    # config validation
    with pytest.raises(L.TtsaError) as ei:
        AudioProcessor(verbose=False, **dict(MAIN_AUDIO, num_freq=1001))._plan(host_only=True)   # n_fft 2000: not a power of two
    assert ei.value.code == L.TTSA_ERR_UNSUPPORTED
    with pytest.raises(L.TtsaError) as ei:
        AudioProcessor(verbose=False, **dict(MAIN_AUDIO, num_freq=4097))._plan(host_only=True)   # n_fft 8192 > 4096
    assert ei.value.code == L.TTSA_ERR_UNSUPPORTED
    with pytest.raises(L.TtsaError) as ei:                                                        # win 1102 > n_fft 1024
        AudioProcessor(verbose=False, **dict(MAIN_AUDIO, num_freq=513))._plan(host_only=True)
    assert ei.value.code == L.TTSA_ERR_UNSUPPORTED
    with pytest.raises(AssertionError):     # utils/audio.py:70-71
        AudioProcessor(verbose=False, **dict(MAIN_AUDIO, mel_fmax=12000.0))._build_mel_basis()
    c = L.TtsaConfig()
    h = ctypes.c_void_p()
    assert lib.ttsa_plan_create(ctypes.byref(c), -1, ctypes.byref(h)) == L.TTSA_ERR_BAD_CONFIG
    assert lib.ttsa_plan_create(None, -1, ctypes.byref(h)) == L.TTSA_ERR_BAD_ARG
    bad = np.array([-3], dtype=np.int32)
    hb = ctypes.c_void_p()
    rc = lib.ttsa_batch_from_frames(plan.handle, bad.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), 1, ctypes.byref(hb))
    assert rc == L.TTSA_ERR_BAD_ARG


def test_audio_processor_attributes_and_prints(capsys):
    ap = AudioProcessor(**TEST_AUDIO)
    out = capsys.readouterr().out
    assert " > Setting up Audio Processor..." in out and " | > hop_length:275" in out
    assert (ap.n_fft, ap.hop_length, ap.win_length) == (2048, 275, 1102)
    assert ap.max_norm == 4.0 and isinstance(ap.max_norm, float) and ap.mel_fmin == 95 and ap.clip_norm is True
    ap2 = AudioProcessor(verbose=False, **dict(MAIN_AUDIO, mel_fmin=None, max_norm=None))
    assert ap2.mel_fmin == 0 and ap2.max_norm == 1.0
    with pytest.raises(RuntimeError):       # utils/audio.py:129-130
        AudioProcessor(verbose=False, **dict(MAIN_AUDIO, preemphasis=0.0)).apply_preemphasis(np.zeros(8))


def test_compute_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    ap = AudioProcessor(verbose=False, **MAIN_AUDIO)
    with pytest.raises(RuntimeError, match="no CPU"):
        ap.spectrogram(np.zeros(4000, dtype=np.float32))


def test_host_helpers_match_reference_formulas(golden):
    ap = AudioProcessor(verbose=False, **MAIN_AUDIO)
    orc = OracleAudioProcessor(**MAIN_AUDIO)
    x = np.linspace(-1, 1, 101)
    np.testing.assert_array_equal(ap.mulaw_encode(x, 9), orc.mulaw_encode(x, 9))
    np.testing.assert_allclose(ap.mulaw_decode(x, 9), orc.mulaw_decode(x, 9))
    assert ap.encode_16bits(np.array([0.5, -1.0, 1.0])).tolist() == [16384, -32768, 32767]
    np.testing.assert_allclose(ap.dequantize(ap.quantize(x, 9), 9), x, atol=1e-12)
    wav = golden["wav_i16"].astype(np.float64) / 32768.0
    assert ap.find_endpoint(wav) == orc.find_endpoint(wav)


def test_strided_batch_layout():
    """Padded [B, frame_stride, D] layouts (model outputs, collate tensors): rows of utterance u start at u*stride."""
    ap, plan = _host_plan(MAIN_AUDIO)
    lay = pkg.BatchLayout(plan, n_frames=[23, 9, 40], frame_stride=48)
    assert list(lay.frame_off) == [0, 48, 96, 144] and lay.total_frames == 144
    assert list(lay.n_frames) == [23, 9, 40] and list(lay.wav_len) == [275 * 22, 275 * 8, 275 * 39]
    lay2 = pkg.BatchLayout(plan, wav_lengths=[5000, 2750], frame_stride=20)
    assert list(lay2.n_frames) == [19, 11] and list(lay2.frame_off) == [0, 20, 40]
    with pytest.raises(L.TtsaError) as ei:
        pkg.BatchLayout(plan, n_frames=[23, 50], frame_stride=48)
    assert ei.value.code == L.TTSA_ERR_BAD_ARG


def test_load_wav_resamples_like_the_reference_call(tmp_path):
    """utils/audio.py:235-246: load_wav(filename) asserts the configured rate; load_wav(filename, sr=...) resamples
    (the reference through librosa.load; here a polyphase resampler -- same rate, length and spectrum)."""
    from scipy.io import wavfile
    sr_file, f0 = 16000, 440.0
    t = np.arange(sr_file) / sr_file
    wavfile.write(tmp_path / "a.wav", sr_file, (0.5 * np.sin(2 * np.pi * f0 * t) * 32767).astype(np.int16))
    ap = AudioProcessor(verbose=False, **dict(MAIN_AUDIO, do_trim_silence=False))
    with pytest.raises(AssertionError):
        ap.load_wav(str(tmp_path / "a.wav"))                       # 16000 vs the configured 22050
    x = ap.load_wav(str(tmp_path / "a.wav"), sr=22050)
    assert abs(len(x) - 22050) <= 1
    spec = np.abs(np.fft.rfft(x * np.hanning(len(x))))
    assert abs(np.argmax(spec) * 22050.0 / len(x) - f0) < 2.0      # the tone is where it was
    assert 0.45 < np.abs(x[2000:-2000]).max() < 0.55


def test_pinv_follows_numpy_svd_and_moore_penrose():
    """The pseudo-inverse comes from a one-sided Jacobi SVD of the basis itself (numpy's cut-off, rcond 1e-15), not from
    its Gram matrix (which squares the condition number): on a heavily rank-deficient basis and on a dense one it
    follows np.linalg.pinv (utils/audio.py:65) and satisfies the Moore-Penrose identities."""
    for mels, fmin, fmax in ((128, 0.0, 900.0), (128, 500.0, 3000.0), (80, 0.0, 8000.0)):
        M = lr_mel(22050, 2048, mels, fmin, fmax)
        ap = AudioProcessor(verbose=False, **dict(MAIN_AUDIO, num_mels=mels, mel_fmin=fmin, mel_fmax=fmax))
        P, Pref = ap._inv_mel_basis(), np.linalg.pinv(M)
        np.testing.assert_allclose(P, Pref, atol=1e-9 * np.abs(Pref).max())
        np.testing.assert_allclose(M @ P @ M, M, atol=1e-10 * np.abs(M).max())
        np.testing.assert_allclose(P @ M @ P, P, atol=1e-10 * np.abs(P).max())


def _run_mel_schedule(pairs, words, mag):
    """Numpy restatement of the feature kernel's mel stage on the segment schedule (feat_stream.cuh): three slots of
    (falling, rising) accumulators per lane, then <= 4 partial sums per filter through the 193-entry scratch row."""
    npairs = int(sum(pairs))
    w = words[:128 * npairs].view(np.float32).reshape(npairs, 32, 4).astype(np.float64)
    ix = words[128 * npairs:160 * npairs].reshape(npairs, 32)
    comb = words[160 * npairs:160 * npairs + 96].reshape(3, 32)
    assert np.all((ix & 0xffff) % 4 == 0) and np.all((ix >> 16) % 4 == 0)
    ba, bb = (ix & 0xffff) // 4, (ix >> 16) // 4
    assert ba.max() < mag.size and bb.max() < mag.size
    part = np.zeros(193)
    pr = 0
    conflicts = 0
    for sl in range(3):
        accd, accu = np.zeros(32), np.zeros(32)
        for _ in range(int(pairs[sl])):
            for b in (ba[pr], bb[pr]):
                conflicts += 32 - len(set((b % 32).tolist()))
            accd += w[pr, :, 0] * mag[ba[pr]] + w[pr, :, 2] * mag[bb[pr]]
            accu += w[pr, :, 1] * mag[ba[pr]] + w[pr, :, 3] * mag[bb[pr]]
            pr += 1
        part[2 * (32 * sl + np.arange(32))] = accd
        part[2 * (32 * sl + np.arange(32)) + 1] = accu
    out = np.zeros(96)
    for r in range(3):
        c = comb[r]
        out[32 * r:32 * r + 32] = part[c & 255] + part[(c >> 8) & 255] + part[(c >> 16) & 255] + part[c >> 24]
    return out, conflicts


@pytest.mark.parametrize("audio", [MAIN_AUDIO, TEST_AUDIO, dict(MAIN_AUDIO, sample_rate=16000), dict(MAIN_AUDIO, num_mels=90),
                                   dict(MAIN_AUDIO, num_mels=40, mel_fmin=95.0, mel_fmax=7600.0),
                                   dict(MAIN_AUDIO, sample_rate=24000, mel_fmax=None, mel_fmin=50.0)])
def test_mel_segment_schedule_reproduces_the_basis(audio):
    """The schedule the warp-stream feature kernel walks (ttsa_plan_mel_schedule; _linear_to_mel inside melspectrogram,
    utils/audio.py:60-62) is the float32 mel basis exactly: every tap appears once, filters are sums of their partial sums."""
    ap, plan = _host_plan(audio)
    pairs = (ctypes.c_int32 * 3)()
    n = plan.lib.ttsa_plan_mel_schedule(plan.handle, pairs, None, 0)
    assert n > 0, "a Slaney triangular bank must get the segment schedule"
    words = np.zeros(n, dtype=np.uint32)
    assert plan.lib.ttsa_plan_mel_schedule(plan.handle, pairs, words.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), n) == n
    pairs = [int(x) for x in pairs]
    assert n == 160 * sum(pairs) + 96 and pairs[0] >= pairs[1] >= pairs[2] >= 0
    basis32 = ap._build_mel_basis().astype(np.float32).astype(np.float64)
    nm = audio["num_mels"]
    # unit magnitudes recover every row sum, one-hot magnitudes every single tap
    rng = np.random.default_rng(3)
    for mag in (np.ones(1025), rng.random(1025) * 10.0 ** rng.uniform(-4, 2, 1025)):
        out, conflicts = _run_mel_schedule(pairs, words, mag)
        ref = basis32 @ mag
        np.testing.assert_allclose(out[:nm], ref, rtol=1e-12, atol=1e-300)
        assert np.all(out[nm:] == 0.0)
        assert conflicts <= 0.05 * 64 * sum(pairs)      # bank conflicts among the 32 magnitude reads of a step
    for k in rng.integers(0, 1025, 40):
        mag = np.zeros(1025); mag[k] = 1.0
        out, _ = _run_mel_schedule(pairs, words, mag)
        np.testing.assert_array_equal(out[:nm], basis32[:, k])
    # the schedule is shorter than a per-filter walk: about one step per bin of the bank and lane
    nz_bins = int(np.count_nonzero(basis32.any(axis=0)))
    assert 2 * sum(pairs) <= 1.5 * nz_bins / 32 + 8, (pairs, nz_bins)


def test_mel_segment_schedule_declines_what_it_cannot_hold():
    """More segments than the 96 (slot, lane) cells: no schedule, the feature kernel keeps the per-filter lane schedule."""
    ap, plan = _host_plan(dict(MAIN_AUDIO, num_mels=96))
    pairs = (ctypes.c_int32 * 3)(7, 7, 7)
    assert plan.lib.ttsa_plan_mel_schedule(plan.handle, pairs, None, 0) == 0
    assert list(pairs) == [0, 0, 0]
