"""CPU tests that pin the oracle (oracle/audio_oracle.py).

(1) against fixtures produced by the reference's own utils/audio.py run over a
    torch-backed librosa shim (tests/golden/make_golden.py),
(2) against independent implementations (torch.stft/istft, scipy lfilter),
(3) against the reference's own numeric contract, tests/test_audio.py:57-144.
"""
import numpy as np
import pytest
import torch
from scipy import signal

from conftest import MAIN_AUDIO, TEST_AUDIO, run_reference_test_normalize, snr_db, synth_speech_like
from oracle.audio_oracle import (OracleAudioProcessor, lr_hann_padded, lr_istft, lr_mel, lr_stft,
                                 lr_window_sumsquare, lfilter_fir2, lfilter_iir1, reflect_index)


def _wav(golden):
    return golden["wav_i16"].astype(np.float64) / 32768.0


@pytest.mark.parametrize("name", ["test", "main"])
def test_forward_matches_reference_shim(golden, golden_audio_cfgs, name):
    ap = OracleAudioProcessor(**golden_audio_cfgs[name])
    wav = _wav(golden)
    lin = ap.spectrogram(wav)
    mel = ap.melspectrogram(wav)
    assert lin.shape == (1025, 153) and mel.shape == (80, 153)
    np.testing.assert_allclose(lin[:, ::4], golden[f"{name}_lin_sub4"], atol=2e-6 * ap.max_norm, rtol=0)
    # mel shim uses torchaudio's float32 filterbank -> 1e-5 level
    np.testing.assert_allclose(mel, golden[f"{name}_mel"], atol=2e-5 * ap.max_norm, rtol=0)
    l2m = ap.out_linear_to_mel(lin.astype(np.float32))
    np.testing.assert_allclose(l2m, golden[f"{name}_lin2mel"], atol=2e-5 * ap.max_norm, rtol=0)


@pytest.mark.parametrize("name", ["test", "main"])
def test_stft_istft_match_reference_shim(golden, golden_audio_cfgs, name):
    ap = OracleAudioProcessor(**golden_audio_cfgs[name])
    D = ap._stft(golden[f"{name}_stft_in"])
    np.testing.assert_allclose(D.real, golden[f"{name}_stft_re"], atol=1e-12)
    np.testing.assert_allclose(D.imag, golden[f"{name}_stft_im"], atol=1e-12)
    y = ap._istft(D)
    np.testing.assert_allclose(y, golden[f"{name}_istft"], atol=1e-13)
    np.testing.assert_allclose(ap._build_mel_basis(), golden[f"{name}_mel_basis"], atol=2e-7)


@pytest.mark.parametrize("name", ["test", "main"])
def test_griffin_lim_matches_reference_shim(golden, golden_audio_cfgs, name):
    """Reference _griffin_lim with np.random.seed(1234) phases == oracle with the same phases injected."""
    ap = OracleAudioProcessor(**golden_audio_cfgs[name])
    lin_x = golden[f"{name}_gl_lin_in"]
    angles = 2.0 * np.pi * np.random.RandomState(1234).rand(*lin_x.shape)
    y = ap.inv_spectrogram(lin_x, init_angles=angles)
    ref = golden[f"{name}_inv_spectrogram"]
    assert y.shape == ref.shape == (ap.hop_length * (lin_x.shape[1] - 1),)
    # the reference keeps S in float32 (float32 model output through _denormalize/_db_to_amp/**power,
    # utils/audio.py:156-160); the strict-float64 oracle differs from it only by that rounding (~124 dB)
    assert snr_db(ref, y) > 110.0
    y2 = ap.inv_mel_spectrogram(golden[f"{name}_gl_mel_in"], init_angles=angles)
    # pinv of the float32-accurate shim filterbank vs float64 basis: ~1e-6 relative on S
    assert snr_db(golden[f"{name}_inv_mel_spectrogram"], y2) > 70.0


def test_mel_tables_other_geometries(golden):
    for key, (sr, fmin, fmax) in {"melbasis_16000_0_8000": (16000, 0.0, 8000.0),
                                  "melbasis_24000_0_8000": (24000, 0.0, 8000.0),
                                  "melbasis_22050_50_none": (22050, 50.0, None)}.items():
        np.testing.assert_allclose(lr_mel(sr, 2048, 80, fmin, fmax), golden[key], atol=2e-7)


@pytest.mark.parametrize("hop,win", [(275, 1102), (200, 800), (300, 1200)])
def test_stft_istft_vs_torch(hop, win):
    rng = np.random.default_rng(0)
    y = rng.standard_normal(7000)
    n_fft = 2048
    D = lr_stft(y, n_fft, hop, win)
    w = torch.hann_window(win, periodic=True, dtype=torch.float64)
    Dt = torch.stft(torch.from_numpy(y), n_fft, hop_length=hop, win_length=win, window=w, center=True,
                    pad_mode="reflect", return_complex=True).numpy()
    assert D.shape == Dt.shape == (1025, 1 + len(y) // hop)
    np.testing.assert_allclose(D, Dt, atol=1e-11)
    yi = lr_istft(D, hop, win)
    yt = torch.istft(torch.from_numpy(Dt), n_fft, hop_length=hop, win_length=win, window=w, center=True).numpy()
    np.testing.assert_allclose(yi, yt, atol=1e-12)
    # analysis/synthesis round trip over the kept region
    np.testing.assert_allclose(yi, y[:hop * (D.shape[1] - 1)], atol=1e-12)


def test_reflect_index_equals_numpy_pad():
    for L in (2, 3, 5, 40, 700, 1024, 1025, 5000):
        y = np.arange(L, dtype=np.float64)
        yp = np.pad(y, 1024, mode="reflect")
        idx = reflect_index(np.arange(-1024, L + 1024), L)
        np.testing.assert_array_equal(yp, y[idx])


def test_window_sumsquare_range_in_kept_region():
    for hop, win in [(275, 1102), (200, 800), (300, 1200)]:
        for T in (2, 3, 5, 9, 153, 482):
            wss = lr_window_sumsquare(T, hop, win, 2048)[1024:-1024]
            assert wss.shape == (hop * (T - 1),)
            assert wss.min() > 1.2 and wss.max() < 1.51
    w = lr_hann_padded(1102, 2048)
    assert w[472] == 0 and w[473] == 0 and w[474] > 0 and w[473 + 1101] > 0 and w[473 + 1102] == 0


def test_lfilter_restatements_vs_scipy():
    x = np.random.default_rng(1).standard_normal(5000)
    for p in (0.97, 0.98):
        np.testing.assert_allclose(lfilter_fir2(x, p), signal.lfilter([1, -p], [1], x), atol=1e-13)
        np.testing.assert_allclose(lfilter_iir1(x, p), signal.lfilter([1], [1, -p], x), atol=1e-10)


def test_preemphasis_zero_raises():
    ap = OracleAudioProcessor(**dict(MAIN_AUDIO, preemphasis=0.0))
    with pytest.raises(RuntimeError):
        ap.apply_preemphasis(np.zeros(4))
    with pytest.raises(RuntimeError):
        ap.apply_inv_preemphasis(np.zeros(4))


def test_stft_parameters_truncation():
    ap = OracleAudioProcessor(**MAIN_AUDIO)
    assert (ap.n_fft, ap.hop_length, ap.win_length) == (2048, 275, 1102)
    ap = OracleAudioProcessor(**dict(MAIN_AUDIO, sample_rate=16000))
    assert (ap.hop_length, ap.win_length) == (200, 800)


def test_reference_normalize_contract(golden):
    """Mirror of the reference's tests/test_audio.py:57-144 (test_normalize)."""
    ap = OracleAudioProcessor(**TEST_AUDIO)
    run_reference_test_normalize(ap, _wav(golden))


def test_gl_sc_log_and_fp32_headroom():
    """SC decreases; a float32 end-to-end run of the same algorithm stays > 90 dB from the float64 oracle."""
    ap = OracleAudioProcessor(**dict(MAIN_AUDIO, griffin_lim_iters=12))
    wav = synth_speech_like(1234, n_samples=275 * 24)
    spec = ap.spectrogram(wav).astype(np.float32)
    angles = (2 * np.pi * np.random.default_rng(7).random(spec.shape)).astype(np.float32)
    y, sc = ap.inv_spectrogram(spec, init_angles=angles, return_sc=True)
    assert y.shape == (275 * (spec.shape[1] - 1),)
    assert sc.shape == (12,) and sc[-1] < sc[0]
    assert np.all(np.isfinite(y))


def test_save_wav_int16_matches_reference_formula_and_scipy_container(tmp_path):
    """Oracle save_wav_int16 == the reference expression (utils/audio.py:57) for a float64 waveform; the product's
    RIFF writer is byte-identical to scipy.io.wavfile.write (utils/audio.py:58)."""
    import io
    from scipy.io import wavfile
    from your_voice_tts_b200 import AudioProcessor
    orc = OracleAudioProcessor(**MAIN_AUDIO)
    rng = np.random.default_rng(3)
    for scale in (1.0, 0.003, 40.0):
        wav = rng.standard_normal(5000) * scale
        ref = (wav * (32767 / max(0.01, np.max(np.abs(wav))))).astype(np.int16)
        np.testing.assert_array_equal(orc.save_wav_int16(wav), ref)
    w32 = (rng.standard_normal(5000) * 0.2).astype(np.float32)
    got = orc.save_wav_int16(w32)
    assert np.abs(got.astype(int) - orc.save_wav_int16(w32.astype(np.float64)).astype(int)).max() <= 1
    cat = orc.server_concat([[0.5, -0.25], [1.0]], gap=3)
    np.testing.assert_array_equal(cat, [0.5, -0.25, 0, 0, 0, 1.0, 0, 0, 0])
    ap = AudioProcessor(verbose=False, **MAIN_AUDIO)
    pcm = rng.integers(-32768, 32767, size=4321).astype(np.int16)
    ref = io.BytesIO()
    wavfile.write(ref, MAIN_AUDIO["sample_rate"], pcm)
    assert ap.wav_file_bytes(pcm) == ref.getvalue()


def test_tacotron2_fixture_shape_and_range(tacotron2_postnet):
    """The configs[3] fixture is what SURVEY section 8d predicts for the random-init reference model: 2 N + 22 = 482
    frames for N = 230 tokens, 80 mels, values around zero that _denormalize clips into [0, 1]."""
    mel = tacotron2_postnet
    assert mel.shape == (482, 80) and mel.dtype == np.float32 and np.isfinite(mel).all()
    orc = OracleAudioProcessor(**MAIN_AUDIO)
    d = orc._denormalize(mel.astype(np.float64))
    assert d.min() >= orc.min_level_db - 1e-9 and d.max() <= 0.0 + 1e-9
