"""GPU parity tests: the CUDA path (through the C ABI, via the drop-in AudioProcessor) against the float64 oracle
and the committed fixtures.  Tolerances are the north-star's: forward STFT/mel within 1e-4, Griffin-Lim >= 60 dB
waveform SNR with identically injected phases, spectral convergence matched per iteration (1e-3 relative)."""
import numpy as np
import pytest

from conftest import MAIN_AUDIO, TEST_AUDIO, run_reference_test_normalize, snr_db, synth_speech_like
from oracle.audio_oracle import OracleAudioProcessor, lfilter_fir2, lfilter_iir1, lr_istft, lr_stft
from your_voice_tts_b200 import _lib as L

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

FWD_TOL = 1e-4       # north_star: forward STFT / mel within 1e-4 (max-abs on normalised output, per unit max_norm)
GL_SNR_DB = 60.0     # north_star: >= 60 dB waveform SNR vs the float64 reference
SC_RTOL = 1e-3


def _ap(audio):
    from your_voice_tts_b200 import AudioProcessor
    return AudioProcessor(verbose=False, **audio)


def _wav(golden):
    return golden["wav_i16"].astype(np.float64) / 32768.0


@pytest.fixture(scope="module", autouse=True)
def _need_cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


# ------------------------------------------------------------------------------------------------ forward
def _db_tol(ap, amp_ref, frame_peak):
    """Tolerance on a normalised-dB value: the north-star's 1e-4 (per unit max_norm) where log|X| is well conditioned
    in float32, widened by the float32 conditioning of the logarithm for bins far below their frame's peak
    (an exact float32 FFT has absolute error ~1e-7 * peak per bin, i.e. relative error 1e-7 * peak / |X_k|;
    pocketfft in float32 shows 3e-4 max-abs on this very fixture)."""
    scale = ap.max_norm * (2.0 if ap.symmetric_norm else 1.0) / -ap.min_level_db if ap.signal_norm else 1.0
    d_db = 8.686 * 3e-6 * frame_peak[None, :] / np.maximum(amp_ref, 10.0 ** (ap.min_level_db / 20.0))
    return FWD_TOL * ap.max_norm + scale * d_db


@pytest.mark.parametrize("name", ["test", "main"])
def test_spectrogram_and_mel_vs_oracle_and_golden(golden, golden_audio_cfgs, name):
    audio = golden_audio_cfgs[name]
    ap, orc = _ap(audio), OracleAudioProcessor(**audio)
    wav = _wav(golden)
    lin, mel = ap.spectrogram(wav), ap.melspectrogram(wav)
    w32 = wav.astype(np.float32)
    lin_o, mel_o = orc.spectrogram(w32), orc.melspectrogram(w32)
    assert lin.shape == (1025, 153) and mel.shape == (80, 153) and lin.dtype == np.float32
    D = np.abs(orc._stft(orc.apply_preemphasis(w32)))
    peak = D.max(axis=0)
    assert np.all(np.abs(lin - lin_o) <= _db_tol(ap, D, peak))
    assert np.mean(np.abs(lin - lin_o) <= FWD_TOL * ap.max_norm) >= 0.995      # plain 1e-4 on >= 99.5 % of the bins
    Dm = orc._linear_to_mel(D)
    assert np.all(np.abs(mel - mel_o) <= _db_tol(ap, Dm, Dm.max(axis=0)))
    assert np.mean(np.abs(mel - mel_o) <= FWD_TOL * ap.max_norm) >= 0.995
    # fixtures produced by the reference's own utils/audio.py (librosa shimmed by torch)
    assert np.all(np.abs(lin[:, ::4] - golden[f"{name}_lin_sub4"]) <= _db_tol(ap, D[:, ::4], peak[::4]))
    assert np.all(np.abs(mel - golden[f"{name}_mel"]) <= _db_tol(ap, Dm, Dm.max(axis=0)) + 2e-5 * ap.max_norm)


@pytest.mark.parametrize("hop_win_sr", [(275, 1102, 22050), (200, 800, 16000), (300, 1200, 24000)])
def test_stft_istft_vs_oracle(hop_win_sr):
    hop, win, sr = hop_win_sr
    ap = _ap(dict(MAIN_AUDIO, sample_rate=sr))
    assert (ap.hop_length, ap.win_length) == (hop, win)
    y = np.random.default_rng(3).standard_normal(hop * 37 + 11).astype(np.float32)
    D = ap._stft(y)
    Do = lr_stft(y, 2048, hop, win)
    assert D.shape == Do.shape and D.dtype == np.complex64
    assert np.linalg.norm(D - Do) / np.linalg.norm(Do) <= 1e-5          # normwise (1e-4 bar, fp32 gives ~1e-7)
    assert np.abs(D - Do).max() <= 1e-4 * np.abs(Do).max()
    yi = ap._istft(Do.astype(np.complex64))
    yo = lr_istft(Do.astype(np.complex64), hop, win)
    assert yi.shape == yo.shape == (hop * (Do.shape[1] - 1),)
    assert snr_db(yo, yi) >= 100.0
    # analysis -> synthesis round trip (size-independent property)
    assert snr_db(y[:len(yi)], ap._istft(D)) >= 100.0


def test_stft_edge_lengths():
    ap = _ap(MAIN_AUDIO)
    for n in (1, 2, 3, 274, 275, 276, 1023, 1025, 2049):
        y = np.random.default_rng(n).standard_normal(n).astype(np.float32)
        D = ap._stft(y)
        Do = lr_stft(y, 2048, 275, 1102) if n >= 2 else None
        assert D.shape == (1025, 1 + n // 275)
        if Do is not None:
            assert np.abs(D - Do).max() <= 1e-4 * max(1.0, np.abs(Do).max()), n


# ------------------------------------------------------------------------------------------------ Griffin-Lim
@pytest.mark.parametrize("name", ["test", "main"])
def test_griffin_lim_vs_oracle_and_golden(golden, golden_audio_cfgs, name):
    audio = golden_audio_cfgs[name]
    ap, orc = _ap(audio), OracleAudioProcessor(**audio)
    lin_x, mel_x = golden[f"{name}_gl_lin_in"], golden[f"{name}_gl_mel_in"]
    angles = (2.0 * np.pi * np.random.RandomState(1234).rand(*lin_x.shape)).astype(np.float32)
    y, sc = ap.inv_spectrogram(lin_x, init_angles=angles, return_sc=True)
    yo, sco = orc.inv_spectrogram(lin_x, init_angles=angles, return_sc=True)
    assert y.shape == yo.shape == (275 * 39,)
    assert snr_db(yo, y) >= GL_SNR_DB, snr_db(yo, y)
    np.testing.assert_allclose(sc, sco, rtol=SC_RTOL)
    # the reference's own run (np.random.seed(1234) phases in float64; ours are the float32 rounding of the same draw)
    assert snr_db(golden[f"{name}_inv_spectrogram"], y) >= GL_SNR_DB
    ym, scm = ap.inv_mel_spectrogram(mel_x, init_angles=angles, return_sc=True)
    ymo, scmo = orc.inv_mel_spectrogram(mel_x, init_angles=angles, return_sc=True)
    assert snr_db(ymo, ym) >= GL_SNR_DB, snr_db(ymo, ym)
    np.testing.assert_allclose(scm, scmo, rtol=SC_RTOL)
    assert snr_db(golden[f"{name}_inv_mel_spectrogram"], ym) >= GL_SNR_DB


def test_griffin_lim_reference_rng_draw(golden):
    """Without injected phases the reference's np.random.rand draw is reproduced (utils/audio.py:183)."""
    ap, orc = _ap(MAIN_AUDIO), OracleAudioProcessor(**MAIN_AUDIO)
    lin_x = golden["main_gl_lin_in"]
    np.random.seed(99)
    y = ap.inv_spectrogram(lin_x)
    np.random.seed(99)
    yo = orc.inv_spectrogram(lin_x)
    assert snr_db(yo, y) >= GL_SNR_DB


def test_griffin_lim_magnitude_input_and_full_length_utterance():
    """_griffin_lim(S) on a 6 s LJSpeech-shape spectrogram (T = 482), configs[0] of BASELINE.json at 8 iterations."""
    audio = dict(MAIN_AUDIO, griffin_lim_iters=8)
    ap, orc = _ap(audio), OracleAudioProcessor(**audio)
    wav = synth_speech_like(1234)
    spec = orc.spectrogram(wav).astype(np.float32)
    assert spec.shape == (1025, 482)
    S = (orc._db_to_amp(orc._denormalize(spec.astype(np.float64)) + orc.ref_level_db) ** orc.power).astype(np.float32)
    angles = (2 * np.pi * np.random.default_rng(5).random(S.shape)).astype(np.float32)
    y, sc = ap._griffin_lim(S, init_angles=angles, return_sc=True)
    yo, sco = orc._griffin_lim(S, init_angles=angles, return_sc=True)
    assert y.shape == (132275,)
    assert snr_db(yo, y) >= GL_SNR_DB, snr_db(yo, y)
    np.testing.assert_allclose(sc, sco, rtol=SC_RTOL)


def test_ragged_batch_matches_per_utterance_oracle():
    """Variable-length packed batch (T in {2, 5, 9, 40, 61, 153}): every utterance equals its own oracle run."""
    audio = dict(MAIN_AUDIO, griffin_lim_iters=5)
    ap, orc = _ap(audio), OracleAudioProcessor(**audio)
    Ts = [40, 2, 153, 5, 9, 61, 1]
    rng = np.random.default_rng(11)
    specs = [rng.random((T, 1025)).astype(np.float32) for T in Ts]         # packed layout is frame-major
    angs = [(2 * np.pi * rng.random((T, 1025))).astype(np.float32) for T in Ts]
    lay = ap.layout(n_frames=Ts)
    dev = torch.device("cuda")
    out, sc = ap.inv_spectrogram_batch(torch.from_numpy(np.concatenate(specs)).to(dev), lay,
                                       init_angles=torch.from_numpy(np.concatenate(angs)).to(dev), return_sc=True)
    outs = [o.cpu().numpy() for o in lay.split_wav(out)]
    sc = sc.cpu().numpy()
    for u, T in enumerate(Ts):
        assert outs[u].shape == (275 * max(0, T - 1),)
        if T < 2:
            continue
        yo, sco = orc.inv_spectrogram(specs[u].T, init_angles=angs[u].T, return_sc=True)
        assert snr_db(yo, outs[u]) >= GL_SNR_DB, (T, snr_db(yo, outs[u]))
        np.testing.assert_allclose(sc[:, u], sco, rtol=SC_RTOL)


def test_large_batch_segments_deterministic_and_match_oracle():
    """64 x T=482 (BASELINE configs[1] shape): CTAs own multi-tile segments with warm-up frames.  Checked by
    determinism, by comparing two utterances against the oracle, and by batch-composition independence."""
    audio = dict(MAIN_AUDIO, griffin_lim_iters=3)
    ap, orc = _ap(audio), OracleAudioProcessor(**audio)
    B, T = 64, 482
    g = torch.Generator(device="cuda").manual_seed(0)
    spec = torch.rand((B * T, 1025), device="cuda", generator=g)
    ang = torch.rand((B * T, 1025), device="cuda", generator=g) * (2 * np.pi)
    lay = ap.layout(n_frames=[T] * B)
    y1 = ap.inv_spectrogram_batch(spec, lay, init_angles=ang).clone()
    y2 = ap.inv_spectrogram_batch(spec, lay, init_angles=ang)
    assert torch.equal(y1, y2)                                         # no atomics on the data path
    for u in (0, 37):
        yo = orc.inv_spectrogram(spec[u * T:(u + 1) * T].cpu().numpy().T, init_angles=ang[u * T:(u + 1) * T].cpu().numpy().T)
        yu = lay.split_wav(y1)[u].cpu().numpy()
        assert snr_db(yo, yu) >= GL_SNR_DB, (u, snr_db(yo, yu))
    # the same utterance inside a different batch gives the same samples (different CTA partition) to fp32 rounding
    lay1 = ap.layout(n_frames=[T])
    y_single = ap.inv_spectrogram_batch(spec[37 * T:38 * T].contiguous(), lay1, init_angles=ang[37 * T:38 * T].contiguous())
    assert snr_db(lay.split_wav(y1)[37].cpu().numpy(), y_single[:275 * (T - 1)].cpu().numpy()) >= 120.0


def test_small_batch_fine_segments(monkeypatch):
    """Batches too small to fill the GPU (the server synthesises one sentence at a time, server/synthesizer.py:147-157)
    run the iteration kernel with one CTA per fine segment (8 frames, then 4 owned + 4 recomputed frames each).  Checked
    against the oracle and against the ordinary tile partition (TTSA_GL_FINE=0), for frame counts around every boundary."""
    from your_voice_tts_b200 import audio as A
    audio = dict(MAIN_AUDIO, griffin_lim_iters=5)
    orc = OracleAudioProcessor(**audio)
    rng = np.random.default_rng(17)
    cases = [[1], [2], [7], [8], [9], [12], [13], [61], [482], [100, 3, 57]]
    data = [([rng.random((T, 1025)).astype(np.float32) for T in Ts],
             [(2 * np.pi * rng.random((T, 1025))).astype(np.float32) for T in Ts]) for Ts in cases]
    outs = {}
    for fine in ("1", "0"):
        monkeypatch.setenv("TTSA_GL_FINE", fine)
        A._PLAN_CACHE.clear()
        ap = _ap(audio)
        for Ts, (specs, angs) in zip(cases, data):
            lay = ap.layout(n_frames=Ts)
            y = ap.inv_spectrogram_batch(torch.from_numpy(np.concatenate(specs)).cuda(), lay,
                                         init_angles=torch.from_numpy(np.concatenate(angs)).cuda())
            outs[(fine, tuple(Ts))] = [w.cpu().numpy().copy() for w in lay.split_wav(y)]
    monkeypatch.delenv("TTSA_GL_FINE")
    A._PLAN_CACHE.clear()
    for Ts, (specs, angs) in zip(cases, data):
        for u, T in enumerate(Ts):
            if T < 2:
                continue
            yf, yt = outs[("1", tuple(Ts))][u], outs[("0", tuple(Ts))][u]
            yo = orc.inv_spectrogram(specs[u].T, init_angles=angs[u].T)
            assert yf.shape == yo.shape
            assert snr_db(yo, yf) >= GL_SNR_DB, (Ts, u, snr_db(yo, yf))
            assert snr_db(yt, yf) >= 110.0, (Ts, u, snr_db(yt, yf))


@pytest.mark.parametrize("sr", [22050, 16000, 24000])
def test_stream_kernel_other_geometries_and_convergence(monkeypatch, sr):
    """The warp-stream iteration kernel on every shipped geometry it is built for (275/1102 with an odd hop, 200/800 with an
    even one; 300/1200 does not fit its shared-memory ring and must fall back to the tile kernel), on a ragged batch
    forced onto a 2-CTA partition so that runs are short, cross utterance boundaries and hand zones over between CTAs;
    normalised-dB input with spectral-convergence sums.  Against the oracle and against the tile kernel."""
    from your_voice_tts_b200 import audio as A
    audio = dict(MAIN_AUDIO, sample_rate=sr, griffin_lim_iters=5)
    orc = OracleAudioProcessor(**audio)
    Ts = [150, 90, 6, 130, 1, 7, 60, 33]
    rng = np.random.default_rng(sr)
    specs = [rng.random((T, 1025)).astype(np.float32) for T in Ts]
    angs = [(2 * np.pi * rng.random((T, 1025))).astype(np.float32) for T in Ts]
    res = {}
    for kernel in ("stream", "tile"):
        monkeypatch.setenv("TTSA_WPS_GRID", "2")
        monkeypatch.setenv("TTSA_GL_KERNEL", kernel)
        A._PLAN_CACHE.clear()
        ap = _ap(audio)
        lay = ap.layout(n_frames=Ts)
        y, sc = ap.inv_spectrogram_batch(torch.from_numpy(np.concatenate(specs)).cuda(), lay,
                                         init_angles=torch.from_numpy(np.concatenate(angs)).cuda(), return_sc=True)
        res[kernel] = ([w.cpu().numpy().copy() for w in lay.split_wav(y)], sc.cpu().numpy().copy())
    monkeypatch.delenv("TTSA_WPS_GRID"); monkeypatch.delenv("TTSA_GL_KERNEL")
    A._PLAN_CACHE.clear()
    for u, T in enumerate(Ts):
        if T < 2:
            continue
        yo, sco = orc.inv_spectrogram(specs[u].T, init_angles=angs[u].T, return_sc=True)
        ys, yt = res["stream"][0][u], res["tile"][0][u]
        assert ys.shape == yo.shape
        assert snr_db(yo, ys) >= GL_SNR_DB, (sr, u, snr_db(yo, ys))
        assert snr_db(yt, ys) >= 100.0, (sr, u, snr_db(yt, ys))
        np.testing.assert_allclose(res["stream"][1][:, u], sco, rtol=SC_RTOL)


@pytest.mark.parametrize("sr", [22050, 16000])
def test_stream_kernel_fused_iterations_bit_identical(monkeypatch, sr):
    """TTSA_GL_FUSE=n runs n Griffin-Lim iterations per launch of the warp-stream kernel (neighbour-only synchronisation
    through global flags instead of a kernel boundary, ping-pong inside the kernel).  The arithmetic and its order are the
    same, so waveforms must be BIT-identical to one launch per iteration (the convergence sums equal up to summation order) -- on the full 148-CTA
    partition of a long ragged batch and on a 2-CTA partition whose runs cross utterance boundaries; 7 iterations as 7, as
    3 + 3 + 1 and as 7 x 1; repeated to catch an ordering race."""
    from your_voice_tts_b200 import audio as A
    audio = dict(MAIN_AUDIO, sample_rate=sr, griffin_lim_iters=7)
    rng = np.random.default_rng(7 + sr)
    for grid, Ts in ((None, [482] * 30 + [300, 77, 5, 1, 640]), ("2", [150, 90, 6, 130, 1, 7, 60, 33])):
        spec = torch.from_numpy(rng.random((sum(Ts), 1025)).astype(np.float32)).cuda()
        ang = torch.from_numpy((2 * np.pi * rng.random((sum(Ts), 1025))).astype(np.float32)).cuda()
        res = {}
        for fuse in ("1", "7", "3", "7"):
            monkeypatch.setenv("TTSA_GL_FUSE", fuse)
            if grid:
                monkeypatch.setenv("TTSA_WPS_GRID", grid)
            A._PLAN_CACHE.clear()
            ap = _ap(audio)
            lay = ap.layout(n_frames=Ts)
            y, sc = ap.inv_spectrogram_batch(spec, lay, init_angles=ang, return_sc=True)
            out = (y.cpu().numpy().copy(), sc.cpu().numpy().copy())
            if "1" in res:
                assert np.array_equal(out[0], res["1"][0]), (sr, grid, fuse)
                np.testing.assert_allclose(out[1], res["1"][1], rtol=1e-5)     # atomicAdd over warps: order varies
            res[fuse] = out
        assert np.isfinite(res["1"][0]).all() and float(np.abs(res["1"][0]).max()) > 0
    monkeypatch.delenv("TTSA_GL_FUSE")
    monkeypatch.delenv("TTSA_WPS_GRID", raising=False)
    A._PLAN_CACHE.clear()


def test_generic_geometry_kernel_class():
    """win 1764 / hop 275 (80 ms window): more than 5 taps per hop residue and 28 non-zero packed rows -> the generic
    (NZ = 32) kernel class; and win 2048 == n_fft."""
    for flm in (80.0, 92.88):
        audio = dict(MAIN_AUDIO, frame_length_ms=flm, griffin_lim_iters=4)
        ap, orc = _ap(audio), OracleAudioProcessor(**audio)
        assert ap.win_length in (1764, 2048)
        y = synth_speech_like(3, n_samples=275 * 21)
        spec, spec_o = ap.spectrogram(y), orc.spectrogram(y)
        assert np.mean(np.abs(spec - spec_o) <= FWD_TOL) >= 0.995
        ang = (2 * np.pi * np.random.default_rng(1).random(spec_o.shape)).astype(np.float32)
        w, sc = ap.inv_spectrogram(spec_o.astype(np.float32), init_angles=ang, return_sc=True)
        wo, sco = orc.inv_spectrogram(spec_o.astype(np.float32), init_angles=ang, return_sc=True)
        assert snr_db(wo, w) >= GL_SNR_DB, snr_db(wo, w)
        np.testing.assert_allclose(sc, sco, rtol=SC_RTOL)


def test_fixed_and_runtime_geometry_kernels_agree(monkeypatch):
    """The shipped (hop, win) pairs run kernels with the geometry as compile-time constants; TTSA_GENERIC_GEO=1
    forces the run-time-geometry kernels of the same class.  Both must meet the oracle bar, on a ragged batch with
    multi-tile segments; and a class-20 geometry that is not shipped (20 kHz: hop 250, win 1000) runs the latter."""
    from your_voice_tts_b200 import audio as A
    for sr in (22050, 16000, 24000, 20000):
        audio = dict(MAIN_AUDIO, sample_rate=sr, griffin_lim_iters=4, preemphasis=0.97)
        orc = OracleAudioProcessor(**audio)
        hop = orc.hop_length
        ys = [synth_speech_like(5 + i, n_samples=hop * n) for i, n in enumerate((37, 9, 64))]
        specs_o = [orc.spectrogram(y).astype(np.float32) for y in ys]
        angs = [(2 * np.pi * np.random.default_rng(i).random(s.shape)).astype(np.float32) for i, s in enumerate(specs_o)]
        res = {}
        for generic in ("0", "1"):
            monkeypatch.setenv("TTSA_GENERIC_GEO", generic)
            A._PLAN_CACHE.clear()
            ap = _ap(audio)
            spec0 = ap.spectrogram(ys[0])
            assert np.mean(np.abs(spec0 - orc.spectrogram(ys[0])) <= FWD_TOL) >= 0.995
            lay = ap.layout(n_frames=[s.shape[1] for s in specs_o])
            dev = torch.device("cuda")
            packed = lambda xs: torch.from_numpy(np.ascontiguousarray(np.concatenate([x.T for x in xs]))).to(dev)
            with pytest.raises(ValueError):                  # strided views are refused, not misread
                ap.inv_spectrogram_batch(packed(specs_o).t(), lay)
            out, sc = ap.inv_spectrogram_batch(packed(specs_o), lay, init_angles=packed(angs), return_sc=True)
            res[generic] = [o.cpu().numpy() for o in lay.split_wav(out)]
            for u, s in enumerate(specs_o):
                wo, sco = orc.inv_spectrogram(s, init_angles=angs[u], return_sc=True)
                assert snr_db(wo, res[generic][u]) >= GL_SNR_DB, (sr, generic, u, snr_db(wo, res[generic][u]))
                np.testing.assert_allclose(sc[:, u].cpu().numpy(), sco, rtol=SC_RTOL)
        for a, b in zip(res["0"], res["1"]):
            assert snr_db(a, b) >= 100.0
    monkeypatch.delenv("TTSA_GENERIC_GEO")
    A._PLAN_CACHE.clear()


def test_device_rng_phases():
    audio = dict(MAIN_AUDIO, griffin_lim_iters=6)
    ap = _ap(audio)
    T = 40
    spec = torch.rand((T, 1025), device="cuda")
    lay = ap.layout(n_frames=[T])
    a, sc = ap.inv_spectrogram_batch(spec, lay, seed=7, return_sc=True)
    b, _ = ap.inv_spectrogram_batch(spec, lay, seed=7, return_sc=True)
    c, _ = ap.inv_spectrogram_batch(spec, lay, seed=8, return_sc=True)
    assert torch.isfinite(a).all() and torch.equal(a, b) and not torch.equal(a, c)
    sc = sc.cpu().numpy()[:, 0]
    assert sc[-1] < sc[0]


# ------------------------------------------------------------------------------------------------ mel <-> linear
def test_mel_linear_projections(golden):
    ap, orc = _ap(MAIN_AUDIO), OracleAudioProcessor(**MAIN_AUDIO)
    rng = np.random.default_rng(2)
    S = rng.random((1025, 33)).astype(np.float32) * 3.0
    m, mo = ap._linear_to_mel(S), orc._linear_to_mel(S.astype(np.float64))
    assert m.shape == (80, 33)
    np.testing.assert_allclose(m, mo, rtol=2e-6, atol=1e-7)
    mel_amp = (rng.random((80, 33)) * 2.0).astype(np.float32)
    lin, lino = ap._mel_to_linear(mel_amp), orc._mel_to_linear(mel_amp.astype(np.float64))
    assert lin.shape == (1025, 33) and lin.min() >= 1e-10
    np.testing.assert_allclose(lin, lino, rtol=1e-4, atol=2e-5)
    spec = orc.spectrogram(_wav(golden)).astype(np.float32)
    l2m = ap.out_linear_to_mel(spec)
    assert np.abs(l2m - orc.out_linear_to_mel(spec)).max() <= FWD_TOL
    assert np.abs(l2m - golden["main_lin2mel"]).max() <= FWD_TOL


# ------------------------------------------------------------------------------------------------ filters / elementwise
@pytest.mark.parametrize("p", [0.97, 0.98])
def test_pre_and_de_emphasis(p):
    ap = _ap(dict(MAIN_AUDIO, preemphasis=p))
    for n in (1, 7, 2048, 2049, 132275):
        x = np.random.default_rng(n).standard_normal(n).astype(np.float32)
        np.testing.assert_allclose(ap.apply_preemphasis(x), lfilter_fir2(x, p), atol=2e-6)
        yo = lfilter_iir1(x, p)
        y = ap.apply_inv_preemphasis(x)
        assert y.shape == yo.shape
        assert snr_db(yo, y) >= 100.0, (n, snr_db(yo, y))


def test_reference_test_normalize_on_gpu(golden):
    """The reference's tests/test_audio.py:57-144, run against the drop-in class (attributes mutated live)."""
    run_reference_test_normalize(_ap(TEST_AUDIO), _wav(golden))


def test_pointwise_vs_oracle():
    x = (np.random.default_rng(0).random((80, 50)) * 140.0 - 120.0).astype(np.float32)
    for audio in (MAIN_AUDIO, TEST_AUDIO, dict(TEST_AUDIO, clip_norm=False), dict(MAIN_AUDIO, signal_norm=False)):
        ap, orc = _ap(audio), OracleAudioProcessor(**audio)
        n = ap._normalize(x)
        np.testing.assert_allclose(n, orc._normalize(x.astype(np.float64)), atol=2e-6 * ap.max_norm)
        np.testing.assert_allclose(ap._denormalize(n), orc._denormalize(n.astype(np.float64)), atol=2e-4)
        amp = ap._db_to_amp(x * 0.5)
        np.testing.assert_allclose(amp, orc._db_to_amp(x.astype(np.float64) * 0.5), rtol=5e-6)
        np.testing.assert_allclose(ap._amp_to_db(amp), orc._amp_to_db(amp.astype(np.float64)), atol=2e-5)


def test_reference_audio_synthesis_matrix(golden, tmp_path):
    """Mirror of tests/test_audio.py:23-55 (wav -> mel -> wav for 10 normalisation settings, 30 GL iterations)."""
    ap = _ap(TEST_AUDIO)
    wav = _wav(golden)[:275 * 60]
    for max_norm in (1.0, 4.0):
        for signal_norm, symmetric_norm, clip_norm in [(False, False, False), (True, False, False), (True, True, False),
                                                       (True, False, True), (True, True, True)]:
            ap.max_norm, ap.signal_norm, ap.symmetric_norm, ap.clip_norm = max_norm, signal_norm, symmetric_norm, clip_norm
            mel = ap.melspectrogram(wav)
            wav_ = ap.inv_mel_spectrogram(mel)
            assert wav_.shape == (275 * (mel.shape[1] - 1),) and np.isfinite(wav_).all() and np.abs(wav_).max() > 1e-3
            ap.save_wav(wav_, str(tmp_path / "out.wav"))


def test_tensor_in_tensor_out():
    ap = _ap(MAIN_AUDIO)
    y = torch.randn(275 * 20, device="cuda")
    S = ap.spectrogram(y)
    assert isinstance(S, torch.Tensor) and S.is_cuda and S.shape == (1025, 21)
    w = ap.inv_spectrogram(S)
    assert isinstance(w, torch.Tensor) and w.shape == (275 * 20,)


# ------------------------------------------------------------------------------------------------ callers either side
def _ref_prepare_tensor(inputs, out_steps):
    """utils/data.py:25-31 restated (pad [D, T] arrays to the longest + a zero frame, multiple of out_steps)."""
    max_len = max(x.shape[1] for x in inputs) + 1
    rem = max_len % out_steps
    pad_len = max_len + (out_steps - rem) if rem > 0 else max_len
    return np.stack([np.pad(x, [[0, 0], [0, pad_len - x.shape[1]]]) for x in inputs])


@pytest.mark.parametrize("r", [1, 5])
def test_collate_features_matches_reference_collate(golden, r):
    """GPU-side feature half of MyDataset.collate_fn (datasets/TTSDataset.py:191-217) vs the oracle per utterance plus
    the reference's padding rules."""
    ap, orc = _ap(TEST_AUDIO), OracleAudioProcessor(**TEST_AUDIO)
    wav = _wav(golden).astype(np.float32)
    wavs = [wav[:41885], wav[3000:3000 + 275 * 31 + 7], wav[10000:10000 + 5001]]
    linear, mel, mel_lengths, stop = ap.collate_features(wavs, outputs_per_step=r)
    mel_o = [orc.melspectrogram(w) for w in wavs]
    lin_o = [orc.spectrogram(w) for w in wavs]
    mel_ref = _ref_prepare_tensor(mel_o, r).transpose(0, 2, 1)
    lin_ref = _ref_prepare_tensor(lin_o, r).transpose(0, 2, 1)
    assert tuple(mel.shape) == mel_ref.shape and tuple(linear.shape) == lin_ref.shape and mel.shape[1] % r == 0
    assert mel_lengths.tolist() == [m.shape[1] + 1 for m in mel_o]
    m, l = mel.cpu().numpy(), linear.cpu().numpy()
    assert np.mean(np.abs(m - mel_ref) <= FWD_TOL * ap.max_norm) >= 0.995
    assert np.mean(np.abs(l - lin_ref) <= FWD_TOL * ap.max_norm) >= 0.995
    for u, w in enumerate(wavs):            # every bin within the float32 conditioning of its logarithm (see _db_tol)
        D = np.abs(orc._stft(orc.apply_preemphasis(w)))
        Dm = orc._linear_to_mel(D)
        Tu = D.shape[1]
        assert np.all(np.abs(l[u, :Tu].T - lin_o[u]) <= _db_tol(ap, D, D.max(axis=0)))
        assert np.all(np.abs(m[u, :Tu].T - mel_o[u]) <= _db_tol(ap, Dm, Dm.max(axis=0)))
    for u, mo in enumerate(mel_o):                                     # zero frame and padding are exact zeros
        assert not m[u, mo.shape[1]:].any() and not l[u, mo.shape[1]:].any()
        assert stop[u, :mo.shape[1]].sum() == 0 and bool((stop[u, mo.shape[1]:] == 1).all())


def test_inv_mel_on_padded_model_output():
    """Griffin-Lim straight from a padded [B, T_max, 80] device tensor (the layout models/tacotron2.py:62-73 returns)."""
    audio = dict(MAIN_AUDIO, griffin_lim_iters=4)
    ap, orc = _ap(audio), OracleAudioProcessor(**audio)
    Ts, t_max = [23, 9, 40], 48
    rng = np.random.default_rng(4)
    mel = rng.random((3, t_max, 80)).astype(np.float32)
    ang = (2 * np.pi * rng.random((3 * t_max, 1025))).astype(np.float32)
    wavs = ap.inv_mel_spectrogram_padded(torch.from_numpy(mel).cuda(), Ts, init_angles=torch.from_numpy(ang).cuda())
    for u, T in enumerate(Ts):
        yo = orc.inv_mel_spectrogram(mel[u, :T].T, init_angles=ang[u * t_max:u * t_max + T].T)
        assert wavs[u].shape == (275 * (T - 1),)
        assert snr_db(yo, wavs[u].cpu().numpy()) >= GL_SNR_DB


def test_host_pipeline_matches_direct_calls():
    """Double-buffered host-to-host path (pinned host mel in, pinned host waveform out) == the direct batched call."""
    from your_voice_tts_b200 import HostPipeline
    audio = dict(MAIN_AUDIO, griffin_lim_iters=3)
    ap = _ap(audio)
    Ts = [30, 12, 45]
    lay = ap.layout(n_frames=Ts)
    rng = np.random.default_rng(8)
    mels = [torch.from_numpy(rng.random((sum(Ts), 80)).astype(np.float32)).pin_memory() for _ in range(4)]
    outs = [torch.empty((lay.total_samples,), dtype=torch.float32).pin_memory() for _ in range(4)]
    pipe = HostPipeline(ap, lay)
    for i in range(4):
        pipe.submit(mels[i], outs[i], seed=10 + i)
    pipe.drain()
    for i in range(4):
        ref = ap.inv_mel_spectrogram_batch(mels[i].cuda(), lay, seed=10 + i).cpu()
        for u in range(len(Ts)):
            a, b = lay.split_wav(outs[i])[u], lay.split_wav(ref)[u]
            assert torch.equal(a, b)
    # the same pipeline replaying one CUDA graph per buffer set: the seed of each set is the one of its first batch
    outs_g = [torch.empty((lay.total_samples,), dtype=torch.float32).pin_memory() for _ in range(4)]
    pipe_g = HostPipeline(ap, lay, graph=True)
    for i in range(4):
        pipe_g.submit(mels[i], outs_g[i], seed=10 + i)
    pipe_g.drain()
    for i in range(4):
        ref = ap.inv_mel_spectrogram_batch(mels[i].cuda(), lay, seed=10 + (i & 1)).cpu()
        for u in range(len(Ts)):
            assert torch.equal(lay.split_wav(outs_g[i])[u], lay.split_wav(ref)[u])


def test_bench_step_matches_oracle_cfg2():
    """BASELINE configs[1], exactly bench.py's step at its own size: 64 x 482 normalised mel -> ttsa_mel_to_linear (tcgen05
    GEMM, |S|**power) -> ttsa_griffin_lim 60 iterations + GL_DEEMPHASIS, with injected phases; the first, a middle and
    the last utterance against oracle.inv_mel_spectrogram (utils/audio.py:164-172): >= 60 dB, SC matched per iteration."""
    import bench
    audio = {k: v for k, v in bench.AUDIO.items() if k != "do_trim_silence"}
    assert audio["griffin_lim_iters"] == 60
    ap, orc = _ap(audio), OracleAudioProcessor(**audio)
    B, T = 64, bench.T_FRAMES
    dev = torch.device("cuda")
    mel = bench.make_inputs(ap, B, 1234, dev)                              # the bench's own synthetic input
    lay = ap.layout(n_frames=[T] * B)
    g = torch.Generator(device="cuda").manual_seed(7)
    ang = torch.rand((B * T, 1025), device="cuda", generator=g) * (2 * np.pi)
    y, sc = ap.inv_mel_spectrogram_batch(mel, lay, init_angles=ang, return_sc=True)   # the calls of bench.step_device
    wavs, sc = lay.split_wav(y), sc.cpu().numpy()
    assert sc.shape == (60, B)
    for u in (0, 31, 63):
        mel_u = mel[u * T:(u + 1) * T].cpu().numpy().T.copy()
        ang_u = ang[u * T:(u + 1) * T].cpu().numpy().T.copy()
        yo, sco = orc.inv_mel_spectrogram(mel_u, init_angles=ang_u, return_sc=True)
        yu = wavs[u].cpu().numpy()
        assert yu.shape == yo.shape == (bench.L_OUT,)
        assert snr_db(yo, yu) >= GL_SNR_DB, (u, snr_db(yo, yu))
        np.testing.assert_allclose(sc[:, u], sco, rtol=SC_RTOL)


def test_host_pipeline_matches_oracle():
    """HostPipeline (pinned host mel in, pinned host waveform out, both the direct and the CUDA-graph form) against the
    ORACLE with injected phases, not against the repo's own direct call."""
    from your_voice_tts_b200 import HostPipeline
    audio = dict(MAIN_AUDIO, griffin_lim_iters=6)
    ap, orc = _ap(audio), OracleAudioProcessor(**audio)
    Ts = [33, 7, 52]
    lay = ap.layout(n_frames=Ts)
    rng = np.random.default_rng(21)
    mels = [rng.random((sum(Ts), 80)).astype(np.float32) for _ in range(3)]
    angs = [(2 * np.pi * rng.random((sum(Ts), 1025))).astype(np.float32) for _ in range(3)]
    for graph in (False, True):
        pipe = HostPipeline(ap, lay, graph=graph)
        n = 2 if graph else 3                       # a graph replays the phases captured per buffer set: one batch per set
        outs = [torch.empty((lay.total_samples,), dtype=torch.float32).pin_memory() for _ in range(n)]
        angs_dev = [torch.from_numpy(a).cuda() for a in angs[:n]]
        for i in range(n):
            pipe.submit(torch.from_numpy(mels[i]).pin_memory(), outs[i], init_angles=angs_dev[i])
        pipe.drain()
        off = np.concatenate(([0], np.cumsum(Ts)))
        for i in range(n):
            for u, T in enumerate(Ts):
                yo = orc.inv_mel_spectrogram(mels[i][off[u]:off[u + 1]].T, init_angles=angs[i][off[u]:off[u + 1]].T)
                yu = lay.split_wav(outs[i])[u].numpy()
                assert snr_db(yo, yu) >= GL_SNR_DB, (graph, i, u, snr_db(yo, yu))
    with pytest.raises(ValueError):                 # strict_seed: a replayed graph must not silently ignore a new seed
        pipe.submit(torch.from_numpy(mels[0]).pin_memory(), outs[0], seed=99, strict_seed=True)
    pipe.drain()


def test_async_audio_logger_side_stream():
    """Training-time audio samples (train.py:225-232, 379-384) on a side stream: the caller's stream is never blocked, the
    waveform arrives in pinned host memory and equals the oracle's inversion of the same model output."""
    from your_voice_tts_b200 import AsyncAudioLogger
    audio = dict(MAIN_AUDIO, griffin_lim_iters=6)
    ap, orc = _ap(audio), OracleAudioProcessor(**audio)
    rng = np.random.default_rng(31)
    T_pad, T = 70, 57
    mel_out = torch.from_numpy(rng.random((T_pad, 80)).astype(np.float32)).cuda()      # padded batch element, as the model leaves it
    lin_out = torch.from_numpy(rng.random((T, 1025)).astype(np.float32)).cuda()
    ang = (2 * np.pi * rng.random((T, 1025))).astype(np.float32)
    ang_dev = torch.from_numpy(ang).cuda()
    logger = AsyncAudioLogger(ap)
    logger.submit("TrainAudio", mel_out, step=10, kind="mel", n_frames=T, init_angles=ang_dev)
    mel_ref = mel_out[:T].cpu().numpy().copy()
    mel_out.zero_()                                             # the next training step overwrites the tensor: a snapshot was taken
    logger.submit("ValAudio", lin_out, step=11, kind="linear", init_angles=ang_dev)
    got = logger.flush()
    assert [(t, s) for t, s, _ in got] == [("TrainAudio", 10), ("ValAudio", 11)] and logger.poll() == []
    yo_m = orc.inv_mel_spectrogram(mel_ref.T, init_angles=ang.T)
    yo_l = orc.inv_spectrogram(lin_out.cpu().numpy().T, init_angles=ang.T)
    assert got[0][2].shape == yo_m.shape and snr_db(yo_m, got[0][2]) >= GL_SNR_DB
    assert got[1][2].shape == yo_l.shape and snr_db(yo_l, got[1][2]) >= GL_SNR_DB


def test_feature_extraction_training_batch_cfg3():
    """BASELINE configs[2]: spectrogram + melspectrogram of a 32-utterance batch of 6 s waves in one pass."""
    ap, orc = _ap(MAIN_AUDIO), OracleAudioProcessor(**MAIN_AUDIO)
    B, n = 32, 132300
    waves = [synth_speech_like(2000 + i, n_samples=n) for i in (0, 17)]
    g = torch.Generator(device="cuda").manual_seed(3)
    packed = torch.randn((B, n), device="cuda", generator=g) * 0.1
    packed[0] = torch.from_numpy(waves[0]).cuda()
    packed[17] = torch.from_numpy(waves[1]).cuda()
    lay = ap.layout(wav_lengths=[n] * B)
    assert lay.total_samples == B * n and lay.total_frames == B * 482
    lin, mel = ap.features_batch(packed.reshape(-1), lay)
    lin2, mel2 = ap.features_batch(packed.reshape(-1), lay)
    assert torch.equal(lin, lin2) and torch.equal(mel, mel2)
    for u, w in zip((0, 17), waves):
        lo, mo = orc.spectrogram(w).T, orc.melspectrogram(w).T
        l, m = lay.split_frames(lin)[u].cpu().numpy(), lay.split_frames(mel)[u].cpu().numpy()
        assert l.shape == (482, 1025) and m.shape == (482, 80)
        D = np.abs(orc._stft(orc.apply_preemphasis(w.astype(np.float32)))).T
        Dm = orc._linear_to_mel(D.T).T
        assert np.mean(np.abs(l - lo) <= FWD_TOL) >= 0.995 and np.all(np.abs(l - lo) <= _db_tol(ap, D.T, D.max(axis=1)).T)
        assert np.mean(np.abs(m - mo) <= FWD_TOL) >= 0.995 and np.all(np.abs(m - mo) <= _db_tol(ap, Dm.T, Dm.max(axis=1)).T)


@pytest.mark.parametrize("sr,pre", [(22050, 0.98), (22050, 0.0), (16000, 0.97)])
def test_feature_stream_kernel_ragged_vs_oracle_and_tile(monkeypatch, sr, pre):
    """The warp-stream feature kernel (feat_stream.cuh) forced onto a small ragged batch (TTSA_FEAT_MINFRAMES=0) with a
    2-CTA partition, so that runs cross utterance boundaries: utterances shorter than two spans (index-mapped path with
    several folds of the reflection), first / last frames of long ones (bulk copy + mirrored fill), interior frames; with
    and without pre-emphasis; odd (275) and even (200) hop; linear only, mel only and both.  Against the oracle per
    utterance and against the tile kernel."""
    from your_voice_tts_b200 import audio as A
    audio = dict(MAIN_AUDIO, sample_rate=sr, preemphasis=pre)
    orc = OracleAudioProcessor(**audio)
    lens = [9000, 1, 551, 2300, 40000, 275, 1103, 6000, 2209]
    wavs = [synth_speech_like(300 + i, n_samples=n) for i, n in enumerate(lens)]
    res = {}
    for kernel in ("stream", "lane", "tile"):     # "lane": the stream kernel with the per-filter lane schedule of the mel basis
        monkeypatch.setenv("TTSA_FEAT_KERNEL", "tile" if kernel == "tile" else "stream")
        monkeypatch.setenv("TTSA_FEAT_MEL", "lane" if kernel == "lane" else "seg")
        monkeypatch.setenv("TTSA_FEAT_MINFRAMES", "0")
        monkeypatch.setenv("TTSA_WPS_GRID", "2")
        A._PLAN_CACHE.clear()
        ap = _ap(audio)
        buf, lay = _packed_wavs(ap, wavs)
        lin, mel = ap.features_batch(buf, lay)
        lin_only, _ = ap.features_batch(buf, lay, want_mel=False)
        _, mel_only = ap.features_batch(buf, lay, want_linear=False)
        assert torch.equal(lin, lin_only) and torch.equal(mel, mel_only)
        res[kernel] = ([x.cpu().numpy().copy() for x in lay.split_frames(lin)], [x.cpu().numpy().copy() for x in lay.split_frames(mel)])
    for k in ("TTSA_FEAT_KERNEL", "TTSA_FEAT_MEL", "TTSA_FEAT_MINFRAMES", "TTSA_WPS_GRID"):
        monkeypatch.delenv(k)
    A._PLAN_CACHE.clear()
    ap = _ap(audio)
    for u, w in enumerate(wavs):
        lo, mo = orc.spectrogram(w).T, orc.melspectrogram(w).T
        wp = orc.apply_preemphasis(w.astype(np.float32)) if pre != 0 else w.astype(np.float32)
        D = np.abs(orc._stft(wp)).T
        Dm = orc._linear_to_mel(D.T).T
        l, m = res["stream"][0][u], res["stream"][1][u]
        assert l.shape == lo.shape and m.shape == mo.shape, (u, l.shape, lo.shape)
        assert np.mean(np.abs(l - lo) <= FWD_TOL) >= 0.99 and np.all(np.abs(l - lo) <= _db_tol(ap, D.T, D.max(axis=1)).T), (sr, pre, u)
        assert np.mean(np.abs(m - mo) <= FWD_TOL) >= 0.99 and np.all(np.abs(m - mo) <= _db_tol(ap, Dm.T, Dm.max(axis=1)).T), (sr, pre, u)
        lt, mt = res["tile"][0][u], res["tile"][1][u]
        assert np.mean(np.abs(l - lt) <= FWD_TOL) >= 0.99 and np.mean(np.abs(m - mt) <= FWD_TOL) >= 0.99, (sr, pre, u)
        # the two mel schedules of the stream kernel: the same taps in another order; the linear output is the same code
        ll, ml = res["lane"][0][u], res["lane"][1][u]
        assert np.array_equal(l, ll), (sr, pre, u)
        assert np.all(np.abs(ml - mo) <= _db_tol(ap, Dm.T, Dm.max(axis=1)).T) and np.mean(np.abs(m - ml) <= FWD_TOL) >= 0.99, (sr, pre, u)


# ------------------------------------------------------------------------------------------------ post-processing
def _packed_wavs(ap, wavs):
    lay = ap.layout(wav_lengths=[len(w) for w in wavs])
    buf = torch.zeros((max(1, lay.total_samples),), dtype=torch.float32, device="cuda")
    for u, w in enumerate(wavs):
        buf[int(lay.wav_off[u]):int(lay.wav_off[u]) + len(w)] = torch.from_numpy(np.asarray(w, dtype=np.float32)).cuda()
    return buf, lay


def test_pcm16_conversion_bit_exact_vs_oracle():
    """save_wav's int16 samples (utils/audio.py:56-58): bit-exact against the oracle on the same float32 waveform,
    in the float64 arithmetic of the reference's default path and in its float32 variant; quiet utterances hit the
    max(0.01, peak) floor; the server layout (10 000-sample gaps, ONE peak) equals save_wav of the concatenation."""
    ap, orc = _ap(MAIN_AUDIO), OracleAudioProcessor(**MAIN_AUDIO)
    rng = np.random.default_rng(5)
    wavs = [synth_speech_like(40 + i, n_samples=n) * a for i, (n, a) in enumerate([(33000, 1.0), (7001, 0.004), (1, 1.0), (52345, 3.7)])]
    wavs.append((rng.standard_normal(20000) * 0.3).astype(np.float32))
    buf, lay = _packed_wavs(ap, wavs)
    for f32 in (False, True):
        pcm, off = ap.pcm16_batch(buf, lay, float32_arith=f32)
        off = off.cpu().numpy()
        assert off[-1] == sum(len(w) for w in wavs)
        for u, w in enumerate(wavs):
            want = orc.save_wav_int16(w if f32 else w.astype(np.float64))
            got = pcm[off[u]:off[u + 1]].cpu().numpy()
            assert got.dtype == np.int16 and np.array_equal(got, want), (f32, u, np.abs(got.astype(int) - want).max())
    pcm, off = ap.pcm16_batch(buf, lay, joint_peak=True, gap_samples=10000)
    want = orc.save_wav_int16(orc.server_concat([w.astype(np.float64) for w in wavs]))
    n = int(off[-1].item())
    assert n == len(want) and np.array_equal(pcm[:n].cpu().numpy(), want)
    peaks = ap.wav_peaks_batch(buf, lay).cpu().numpy()
    np.testing.assert_array_equal(peaks, [np.max(np.abs(w)) for w in wavs])


def test_find_endpoint_vs_oracle():
    """find_endpoint (utils/audio.py:203-210) incl. its quirks: the SIGNED maximum is compared, candidates start at
    hop, utterances shorter than window + hop return their length."""
    ap, orc = _ap(MAIN_AUDIO), OracleAudioProcessor(**MAIN_AUDIO)
    rng = np.random.default_rng(9)
    sr = 22050
    speech = lambda n, s: synth_speech_like(s, n_samples=n)
    hush = lambda n: (rng.standard_normal(n) * 0.002).astype(np.float32)
    wavs = [np.concatenate([speech(int(1.3 * sr), 1), hush(int(1.7 * sr)), speech(sr, 2)]),      # silence in the middle
            speech(3 * sr, 3),                                                                      # none
            np.concatenate([speech(sr, 4), hush(3 * sr)]),                                         # trailing silence
            hush(2 * sr),                                                                           # silent from the start
            speech(9000, 5),                                                                        # shorter than a window
            -np.abs(speech(2 * sr, 6)) - 0.1,                                                       # negative-only: "silent" for the reference
            np.concatenate([speech(int(0.9 * sr), 7), hush(int(0.82 * sr)), speech(sr, 8)])]       # barely long enough
    buf, lay = _packed_wavs(ap, wavs)
    got = ap.find_endpoint_batch(buf, lay).cpu().numpy()
    want = [orc.find_endpoint(w) for w in wavs]
    assert got.tolist() == want, (got.tolist(), want)
    assert len(set(want)) >= 5
    got2 = ap.find_endpoint_batch(buf, lay, threshold_db=-30, min_silence_sec=0.4).cpu().numpy()
    assert got2.tolist() == [orc.find_endpoint(w, -30, 0.4) for w in wavs]
    assert ap.find_endpoint(torch.from_numpy(wavs[0]).cuda()) == want[0]
    # trimmed conversion: the endpoints as length override
    ends = ap.find_endpoint_batch(buf, lay)
    pcm, off = ap.pcm16_batch(buf, lay, lens=ends)
    off = off.cpu().numpy()
    for u, w in enumerate(wavs):
        assert np.array_equal(pcm[off[u]:off[u + 1]].cpu().numpy(), orc.save_wav_int16(w[:want[u]].astype(np.float64)))


def test_any_size_path_is_deterministic():
    """The any-size path (num_freq != 1025) overlap-adds in ceil(win/hop) launch phases of non-overlapping frames instead of
    atomicAdd: repeated runs are bit-identical (and still match the oracle: test_other_transform_sizes)."""
    audio = dict(MAIN_AUDIO, num_freq=513, frame_length_ms=40.0, frame_shift_ms=10.0, griffin_lim_iters=4)
    ap, orc = _ap(audio), OracleAudioProcessor(**audio)
    Ts = [70, 9, 33]
    rng = np.random.default_rng(3)
    spec = torch.from_numpy(rng.random((sum(Ts), 513)).astype(np.float32)).cuda()
    ang = torch.from_numpy((2 * np.pi * rng.random((sum(Ts), 513))).astype(np.float32)).cuda()
    lay = ap.layout(n_frames=Ts)
    ys = [ap.inv_spectrogram_batch(spec, lay, init_angles=ang).clone() for _ in range(6)]
    assert all(torch.equal(ys[0], y) for y in ys[1:])
    yo = orc.inv_spectrogram(spec[:70].cpu().numpy().T, init_angles=ang[:70].cpu().numpy().T)
    assert snr_db(yo, lay.split_wav(ys[0])[0].cpu().numpy()) >= GL_SNR_DB


def test_save_wav_from_device_and_server_sentences(golden, tmp_path):
    """save_wav of a CUDA waveform writes the bytes scipy writes for the oracle's int16 samples; the server's
    sentence loop (server/synthesizer.py:133-162) as one batch."""
    import io
    from scipy.io import wavfile
    audio = dict(MAIN_AUDIO, griffin_lim_iters=8)
    ap, orc = _ap(audio), OracleAudioProcessor(**audio)
    w = synth_speech_like(77, n_samples=30000)
    path = str(tmp_path / "dev.wav")
    ap.save_wav(torch.from_numpy(w).cuda(), path)
    ref = io.BytesIO()
    wavfile.write(ref, audio["sample_rate"], orc.save_wav_int16(w.astype(np.float64)))
    assert open(path, "rb").read() == ref.getvalue()
    # three "sentences": normalised mel spectrograms as a Tacotron2 postnet would leave them on the device
    mels = [orc.melspectrogram(synth_speech_like(80 + i, n_samples=275 * n)).astype(np.float32) for i, n in enumerate((40, 75, 23))]
    angs = [(2 * np.pi * np.random.default_rng(i).random((1025, m.shape[1]))).astype(np.float32) for i, m in enumerate(mels)]
    ang_packed = torch.from_numpy(np.ascontiguousarray(np.concatenate([a.T for a in angs]))).cuda()
    data = ap.sentences_to_wav_bytes([torch.from_numpy(np.ascontiguousarray(m.T)).cuda() for m in mels], init_angles=ang_packed)
    sr, pcm = wavfile.read(io.BytesIO(data))
    assert sr == audio["sample_rate"] and pcm.dtype == np.int16
    wavs_o = [orc.inv_mel_spectrogram(m, init_angles=a) for m, a in zip(mels, angs)]
    want = orc.save_wav_int16(orc.server_concat(wavs_o))
    assert len(pcm) == len(want) == sum(len(x) + 10000 for x in wavs_o)
    assert np.abs(pcm.astype(int) - want.astype(int)).max() <= 2             # float32 Griffin-Lim vs float64: +-1 LSB
    assert snr_db(want.astype(np.float64), pcm.astype(np.float64)) >= 55.0


# ------------------------------------------------------------------------------------------------ other transform sizes
@pytest.mark.parametrize("nf_flm_fsm", [(257, 20.0, 5.0), (513, 40.0, 10.0), (513, 46.4, 11.6), (2049, 50.0, 12.5), (2049, 185.0, 30.0)])
def test_other_num_freq_any_size_path(nf_flm_fsm):
    """num_freq != 1025 (n_fft 512 / 1024 / 4096) runs the any-size kernels: same bars as the main path -- forward
    1e-4, stft/istft vs the oracle, Griffin-Lim >= 60 dB with injected phases, spectral convergence per iteration."""
    nf, flm, fsm = nf_flm_fsm
    audio = dict(MAIN_AUDIO, num_freq=nf, frame_length_ms=flm, frame_shift_ms=fsm, griffin_lim_iters=6)
    ap, orc = _ap(audio), OracleAudioProcessor(**audio)
    hop = orc.hop_length
    y = synth_speech_like(21, n_samples=hop * 57 + 13)
    spec, spec_o = ap.spectrogram(y), orc.spectrogram(y)
    mel, mel_o = ap.melspectrogram(y), orc.melspectrogram(y)
    assert spec.shape == spec_o.shape == (nf, 58) and mel.shape == mel_o.shape
    assert np.mean(np.abs(spec - spec_o) <= FWD_TOL) >= 0.995
    assert np.mean(np.abs(mel - mel_o) <= FWD_TOL) >= 0.995
    D, Do = ap._stft(y), lr_stft(y.astype(np.float64), orc.n_fft, orc.hop_length, orc.win_length)
    assert np.abs(D - Do).max() <= 2e-6 * np.abs(Do).max()
    yi, yio = ap._istft(Do.astype(np.complex64)), lr_istft(Do.astype(np.complex64).astype(np.complex128), orc.hop_length, orc.win_length)
    assert snr_db(yio, yi) >= 100.0
    ang = (2 * np.pi * np.random.default_rng(2).random(spec_o.shape)).astype(np.float32)
    w, sc = ap.inv_spectrogram(spec_o.astype(np.float32), init_angles=ang, return_sc=True)
    wo, sco = orc.inv_spectrogram(spec_o.astype(np.float32), init_angles=ang, return_sc=True)
    assert w.shape == wo.shape and snr_db(wo, w) >= GL_SNR_DB, snr_db(wo, w)
    np.testing.assert_allclose(sc, sco, rtol=SC_RTOL)
    wm = ap.inv_mel_spectrogram(mel_o.astype(np.float32), init_angles=ang)
    wmo = orc.inv_mel_spectrogram(mel_o.astype(np.float32), init_angles=ang)
    assert snr_db(wmo, wm) >= GL_SNR_DB, snr_db(wmo, wm)
    lm, lmo = ap.out_linear_to_mel(spec_o.astype(np.float32)), orc.out_linear_to_mel(spec_o.astype(np.float32))
    assert np.mean(np.abs(lm - lmo) <= FWD_TOL) >= 0.995
    # ragged batch, device RNG, de-emphasis
    Ts = [17, 1, 40, 3]
    lay = ap.layout(n_frames=Ts)
    specs = torch.rand((sum(Ts), nf), device="cuda")
    out = ap.inv_spectrogram_batch(specs, lay, seed=3)
    assert all(torch.isfinite(o).all() for o in lay.split_wav(out))
    assert [int(o.numel()) for o in lay.split_wav(out)] == [hop * max(0, t - 1) for t in Ts]


# ------------------------------------------------------------------------------------------------ full-size properties
def test_full_size_properties_64x6s():
    """BASELINE configs[1]/[2] sizes (64 utterances x 132 300 samples, 482 frames), checked through size-independent
    properties: istft(stft(y)) == y on the samples librosa returns; the STFT is linear; a consistent spectrogram with
    its own phases is a fixed point of the Griffin-Lim iteration (the projection leaves it unchanged), whatever the
    number of iterations; and the batched features equal the per-utterance call."""
    ap = _ap(dict(MAIN_AUDIO, griffin_lim_iters=3))
    B, Lw = 64, 132300
    lay = ap.layout(wav_lengths=[Lw] * B)
    g = torch.Generator(device="cuda").manual_seed(7)
    wav = torch.zeros((lay.total_samples,), device="cuda")
    t = torch.arange(Lw, device="cuda", dtype=torch.float32) / 22050.0
    for u in range(B):
        f0 = 90.0 + 3.0 * u
        wav[int(lay.wav_off[u]):int(lay.wav_off[u]) + Lw] = 0.4 * torch.sin(2 * np.pi * f0 * t) * (0.6 + 0.4 * torch.sin(2 * np.pi * 1.7 * t)) \
            + 0.01 * torch.randn((Lw,), device="cuda", generator=g)
    D = ap.stft_batch(wav, lay)                                   # [sum_T, 1025, 2]
    assert D.shape == (B * 482, 1025, 2)
    y = ap.istft_batch(D, lay)
    for u in (0, 31, 63):
        a = wav[int(lay.wav_off[u]):int(lay.wav_off[u]) + 275 * 481]
        b = y[int(lay.wav_off[u]):int(lay.wav_off[u]) + 275 * 481]
        assert snr_db(a.cpu().numpy(), b.cpu().numpy()) >= 100.0
    # linearity
    wav2 = torch.roll(wav, 12345) * 0.5
    D2 = ap.stft_batch(wav2, lay)
    D12 = ap.stft_batch(wav + wav2, lay)
    assert float((D12 - (D + D2)).abs().max()) <= 2e-5 * float(D12.abs().max())
    # fixed point: take y' = istft(D) (hop*(T-1) samples, so that stft and istft are exact inverses on it); |stft(y')|
    # with the phases of stft(y') reproduces y' after any number of iterations, with spectral convergence ~ 0
    Lf = 275 * 481
    lay_f = ap.layout(n_frames=[482] * B)
    lay_w = ap.layout(wav_lengths=[Lf] * B)
    yp = torch.zeros((lay_w.total_samples,), device="cuda")
    for u in range(B):
        yp[int(lay_w.wav_off[u]):int(lay_w.wav_off[u]) + Lf] = y[int(lay.wav_off[u]):int(lay.wav_off[u]) + Lf]
    Dp = ap.stft_batch(yp, lay_w)
    mag = torch.sqrt(Dp[..., 0] ** 2 + Dp[..., 1] ** 2).contiguous()
    ang = torch.atan2(Dp[..., 1], Dp[..., 0]).contiguous()
    yg, sc = ap.griffin_lim_batch(mag, lay_f, L.SPEC_MAGNITUDE, init_angles=ang, return_sc=True)
    assert float(sc.max()) <= 1e-4, float(sc.max())
    for u in (0, 40, 63):
        a = yp[int(lay_w.wav_off[u]):int(lay_w.wav_off[u]) + Lf].cpu().numpy()
        b = lay_f.split_wav(yg)[u].cpu().numpy()
        assert snr_db(a, b) >= 80.0, snr_db(a, b)
    # batched features == single-utterance features
    # (the batch runs the warp-stream feature kernel, the single call the tile kernel: both within the forward bar of the
    # oracle, and of each other within twice the float32 conditioning of log|X|)
    lin, mel = ap.features_batch(wav, lay)
    one = wav[int(lay.wav_off[5]):int(lay.wav_off[5]) + Lw].cpu().numpy()
    orc = OracleAudioProcessor(**dict(MAIN_AUDIO, griffin_lim_iters=3))
    Do = np.abs(orc._stft(orc.apply_preemphasis(one.astype(np.float32))))          # [1025, 482]
    Dmo = orc._linear_to_mel(Do)
    for got_b, got_1, ref, amp in ((lin[5 * 482:6 * 482].cpu().numpy().T, ap.spectrogram(one), orc.spectrogram(one), Do),
                                   (mel[5 * 482:6 * 482].cpu().numpy().T, ap.melspectrogram(one), orc.melspectrogram(one), Dmo)):
        tol = _db_tol(ap, amp, amp.max(axis=0))
        assert np.mean(np.abs(got_b - ref) <= FWD_TOL) >= 0.995 and np.all(np.abs(got_b - ref) <= tol)
        assert np.mean(np.abs(got_1 - ref) <= FWD_TOL) >= 0.995 and np.all(np.abs(got_1 - ref) <= tol)
        assert np.all(np.abs(got_b - got_1) <= 2 * tol)


def test_fast_griffin_lim_momentum_opt_in():
    """Fast Griffin-Lim (momentum; not in the reference, opt-in): the waveform-domain momentum of the kernels equals
    the oracle's STFT-domain restatement of librosa's griffinlim(momentum=...), momentum 0 is the reference algorithm
    bit for bit, and with momentum the spectral error after the same number of iterations is lower."""
    audio = dict(MAIN_AUDIO, griffin_lim_iters=12)
    ap, orc = _ap(audio), OracleAudioProcessor(**audio)
    y = synth_speech_like(31, n_samples=275 * 60)
    S = np.abs(lr_stft(y.astype(np.float64), 2048, 275, 1102)).astype(np.float32)
    ang = (2 * np.pi * np.random.default_rng(4).random(S.shape)).astype(np.float32)
    dev = torch.device("cuda")
    St = torch.from_numpy(np.ascontiguousarray(S.T)).to(dev)
    At = torch.from_numpy(np.ascontiguousarray(ang.T)).to(dev)
    lay = ap.layout(n_frames=[S.shape[1]])
    n = 275 * (S.shape[1] - 1)
    w0 = ap.griffin_lim_batch(St, lay, L.SPEC_MAGNITUDE, init_angles=At)[:n].clone()
    w00 = ap.griffin_lim_batch(St, lay, L.SPEC_MAGNITUDE, init_angles=At, momentum=0.0)[:n].clone()
    assert torch.equal(w0, w00)
    for mom in (0.5, 0.99):
        w, sc = ap.griffin_lim_batch(St, lay, L.SPEC_MAGNITUDE, init_angles=At, momentum=mom, return_sc=True)
        wo, sco = orc._griffin_lim_fast(S, mom, init_angles=ang, return_sc=True)
        assert snr_db(wo, w[:n].cpu().numpy()) >= GL_SNR_DB, (mom, snr_db(wo, w[:n].cpu().numpy()))
        np.testing.assert_allclose(sc[:, 0].cpu().numpy(), sco, rtol=SC_RTOL)
    # faster convergence: consistency error of the result after 12 iterations
    err = lambda wav: spectral_err(np.abs(lr_stft(np.asarray(wav, dtype=np.float64), 2048, 275, 1102)), S)
    def spectral_err(a, b):
        return float(np.linalg.norm(a - b) / np.linalg.norm(b))
    e_plain = err(w0.cpu().numpy())
    e_fast = err(ap.griffin_lim_batch(St, lay, L.SPEC_MAGNITUDE, init_angles=At, momentum=0.99)[:n].cpu().numpy())
    assert e_fast < e_plain, (e_fast, e_plain)
    # the attribute switches the drop-in methods, the any-size path follows the same definition
    ap_f = _ap(dict(audio, griffin_lim_momentum=0.99))
    spec_n = orc.spectrogram(y).astype(np.float32)
    wf = ap_f.inv_spectrogram(spec_n, init_angles=ang)
    Sn = orc._db_to_amp(orc._denormalize(spec_n.astype(np.float64)) + orc.ref_level_db) ** orc.power
    wfo = orc.apply_inv_preemphasis(orc._griffin_lim_fast(Sn, 0.99, init_angles=ang))
    assert snr_db(wfo, wf) >= GL_SNR_DB
    audio_g = dict(MAIN_AUDIO, num_freq=513, frame_length_ms=40.0, frame_shift_ms=10.0, griffin_lim_iters=8)
    apg, orcg = _ap(audio_g), OracleAudioProcessor(**audio_g)
    yg = synth_speech_like(32, n_samples=orcg.hop_length * 40)
    Sg = np.abs(lr_stft(yg.astype(np.float64), orcg.n_fft, orcg.hop_length, orcg.win_length)).astype(np.float32)
    angg = (2 * np.pi * np.random.default_rng(5).random(Sg.shape)).astype(np.float32)
    layg = apg.layout(n_frames=[Sg.shape[1]])
    wg = apg.griffin_lim_batch(torch.from_numpy(np.ascontiguousarray(Sg.T)).to(dev), layg, L.SPEC_MAGNITUDE,
                               init_angles=torch.from_numpy(np.ascontiguousarray(angg.T)).to(dev), momentum=0.9)
    wgo = orcg._griffin_lim_fast(Sg, 0.9, init_angles=angg)
    assert snr_db(wgo, wg[:len(wgo)].cpu().numpy()) >= GL_SNR_DB


def test_tacotron2_output_to_waveform_cfg4(tacotron2_postnet):
    """BASELINE configs[3]: the postnet output of the reference's Tacotron2 (fixture produced from models/tacotron2.py)
    stays on the device as [1, T, 80] and goes through inv_mel_spectrogram (60 iterations, config.json audio) end to
    end; the reference's own path for the same tensor (utils/synthesis.py:53-67: .cpu().numpy(), transpose,
    ap.inv_mel_spectrogram) is the oracle, with the initial phases injected identically."""
    ap, orc = _ap(MAIN_AUDIO), OracleAudioProcessor(**MAIN_AUDIO)
    mel = tacotron2_postnet                                     # [482, 80], roughly in [-0.35, 0.26]: _denormalize clips
    assert mel.shape == (482, 80)
    T = mel.shape[0]
    ang = (2 * np.pi * np.random.default_rng(42).random((1025, T))).astype(np.float32)
    wo, sco = orc.inv_mel_spectrogram(mel.T, init_angles=ang, return_sc=True)
    dev = torch.device("cuda")
    postnet = torch.from_numpy(mel).to(dev).unsqueeze(0)       # [1, T, 80] as models/tacotron2.py:62-73 returns it
    ang_dev = torch.from_numpy(np.ascontiguousarray(ang.T)).to(dev)
    wavs = ap.inv_mel_spectrogram_padded(postnet, [T], init_angles=ang_dev)
    w = wavs[0].cpu().numpy()
    assert w.shape == wo.shape == (275 * (T - 1),)
    assert snr_db(wo, w) >= GL_SNR_DB, snr_db(wo, w)
    w2, sc = ap.inv_mel_spectrogram(mel.T, init_angles=ang, return_sc=True)      # the drop-in numpy call
    assert snr_db(wo, w2) >= GL_SNR_DB
    np.testing.assert_allclose(sc, sco, rtol=SC_RTOL)
    # and straight to a WAV file as the server does (one sentence)
    data = ap.sentences_to_wav_bytes([postnet[0]], init_angles=ang_dev)
    import io
    from scipy.io import wavfile
    sr, pcm = wavfile.read(io.BytesIO(data))
    want = orc.save_wav_int16(orc.server_concat([wo]))
    assert sr == 22050 and len(pcm) == len(want) and np.abs(pcm.astype(int) - want.astype(int)).max() <= 2


def test_mel_gemm_variants_agree(monkeypatch):
    """mel <-> linear: the transposed warp-specialised tcgen05 kernel (default), the round-1 pipelined kernel
    (TTSA_MEL_GEMM=tc96), the un-pipelined tensor-core kernel (tc_simple) and the fp32 SIMT kernels (simt) all meet the
    float64 product within the same bound; T = 300 frames is three frame tiles with a ragged last one."""
    from your_voice_tts_b200 import audio as A
    orc = OracleAudioProcessor(**MAIN_AUDIO)
    rng = np.random.default_rng(12)
    T = 300
    mel_amp = (rng.random((80, T)) ** 3).astype(np.float32) * 5.0
    lin_amp = (rng.random((1025, T)) ** 3).astype(np.float32) * 5.0
    want_lin = np.maximum(1e-10, np.linalg.pinv(orc._build_mel_basis()) @ mel_amp.astype(np.float64))
    want_mel = orc._build_mel_basis() @ lin_amp.astype(np.float64)
    for mode in ("", "tc96", "tc_simple", "simt"):
        if mode:
            monkeypatch.setenv("TTSA_MEL_GEMM", mode)
        A._PLAN_CACHE.clear()
        ap = _ap(MAIN_AUDIO)
        got_lin = ap._mel_to_linear(mel_amp)
        got_mel = ap._linear_to_mel(lin_amp)
        assert np.abs(got_lin - want_lin).max() <= 3e-6 * np.abs(want_lin).max(), (mode, np.abs(got_lin - want_lin).max() / np.abs(want_lin).max())
        assert np.abs(got_mel - want_mel).max() <= 3e-6 * np.abs(want_mel).max(), (mode, np.abs(got_mel - want_mel).max() / np.abs(want_mel).max())
    monkeypatch.delenv("TTSA_MEL_GEMM")
    A._PLAN_CACHE.clear()


def test_mel_basis_that_does_not_fit_shared_memory():
    """mel_fmax=None (filters up to sr/2: 2 050 taps) does not fit beside the frame buffers, so the feature kernel
    reads the banded basis from global memory; 128 mels likewise.  Same forward bar."""
    for audio in (dict(MAIN_AUDIO, mel_fmax=None), dict(MAIN_AUDIO, num_mels=128, mel_fmax=11000.0)):
        ap, orc = _ap(audio), OracleAudioProcessor(**audio)
        y = synth_speech_like(17, n_samples=275 * 50 + 100)
        mel, mel_o = ap.melspectrogram(y), orc.melspectrogram(y)
        assert mel.shape == mel_o.shape
        assert np.mean(np.abs(mel - mel_o) <= FWD_TOL) >= 0.995
        ang = (2 * np.pi * np.random.default_rng(3).random((1025, mel_o.shape[1]))).astype(np.float32)
        a2 = dict(audio, griffin_lim_iters=4)
        w = _ap(a2).inv_mel_spectrogram(mel_o.astype(np.float32), init_angles=ang)
        wo = OracleAudioProcessor(**a2).inv_mel_spectrogram(mel_o.astype(np.float32), init_angles=ang)
        assert snr_db(wo, w) >= GL_SNR_DB, snr_db(wo, w)


def test_silent_and_tiny_spectrograms_stay_finite():
    """All-zero magnitudes give an all-zero waveform (np.angle(0) = 0 path), magnitudes far below float32's useful
    range stay finite, and a quiet-but-real signal (-120 dB) still meets the bar: the select-free zero-phase offset
    (1e-18) only matters for frames whose whole spectrum lies below ~1e-13."""
    audio = dict(MAIN_AUDIO, griffin_lim_iters=5)
    ap, orc = _ap(audio), OracleAudioProcessor(**audio)
    T = 37
    ang = (2 * np.pi * np.random.default_rng(0).random((1025, T))).astype(np.float32)
    w = ap._griffin_lim(np.zeros((1025, T), dtype=np.float32), init_angles=ang)
    assert w.shape == (275 * (T - 1),) and np.all(w == 0.0)
    w = ap._griffin_lim(np.full((1025, T), 1e-25, dtype=np.float32), init_angles=ang)
    assert np.isfinite(w).all() and np.abs(w).max() < 1e-20
    y = synth_speech_like(3, n_samples=275 * (T - 1) + 50) * 1e-6
    S = np.abs(lr_stft(y.astype(np.float64), 2048, 275, 1102)).astype(np.float32)
    wq, wqo = ap._griffin_lim(S, init_angles=ang), orc._griffin_lim(S, init_angles=ang)
    assert snr_db(wqo, wq) >= GL_SNR_DB, snr_db(wqo, wq)


def test_device_phases_opt_in_for_the_numpy_api():
    """AudioProcessor(device_phases=True): the drop-in numpy calls draw their initial phases on the device (seeded from
    numpy's global RNG, so np.random.seed still repeats a run) instead of uploading 2*pi*np.random.rand(F, T)."""
    audio = dict(MAIN_AUDIO, griffin_lim_iters=30)
    ap, ap_d, orc = _ap(audio), _ap(dict(audio, device_phases=True)), OracleAudioProcessor(**audio)
    assert ap.device_phases is False and ap_d.device_phases is True
    y = synth_speech_like(9, n_samples=275 * 80)
    spec = orc.spectrogram(y).astype(np.float32)
    np.random.seed(5); w1 = ap_d.inv_spectrogram(spec)
    np.random.seed(5); w2 = ap_d.inv_spectrogram(spec)
    np.random.seed(6); w3 = ap_d.inv_spectrogram(spec)
    assert np.array_equal(w1, w2) and not np.array_equal(w1, w3)
    # same quality as with host-drawn phases: spectral error of the result against the target magnitudes
    S = orc._db_to_amp(orc._denormalize(spec.astype(np.float64)) + orc.ref_level_db) ** orc.power
    def err(w):
        x = orc.apply_preemphasis(np.asarray(w, dtype=np.float64))      # undo the de-emphasis of inv_spectrogram
        return np.linalg.norm(np.abs(lr_stft(x, 2048, 275, 1102)) - S) / np.linalg.norm(S)
    np.random.seed(7)
    e_host, e_dev = err(ap.inv_spectrogram(spec)), err(w1)
    assert abs(e_dev - e_host) <= 0.15 * e_host, (e_dev, e_host)


@pytest.mark.parametrize("seed", list(range(12)))
def test_randomised_configs_vs_oracle(seed):
    """Seeded random audio blocks (sample rate, frame shift / length, normalisation flags, power, pre-emphasis, mel
    range, ragged frame counts incl. utterance edges): forward features, stft/istft and 3 Griffin-Lim iterations with
    injected phases against the oracle -- exercises both kernel classes, compile-time and run-time geometries and the
    any-size path on configurations nobody hand-picked."""
    rng = np.random.default_rng(1000 + seed)
    sr = int(rng.choice([16000, 22050, 24000, 20000, 11025]))
    num_freq = int(rng.choice([1025, 1025, 1025, 513, 2049]))
    n_fft = (num_freq - 1) * 2
    shift = float(rng.choice([5.0, 10.0, 12.5, 15.0]))
    hop = int(shift / 1000.0 * sr)
    ratio = float(rng.choice([2.0, 3.0, 4.0, 4.5, 6.5]))
    win = min(n_fft, int(ratio * hop))
    length = 1000.0 * (win + 0.5) / sr
    audio = dict(MAIN_AUDIO, sample_rate=sr, num_freq=num_freq, frame_shift_ms=shift, frame_length_ms=length,
                 symmetric_norm=bool(rng.integers(2)), max_norm=float(rng.choice([1.0, 4.0])),
                 clip_norm=bool(rng.integers(2)), power=float(rng.choice([1.0, 1.2, 1.5])),
                 preemphasis=float(rng.choice([0.0, 0.9, 0.97, 0.98])), ref_level_db=float(rng.choice([0.0, 20.0])),
                 mel_fmin=float(rng.choice([0.0, 50.0, 95.0])), mel_fmax=float(rng.choice([3800.0, 5000.0, sr / 2.0 - 12.5])),
                 griffin_lim_iters=3)
    orc = OracleAudioProcessor(**audio)
    assert (orc.hop_length, orc.win_length) == (hop, win), (orc.hop_length, orc.win_length, hop, win)
    if win - hop > 8 * hop:
        pytest.skip("win > 9 hop is outside the frame kernels' range")
    ap = _ap(audio)
    n = hop * int(rng.integers(3, 70)) + int(rng.integers(0, hop))
    y = synth_speech_like(seed, n_samples=n, sr=sr)
    spec, spec_o = ap.spectrogram(y), orc.spectrogram(y)
    mel, mel_o = ap.melspectrogram(y), orc.melspectrogram(y)
    tol = FWD_TOL * audio["max_norm"] * (2.0 if audio["symmetric_norm"] else 1.0)
    assert spec.shape == spec_o.shape and np.mean(np.abs(spec - spec_o) <= tol) >= 0.99, (audio, np.mean(np.abs(spec - spec_o) <= tol))
    assert mel.shape == mel_o.shape and np.mean(np.abs(mel - mel_o) <= tol) >= 0.99
    Do = lr_stft(y.astype(np.float64), orc.n_fft, hop, win)
    assert np.abs(ap._stft(y) - Do).max() <= 3e-6 * np.abs(Do).max()
    # ragged Griffin-Lim batch from random magnitudes (T = 1 gives an empty waveform)
    Ts = [int(t) for t in rng.integers(1, 45, size=4)]
    mags = [rng.random((T, num_freq)).astype(np.float32) ** 2 for T in Ts]
    angs = [(2 * np.pi * rng.random((T, num_freq))).astype(np.float32) for T in Ts]
    lay = ap.layout(n_frames=Ts)
    dev = torch.device("cuda")
    out, sc = ap.griffin_lim_batch(torch.from_numpy(np.concatenate(mags)).to(dev), lay, L.SPEC_MAGNITUDE,
                                   init_angles=torch.from_numpy(np.concatenate(angs)).to(dev), return_sc=True)
    for u, T in enumerate(Ts):
        w = lay.split_wav(out)[u].cpu().numpy()
        assert w.shape == (hop * max(0, T - 1),)
        if T < 2:
            continue
        wo, sco = orc._griffin_lim(mags[u].T, init_angles=angs[u].T, return_sc=True)
        assert snr_db(wo, w) >= GL_SNR_DB, (audio, T, snr_db(wo, w))
        np.testing.assert_allclose(sc[:, u].cpu().numpy(), sco, rtol=SC_RTOL)
