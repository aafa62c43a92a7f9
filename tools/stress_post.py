"""One-off stress of the waveform post-processing kernels: random waveforms (lengths, scales, silences) against the oracle;
int16 samples and endpoints must be bit-exact."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import MAIN_AUDIO
from oracle.audio_oracle import OracleAudioProcessor
from your_voice_tts_b200 import AudioProcessor
ap, orc = AudioProcessor(verbose=False, **MAIN_AUDIO), OracleAudioProcessor(**MAIN_AUDIO)
rng = np.random.default_rng(0)
bad = 0
for trial in range(40):
    B = int(rng.integers(1, 9))
    wavs = []
    for _ in range(B):
        n = int(rng.integers(1, 90000))
        w = (rng.standard_normal(n) * 10.0 ** rng.uniform(-4, 0.5)).astype(np.float32)
        if rng.random() < 0.6 and n > 30000:                       # a silent stretch
            a = int(rng.integers(0, n - 20000)); b = a + int(rng.integers(5000, 40000))
            w[a:b] *= np.float32(10.0 ** rng.uniform(-6, -1.5))
        if rng.random() < 0.2:
            w = -np.abs(w) - np.float32(0.02)
        wavs.append(w)
    lay = ap.layout(wav_lengths=[len(w) for w in wavs])
    buf = torch.zeros((max(1, lay.total_samples),), device="cuda")
    for u, w in enumerate(wavs):
        buf[int(lay.wav_off[u]):int(lay.wav_off[u]) + len(w)] = torch.from_numpy(w).cuda()
    thr, sec = float(rng.choice([-40.0, -30.0, -55.0])), float(rng.choice([0.8, 0.4, 0.25]))
    ends = ap.find_endpoint_batch(buf, lay, thr, sec)
    want_e = [orc.find_endpoint(w, thr, sec) for w in wavs]
    if ends.cpu().numpy().tolist() != want_e:
        bad += 1; print(trial, "endpoint mismatch", ends.cpu().numpy().tolist(), want_e)
    for f32 in (False, True):
        gap = int(rng.choice([0, 0, 10000, 17]))
        joint = bool(rng.integers(2))
        pcm, off = ap.pcm16_batch(buf, lay, lens=ends if rng.random() < 0.5 else None, joint_peak=joint, gap_samples=gap, float32_arith=f32)
        off = off.cpu().numpy(); pcm = pcm.cpu().numpy()
        lens = [int(off[u + 1] - off[u] - gap) for u in range(B)]
        srcs = [w[:l] for w, l in zip(wavs, lens)]
        peak = max([float(np.max(np.abs(s))) if len(s) else 0.0 for s in srcs]) if joint else None
        for u, s in enumerate(srcs):
            pk = peak if joint else (float(np.max(np.abs(s))) if len(s) else 0.0)
            scale = 32767.0 / max(0.01, pk)
            want = (s * np.float32(scale)).astype(np.int16) if f32 else (s.astype(np.float64) * scale).astype(np.int16)
            got = pcm[off[u]:off[u] + len(s)]
            if not np.array_equal(got, want):
                bad += 1; print(trial, u, f32, joint, "pcm mismatch", np.abs(got.astype(int) - want.astype(int)).max())
            if gap and np.any(pcm[off[u] + len(s):off[u + 1]] != 0):
                bad += 1; print(trial, u, "gap not zero")
print("failures", bad)
