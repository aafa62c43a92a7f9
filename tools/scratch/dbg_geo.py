import os, sys, numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from conftest import MAIN_AUDIO, snr_db, synth_speech_like
from oracle.audio_oracle import OracleAudioProcessor
from your_voice_tts_b200 import AudioProcessor, audio as A

def run(audio, Ts, generic, pre_spec, clear):
    os.environ["TTSA_GENERIC_GEO"] = generic
    if clear: A._PLAN_CACHE.clear()
    orc = OracleAudioProcessor(**audio); hop = orc.hop_length
    ys = [synth_speech_like(5 + i, n_samples=hop * n) for i, n in enumerate(Ts)]
    specs_o = [orc.spectrogram(y).astype(np.float32) for y in ys]
    angs = [(2 * np.pi * np.random.default_rng(i).random(s.shape)).astype(np.float32) for i, s in enumerate(specs_o)]
    ap = AudioProcessor(verbose=False, **audio)
    if pre_spec: ap.spectrogram(ys[0])
    lay = ap.layout(n_frames=[s.shape[1] for s in specs_o])
    dev = torch.device("cuda")
    out, sc = ap.inv_spectrogram_batch(torch.from_numpy(np.concatenate([s.T for s in specs_o])).to(dev), lay,
                                       init_angles=torch.from_numpy(np.concatenate([a.T for a in angs])).to(dev), return_sc=True)
    res = [o.cpu().numpy() for o in lay.split_wav(out)]
    snrs = []
    for u, s in enumerate(specs_o):
        wo, sco = orc.inv_spectrogram(s, init_angles=angs[u], return_sc=True)
        snrs.append(round(float(snr_db(wo, res[u])), 1))
        # single-utterance API
        w1 = ap.inv_spectrogram(s, init_angles=angs[u])
        snrs.append(('single', round(float(snr_db(wo, w1)), 1)))
    return snrs

base = dict(MAIN_AUDIO, griffin_lim_iters=4)
print('A main p=.98 fixed', run(base, (37, 9, 64), "0", False, False))
print('B main p=.97 fixed', run(dict(base, preemphasis=0.97), (37, 9, 64), "0", False, False))
print('C main p=.97 fixed prespec', run(dict(base, preemphasis=0.97), (37, 9, 64), "0", True, False))
print('D main p=.97 generic clear', run(dict(base, preemphasis=0.97), (37, 9, 64), "1", False, True))
print('E main p=.97 fixed clear', run(dict(base, preemphasis=0.97), (37, 9, 64), "0", False, True))
print('F main p=.98 fixed T482', run(base, (481,), "0", False, False))
