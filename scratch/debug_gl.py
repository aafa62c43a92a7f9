import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from conftest import MAIN_AUDIO, snr_db
from your_voice_tts_b200 import AudioProcessor
from oracle.audio_oracle import OracleAudioProcessor
quick = len(sys.argv) > 1
Ts = [40] if quick else [8, 9, 16, 24, 40, 41, 48, 100]
for T in Ts:
    for iters in ([1] if quick else [0, 1, 2, 5]):
        audio = dict(MAIN_AUDIO, griffin_lim_iters=iters, preemphasis=0.0)
        ap = AudioProcessor(verbose=False, **audio); orc = OracleAudioProcessor(**audio)
        rng = np.random.default_rng(T)
        spec = rng.random((T, 1025)).astype(np.float32)
        ang = (2*np.pi*rng.random((T, 1025))).astype(np.float32)
        lay = ap.layout(n_frames=[T])
        sd = torch.from_numpy(spec).cuda(); ad = torch.from_numpy(ang).cuda()
        y1 = ap.griffin_lim_batch(sd, lay, 1, init_angles=ad).clone()
        y2 = ap.griffin_lim_batch(sd, lay, 1, init_angles=ad).clone()
        torch.cuda.synchronize()
        if quick: continue
        S = orc._db_to_amp(orc._denormalize(spec.astype(np.float64).T) + 20) ** 1.5
        yo = orc._griffin_lim(S, init_angles=ang.T.astype(np.float64))
        y = y1[:275*(T-1)].cpu().numpy()
        err = np.abs(y - yo)
        print(f"T={T} iters={iters} equal={torch.equal(y1,y2)} snr={snr_db(yo,y):.1f} worst_idx={err.argmax()} of {len(y)} maxerr={err.max():.3g} rms={np.sqrt((yo**2).mean()):.3g}")
