"""Reduced GPU suite for compute-sanitizer (memcheck / racecheck / initcheck): every kernel family once, small sizes.
    compute-sanitizer --tool racecheck python tools/sanitizer_suite.py"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import MAIN_AUDIO
from your_voice_tts_b200 import AudioProcessor

dev = torch.device("cuda")
g = torch.Generator(device="cuda").manual_seed(0)

def section(name):
    torch.cuda.synchronize(); print("== " + name, flush=True)

# 1. ragged batch, tile kernel (small batch -> frame_kernel<GL_ITER>), injected phases, spectral convergence
ap = AudioProcessor(verbose=False, **dict(MAIN_AUDIO, griffin_lim_iters=2))
Ts = [40, 2, 153, 5, 9, 61, 1]
lay = ap.layout(n_frames=Ts)
spec = torch.rand((sum(Ts), 1025), device=dev, generator=g)
ang = torch.rand((sum(Ts), 1025), device=dev, generator=g) * 6.28
section("ragged batch: synthesis + tile GL iterations + de-emphasis")
y, sc = ap.inv_spectrogram_batch(spec, lay, init_angles=ang, return_sc=True)
# 2. 64 x 482 at 2 iterations: warp-stream kernel (cross-CTA head-zone hand-over), tcgen05 mel GEMM, device RNG phases
B, T = 64, 482
lay2 = ap.layout(n_frames=[T] * B)
mel = torch.rand((B * T, 80), device=dev, generator=g)
section("64 x 482: tcgen05 mel->linear GEMM, synthesis (Philox), gl_stream iterations, de-emphasis")
y2 = ap.inv_mel_spectrogram_batch(mel, lay2, seed=3)
# 2b. ragged large batch through the stream kernel
Ts3 = [int(t) for t in np.random.default_rng(1).integers(1, 700, size=180)]
lay3 = ap.layout(n_frames=Ts3)
spec3 = torch.rand((sum(Ts3), 1025), device=dev, generator=g)
section("ragged 180-utterance batch: gl_stream with runs that cross utterance boundaries")
y3 = ap.inv_spectrogram_batch(spec3, lay3, seed=5)
# 3. features (analysis kernel, mel contraction), stft / istft, pre-emphasis
wavs = torch.randn((4 * 20000,), device=dev, generator=g) * 0.1
layw = ap.layout(wav_lengths=[20000] * 4)
section("features: analysis kernel (linear + mel), stft, istft")
lin, melf = ap.features_batch(wavs, layw)
w = (np.random.default_rng(0).standard_normal(275 * 30 + 3) * 0.1).astype(np.float32)
D = ap._stft(w); yi = ap._istft(D)
section("linear -> mel GEMM (tcgen05), out_linear_to_mel")
m2 = ap.out_linear_to_mel(ap.spectrogram(w))
# 4. any-size path (n_fft 1024), fast Griffin-Lim (momentum), post-processing
apg = AudioProcessor(verbose=False, **dict(MAIN_AUDIO, num_freq=513, frame_length_ms=40.0, frame_shift_ms=10.0, griffin_lim_iters=2))
section("any-size path: num_freq 513")
sg = apg.spectrogram(w); yg = apg.inv_spectrogram(sg)
apm = AudioProcessor(verbose=False, **dict(MAIN_AUDIO, griffin_lim_iters=3, griffin_lim_momentum=0.99))
section("momentum Griffin-Lim")
ym = apm.inv_spectrogram(ap.spectrogram(w))
section("post-processing: peak / endpoint / pcm16")
pcm = ap.sentences_to_wav_bytes([torch.rand((30, 80), device=dev), torch.rand((12, 80), device=dev)], seed=1)
ap.find_endpoint(torch.from_numpy(yi).cuda())
torch.cuda.synchronize()
print("suite done", flush=True)
