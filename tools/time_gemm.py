"""Live timing of the mel -> linear tcgen05 GEMM (ttsa_mel_to_linear, 64 x 482 frames, |S|**1.5 epilogue) and of the
feature kernel (64 x 6 s): CUDA events, output buffers rotated so that the writes stream to HBM."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import MAIN_AUDIO
from your_voice_tts_b200 import AudioProcessor, _lib as L
ap = AudioProcessor(verbose=False, **MAIN_AUDIO)
B, T = 64, 482
lay = ap.layout(n_frames=[T] * B)
plan = lay.plan
mel = torch.rand((B * T, 80), device="cuda")
outs = [torch.empty((B * T, 1025), device="cuda") for _ in range(4)]
def run(n):
    for i in range(n):
        L.check(plan.lib.ttsa_mel_to_linear(plan.handle, lay.handle, ap._ptr(mel), L.MEL_IN_NORM_DB, ap._ptr(outs[i & 3]), L.MEL_OUT_POWER, ap._stream()))
run(8); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); run(40); b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 40
print("mel_to_linear 64 x 482: %.1f us per call (HBM write bound %.1f us at %.0f GB/s)" % (1e3 * ms, 1e6 * B * T * 1025 * 4 / 6550.7e9, 6550.7))
wl = 132300
layw = ap.layout(wav_lengths=[wl] * B)
wav = torch.randn((B * wl,), device="cuda") * 0.1
lins = [torch.empty((layw.total_frames, 1025), device="cuda") for _ in range(3)]
mels = [torch.empty((layw.total_frames, 80), device="cuda") for _ in range(3)]
def runf(n):
    for i in range(n):
        ap.features_batch(wav, layw, lin_out=lins[i % 3], mel_out=mels[i % 3])
runf(6); torch.cuda.synchronize()
a.record(); runf(30); b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 30
byt = B * (4 * wl + 4 * 1025 * 482 + 4 * 80 * 482)
print("features 64 x 6 s: %.1f us per call = %.0f GB/s algorithmic = %.3f of the HBM roofline" % (1e3 * ms, byt / ms / 1e6, byt / ms / 1e6 / 6550.7))

# any-size path (K7): num_freq 513 (n_fft 1024, hop 220, win 882), 64 utterances x 601 frames, 10 Griffin-Lim iterations
apg = AudioProcessor(verbose=False, **dict(MAIN_AUDIO, num_freq=513, frame_length_ms=40.0, frame_shift_ms=10.0, griffin_lim_iters=10))
Tg = 601
layg = apg.layout(n_frames=[Tg] * B)
specg = torch.rand((B * Tg, 513), device="cuda")
outg = apg.inv_spectrogram_batch(specg, layg, seed=1)
torch.cuda.synchronize()
a.record()
for _ in range(3):
    outg = apg.inv_spectrogram_batch(specg, layg, seed=1, out=outg)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 3
print("any-size path, num_freq 513, 64 x 601 frames, 10 iterations + initial synthesis + de-emphasis: %.2f ms (%.0f audio-s/s at 10 iterations)"
      % (ms, B * 220 * (Tg - 1) / 22050 / (ms * 1e-3)))
