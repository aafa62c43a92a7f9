"""BASELINE configs[0]: latency of AudioProcessor.inv_spectrogram (Griffin-Lim 60 iterations) on ONE synthetic 6 s
LJSpeech-shape linear spectrogram -- through the drop-in numpy call (host in / host out) and device-resident."""
import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from conftest import MAIN_AUDIO, synth_speech_like
from oracle.audio_oracle import OracleAudioProcessor
from your_voice_tts_b200 import AudioProcessor
ap = AudioProcessor(verbose=False, **MAIN_AUDIO)
orc = OracleAudioProcessor(**MAIN_AUDIO)
spec = orc.spectrogram(synth_speech_like(1234)).astype(np.float32)      # [1025, 482]
for _ in range(3):
    w = ap.inv_spectrogram(spec)
torch.cuda.synchronize()
t0 = time.perf_counter()
n = 20
for _ in range(n):
    w = ap.inv_spectrogram(spec)
dt = (time.perf_counter() - t0) / n
print("numpy in / numpy out: %.3f ms per call (%.0f x real time)" % (dt * 1e3, 6.0 / dt))
apd = AudioProcessor(verbose=False, **dict(MAIN_AUDIO, device_phases=True))
for _ in range(3):
    w = apd.inv_spectrogram(spec)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(n):
    w = apd.inv_spectrogram(spec)
dtd = (time.perf_counter() - t0) / n
print("numpy in / numpy out, device_phases=True: %.3f ms per call" % (dtd * 1e3))
lay = ap.layout(n_frames=[482])
st = torch.from_numpy(np.ascontiguousarray(spec.T)).cuda()
out = ap.inv_spectrogram_batch(st, lay)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(n):
    out = ap.inv_spectrogram_batch(st, lay, out=out)
b.record(); torch.cuda.synchronize()
print("device resident: %.3f ms per call" % (a.elapsed_time(b) / n))
t0 = time.perf_counter(); wo = orc.inv_spectrogram(spec); dtc = time.perf_counter() - t0
print("CPU oracle port (one host thread pool as numpy runs it): %.2f s per call" % dtc)
