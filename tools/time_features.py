"""Time the feature kernel (BASELINE configs[2]: preemphasis + STFT + mel + amp_to_db + normalize) on synthetic waves."""
import sys, time, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from conftest import MAIN_AUDIO
from your_voice_tts_b200 import AudioProcessor
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
ap = AudioProcessor(verbose=False, **MAIN_AUDIO)
lay = ap.layout(wav_lengths=[132300] * B)
wav = torch.randn((lay.total_samples,), device="cuda") * 0.1
lin, mel = ap.features_batch(wav, lay)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
lin = torch.empty_like(lin); mel = torch.empty_like(mel)
what = sys.argv[2] if len(sys.argv) > 2 else "both"          # both | lin | mel
g = torch.cuda.CUDAGraph()                                    # graph replay: no host overhead between the launches
def call():
    ap.features_batch(wav, lay, want_linear=what != "mel", want_mel=what != "lin", lin_out=lin if what != "mel" else None,
                      mel_out=mel if what != "lin" else None)
call(); torch.cuda.synchronize()
with torch.cuda.graph(g):
    for _ in range(20):
        call()
g.replay(); torch.cuda.synchronize()
a.record()
g.replay()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 20
bytes_ = B * (132300 * 4 + 482 * 1025 * 4 + 482 * 80 * 4)
print("features[%s] B=%d: %.4f ms per batch, %.1f GB/s algorithmic (%.1f %% of 6550.7), %.0f audio-s/s" % (what, B, ms, bytes_ / ms / 1e6, 100 * bytes_ / ms / 1e6 / 6550.7, B * 6.0 / ms * 1e3))
