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
a.record()
for _ in range(20):
    ap.features_batch(wav, lay, lin_out=lin, mel_out=mel)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 20
bytes_ = B * (132300 * 4 + 482 * 1025 * 4 + 482 * 80 * 4)
print("features B=%d: %.4f ms per batch, %.1f GB/s algorithmic (%.1f %% of 6550.7), %.0f audio-s/s" % (B, ms, bytes_ / ms / 1e6, 100 * bytes_ / ms / 1e6 / 6550.7, B * 6.0 / ms * 1e3))
