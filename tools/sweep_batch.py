"""Griffin-Lim iteration time versus batch size (6 s utterances, 482 frames): where the fixed per-launch costs and the
tile imbalance stop mattering."""
import sys, ctypes, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from conftest import MAIN_AUDIO
from your_voice_tts_b200 import AudioProcessor, _lib as L
ap = AudioProcessor(verbose=False, **MAIN_AUDIO)
lib = L.load()
print("utts  iter_ms  per64_ms  %roofline  audio-s/s(60 it)")
BATCHES = [int(x) for x in sys.argv[1].split(',')] if len(sys.argv) > 1 else (1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024)
for B in BATCHES:
    lay = ap.layout(n_frames=[482] * B)
    plan = lay.plan
    S = torch.rand((lay.total_frames, 1025), device="cuda")
    wav = torch.zeros((lay.total_samples,), device="cuda")
    ws = torch.empty((int(lib.ttsa_griffin_lim_workspace_bytes(plan.handle, lay.handle)),), dtype=torch.uint8, device="cuda")
    def run(iters, reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        for _ in range(reps):
            L.check(lib.ttsa_griffin_lim(plan.handle, lay.handle, ap._ptr(S), L.SPEC_MAGNITUDE, iters, None, ctypes.c_uint64(1), 0,
                                         ap._ptr(wav), None, ap._ptr(ws), ws.numel(), ap._stream()))
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps
    g = torch.cuda.CUDAGraph()
    run(60, 1)
    with torch.cuda.graph(g):
        L.check(lib.ttsa_griffin_lim(plan.handle, lay.handle, ap._ptr(S), L.SPEC_MAGNITUDE, 60, None, ctypes.c_uint64(1), 0,
                                     ap._ptr(wav), None, ap._ptr(ws), ws.numel(), ap._stream()))
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5 if B <= 64 else 2
    a.record()
    for _ in range(reps): g.replay()
    b.record(); torch.cuda.synchronize()
    t60 = a.elapsed_time(b) / reps
    t0 = run(0, 3)
    it = (t60 - t0) / 60
    print("%5d  %7.4f  %7.4f  %6.1f  %9.0f" % (B, it, it * 64 / B, 100 * B * 3034400 / (it * 1e-3) / 1e9 / 6550.7, B * 481 * 275 / 22050 / (t60 * 1e-3)))
