"""Per-warp phase stamps of the fused warp-stream Griffin-Lim kernel (experiment build -DTTSA_WPS_TRACE):
    TTSA_BUILD_TAG=_trace TTSA_NVCC_EXTRA=-DTTSA_WPS_TRACE python your-voice-tts_b200/build.py
    TTSA_LIB=$PWD/your-voice-tts_b200/libttsa_b200_trace.so python tools/wps_trace.py gpurun_out/trace.npz
Stamps (global timer, ns) per warp and iteration: 0 top, 1 after the start wait, 2 frames done, 3 after the zone wait,
4 zone finished, 5 done flag published."""
import sys, ctypes, numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from conftest import MAIN_AUDIO
from your_voice_tts_b200 import AudioProcessor, _lib as L
ap = AudioProcessor(verbose=False, **MAIN_AUDIO)
lib = L.load()
B, ITERS = 64, 60
lay = ap.layout(n_frames=[482] * B)
plan = lay.plan
S = torch.rand((lay.total_frames, 1025), device="cuda")
wav = torch.zeros((lay.total_samples,), device="cuda")
ws = torch.zeros((int(lib.ttsa_griffin_lim_workspace_bytes(plan.handle, lay.handle)),), dtype=torch.uint8, device="cuda")
def run():
    L.check(lib.ttsa_griffin_lim(plan.handle, lay.handle, ap._ptr(S), L.SPEC_MAGNITUDE, ITERS, None, ctypes.c_uint64(1), 0,
                                 ap._ptr(wav), None, ap._ptr(ws), ws.numel(), ap._stream()))
for _ in range(3):
    run()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); run(); b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b)
nw = torch.cuda.get_device_properties(0).multi_processor_count * 16
tr = ws[ws.numel() - nw * 64 * 8 * 8:].cpu().numpy().view(np.uint64).reshape(nw, 64, 8)
np.savez_compressed(sys.argv[1], trace=tr, ms=ms)
t = tr[:, :ITERS, :6].astype(np.int64)
t0 = t[:, 0, 0].min()
print("call ms", ms, "kernel span us", (t[:, ITERS - 1, 5].max() - t0) / 1e3)
per_it = (t[:, ITERS - 1, 5].max() - t0) / 1e3 / ITERS
print("per iteration us", per_it)
d = np.diff(t, axis=2) / 1e3
print("mean us per phase [start wait, frames, zone wait, zone finish, publish]:", d[:, 1:].mean(axis=(0, 1)))
print("iteration period per warp (us): mean", np.diff(t[:, :, 0], axis=1).mean() / 1e3)
