"""One-off: the same plan used from several host threads on their own CUDA streams at once (the ABI's threading
contract: work calls are stream-ordered and re-entrant, a plan is immutable) -- results must equal the serial run."""
import os, sys, threading
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import MAIN_AUDIO
from your_voice_tts_b200 import AudioProcessor
ap = AudioProcessor(verbose=False, **dict(MAIN_AUDIO, griffin_lim_iters=20))
N = 6
g = torch.Generator(device="cuda").manual_seed(0)
jobs = []
for i in range(N):
    Ts = [int(t) for t in torch.randint(20, 200, (3 + i,), generator=torch.Generator().manual_seed(i))]
    lay = ap.layout(n_frames=Ts)
    mel = torch.rand((sum(Ts), 80), device="cuda", generator=g)
    jobs.append((lay, mel, 100 + i))
serial = [ap.inv_mel_spectrogram_batch(m, l, seed=s).clone() for l, m, s in jobs]
torch.cuda.synchronize()
out = [None] * N
def work(i):
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for _ in range(5):
            o = ap.inv_mel_spectrogram_batch(jobs[i][1], jobs[i][0], seed=jobs[i][2])
        out[i] = o.clone()
    st.synchronize()
ths = [threading.Thread(target=work, args=(i,)) for i in range(N)]
[t.start() for t in ths]; [t.join() for t in ths]
torch.cuda.synchronize()
print("identical to serial:", [bool(torch.equal(a, b)) for a, b in zip(serial, out)])
