import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from conftest import MAIN_AUDIO
from your_voice_tts_b200 import AudioProcessor
from oracle.audio_oracle import OracleAudioProcessor
ap = AudioProcessor(verbose=False, **MAIN_AUDIO); orc = OracleAudioProcessor(**MAIN_AUDIO)
rng = np.random.default_rng(0)
mel_amp = (rng.random((80, 300)) * 2.0).astype(np.float32)
lin = ap._mel_to_linear(mel_amp); torch.cuda.synchronize()
lino = orc._mel_to_linear(mel_amp.astype(np.float64))
print("mel_to_linear max rel err", np.abs(lin - lino).max() / np.abs(lino).max(), "max abs", np.abs(lin-lino).max())
S = (rng.random((1025, 300)) * 3).astype(np.float32)
m = ap._linear_to_mel(S); mo = orc._linear_to_mel(S.astype(np.float64))
print("linear_to_mel max rel err", np.abs(m - mo).max() / np.abs(mo).max())
import os
os.environ["TTSA_MEL_GEMM"] = "simt"
lin2 = ap._mel_to_linear(mel_amp)
print("simt mel_to_linear max rel err", np.abs(lin2 - lino).max() / np.abs(lino).max(), "tc vs simt", np.abs(lin - lin2).max())
