"""BASELINE configs[3] fixture: the reference's own Tacotron2 (models/tacotron2.py, random-init weights) run in THIS
container on CPU over synthetic text; its postnet output is what the synthesis path receives
(utils/synthesis.py:53-67, server/synthesizer.py:146-156).  The reference tree does not travel to the GPU box, so the
output is committed as a fixture (tests/golden/tacotron2_cfg4.npz) together with this script.

    python tests/golden/make_tacotron2_golden.py        # needs /root/reference

Third-party imports of the reference's text front-end that are absent here (phonemizer, unidecode, inflect, librosa,
soundfile, tensorboardX, matplotlib) are stubbed: the model code path never calls them.
Model flags = config.json:42-51."""
import importlib
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


class _Stub(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return lambda *a, **kw: ""


def main():
    sys.path.insert(0, REF)
    for name in ("phonemizer", "phonemizer.phonemize", "unidecode", "inflect", "librosa", "soundfile", "librosa.filters",
                 "librosa.effects", "tensorboardX", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = _Stub(name)
    from models.tacotron2 import Tacotron2

    torch.manual_seed(0)
    model = Tacotron2(num_chars=130, num_speakers=0, r=1, attn_norm="sigmoid", forward_attn=True, forward_attn_mask=True,
                      location_attn=False, separate_stopnet=True)
    model.eval()
    tokens = torch.randint(3, 130, (1, 230), generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        decoder_out, postnet_out, alignments, stop_tokens = model.inference(tokens)
    mel = postnet_out[0].cpu().numpy().astype(np.float32)          # [T, 80], as utils/synthesis.py:53 takes it
    print("postnet output", mel.shape, "min %.3f max %.3f mean %.3f" % (mel.min(), mel.max(), mel.mean()))
    np.savez_compressed(os.path.join(HERE, "tacotron2_cfg4.npz"), postnet_out=mel, tokens=tokens.numpy().astype(np.int32))


if __name__ == "__main__":
    main()
