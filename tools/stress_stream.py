"""Race / ordering stress of the warp-stream Griffin-Lim kernel (gl_stream.cuh) -- compute-sanitizer is closed on this
GPU pool, so the hand-over protocol (raw head zones, release/acquire flags across CTAs, one-pass ring) is checked by
what a race would break: bit-identical repeats under a perturbing side stream, agreement between different partitions
(TTSA_WPS_GRID) and with the tile kernel, and the oracle on sampled utterances.
    python tools/stress_stream.py [repeats]"""
import os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = {"64x482": [482] * 64, "ragged180": None, "short300": None, "long12": [3000, 17, 2900, 5, 1, 2500, 2600, 40, 2800, 9, 3100, 2700]}


def frames(name):
    if CASES[name] is not None:
        return CASES[name]
    rng = np.random.default_rng(len(name))
    return [int(t) for t in (rng.integers(1, 700, size=180) if name == "ragged180" else rng.integers(1, 120, size=300))]


def child(name, repeats):
    import torch
    from conftest import MAIN_AUDIO, snr_db
    from oracle.audio_oracle import OracleAudioProcessor
    from your_voice_tts_b200 import AudioProcessor
    audio = dict(MAIN_AUDIO, griffin_lim_iters=4)
    ap, orc = AudioProcessor(verbose=False, **audio), OracleAudioProcessor(**audio)
    Ts = frames(name)
    lay = ap.layout(n_frames=Ts)
    g = torch.Generator(device="cuda").manual_seed(1)
    spec = torch.rand((sum(Ts), 1025), device="cuda", generator=g)
    ang = torch.rand((sum(Ts), 1025), device="cuda", generator=g) * 6.2831853
    side = torch.cuda.Stream()
    junk = torch.empty((64 << 20,), device="cuda")
    first, bad = None, 0
    for r in range(repeats):
        with torch.cuda.stream(side):                      # perturb the timing: copies and a GEMM on another stream
            junk.copy_(junk.roll(1)) if r % 3 == 0 else torch.mm(junk[:1 << 22].view(2048, 2048), junk[:1 << 22].view(2048, 2048))
        y = ap.inv_spectrogram_batch(spec, lay, init_angles=ang).clone()
        torch.cuda.synchronize()
        if first is None:
            first = y
        elif not torch.equal(first, y):
            bad += 1
    worst = 1e9
    off = np.concatenate(([0], np.cumsum(Ts)))
    for u in sorted(set([0, len(Ts) // 2, len(Ts) - 1])):
        if Ts[u] < 2:
            continue
        yo = orc.inv_spectrogram(spec[off[u]:off[u + 1]].cpu().numpy().T, init_angles=ang[off[u]:off[u + 1]].cpu().numpy().T)
        worst = min(worst, snr_db(yo, lay.split_wav(first)[u].cpu().numpy()))
    np.save(os.path.join(ROOT, "gpurun_out", "stress_%s_%s_%s.npy" % (name, os.environ.get("TTSA_GL_KERNEL", "stream"),
                                                                     os.environ.get("TTSA_WPS_GRID", "all"))), first.cpu().numpy())
    print("%-10s kernel=%-6s grid=%-4s repeats=%d nonidentical=%d worst SNR vs oracle %.1f dB" % (
        name, os.environ.get("TTSA_GL_KERNEL", "stream"), os.environ.get("TTSA_WPS_GRID", "all"), repeats, bad, worst), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        child(sys.argv[2], int(sys.argv[3]))
        sys.exit(0)
    repeats = int(sys.argv[1]) if len(sys.argv) > 1 else 25
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    for name in CASES:
        outs = {}
        for env in ({"TTSA_GL_KERNEL": "tile"}, {}, {"TTSA_WPS_GRID": "37"}, {"TTSA_WPS_GRID": "3"}, {"TTSA_WPS_GRID": "1"}):
            e = dict(os.environ); e.update(env)
            subprocess.run([sys.executable, __file__, "--child", name, str(repeats)], env=e, check=False)
            key = "%s_%s" % (env.get("TTSA_GL_KERNEL", "stream"), env.get("TTSA_WPS_GRID", "all"))
            path = os.path.join(ROOT, "gpurun_out", "stress_%s_%s.npy" % (name, key))
            if os.path.exists(path):
                outs[key] = np.load(path); os.remove(path)
        ref = outs.get("tile_all")
        for k, v in outs.items():
            if ref is not None and k != "tile_all":
                err = float(np.sum((ref - v) ** 2)); sig = float(np.sum(ref ** 2))
                print("%-10s %-10s vs tile kernel: %.1f dB" % (name, k, 10 * np.log10(sig / max(err, 1e-300))), flush=True)
