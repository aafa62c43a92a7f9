#!/usr/bin/env python
"""Headline benchmark: Griffin-Lim audio-seconds synthesised per second (60 iterations, n_fft 2048).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): batched inv_mel_spectrogram -- 80 mels -> linear through the pseudo-inverse
mel basis -> 60 Griffin-Lim iterations -> de-emphasis -- over 64 synthetic 6 s utterances of LJSpeech shape
(22050 Hz, num_freq 1025, hop 275, win 1102, T = 482) per GPU.  One step = one pass over one batch.

  value   whole-job audio-s/s with the mel spectrograms already resident in HBM (CUDA events, max over ranks)
  e2e     same metric through the public AudioProcessor call with HOST buffers: pinned host mel -> device,
          compute, device -> pinned host waveform, all inside the timed region
  roofline  Griffin-Lim iteration kernel: algorithmic bytes per launch / measured launch duration vs the measured
          HBM copy bandwidth (MEASURED_PEAKS.json); one launch = one iteration over the batch (with TTSA_GL_FUSE=n a
          launch runs n iterations and the figure is per iteration); roofline_fp32: the FP32 pipe, one of the three
          resources that bind
  gpu_launches  kernel launches of the library inside the timed region, counted by the library (ttsa_launch_count)
  e2e_dropin  the same 64 utterances through the reference-signature call, one ap.inv_mel_spectrogram(np.ndarray)
          per utterance, host arrays in and out, host-drawn random phases as the reference
  latency_single_ms  BASELINE configs[0]: one 6 s linear spectrogram through inv_spectrogram, device resident
  cpu_baseline  the float64 oracle port of the reference (oracle/audio_oracle.py) on the host cores, bounded sample

--impl reference runs that CPU implementation alone (rank 0 only) and prints the same line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

AUDIO = dict(num_mels=80, num_freq=1025, sample_rate=22050, frame_length_ms=50, frame_shift_ms=12.5,
             preemphasis=0.98, min_level_db=-100, ref_level_db=20, power=1.5, griffin_lim_iters=60,
             signal_norm=True, symmetric_norm=False, max_norm=1, clip_norm=True, mel_fmin=0.0, mel_fmax=8000.0,
             do_trim_silence=True)           # config.json:5-25
SR, HOP, F, MELS, T_FRAMES, WAV_LEN = 22050, 275, 1025, 80, 482, 132300
L_OUT = HOP * (T_FRAMES - 1)                 # 132275 samples = 5.9989 s per utterance
ITERS = 60
# algorithmic bytes: one GL iteration reads y (4*L_OUT), reads |S| (4*F*T), writes y (4*L_OUT)
BYTES_PER_UTT_ITER = 4 * L_OUT + 4 * F * T_FRAMES + 4 * L_OUT      # 3 034 400


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------------------
# synthetic inputs
# ------------------------------------------------------------------------------------------------------------
def synth_waves_torch(batch, seed, device):
    """Seeded speech-like waves [batch, WAV_LEN]: harmonic stack with vibrato, AM envelope and a noise floor."""
    import math
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    t = torch.arange(WAV_LEN, device=device, dtype=torch.float64) / SR
    ph = torch.rand((batch, 3 + 30), device=device, generator=g, dtype=torch.float64) * (2 * math.pi)
    f0 = 120.0 + 30.0 * torch.sin(2 * math.pi * 0.7 * t[None, :] + ph[:, 0:1]) + 40.0 * (ph[:, 1:2] / math.pi - 1.0)
    phi = 2 * math.pi * torch.cumsum(f0, dim=1) / SR
    y = torch.zeros((batch, WAV_LEN), device=device, dtype=torch.float64)
    for k in range(1, 31):
        y += torch.sin(k * phi + ph[:, 2 + k:3 + k]) / k
    env = 0.55 + 0.45 * torch.sin(2 * math.pi * 2.3 * t[None, :] + ph[:, 2:3])
    noise = torch.randn((batch, WAV_LEN), device=device, generator=g, dtype=torch.float64)
    return (0.25 * y * env + 0.003 * noise).float()


def make_inputs(ap, batch, seed, device):
    """Normalised mel spectrograms [batch*T, 80] of the synthetic waves, produced by the product's own forward path."""
    waves = synth_waves_torch(batch, seed, device)
    lay_w = ap.layout(wav_lengths=[WAV_LEN] * batch)
    _, mel = ap.features_batch(waves.reshape(-1).contiguous(), lay_w, want_linear=False, want_mel=True)
    assert mel.shape == (batch * T_FRAMES, MELS)
    return mel


# ------------------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines, self.windows = index, None, [], []

    def start(self):
        """Started before the warm-up (nvidia-smi needs ~0.1 s to produce its first line); only samples that fall
        inside a begin()/end() window -- the timed regions -- are reported."""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def begin(self):
        self.windows.append([time.perf_counter(), None])

    def end(self):
        if self.windows and self.windows[-1][1] is None:
            self.windows[-1][1] = time.perf_counter()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def collect(pred):
            sm, mx, reasons = [], None, set()
            for t, ln in self.lines:
                if not pred(t):
                    continue
                parts = [p.strip() for p in ln.split(",")]
                if len(parts) < 6:
                    continue
                try:
                    sm.append(float(parts[0])); mx = float(parts[1])
                except ValueError:
                    continue
                for n, v in zip(names, parts[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            return sm, mx, reasons

        wins = [(a, b if b is not None else float("inf")) for a, b in self.windows]
        sm, mx, reasons = collect(lambda t: any(a <= t <= b + 0.03 for a, b in wins))
        window = "timed regions"
        if not sm and wins:            # timed regions shorter than the sampling period: the GPU has been under the same
            sm, mx, reasons = collect(lambda t: t >= wins[0][0] - 1.0)    # load since the warm-up; use those samples
            window = "warm-up + timed regions"
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "window": window}


# ------------------------------------------------------------------------------------------------------------
# CPU implementation of the reference path (oracle port), all host cores, bounded sample
# ------------------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    import numpy as np
    from oracle.audio_oracle import OracleAudioProcessor
    mel_dt, seed, iters = args
    orc = OracleAudioProcessor(**dict(AUDIO, griffin_lim_iters=iters))
    np.random.seed(seed)
    y = orc.inv_mel_spectrogram(mel_dt)          # [80, T] -> wav, 60 iterations, exactly the reference's call
    return len(y)


def cpu_reference_run(n_utts, cores):
    """Time the oracle's inv_mel_spectrogram on `n_utts` utterances spread over `cores` processes."""
    import multiprocessing as mp
    import numpy as np
    from oracle.audio_oracle import OracleAudioProcessor
    orc = OracleAudioProcessor(**AUDIO)
    rng = np.random.default_rng(1234)
    t = np.arange(WAV_LEN) / SR
    wav = (0.25 * sum(np.sin(2 * np.pi * 120 * k * t + rng.uniform(0, 6.28)) / k for k in range(1, 31))
           * (0.55 + 0.45 * np.sin(2 * np.pi * 2.3 * t)) + 0.003 * rng.standard_normal(WAV_LEN))
    mel = orc.melspectrogram(wav).astype(np.float32)      # [80, 482]
    for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = "1"                                  # one thread per worker process, no oversubscription
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, [(mel[:, :8], 0, 2)] * cores)   # process start-up and imports, untimed
        t0 = time.perf_counter()
        lens = pool.map(_cpu_worker, [(mel, 100 + i, ITERS) for i in range(n_utts)])
        dt = time.perf_counter() - t0
    return sum(lens) / SR / dt, dt


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------------------------------------
def run_reference(args, rank):
    if rank != 0:
        return
    cores = host_cores()
    per_step = cores                       # one utterance per core per step: ~5-10 s of CPU work per step
    vals = []
    total = args.warmup + args.steps
    # every step is a bounded sample (one utterance per host thread, ~3-5 s of wall time); the requested step counts
    # are honoured up to a bound that keeps the whole run within a few minutes
    n_warm, n_steps = max(0, args.warmup), max(1, min(args.steps, 30))
    for i in range(n_warm + n_steps):
        v, dt = cpu_reference_run(per_step, cores)
        if i >= n_warm:
            vals.append((v, dt))
    value = sum(v for v, _ in vals) / len(vals)
    ms = 1e3 * sum(dt for _, dt in vals) / len(vals)
    sample = "%d utterances (one per host thread) x inv_mel_spectrogram %d iters per step, %d of %d requested steps" % (
        per_step, ITERS, n_steps, args.steps)
    line = {"impl": "reference", "metric": "griffin_lim_audio_sec_per_sec", "value": value, "unit": "audio-s/s",
            "n_gpus": args.gpus, "steps": n_steps, "warmup": n_warm, "steps_requested": args.steps,
            "warmup_requested": args.warmup, "steps_clamped": n_steps != args.steps,
            "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args, 64),
            "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def config_dict(args, batch):
    which = "configs[1]" if (batch == 64 and ITERS == 60) else "configs[4] shard (throughput sweep)"
    return {"workload": "%s: batched inv_mel_spectrogram, 80 mels -> pinv mel basis -> Griffin-Lim %d iters "
                        "-> de-emphasis, %d synthetic 6 s utterances per GPU" % (which, ITERS, batch),
            "sample_rate": SR, "num_freq": F, "n_fft": 2048, "hop_length": HOP, "win_length": 1102,
            "frames_per_utt": T_FRAMES, "utts_per_gpu": batch, "griffin_lim_iters": ITERS,
            "parallelism": "utterances sharded across GPUs, no data-path collective",
            "l2": "inputs exceed L2: per iteration 126 MB of |S| + 2 x 34 MB of waveform stream through a 126 MB L2"}


def main():
    ap_ = argparse.ArgumentParser()
    ap_.add_argument("--gpus", type=int, default=1)
    ap_.add_argument("--steps", type=int, default=20)
    ap_.add_argument("--warmup", type=int, default=3)
    ap_.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap_.add_argument("--batch", type=int, default=64, help="utterances per GPU")
    ap_.add_argument("--iters", type=int, default=ITERS, help="Griffin-Lim iterations (BASELINE configs[4] sweeps 30 and 60)")
    ap_.add_argument("--no-cpu-baseline", action="store_true")
    ap_.add_argument("--no-graph", action="store_true", help="launch kernels directly instead of replaying a CUDA graph")
    ap_.add_argument("--gather", action="store_true", help="also time the optional final gather of the waveforms (NCCL)")
    ap_.add_argument("--no-extras", action="store_true", help="skip e2e_dropin / latency_single_ms (profiling runs)")
    args = ap_.parse_args()
    globals()["ITERS"] = int(args.iters)
    AUDIO["griffin_lim_iters"] = int(args.iters)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback "
                         "(use --impl reference for the CPU reference arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group(backend="nccl", device_id=dev)
    if args.warmup < 3:
        args.warmup = 3

    import __graft_entry__
    __graft_entry__.build()
    from your_voice_tts_b200 import AudioProcessor, _lib as L
    lib = L.load()
    ap = AudioProcessor(verbose=False, **AUDIO)
    B = args.batch
    mel = make_inputs(ap, B, 1234 + 1000 * rank, dev)
    lay = ap.layout(n_frames=[T_FRAMES] * B)
    audio_sec_per_step = B * L_OUT / SR
    plan = lay.plan

    wav_out = torch.zeros((lay.total_samples,), dtype=torch.float32, device=dev)
    ws = torch.empty((int(lib.ttsa_griffin_lim_workspace_bytes(plan.handle, lay.handle)),), dtype=torch.uint8, device=dev)
    S_buf = torch.empty((lay.total_frames, F), dtype=torch.float32, device=dev)

    import ctypes

    def step_device(seed):
        """inv_mel_spectrogram over the resident batch (mel -> |S|^power -> GL 60 -> de-emphasis)."""
        st = ap._stream()
        L.check(lib.ttsa_mel_to_linear(plan.handle, lay.handle, ap._ptr(mel), L.MEL_IN_NORM_DB, ap._ptr(S_buf),
                                       L.MEL_OUT_POWER, st))
        L.check(lib.ttsa_griffin_lim(plan.handle, lay.handle, ap._ptr(S_buf), L.SPEC_MAGNITUDE, ITERS, None,
                                     ctypes.c_uint64(seed), L.GL_DEEMPHASIS, ap._ptr(wav_out), None, ap._ptr(ws),
                                     ws.numel(), st))

    # optional CUDA graph of one step (the library is capture-safe: kernel launches only)
    graph = None
    n0 = int(lib.ttsa_launch_count())
    step_device(1)
    launches_per_step = int(lib.ttsa_launch_count()) - n0       # counted by the library, not assumed
    torch.cuda.synchronize()
    if not args.no_graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            step_device(1)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            step_device(1)

    def run_step(i):
        if graph is not None:
            graph.replay()
        else:
            step_device(1 + i)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ("value") ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    for i in range(args.warmup):
        run_step(i)
    barrier()
    sampler.begin()
    launches0 = int(lib.ttsa_launch_count())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        run_step(i)
    e1.record()
    barrier()
    launches = int(lib.ttsa_launch_count()) - launches0
    if graph is not None:
        launches = launches_per_step * args.steps          # replayed from the captured graph
    ms_total = e0.elapsed_time(e1)
    sampler.end()
    t_ms = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(t_ms.item()) / args.steps
    value = world * audio_sec_per_step / (ms_per_step * 1e-3)

    # ---- end-to-end through the public API with host buffers ("e2e") ----
    # every step: pinned host mel -> device, inv_mel_spectrogram, device waveform -> pinned host, all inside the
    # timed region; consecutive steps are double-buffered (HostPipeline) so the copies overlap neighbouring compute
    from your_voice_tts_b200 import HostPipeline
    mel_host = mel.cpu().pin_memory()
    wav_hosts = [torch.empty((lay.total_samples,), dtype=torch.float32).pin_memory() for _ in range(2)]
    h2d, d2h = mel_host.numel() * 4, wav_hosts[0].numel() * 4
    pipe = HostPipeline(ap, lay, graph=not args.no_graph)
    for i in range(args.warmup):
        pipe.submit(mel_host, wav_hosts[i & 1], seed=1 + i)
    pipe.drain()
    barrier()
    sampler.begin()
    t0 = time.perf_counter()
    with torch.cuda.stream(pipe.h2d):
        e0.record(pipe.h2d)                                 # before the first host -> device copy
    for i in range(args.steps):
        pipe.submit(mel_host, wav_hosts[i & 1], seed=1 + i)
    with torch.cuda.stream(pipe.copy):
        e1.record(pipe.copy)                                # after the last device -> host copy
    pipe.drain()
    barrier()
    wall = time.perf_counter() - t0
    sampler.end()
    clocks = sampler.stop()
    e2e_ms = max(e0.elapsed_time(e1), 0.0)
    if not os.environ.get("TTSA_DEBUG"):                              # (profiling builds skip phases of the kernel)
        assert float(wav_hosts[(args.steps - 1) & 1].abs().max()) > 0.0   # the waveform really arrived on the host
    t_e = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = world * audio_sec_per_step * args.steps / (float(t_e.item()) * 1e-3)

    # ---- roofline of the dominant kernel: the Griffin-Lim iteration ----
    def time_gl(iters, reps):
        st = ap._stream()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            L.check(lib.ttsa_griffin_lim(plan.handle, lay.handle, ap._ptr(S_buf), L.SPEC_MAGNITUDE, iters, None,
                                         ctypes.c_uint64(3), 0, ap._ptr(wav_out), None, ap._ptr(ws), ws.numel(), st))
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps
    time_gl(ITERS, 1)
    t_full = time_gl(ITERS, 3)
    t_init = time_gl(0, 3)
    iter_ms = (t_full - t_init) / ITERS
    hbm_peak, peak_src = peaks()
    algo_bytes = B * BYTES_PER_UTT_ITER
    achieved = algo_bytes / (iter_ms * 1e-3) / 1e9
    traffic, prof = None, {}
    tpath = os.path.join(ROOT, "profiles", "gl_iter_traffic.json")
    if os.path.exists(tpath):
        try:
            prof = json.load(open(tpath))
            traffic = prof.get("dram_bytes_per_launch") if (B == 64) else None
        except Exception:
            traffic, prof = None, {}
    stream_kernel = os.environ.get("TTSA_GL_KERNEL", "") != "tile"
    roofline = {"kernel": ("gl_stream_kernel" if stream_kernel else "frame_kernel<MODE_GL_ITER>") +
                          " (one Griffin-Lim iteration over the batch)", "bound": "hbm",
                "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic,
                "traffic_source": (None if traffic is None else
                                   "profiles/gl_iter_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of one "
                                   "`ncu --set full` capture of this kernel at this batch (%s); committed constant, NOT "
                                   "measured in this run" % prof.get("report", "?")),
                "frac_of_nominal_8tbs": achieved / 8000.0,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": algo_bytes, "launch_ms": iter_ms,
                "share_of_step": ITERS * iter_ms / ms_per_step,
                "init_synthesis_ms": t_init}
    # the roof that binds: the kernel executes a fixed number of FP32-pipe cycles per launch (ncu: sm__pipe_fma_cycles_active
    # per SM, committed with the profile); 100 % pipe utilisation would finish them in cycles / clock
    roofline_fp32 = None
    if B == 64 and stream_kernel and prof.get("fma_pipe_cycles_per_sm"):
        clk = 1e6 * float(clocks.get("sm_mhz") or prof.get("sm_mhz_nominal", 1965.0))
        busy = float(prof["fma_pipe_cycles_per_sm"])
        roofline_fp32 = {"bound": "fp32 pipe (FFMA2/FADD2/FMUL2 hold it two cycles, IMAD shares it)",
                         "fma_pipe_cycles_per_sm_per_launch": busy, "sm_clock_hz": clk,
                         "min_ms_at_full_pipe": 1e3 * busy / clk, "launch_ms": iter_ms,
                         "frac": (busy / clk) / (iter_ms * 1e-3),
                         "source": "profiles/gl_iter_traffic.json (ncu sm__pipe_fma_cycles_active.avg of the same capture); "
                                   "the pipe-cycle count is a property of the instruction stream, the fraction uses this "
                                   "run's live launch time"}

    line = {"metric": "griffin_lim_audio_sec_per_sec", "value": value, "unit": "audio-s/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args, B), "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "audio-s/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "wall_ms_per_step": 1e3 * wall / args.steps},
            "gpu_launches": launches, "cuda_graph": graph is not None, "roofline": roofline,
            "roofline_fp32": roofline_fp32}
    if not (B == 64 and ITERS == 60):
        line["configs4"] = {"total_utterances": world * B, "utterances_per_gpu": B, "griffin_lim_iters": ITERS, "gpus": world}

    # ---- the literal drop-in call: one numpy [80, T] mel in, one numpy waveform out, per utterance (rank 0) ----
    if rank == 0 and not args.no_extras:
        mel_np = [np.ascontiguousarray(m.T) for m in mel.reshape(B, T_FRAMES, MELS)[:min(B, 64)].cpu().numpy()]
        np.random.seed(0)
        for m in mel_np[:3]:
            ap.inv_mel_spectrogram(m)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n_out = 0
        for m in mel_np:
            n_out += len(ap.inv_mel_spectrogram(m))
        dt = time.perf_counter() - t0
        line["e2e_dropin"] = {"value": n_out / SR / dt, "unit": "audio-s/s", "ms_per_utterance": 1e3 * dt / len(mel_np),
                              "utterances": len(mel_np),
                              "api": "AudioProcessor.inv_mel_spectrogram(np.ndarray [80, T]) -> np.ndarray, one call per "
                                     "utterance, host arrays in and out, np.random phases drawn on the host as the "
                                     "reference does (utils/audio.py:164-172, 183)"}
        # BASELINE configs[0]: one 6 s linear spectrogram, inv_spectrogram, device resident
        lay1 = ap.layout(n_frames=[T_FRAMES])
        S1 = torch.rand((T_FRAMES, F), device=dev)
        out1 = ap.inv_spectrogram_batch(S1, lay1)
        torch.cuda.synchronize()
        a1, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a1.record()
        for _ in range(10):
            out1 = ap.inv_spectrogram_batch(S1, lay1, out=out1)
        b1.record()
        torch.cuda.synchronize()
        line["latency_single_ms"] = a1.elapsed_time(b1) / 10

    # ---- optional final gather of the waveforms (north_star: "an optional final NCCL gather") ----
    if args.gather and world > 1:
        from your_voice_tts_b200.sharding import gather_packed
        lens = torch.full((B,), L_OUT, dtype=torch.int64, device=dev)
        for _ in range(3):                                  # communicator set-up and allocator warm-up, untimed
            gather_packed(wav_out, lens)
        times = []
        for _ in range(7):
            barrier()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            blocks, totals, per_rank = gather_packed(wav_out, lens)
            g1.record()
            torch.cuda.synchronize()
            times.append(g0.elapsed_time(g1))
        t_g = torch.tensor(sorted(times), device=dev, dtype=torch.float64)
        dist.all_reduce(t_g, op=dist.ReduceOp.MAX)          # element-wise max over ranks of the sorted samples
        med = float(t_g[len(times) // 2].item())
        nbytes = int(blocks.numel()) * 4
        line["gather"] = {"ms": med, "ms_min": float(t_g[0].item()), "ms_max": float(t_g[-1].item()), "samples": len(times),
                          "bytes_received_per_rank": nbytes,
                          "GBps_per_rank": nbytes / (med * 1e-3) / 1e9, "backend": "nccl",
                          "what": "sharding.gather_packed: every rank receives every rank's packed waveforms "
                                  "(2 small all_gathers for sizes + 1 all_gather_into_tensor)"}

    # ---- CPU baseline (rank 0, N = 1 only): bounded sample of the same workload ----
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = host_cores()
        v1, dt1 = cpu_reference_run(1, 1)            # one utterance on one host thread (the reference's own use)
        v, dt = cpu_reference_run(cores, cores)
        line["cpu_baseline"] = {"value": v, "unit": "audio-s/s", "cores": cores, "kind": "port",
                                "single_thread_value": v1, "single_thread_s_per_utterance": dt1,
                                "sample": "%d utterances (one per host thread) x inv_mel_spectrogram %d iters, %.1f s wall" % (cores, ITERS, dt)}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
