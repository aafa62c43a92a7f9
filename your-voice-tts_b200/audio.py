"""Drop-in ``AudioProcessor`` (reference: utils/audio.py:11-255) whose arithmetic runs in libttsa_b200.so.

Same constructor keywords (the ``audio`` block of config.json:5-25), same public attributes (mutable between
calls, as tests/test_audio.py:32-35 does), same method names and array conventions: single utterance, numpy in,
numpy out, spectrograms as ``[D, T]``.  Every method that touches signal data launches this repo's CUDA kernels
through the C ABI (include/ttsa.h); PyTorch only owns device memory and the stream.  There is no CPU path: without
the compiled library or without a B200 the compute methods raise.

Additive API (not in the reference): ``*_batch`` methods taking/returning packed frame-major CUDA tensors,
``init_angles=`` / ``seed=`` for reproducible Griffin-Lim phases, ``return_sc=`` for the per-iteration spectral
convergence, ``device=`` and ``verbose=`` constructor keywords.
"""
import ctypes
from collections import OrderedDict

import numpy as np

from . import _lib as L

__all__ = ["AudioProcessor", "BatchLayout", "HostPipeline"]

_PLAN_CACHE = OrderedDict()
_PLAN_CACHE_MAX = 32


def _torch():
    import torch
    return torch


class _Plan(object):
    """Owns one ttsa_plan handle (immutable per (config, device))."""

    def __init__(self, cfg_tuple, device_index):
        lib = L.load()
        c = L.TtsaConfig()
        (c.sample_rate, c.num_mels, c.num_freq, c.n_fft, c.hop_length, c.win_length, c.signal_norm, c.symmetric_norm,
         c.clip_norm, c.griffin_lim_iters, c.min_level_db, c.ref_level_db, c.power, c.preemphasis, c.max_norm,
         c.mel_fmin, c.mel_fmax) = cfg_tuple
        self.cfg = c
        self.device_index = device_index
        h = ctypes.c_void_p()
        L.check(lib.ttsa_plan_create(ctypes.byref(c), device_index, ctypes.byref(h)))
        self.handle = h
        self.lib = lib

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.ttsa_plan_destroy(self.handle)
        except Exception:
            pass


class BatchLayout(object):
    """Packed layout of a batch of utterances (ttsa_batch): frame counts, waveform lengths and offsets."""

    def __init__(self, plan, n_frames=None, wav_lengths=None, frame_stride=0):
        assert (n_frames is None) != (wav_lengths is None)
        self.plan = plan
        lib = plan.lib
        src = n_frames if n_frames is not None else wav_lengths
        arr = np.ascontiguousarray(np.asarray(src, dtype=np.int32).reshape(-1))
        self.n_utts = int(arr.shape[0])
        h = ctypes.c_void_p()
        fn = lib.ttsa_batch_from_frames_strided if n_frames is not None else lib.ttsa_batch_from_wav_lengths_strided
        L.check(fn(plan.handle, arr.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), self.n_utts,
                   ctypes.c_int64(int(frame_stride)), ctypes.byref(h)))
        self.frame_stride = int(frame_stride)
        self.handle = h
        self.total_frames = int(lib.ttsa_batch_total_frames(h))
        self.total_samples = int(lib.ttsa_batch_total_samples(h))
        fo = np.zeros(self.n_utts + 1, dtype=np.int64)
        wo = np.zeros(self.n_utts + 1, dtype=np.int64)
        wl = np.zeros(self.n_utts, dtype=np.int32)
        L.check(lib.ttsa_batch_offsets(h, fo.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                                       wo.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                                       wl.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))))
        self.frame_off, self.wav_off, self.wav_len = fo, wo, wl
        # samples a packed waveform buffer must hold: up to the last utterance's last sample (total_samples also counts
        # the padding that keeps every utterance 16-byte aligned)
        self.need_samples = int(wo[self.n_utts - 1] + wl[self.n_utts - 1]) if self.n_utts > 0 else 0
        self.n_frames = (np.asarray(src, dtype=np.int64).reshape(-1) if n_frames is not None
                         else 1 + wl.astype(np.int64) // plan.cfg.hop_length)

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.plan.lib.ttsa_batch_destroy(self.handle)
        except Exception:
            pass

    def split_wav(self, packed):
        """Views of the packed waveform buffer, one per utterance."""
        return [packed[int(self.wav_off[u]):int(self.wav_off[u]) + int(self.wav_len[u])] for u in range(self.n_utts)]

    def split_frames(self, packed):
        return [packed[int(self.frame_off[u]):int(self.frame_off[u]) + int(self.n_frames[u])] for u in range(self.n_utts)]


class HostPipeline(object):
    """Host-to-host batched inv_mel_spectrogram with the copies hidden behind the compute of neighbouring batches.

    submit(mel_host, wav_host, seed) enqueues: pinned-host mel -> device (upload stream), mel -> linear ->
    Griffin-Lim -> de-emphasis (compute stream), device waveform -> pinned host (download stream).  Two device buffer
    sets alternate, so the upload of batch i+1 and the download of batch i-1 overlap the kernels of batch i.  `wav_host` is valid after the returned
    event (or drain())."""

    def __init__(self, ap, layout, graph=False):
        """graph=True replays the kernels of a batch (mel -> linear, initial synthesis, the Griffin-Lim iterations,
        de-emphasis) as one CUDA graph per buffer set instead of ~65 individual launches; the Philox seed of the initial
        phases is then the one given at the first submit() of each buffer set."""
        torch = _torch()
        self.ap, self.layout = ap, layout
        dev = ap._dev()
        self.comp, self.copy, self.h2d = (torch.cuda.Stream(device=dev) for _ in range(3))
        n = max(1, layout.total_samples)
        self.mel_dev = [torch.empty((layout.total_frames, ap.num_mels), dtype=torch.float32, device=dev) for _ in range(2)]
        self.wav_dev = [torch.zeros((n,), dtype=torch.float32, device=dev) for _ in range(2)]
        self.lin_dev = torch.empty((max(1, layout.total_frames), ap.num_freq), dtype=torch.float32, device=dev)
        ws_bytes = int(layout.plan.lib.ttsa_griffin_lim_workspace_bytes(layout.plan.handle, layout.handle))
        self.ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
        self.comp_done = [torch.cuda.Event() for _ in range(2)]
        self.copy_done = [torch.cuda.Event() for _ in range(2)]
        self.h2d_done = [torch.cuda.Event() for _ in range(2)]
        self.count = 0
        self.use_graph = bool(graph)
        self.graphs = [None, None]
        self.graph_seeds = [None, None]
        # the buffers were allocated (and zero-filled) on the caller's stream but are only ever used on the three side
        # streams: order every side stream after the allocation, and tell the caching allocator about the other users
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(dev))
        for s_ in (self.comp, self.copy, self.h2d):
            s_.wait_event(ready)
            for t_ in self.mel_dev + self.wav_dev + [self.lin_dev, self.ws]:
                t_.record_stream(s_)

    def __del__(self):
        try:
            self.drain()                                   # no side-stream work may outlive the buffers
        except Exception:
            pass

    def _compute(self, b, seed, init_angles=None):
        """mel -> |S|^power -> Griffin-Lim -> de-emphasis of buffer set b on the current stream."""
        ap, lay = self.ap, self.layout
        plan = lay.plan
        st = ap._stream()
        L.check(plan.lib.ttsa_mel_to_linear(plan.handle, lay.handle, ap._ptr(self.mel_dev[b]), L.MEL_IN_NORM_DB,
                                            ap._ptr(self.lin_dev), L.MEL_OUT_POWER, st))
        L.check(plan.lib.ttsa_griffin_lim(plan.handle, lay.handle, ap._ptr(self.lin_dev), L.SPEC_MAGNITUDE,
                                          int(ap.griffin_lim_iters), None if init_angles is None else ap._ptr(init_angles),
                                          ctypes.c_uint64(int(seed) & (2 ** 64 - 1)),
                                          L.GL_DEEMPHASIS if ap.preemphasis != 0 else 0, ap._ptr(self.wav_dev[b]), None,
                                          ap._ptr(self.ws), self.ws.numel(), st))

    def submit(self, mel_host, wav_host, seed=0, init_angles=None, strict_seed=False):
        """init_angles: optional device tensor [sum_T, num_freq] of initial phases (radians) instead of the counter RNG
        (read while the batch computes: keep it alive and unchanged until the returned event).  With graph=True the
        seed / init_angles of a buffer set are those of its first submit(); strict_seed=True raises when a later
        submit() to that set passes a different seed instead of silently replaying the captured one."""
        torch = _torch()
        b = self.count & 1
        if self.use_graph and self.graphs[b] is None:
            with torch.cuda.stream(self.comp):
                self._compute(b, seed, init_angles)             # warm-up outside capture
            self.comp.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=self.comp):
                self._compute(b, seed, init_angles)
            self.graphs[b] = g
            self.graph_seeds[b] = int(seed)
        elif self.use_graph and strict_seed and int(seed) != self.graph_seeds[b]:
            raise ValueError("HostPipeline(graph=True): buffer set %d replays the graph captured with seed %d, not %d"
                             % (b, self.graph_seeds[b], int(seed)))
        with torch.cuda.stream(self.h2d):
            if self.count >= 2:
                self.h2d.wait_event(self.comp_done[b])         # batch i-2 has consumed this mel buffer
            self.mel_dev[b].copy_(mel_host, non_blocking=True)  # overlaps the kernels of batch i-1
            self.h2d_done[b].record(self.h2d)
        with torch.cuda.stream(self.comp):
            if self.count >= 2:
                self.comp.wait_event(self.copy_done[b])        # the waveform buffer of batch i-2 has been drained
            self.comp.wait_event(self.h2d_done[b])
            if self.use_graph:
                self.graphs[b].replay()
            else:
                self._compute(b, seed, init_angles)
            self.comp_done[b].record(self.comp)
        with torch.cuda.stream(self.copy):
            self.copy.wait_event(self.comp_done[b])
            wav_host.copy_(self.wav_dev[b], non_blocking=True)
            self.copy_done[b].record(self.copy)
        self.count += 1
        return self.copy_done[b]

    def drain(self):
        self.h2d.synchronize()
        self.comp.synchronize()
        self.copy.synchronize()


class AsyncAudioLogger(object):
    """Training-time audio samples without stalling the step (train.py:225-232, 379-384 run a CPU Griffin-Lim inside the
    training loop).  submit() snapshots the model output that is already on the device, runs the inversion on a side
    stream and copies the waveform to pinned host memory; the training stream never waits.  poll() hands back what has
    finished, typically one or two steps later:

        logger.submit("TrainAudio", mel_output[0], step=current_step)            # [T, num_mels] CUDA tensor
        for tag, step, wav in logger.poll():
            tb_logger.tb_train_audios(step, {tag: wav}, c.audio["sample_rate"])
    """

    def __init__(self, ap):
        torch = _torch()
        self.ap = ap
        self.stream = torch.cuda.Stream(device=ap._dev())
        self.pending = []

    def submit(self, tag, spec_td, step=0, kind="mel", n_frames=None, seed=None, init_angles=None):
        """spec_td: [T, num_mels] (kind "mel", inv_mel_spectrogram) or [T, num_freq] (kind "linear", inv_spectrogram),
        normalised as the model emits it; n_frames trims the padding of a batch element."""
        torch = _torch()
        ap = self.ap
        if kind not in ("mel", "linear"):
            raise ValueError("kind must be 'mel' or 'linear'")
        T = int(spec_td.shape[0] if n_frames is None else n_frames)
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(spec_td.device))
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)                       # the producer of spec_td has finished; nobody waits for us
            snap = spec_td[:T].detach().to(torch.float32).contiguous().clone()
            spec_td.record_stream(self.stream)
            lay = ap.layout(n_frames=[T])
            sd = int(step) if seed is None else int(seed)
            if kind == "mel":
                wav = ap.inv_mel_spectrogram_batch(snap, lay, init_angles=init_angles, seed=sd)
            else:
                wav = ap.inv_spectrogram_batch(snap, lay, init_angles=init_angles, seed=sd)
            n = int(lay.wav_len[0])
            host = torch.empty((n,), dtype=torch.float32).pin_memory()
            host.copy_(wav[:n], non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.stream)
        self.pending.append((tag, int(step), host, done, wav, snap, lay))
        return done

    def poll(self):
        """Finished items as (tag, step, numpy waveform), oldest first; never blocks."""
        out = []
        while self.pending and self.pending[0][3].query():
            tag, step, host, _, _, _, _ = self.pending.pop(0)
            out.append((tag, step, host.numpy()))
        return out

    def flush(self):
        """Wait for everything submitted and return it (end of an epoch / of training)."""
        self.stream.synchronize()
        return self.poll()


class AudioProcessor(object):
    def __init__(self,
                 sample_rate=None,
                 num_mels=None,
                 min_level_db=None,
                 frame_shift_ms=None,
                 frame_length_ms=None,
                 ref_level_db=None,
                 num_freq=None,
                 power=None,
                 preemphasis=None,
                 signal_norm=None,
                 symmetric_norm=None,
                 max_norm=None,
                 mel_fmin=None,
                 mel_fmax=None,
                 clip_norm=True,
                 griffin_lim_iters=None,
                 do_trim_silence=False,
                 device=None,
                 verbose=True,
                 **kwargs):
        if verbose:
            print(" > Setting up Audio Processor...")
        # utils/audio.py:34-51
        self.sample_rate = sample_rate
        self.num_mels = num_mels
        self.min_level_db = min_level_db
        self.frame_shift_ms = frame_shift_ms
        self.frame_length_ms = frame_length_ms
        self.ref_level_db = ref_level_db
        self.num_freq = num_freq
        self.power = power
        self.preemphasis = preemphasis
        self.griffin_lim_iters = griffin_lim_iters
        self.signal_norm = signal_norm
        self.symmetric_norm = symmetric_norm
        self.mel_fmin = 0 if mel_fmin is None else mel_fmin
        self.mel_fmax = mel_fmax
        self.max_norm = 1.0 if max_norm is None else float(max_norm)
        self.clip_norm = clip_norm
        self.do_trim_silence = do_trim_silence
        self.n_fft, self.hop_length, self.win_length = self._stft_parameters()
        if verbose:
            members = vars(self)
            for key, value in members.items():
                print(" | > {}:{}".format(key, value))
        self._device = device
        self._batch_cache = OrderedDict()
        # extension (not a reference field, so it is set after the member print-out): > 0 switches every inversion to
        # fast Griffin-Lim with that momentum; 0 keeps the reference's algorithm
        self.griffin_lim_momentum = float(kwargs.get("griffin_lim_momentum", 0.0))
        # extension: True = initial Griffin-Lim phases from the device RNG when none are injected (the reference draws
        # 2*pi*np.random.rand(F, T) on the host, utils/audio.py:183, which dominates the latency of a single call)
        self.device_phases = bool(kwargs.get("device_phases", False))
        self._phase_seed = 0

    # ------------------------------------------------------------------------------------------ plumbing
    def _stft_parameters(self):
        """utils/audio.py:114-119"""
        n_fft = (self.num_freq - 1) * 2
        hop_length = int(self.frame_shift_ms / 1000.0 * self.sample_rate)
        win_length = int(self.frame_length_ms / 1000.0 * self.sample_rate)
        return n_fft, hop_length, win_length

    def _cfg_tuple(self):
        # parameters are read at call time: the reference's tests mutate them on a live object
        n_fft = (self.num_freq - 1) * 2
        return (int(self.sample_rate), int(self.num_mels), int(self.num_freq), int(n_fft), int(self.hop_length),
                int(self.win_length), int(bool(self.signal_norm)), int(bool(self.symmetric_norm)),
                int(bool(self.clip_norm)), int(self.griffin_lim_iters or 0), float(self.min_level_db),
                float(self.ref_level_db), float(self.power), float(self.preemphasis or 0.0), float(self.max_norm),
                float(self.mel_fmin or 0.0), float(self.mel_fmax) if self.mel_fmax is not None else -1.0)

    def _device_index(self):
        torch = _torch()
        if not torch.cuda.is_available():
            raise RuntimeError("AudioProcessor needs a CUDA device (B200, sm_100a): the audio hot path has no CPU "
                               "implementation in this package")
        dev = self._device
        if dev is None:
            return torch.cuda.current_device()
        dev = torch.device(dev)
        return dev.index if dev.index is not None else torch.cuda.current_device()

    def _plan(self, host_only=False):
        if self.mel_fmax is not None:
            assert self.mel_fmax <= self.sample_rate // 2          # utils/audio.py:70-71
        key = (self._cfg_tuple(), -1 if host_only else self._device_index())
        plan = _PLAN_CACHE.get(key)
        if plan is None:
            plan = _Plan(*key)
            _PLAN_CACHE[key] = plan
            while len(_PLAN_CACHE) > _PLAN_CACHE_MAX:
                _PLAN_CACHE.popitem(last=False)
        else:
            _PLAN_CACHE.move_to_end(key)
        return plan

    def layout(self, n_frames=None, wav_lengths=None, frame_stride=0):
        """BatchLayout for utterances given by frame counts (spectrogram inputs) or by waveform lengths.
        frame_stride > 0: the spectrogram tensors are padded [B, frame_stride, D] blocks instead of packed rows."""
        plan = self._plan()
        src = n_frames if n_frames is not None else wav_lengths
        key = (id(plan), n_frames is not None, int(frame_stride), tuple(int(v) for v in np.asarray(src).reshape(-1)))
        lay = self._batch_cache.get(key)
        if lay is None:
            lay = BatchLayout(plan, n_frames=n_frames, wav_lengths=wav_lengths, frame_stride=frame_stride)
            self._batch_cache[key] = lay
            while len(self._batch_cache) > 16:
                self._batch_cache.popitem(last=False)
        return lay

    def _dev(self):
        return _torch().device("cuda", self._device_index())

    @staticmethod
    def _stream():
        return ctypes.c_void_p(_torch().cuda.current_stream().cuda_stream)

    @staticmethod
    def _ptr(t):
        if t is None:
            return ctypes.c_void_p(0)
        # the C ABI takes plain dense buffers: refuse strided views instead of reading them as if they were dense
        if not t.is_cuda or not t.is_contiguous():
            raise ValueError("expected a contiguous CUDA tensor, got %s strides %s on %s"
                             % (tuple(t.shape), tuple(t.stride()), t.device))
        torch = _torch()
        if t.dtype not in (torch.float32, torch.uint8, torch.int16, torch.int32, torch.int64):
            raise ValueError("expected a float32 tensor, got %s" % t.dtype)
        return ctypes.c_void_p(t.data_ptr())

    def _to_dev(self, x, dtype=None):
        torch = _torch()
        if isinstance(x, torch.Tensor):
            t = x.to(device=self._dev(), dtype=dtype or torch.float32)
        else:
            t = torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float32))).to(self._dev())
        return t.contiguous()

    def _transpose(self, t):
        """[R, C] -> [C, R] with the library's own kernel (layout shim for the reference's [D, T] arrays)."""
        torch = _torch()
        rows, cols = t.shape
        out = torch.empty((cols, rows), dtype=torch.float32, device=t.device)
        L.check(L.load().ttsa_transpose(self._ptr(t), self._ptr(out), rows, cols, self._stream()))
        return out

    @staticmethod
    def _is_tensor(x):
        torch = _torch()
        return isinstance(x, torch.Tensor)

    def _ret(self, t, like):
        """Return numpy for numpy callers (the reference's contract), tensors for tensor callers."""
        if self._is_tensor(like):
            return t
        return t.cpu().numpy()

    # ------------------------------------------------------------------------------------------ batched API
    @staticmethod
    def _need(t, n, what):
        """The C ABI takes bare pointers: the sizes the kernels will touch are checked here, on the host side."""
        if t is not None and int(t.numel()) < int(n):
            raise ValueError("%s holds %d elements, the batch layout needs %d" % (what, int(t.numel()), int(n)))

    def features_batch(self, wav_packed, layout, want_linear=True, want_mel=True, preemphasis=None, lin_out=None,
                       mel_out=None):
        """spectrogram() and melspectrogram() of every utterance in one pass (utils/audio.py:138-152).

        wav_packed: CUDA float32 [layout.total_samples]; returns (linear [sum_T, num_freq] | None, mel [sum_T, num_mels] | None).
        """
        torch = _torch()
        plan = layout.plan
        lin, mel = lin_out, mel_out
        if lin is None and want_linear:
            lin = torch.empty((layout.total_frames, self.num_freq), dtype=torch.float32, device=wav_packed.device)
        if mel is None and want_mel:
            mel = torch.empty((layout.total_frames, self.num_mels), dtype=torch.float32, device=wav_packed.device)
        pre = (self.preemphasis != 0) if preemphasis is None else bool(preemphasis)
        self._need(wav_packed, layout.need_samples, "wav_packed")
        self._need(lin, layout.total_frames * self.num_freq, "linear output")
        self._need(mel, layout.total_frames * self.num_mels, "mel output")
        L.check(plan.lib.ttsa_stft_features(plan.handle, layout.handle, self._ptr(wav_packed), self._ptr(lin),
                                            self._ptr(mel), L.FEAT_PREEMPHASIS if pre else 0, self._stream()))
        return lin, mel

    def stft_batch(self, wav_packed, layout):
        torch = _torch()
        plan = layout.plan
        out = torch.empty((layout.total_frames, self.num_freq, 2), dtype=torch.float32, device=wav_packed.device)
        self._need(wav_packed, layout.need_samples, "wav_packed")
        L.check(plan.lib.ttsa_stft(plan.handle, layout.handle, self._ptr(wav_packed), self._ptr(out), self._stream()))
        return out

    def istft_batch(self, stft_packed, layout):
        torch = _torch()
        plan = layout.plan
        out = torch.zeros((max(1, layout.total_samples),), dtype=torch.float32, device=stft_packed.device)
        self._need(stft_packed, layout.total_frames * self.num_freq * 2, "stft_packed")
        L.check(plan.lib.ttsa_istft(plan.handle, layout.handle, self._ptr(stft_packed), self._ptr(out), self._stream()))
        return out

    def griffin_lim_batch(self, spec_packed, layout, spec_kind=L.SPEC_MAGNITUDE, init_angles=None, seed=0,
                          deemphasis=False, return_sc=False, iters=None, out=None, workspace=None, momentum=None):
        """_griffin_lim over a packed batch (utils/audio.py:182-189).  Returns the packed waveform buffer
        (and, with return_sc, the [iters, B] spectral-convergence log).
        momentum (default: self.griffin_lim_momentum, 0 = the reference's algorithm) > 0 selects fast Griffin-Lim
        as in librosa >= 0.7 griffinlim(momentum=...): not in the reference, opt-in, changes the result."""
        torch = _torch()
        plan = layout.plan
        iters = int(self.griffin_lim_iters if iters is None else iters)
        dev = spec_packed.device
        if out is None:
            out = torch.zeros((max(1, layout.total_samples),), dtype=torch.float32, device=dev)
        ws_bytes = int(plan.lib.ttsa_griffin_lim_workspace_bytes(plan.handle, layout.handle))
        if workspace is None:
            workspace = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
        sc = torch.empty((max(1, iters), layout.n_utts, 2), dtype=torch.float32, device=dev) if return_sc else None
        momentum = float(getattr(self, "griffin_lim_momentum", 0.0) if momentum is None else momentum)
        self._need(spec_packed, layout.total_frames * self.num_freq, "spec_packed")
        self._need(init_angles, layout.total_frames * self.num_freq, "init_angles")
        self._need(out, layout.need_samples, "waveform output")
        self._need(workspace, ws_bytes, "workspace")
        L.check(plan.lib.ttsa_griffin_lim_fast(plan.handle, layout.handle, self._ptr(spec_packed), int(spec_kind), iters,
                                               self._ptr(init_angles), ctypes.c_uint64(int(seed) & (2 ** 64 - 1)),
                                               L.GL_DEEMPHASIS if deemphasis else 0, momentum, self._ptr(out),
                                               self._ptr(sc), self._ptr(workspace), workspace.numel(), self._stream()))
        if return_sc:
            sc_val = torch.sqrt(sc[:iters, :, 0] / sc[:iters, :, 1].clamp_min(1e-30))
            return out, sc_val
        return out

    def mel_to_linear_batch(self, mel_packed, layout, in_kind=L.MEL_IN_AMPLITUDE, out_kind=L.MEL_OUT_PLAIN):
        torch = _torch()
        plan = layout.plan
        out = torch.empty((layout.total_frames, self.num_freq), dtype=torch.float32, device=mel_packed.device)
        self._need(mel_packed, layout.total_frames * self.num_mels, "mel_packed")
        L.check(plan.lib.ttsa_mel_to_linear(plan.handle, layout.handle, self._ptr(mel_packed), in_kind,
                                            self._ptr(out), out_kind, self._stream()))
        return out

    def linear_to_mel_batch(self, lin_packed, layout, in_kind=L.MEL_IN_AMPLITUDE, out_kind=L.MEL_OUT_PLAIN):
        torch = _torch()
        plan = layout.plan
        out = torch.empty((layout.total_frames, self.num_mels), dtype=torch.float32, device=lin_packed.device)
        self._need(lin_packed, layout.total_frames * self.num_freq, "lin_packed")
        L.check(plan.lib.ttsa_linear_to_mel(plan.handle, layout.handle, self._ptr(lin_packed), in_kind,
                                            self._ptr(out), out_kind, self._stream()))
        return out

    def inv_spectrogram_batch(self, spec_packed, layout, init_angles=None, seed=0, return_sc=False, out=None,
                              workspace=None):
        """inv_spectrogram for a packed batch of normalised linear spectrograms [sum_T, num_freq]."""
        return self.griffin_lim_batch(spec_packed, layout, L.SPEC_NORM_DB, init_angles, seed,
                                      deemphasis=self.preemphasis != 0, return_sc=return_sc, out=out,
                                      workspace=workspace)

    def inv_mel_spectrogram_batch(self, mel_packed, layout, init_angles=None, seed=0, return_sc=False, out=None,
                                  workspace=None):
        """inv_mel_spectrogram for a packed batch of normalised mel spectrograms [sum_T, num_mels]."""
        S = self.mel_to_linear_batch(mel_packed, layout, L.MEL_IN_NORM_DB, L.MEL_OUT_POWER)
        return self.griffin_lim_batch(S, layout, L.SPEC_MAGNITUDE, init_angles, seed,
                                      deemphasis=self.preemphasis != 0, return_sc=return_sc, out=out,
                                      workspace=workspace)

    def collate_features(self, wavs, outputs_per_step=1):
        """The feature half of MyDataset.collate_fn (datasets/TTSDataset.py:191-217 with utils/data.py:25-45) on the
        GPU: ONE STFT pass yields mel and linear for the whole batch, written straight into zero-padded
        [B, T_pad, D] tensors (T_pad = longest + 1 zero frame, rounded up to a multiple of outputs_per_step).

        wavs: list of 1-D arrays / tensors (already ordered by the caller).  Returns
        (linear [B, T_pad, num_freq], mel [B, T_pad, num_mels], mel_lengths [B] (= frames + 1), stop_targets [B, T_pad])."""
        torch = _torch()
        dev = self._dev()
        lens = [int(len(w)) for w in wavs]
        T = [1 + n // self.hop_length for n in lens]
        max_len = max(T) + 1                                          # zero-frame (utils/data.py:26)
        rem = max_len % outputs_per_step
        pad_len = max_len + (outputs_per_step - rem) if rem > 0 else max_len
        lay = self.layout(wav_lengths=lens, frame_stride=pad_len)
        if all(self._is_tensor(w) and w.is_cuda for w in wavs):
            packed = torch.zeros((max(1, lay.total_samples),), dtype=torch.float32, device=dev)
            for u, w in enumerate(wavs):                                  # device-to-device, no host round trip
                packed[int(lay.wav_off[u]):int(lay.wav_off[u]) + lens[u]].copy_(w.reshape(-1).to(torch.float32))
        else:
            # host inputs: pack once into one pinned buffer, ONE host-to-device copy for the whole batch
            host = torch.zeros((max(1, lay.total_samples),), dtype=torch.float32).pin_memory()
            hv = host.numpy()
            for u, w in enumerate(wavs):
                hv[int(lay.wav_off[u]):int(lay.wav_off[u]) + lens[u]] = (w.detach().cpu().numpy() if self._is_tensor(w)
                                                                          else np.asarray(w, dtype=np.float32)).reshape(-1)
            packed = host.to(dev, non_blocking=True)
        B = len(wavs)
        linear = torch.zeros((B, pad_len, self.num_freq), dtype=torch.float32, device=dev)
        mel = torch.zeros((B, pad_len, self.num_mels), dtype=torch.float32, device=dev)
        self.features_batch(packed, lay, lin_out=linear.view(B * pad_len, self.num_freq),
                            mel_out=mel.view(B * pad_len, self.num_mels))
        mel_lengths = torch.tensor([t + 1 for t in T], dtype=torch.long)
        # stop targets: zeros for the real frames, ones from the zero-frame on (utils/data.py:34-45)
        stop_max = max(T) + 1
        srem = stop_max % outputs_per_step
        stop_len = stop_max + (outputs_per_step - srem) if srem > 0 else stop_max
        stop_targets = torch.from_numpy((np.arange(stop_len)[None, :] >= np.asarray(T)[:, None]).astype(np.float32))
        return linear, mel, mel_lengths, stop_targets

    def inv_mel_spectrogram_padded(self, mel_btd, n_frames, init_angles=None, seed=0):
        """inv_mel_spectrogram on a model output that stays on the device: mel_btd [B, T_max, num_mels] (the layout
        models/tacotron2.py:62-73 returns), n_frames[u] valid frames per utterance.  Returns the list of waveforms."""
        torch = _torch()
        mel_btd = self._to_dev(mel_btd)
        B, t_max, _ = mel_btd.shape
        lay = self.layout(n_frames=[int(t) for t in n_frames], frame_stride=t_max)
        out = self.inv_mel_spectrogram_batch(mel_btd.reshape(B * t_max, self.num_mels), lay, init_angles=init_angles,
                                             seed=seed)
        return lay.split_wav(out)

    # ------------------------------------------------------------------------------------------ waveform post-processing
    def wav_peaks_batch(self, wav_packed, layout, lens=None):
        """max |wav_u| per utterance: the peak save_wav normalises by (utils/audio.py:57).  -> [B] float32"""
        torch = _torch()
        plan = layout.plan
        out = torch.empty((max(1, layout.n_utts),), dtype=torch.float32, device=wav_packed.device)
        L.check(plan.lib.ttsa_wav_peaks(plan.handle, layout.handle, self._ptr(wav_packed), self._ptr(lens),
                                        self._ptr(out), self._stream()))
        return out[:layout.n_utts]

    def find_endpoint_batch(self, wav_packed, layout, threshold_db=-40, min_silence_sec=0.8):
        """find_endpoint of every utterance (utils/audio.py:203-210).  -> [B] int32 on the device"""
        torch = _torch()
        plan = layout.plan
        out = torch.empty((max(1, layout.n_utts),), dtype=torch.int32, device=wav_packed.device)
        L.check(plan.lib.ttsa_find_endpoint(plan.handle, layout.handle, self._ptr(wav_packed), float(threshold_db),
                                            float(min_silence_sec), self._ptr(out), self._stream()))
        return out[:layout.n_utts]

    def pcm16_batch(self, wav_packed, layout, lens=None, joint_peak=False, gap_samples=0, float32_arith=False):
        """save_wav's int16 conversion for a packed batch (utils/audio.py:56-58), optionally trimmed to `lens`
        (device int32 [B], e.g. find_endpoint_batch) and laid out as the server concatenates sentences:
        every utterance followed by `gap_samples` zeros, ONE peak for the whole output (joint_peak;
        server/synthesizer.py:157-161).  -> (int16 tensor [capacity], offsets int64 [B+1] on the device;
        offsets[B] = samples written)."""
        torch = _torch()
        plan = layout.plan
        dev = wav_packed.device
        B = layout.n_utts
        cap = int(np.sum(layout.wav_len)) + int(gap_samples) * B
        out = torch.zeros((max(1, cap),), dtype=torch.int16, device=dev)
        off = torch.zeros((B + 1,), dtype=torch.int64, device=dev)
        ws = torch.empty((int(plan.lib.ttsa_pcm16_workspace_bytes(plan.handle, layout.handle)),), dtype=torch.uint8, device=dev)
        flags = (L.PCM_JOINT_PEAK if joint_peak else 0) | (L.PCM_F32_ARITH if float32_arith else 0)
        L.check(plan.lib.ttsa_wav_to_pcm16(plan.handle, layout.handle, self._ptr(wav_packed), self._ptr(lens), flags,
                                           int(gap_samples), self._ptr(off), self._ptr(out), cap, self._ptr(ws),
                                           ws.numel(), self._stream()))
        return out, off

    def wav_file_bytes(self, pcm16):
        """RIFF/WAVE container around mono int16 samples at self.sample_rate -- byte-identical to what
        scipy.io.wavfile.write(path, sample_rate, int16_array) puts in the file (utils/audio.py:58)."""
        import struct
        data = np.ascontiguousarray(np.asarray(pcm16, dtype="<i2")).tobytes()
        sr = int(self.sample_rate)
        hdr = b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVE"
        hdr += b"fmt " + struct.pack("<IHHIIHH", 16, 1, 1, sr, sr * 2, 2, 16)
        hdr += b"data" + struct.pack("<I", len(data))
        return hdr + data

    def sentences_to_wav_bytes(self, mels, init_angles=None, seed=0, gap_samples=10000):
        """The audio half of Synthesizer.tts (server/synthesizer.py:133-162) for a Tacotron2 model without WaveRNN:
        every sentence's postnet output [T_i, num_mels] (CUDA tensors, as the model leaves them) goes through ONE
        batched inv_mel_spectrogram instead of a per-sentence CPU loop, `gap_samples` zeros follow each sentence, and
        the concatenation is peak-normalised to int16 and wrapped as a WAV file.  Returns bytes."""
        torch = _torch()
        dev = self._dev()
        Ts = [int(m.shape[0]) for m in mels]
        lay = self.layout(n_frames=Ts)
        packed = torch.cat([self._to_dev(m) for m in mels], dim=0).contiguous()
        wav = self.inv_mel_spectrogram_batch(packed, lay, init_angles=init_angles, seed=seed)
        pcm, off = self.pcm16_batch(wav, lay, joint_peak=True, gap_samples=gap_samples)
        n = int(off[-1].item())
        return self.wav_file_bytes(pcm[:n].cpu().numpy())

    def save_wav(self, wav, path):
        """utils/audio.py:56-58.  A CUDA waveform is normalised and converted on the device and only the int16
        samples cross to the host; arrays and lists take the reference's own numpy/scipy route."""
        if self._is_tensor(wav) and wav.is_cuda:
            t = wav.reshape(-1).to(_torch().float32).contiguous()
            lay = self.layout(wav_lengths=[int(t.numel())])
            buf = _torch().zeros((max(1, lay.total_samples),), dtype=_torch().float32, device=t.device)
            buf[:t.numel()].copy_(t)
            pcm, off = self.pcm16_batch(buf, lay)
            data = self.wav_file_bytes(pcm[:t.numel()].cpu().numpy())
            if hasattr(path, "write"):
                path.write(data)
            else:
                with open(path, "wb") as f:
                    f.write(data)
            return
        from scipy.io import wavfile
        wav = np.asarray(wav)
        wav_norm = wav * (32767 / max(0.01, np.max(np.abs(wav))))
        wavfile.write(path, self.sample_rate, wav_norm.astype(np.int16))

    def _build_mel_basis(self):
        """utils/audio.py:68-77 (librosa.filters.mel semantics), float64 [num_mels, num_freq]."""
        plan = self._plan(host_only=True)
        out = np.empty((self.num_mels, self.num_freq), dtype=np.float64)
        L.check(plan.lib.ttsa_plan_mel_basis(plan.handle, out.ctypes.data_as(ctypes.POINTER(ctypes.c_double))))
        return out

    def _inv_mel_basis(self):
        plan = self._plan(host_only=True)
        out = np.empty((self.num_freq, self.num_mels), dtype=np.float64)
        L.check(plan.lib.ttsa_plan_inv_mel_basis(plan.handle, out.ctypes.data_as(ctypes.POINTER(ctypes.c_double))))
        return out

    def _dt_in(self, x):
        """[D, T] array/tensor -> packed frame-major CUDA tensor [T, D] + layout."""
        t = self._to_dev(x)
        if t.dim() != 2:
            raise ValueError("expected a [D, T] spectrogram, got shape %s" % (tuple(t.shape),))
        tt = self._transpose(t)
        return tt, self.layout(n_frames=[tt.shape[0]])

    def _linear_to_mel(self, spectrogram):
        """utils/audio.py:60-62: np.dot(mel_basis, S); [num_freq, T] -> [num_mels, T]"""
        tt, lay = self._dt_in(spectrogram)
        out = self.linear_to_mel_batch(tt, lay)
        return self._ret(self._transpose(out), spectrogram)

    def _mel_to_linear(self, mel_spec):
        """utils/audio.py:64-66: max(1e-10, pinv(mel_basis) @ mel); [num_mels, T] -> [num_freq, T]"""
        tt, lay = self._dt_in(mel_spec)
        out = self.mel_to_linear_batch(tt, lay)
        return self._ret(self._transpose(out), mel_spec)

    def _pointwise(self, op, x):
        t = self._to_dev(x)
        torch = _torch()
        out = torch.empty_like(t)
        plan = self._plan()
        L.check(plan.lib.ttsa_pointwise(plan.handle, op, self._ptr(t), self._ptr(out), t.numel(), self._stream()))
        return self._ret(out, x)

    def _normalize(self, S):
        """utils/audio.py:79-94"""
        return self._pointwise(L.PW_NORMALIZE, S)

    def _denormalize(self, S):
        """utils/audio.py:96-112"""
        return self._pointwise(L.PW_DENORMALIZE, S)

    def _amp_to_db(self, x):
        """utils/audio.py:121-123"""
        return self._pointwise(L.PW_AMP_TO_DB, x)

    def _db_to_amp(self, x):
        """utils/audio.py:125-126"""
        if np.isscalar(x):          # find_endpoint passes a python scalar (utils/audio.py:206)
            return np.power(10.0, x * 0.05)
        return self._pointwise(L.PW_DB_TO_AMP, x)

    def _wav_in(self, x):
        t = self._to_dev(x).reshape(-1)
        return t, self.layout(wav_lengths=[t.shape[0]])

    def apply_preemphasis(self, x):
        """utils/audio.py:128-131"""
        if self.preemphasis == 0:
            raise RuntimeError(" !! Preemphasis is applied with factor 0.0. ")
        t, lay = self._wav_in(x)
        torch = _torch()
        out = torch.empty_like(t)
        L.check(lay.plan.lib.ttsa_preemphasis(lay.plan.handle, lay.handle, self._ptr(t), self._ptr(out), self._stream()))
        return self._ret(out, x)

    def apply_inv_preemphasis(self, x):
        """utils/audio.py:133-136"""
        if self.preemphasis == 0:
            raise RuntimeError(" !! Preemphasis is applied with factor 0.0. ")
        t, lay = self._wav_in(x)
        torch = _torch()
        out = torch.empty_like(t)
        lib = lay.plan.lib
        ws = torch.empty((int(lib.ttsa_deemphasis_workspace_bytes(lay.plan.handle, lay.handle)),), dtype=torch.uint8,
                         device=t.device)
        L.check(lib.ttsa_deemphasis(lay.plan.handle, lay.handle, self._ptr(t), self._ptr(out), self._ptr(ws),
                                    ws.numel(), self._stream()))
        return self._ret(out, x)

    def spectrogram(self, y):
        """utils/audio.py:138-144; [L] -> [num_freq, T]"""
        t, lay = self._wav_in(y)
        lin, _ = self.features_batch(t, lay, want_linear=True, want_mel=False)
        return self._ret(self._transpose(lin), y)

    def melspectrogram(self, y):
        """utils/audio.py:146-152; [L] -> [num_mels, T]"""
        t, lay = self._wav_in(y)
        _, mel = self.features_batch(t, lay, want_linear=False, want_mel=True)
        return self._ret(self._transpose(mel), y)

    def _host_angles(self, shape_dt, init_angles):
        """Initial phases as a packed [T, D] CUDA tensor.  Without injected phases the reference's own draw is
        reproduced: 2*pi*np.random.rand(*S.shape) on the [D, T] array (utils/audio.py:183), consuming numpy's
        global RNG exactly as the reference does."""
        if init_angles is None:
            if self.device_phases:
                # opt-in: let the kernel draw the phases (counter RNG keyed by a seed taken from numpy's global RNG, so
                # np.random.seed still makes runs repeatable) instead of generating and uploading F x T floats
                self._phase_seed = int(np.random.randint(0, 2 ** 31 - 1))
                return None
            init_angles = 2.0 * np.pi * np.random.rand(*shape_dt)
        a = self._to_dev(np.asarray(init_angles, dtype=np.float32) if not self._is_tensor(init_angles) else init_angles)
        return self._transpose(a)

    def inv_spectrogram(self, spectrogram, init_angles=None, return_sc=False):
        """utils/audio.py:154-162; normalised [num_freq, T] -> waveform [hop*(T-1)]"""
        tt, lay = self._dt_in(spectrogram)
        ang = self._host_angles(tuple(tt.shape[::-1]), init_angles)
        out = self.inv_spectrogram_batch(tt, lay, init_angles=ang, seed=self._phase_seed, return_sc=return_sc)
        if return_sc:
            return self._ret(out[0][:lay.wav_len[0]], spectrogram), self._ret(out[1][:, 0], spectrogram)
        return self._ret(out[:lay.wav_len[0]], spectrogram)

    def inv_mel_spectrogram(self, mel_spectrogram, init_angles=None, return_sc=False):
        """utils/audio.py:164-172; normalised [num_mels, T] -> waveform [hop*(T-1)]"""
        tt, lay = self._dt_in(mel_spectrogram)
        ang = self._host_angles((self.num_freq, tt.shape[0]), init_angles)
        out = self.inv_mel_spectrogram_batch(tt, lay, init_angles=ang, seed=self._phase_seed, return_sc=return_sc)
        if return_sc:
            return self._ret(out[0][:lay.wav_len[0]], mel_spectrogram), self._ret(out[1][:, 0], mel_spectrogram)
        return self._ret(out[:lay.wav_len[0]], mel_spectrogram)

    def out_linear_to_mel(self, linear_spec):
        """utils/audio.py:174-180; normalised [num_freq, T] -> normalised [num_mels, T]"""
        tt, lay = self._dt_in(linear_spec)
        out = self.linear_to_mel_batch(tt, lay, L.MEL_IN_NORM_DB, L.MEL_OUT_NORM_DB)
        return self._ret(self._transpose(out), linear_spec)

    def _griffin_lim(self, S, init_angles=None, return_sc=False):
        """utils/audio.py:182-189; magnitude [num_freq, T] -> waveform"""
        tt, lay = self._dt_in(S)
        ang = self._host_angles(tuple(tt.shape[::-1]), init_angles)
        out = self.griffin_lim_batch(tt, lay, L.SPEC_MAGNITUDE, init_angles=ang, seed=self._phase_seed, return_sc=return_sc)
        if return_sc:
            return self._ret(out[0][:lay.wav_len[0]], S), self._ret(out[1][:, 0], S)
        return self._ret(out[:lay.wav_len[0]], S)

    def _stft(self, y):
        """utils/audio.py:191-197; [L] -> complex64 [num_freq, T]"""
        t, lay = self._wav_in(y)
        torch = _torch()
        D = self.stft_batch(t, lay)                                   # [T, F, 2]
        re = self._transpose(D[..., 0].contiguous())
        im = self._transpose(D[..., 1].contiguous())
        out = torch.complex(re, im)
        return out if self._is_tensor(y) else out.cpu().numpy()

    def _istft(self, y):
        """utils/audio.py:199-201; complex [num_freq, T] -> waveform [hop*(T-1)]"""
        torch = _torch()
        if self._is_tensor(y):
            re, im = self._to_dev(y.real), self._to_dev(y.imag)
        else:
            y_np = np.asarray(y)
            re, im = self._to_dev(np.ascontiguousarray(y_np.real)), self._to_dev(np.ascontiguousarray(y_np.imag))
        D = torch.stack((self._transpose(re), self._transpose(im)), dim=-1).contiguous()   # [T, F, 2]
        lay = self.layout(n_frames=[D.shape[0]])
        out = self.istft_batch(D, lay)
        return self._ret(out[:lay.wav_len[0]], y)

    # ------------------------------------------------------------------------------------------ host helpers
    # (file I/O and O(L) scalar post-processing: outside the kernel scope, kept for drop-in completeness)
    def find_endpoint(self, wav, threshold_db=-40, min_silence_sec=0.8):
        """utils/audio.py:203-210"""
        if self._is_tensor(wav) and wav.is_cuda:
            t = wav.reshape(-1).to(_torch().float32).contiguous()
            lay = self.layout(wav_lengths=[int(t.numel())])
            buf = _torch().zeros((max(1, lay.total_samples),), dtype=_torch().float32, device=t.device)
            buf[:t.numel()].copy_(t)
            return int(self.find_endpoint_batch(buf, lay, threshold_db, min_silence_sec)[0].item())
        window_length = int(self.sample_rate * min_silence_sec)
        hop_length = int(window_length / 4)
        threshold = np.power(10.0, threshold_db * 0.05)
        for x in range(hop_length, len(wav) - window_length, hop_length):
            if np.max(wav[x:x + window_length]) < threshold:
                return x + hop_length
        return len(wav)

    def trim_silence(self, wav):
        """utils/audio.py:212-217: 0.1 s margin, then librosa.effects.trim(top_db=40, frame_length=1024,
        hop_length=256) restated (centred RMS frames, reference = max RMS)."""
        margin = int(self.sample_rate * 0.1)
        wav = np.asarray(wav)[margin:-margin]
        frame_length, hop = 1024, 256
        ypad = np.pad(wav, frame_length // 2, mode="reflect")
        n = 1 + (len(ypad) - frame_length) // hop
        idx = np.arange(frame_length)[None, :] + hop * np.arange(n)[:, None]
        mse = np.mean(ypad[idx] ** 2, axis=1)
        db = 10.0 * np.log10(np.maximum(1e-10, mse)) - 10.0 * np.log10(np.maximum(1e-10, mse.max()))
        nz = np.flatnonzero(db > -40)
        if nz.size == 0:
            return wav[0:0]
        start, end = int(nz[0] * hop), min(len(wav), int((nz[-1] + 1) * hop))
        return wav[start:end]

    @staticmethod
    def mulaw_encode(wav, qc):
        """utils/audio.py:219-226"""
        mu = 2 ** qc - 1
        signal = np.sign(wav) * np.log(1 + mu * np.abs(wav)) / np.log(1. + mu)
        signal = (signal + 1) / 2 * mu + 0.5
        return np.floor(signal,)

    @staticmethod
    def mulaw_decode(wav, qc):
        """utils/audio.py:228-233"""
        mu = 2 ** qc - 1
        x = np.sign(wav) / mu * ((1 + mu) ** np.abs(wav) - 1)
        return x

    def load_wav(self, filename, sr=None):
        """utils/audio.py:235-246 (soundfile is optional here; PCM wav files are read with scipy)."""
        try:
            import soundfile as sf
            x, file_sr = sf.read(filename)
        except ImportError:
            from scipy.io import wavfile
            file_sr, data = wavfile.read(filename)
            if data.dtype.kind == "i":
                x = data.astype(np.float64) / float(2 ** (8 * data.dtype.itemsize - 1))
            elif data.dtype.kind == "u":
                x = (data.astype(np.float64) - 128.0) / 128.0
            else:
                x = data.astype(np.float64)
        if sr is not None and sr != file_sr:
            # the reference resamples through librosa.load (resampy, kaiser_best); librosa is not a dependency here, so a
            # polyphase resampler with a Kaiser window stands in for it (same rate and length, not bit-identical samples)
            from math import gcd
            from scipy.signal import resample_poly
            if x.ndim > 1:
                x = x.mean(axis=1)                                 # librosa.load(mono=True)
            g = gcd(int(sr), int(file_sr))
            x = resample_poly(x, int(sr) // g, int(file_sr) // g, window=("kaiser", 14.769656459379492))
            file_sr = sr
        if self.do_trim_silence:
            try:
                x = self.trim_silence(x)
            except ValueError:
                print(f' [!] File cannot be trimmed for silence - {filename}')
        assert self.sample_rate == file_sr, "%s vs %s" % (self.sample_rate, file_sr)
        return x

    def encode_16bits(self, x):
        return np.clip(x * 2**15, -2**15, 2**15 - 1).astype(np.int16)

    def quantize(self, x, bits):
        return (x + 1.) * (2**bits - 1) / 2

    def dequantize(self, x, bits):
        return 2 * x / (2**bits - 1) - 1
