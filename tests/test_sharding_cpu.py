"""N > 1 host logic on CPU: frame-balanced utterance partition and the optional final waveform gather
(world_size 2, gloo).  The data path itself has no collective -- utterances are independent."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from your_voice_tts_b200.sharding import gather_waveforms, partition_utterances


def test_partition_covers_and_balances():
    rng = np.random.default_rng(0)
    for world in (1, 2, 3, 4, 8):
        for n in (1, 2, 7, 64, 4096):
            T = rng.integers(2, 900, size=n)
            parts = partition_utterances(T, world)
            assert len(parts) == world
            np.testing.assert_array_equal(np.concatenate(parts), np.arange(n))
            if n >= 8 * world:
                loads = np.array([T[p].sum() for p in parts], dtype=np.float64)
                assert loads.max() <= loads.mean() + T.max()          # within one utterance of perfect balance
    parts = partition_utterances([482] * 4096, 8)
    assert [len(p) for p in parts] == [512] * 8                       # BASELINE configs[4]: 4096 utterances over 8 GPUs


def _worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    T = [5, 9, 2, 40, 7]
    part = partition_utterances(T, world)[rank]
    # stand-in waveforms: utterance u -> hop*(T-1) samples with value u (the GPU path is exercised by the -m gpu tests)
    local = [torch.full((275 * (T[u] - 1),), float(u)) for u in part]
    allw = gather_waveforms(local)
    ok = len(allw) == len(T) and all(w.numel() == 275 * (T[u] - 1) and bool((w == float(u)).all()) for u, w in enumerate(allw))
    with open(os.path.join(tmp, "ok%d" % rank), "w") as f:
        f.write("1" if ok else "0")
    dist.destroy_process_group()


def test_gather_waveforms_gloo_world2(tmp_path):
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert [open(tmp_path / ("ok%d" % r)).read() for r in range(2)] == ["1", "1"]
