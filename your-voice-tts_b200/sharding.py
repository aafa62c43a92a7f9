"""Multi-GPU sharding of the audio hot path: utterances are independent, so a batch is split into contiguous,
frame-balanced shards (one process per GPU) and no collective touches the data path.  The only optional exchange is a
final gather of the synthesised waveforms (NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
import numpy as np

__all__ = ["partition_utterances", "shard_for_rank", "gather_waveforms", "gather_packed"]


def partition_utterances(n_frames, world_size):
    """Contiguous shards of utterance indices with near-equal total frame counts.

    n_frames: per-utterance frame counts (the cost of every kernel on the path is linear in frames).
    Returns a list of `world_size` index arrays (possibly empty) that concatenate to arange(len(n_frames)).
    """
    n_frames = np.asarray(n_frames, dtype=np.int64).reshape(-1)
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    total = int(n_frames.sum())
    csum = np.concatenate(([0], np.cumsum(n_frames)))
    bounds = [0]
    for r in range(1, world_size):
        target = total * r / world_size
        # first utterance boundary whose prefix sum is closest to the target, never moving backwards
        j = int(np.searchsorted(csum, target, side="left"))
        if j > 0 and abs(csum[j - 1] - target) <= abs(csum[min(j, len(csum) - 1)] - target):
            j -= 1
        bounds.append(min(max(j, bounds[-1]), len(n_frames)))
    bounds.append(len(n_frames))
    return [np.arange(bounds[r], bounds[r + 1], dtype=np.int64) for r in range(world_size)]


def shard_for_rank(n_frames, world_size, rank):
    return partition_utterances(n_frames, world_size)[rank]


def gather_packed(packed, lengths, group=None):
    """All-gather packed waveform buffers: the optional final exchange of the sharded path (the server concatenates the
    sentences of a request in order, server/synthesizer.py:157-161).

    packed:  1-D float32 tensor, this rank's waveforms back to back (the layout the kernels write; padding between
             utterances is allowed and travels as it is).
    lengths: 1-D int64 tensor / list with this rank's per-utterance sample counts.
    Two small all_gathers for the sizes and ONE all_gather_into_tensor for the samples (every rank contributes a block
    padded to the largest shard).  Returns (blocks [world, max_total] float32, totals [world] int64, per-rank length
    tensors); rank r's samples are blocks[r, :totals[r]].
    """
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = packed.device
    lengths = torch.as_tensor(lengths, dtype=torch.int64, device=dev).reshape(-1)
    meta = torch.tensor([int(packed.numel()), int(lengths.numel())], dtype=torch.int64, device=dev)
    metas = torch.zeros((world, 2), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(metas.reshape(-1), meta, group=group)
    metas_h = metas.cpu()
    max_total, max_n = int(metas_h[:, 0].max()), int(metas_h[:, 1].max())
    len_block = torch.zeros((max(max_n, 1),), dtype=torch.int64, device=dev)
    len_block[:lengths.numel()] = lengths
    len_blocks = torch.zeros((world, max(max_n, 1)), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(len_blocks.reshape(-1), len_block, group=group)
    if int(packed.numel()) == max_total:
        mine = packed.reshape(-1)
    else:
        mine = torch.zeros((max(max_total, 1),), dtype=torch.float32, device=dev)
        mine[:packed.numel()] = packed.reshape(-1)
    blocks = torch.empty((world, max(max_total, 1)), dtype=torch.float32, device=dev)
    dist.all_gather_into_tensor(blocks.reshape(-1), mine, group=group)
    per_rank = [len_blocks[r, :int(metas_h[r, 1])] for r in range(world)]
    return blocks, metas_h[:, 0].clone(), per_rank


def gather_waveforms(local_wavs, group=None):
    """All-gather variable-length waveforms: every rank receives the rank-ordered list of all waveforms (views into one
    gathered block, no per-utterance copies).  local_wavs: list of 1-D float32 tensors (this rank's shard)."""
    import torch
    import torch.distributed as dist
    if len(local_wavs) > 0:
        dev = local_wavs[0].device
    else:
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    lens = [int(w.numel()) for w in local_wavs]
    packed = torch.cat([w.reshape(-1) for w in local_wavs]) if local_wavs else torch.zeros((0,), dtype=torch.float32, device=dev)
    blocks, totals, per_rank = gather_packed(packed, lens, group=group)
    out = []
    for r in range(blocks.shape[0]):
        off = 0
        for n in per_rank[r].tolist():
            out.append(blocks[r, off:off + n])
            off += n
    return out
