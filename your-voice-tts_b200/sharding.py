"""Multi-GPU sharding of the audio hot path: utterances are independent, so a batch is split into contiguous,
frame-balanced shards (one process per GPU) and no collective touches the data path.  The only optional exchange is a
final gather of the synthesised waveforms (NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
import numpy as np

__all__ = ["partition_utterances", "shard_for_rank", "gather_waveforms"]


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


def gather_waveforms(local_wavs, group=None):
    """All-gather variable-length waveforms: every rank receives the rank-ordered list of all waveforms.

    local_wavs: list of 1-D float32 tensors (this rank's shard, device = the backend's device).
    One all_gather of the lengths and one all_gather of a padded [n_local_max, len_max] block.
    """
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if len(local_wavs) > 0:
        dev = local_wavs[0].device
    else:
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    lens = torch.tensor([int(w.numel()) for w in local_wavs], dtype=torch.int64, device=dev)
    meta = torch.tensor([len(local_wavs), int(lens.max()) if len(local_wavs) else 0], dtype=torch.int64, device=dev)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    n_max = max(int(m[0]) for m in metas)
    l_max = max(int(m[1]) for m in metas)
    if n_max == 0:
        return []
    len_block = torch.zeros((n_max,), dtype=torch.int64, device=dev)
    len_block[:len(local_wavs)] = lens
    block = torch.zeros((n_max, max(l_max, 1)), dtype=torch.float32, device=dev)
    for i, w in enumerate(local_wavs):
        block[i, :w.numel()] = w
    len_blocks = [torch.zeros_like(len_block) for _ in range(world)]
    blocks = [torch.zeros_like(block) for _ in range(world)]
    dist.all_gather(len_blocks, len_block, group=group)
    dist.all_gather(blocks, block, group=group)
    out = []
    for r in range(world):
        for i in range(int(metas[r][0])):
            out.append(blocks[r][i, :int(len_blocks[r][i])].clone())
    return out
