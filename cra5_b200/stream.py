"""Multi-GPU frame streaming: hourly ERA5 frames are independent units (reference: test.py:13 loops timestamps, no
temporal context in the model), so N GPUs run N replicas and frame i goes to rank i mod N. No collective touches the
data path; torch.distributed (NCCL on GPUs, gloo in the CPU tests) is used only to agree on timings and to gather
per-rank byte counts."""
from __future__ import annotations

from typing import Iterable, List, Sequence


def shard_frames(n_frames: int, rank: int, world: int) -> List[int]:
    """indices of the frames rank `rank` of `world` processes (rank-strided, like a DistributedSampler without padding)"""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"invalid rank {rank} for world size {world}")
    return list(range(rank, n_frames, world))


def max_over_ranks(seconds: float, device=None) -> float:
    """wall/device time of the slowest rank (multi-GPU numbers are always the max over ranks)"""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(seconds)
    t = torch.tensor([seconds], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_counts(values: Sequence[int], device=None) -> List[List[int]]:
    """per-rank integer lists (e.g. compressed bytes per frame), padded with -1, gathered to every rank"""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [list(values)]
    n = torch.tensor([len(values)], dtype=torch.int64, device=device)
    dist.all_reduce(n, op=dist.ReduceOp.MAX)
    buf = torch.full((int(n.item()),), -1, dtype=torch.int64, device=device)
    buf[: len(values)] = torch.tensor(list(values), dtype=torch.int64, device=device)
    out = [torch.empty_like(buf) for _ in range(dist.get_world_size())]
    dist.all_gather(out, buf)
    return [[int(v) for v in t.tolist() if v >= 0] for t in out]


def stream_frames(codec, frames: Iterable, rank: int, world: int, n_frames: int):
    """compress+decompress this rank's share of `frames` (an indexable of (C,H,W) tensors); yields (index, strings,
    x_hat). `codec` is a cra5_b200.vaeformer.VAEformer living on this rank's GPU."""
    for i in shard_frames(n_frames, rank, world):
        x = frames[i]
        out = codec.compress(x.unsqueeze(0) if x.dim() == 3 else x)
        rec = codec.decompress(out["strings"], out["z_shape"])
        yield i, out["strings"], rec["x_hat"]
