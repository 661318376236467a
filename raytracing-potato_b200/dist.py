"""Multi-GPU host logic for the hot path (one process per GPU). The path shards without any data-path exchange:
ray batches are sliced contiguously, frames are split by sample range (every rank renders the full tile rectangle for
its own samples, into raw sums), and the only collective is one sum of the accumulation buffers at frame end — the
reference's tile workers never communicate either (main.rs:61-92). `torch.distributed` is plumbing here."""
from typing import Tuple

import numpy as np


def ray_slice(n: int, rank: int, world: int) -> Tuple[int, int]:
    """contiguous [begin, end) slice of an n-ray batch for `rank`; sizes differ by at most one ray"""
    base, extra = divmod(n, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def sample_range(num_samples: int, rank: int, world: int) -> Tuple[int, int]:
    """[sample_begin, sample_end) of `rank` out of num_samples per pixel. Samples are keyed by (pixel, sample) in the random
    stream, so the union over ranks is exactly the sample set of a single-GPU render."""
    return ray_slice(num_samples, rank, world)


def reduce_frame(acc, group=None):
    """sum the raw accumulation buffers (rgb sums, foreground counts) of all ranks in place: torch tensor on this rank's
    device (NCCL over NVLink on GPUs, gloo on CPU)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    return acc


def finish_frame(acc_rgb: np.ndarray, num_samples: int) -> np.ndarray:
    """main.rs:86-87: colour = sum / num_samples, after the cross-rank sum"""
    return acc_rgb / float(num_samples)
