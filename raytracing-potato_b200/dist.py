"""Index helpers for hosts that run one process per GPU. The path shards without any data-path exchange: ray batches are
sliced contiguously; frames are split either by ROWS (row_split: each rank renders all samples of its rows, rank 0 gathers
them — bit-identical to one GPU, the split bench.py uses for strong scaling) or by SAMPLE RANGE (sample_range + raw sums +
one sum at frame end). The reference's tile workers never communicate either (main.rs:61-92). A single process can instead
hand all its GPUs to the library (rtp_scene_create_multi + rtp_render, include/rtp.h): that is where the multi-GPU
implementation lives; `torch.distributed` here is plumbing for the per-process variant."""
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


def row_split(rank: int, world: int) -> Tuple[int, int]:
    """(row_offset, row_stride) of `rank` for rtp_render_params: the rows of the tile rectangle are dealt out round-robin, every
    rank renders ALL samples of its rows, so gathering the rows gives the single-GPU frame bit for bit (no sum across ranks)."""
    return rank, world


def rows_of(height: int, rank: int, world: int) -> int:
    """number of rows of a `height`-row rectangle that fall to `rank` under row_split"""
    return (height - rank + world - 1) // world if rank < height else 0


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
