"""Multi-GPU decomposition of the water path.  The path shards with NO data-path
exchange: every vertex is a pure function of (col, row, frame parameters).

  * frame sharding  (animations):  frame f belongs to rank f mod world_size
  * row sharding    (huge grids):  rank g owns rows [g*L/G, (g+1)*L/G); the +-1-row normal stencil
                                   is recomputed locally by the kernel (no halo exchange)

The only collective is an all_gather of the per-frame / per-shard 40-byte digests (and timings) at
the END of a run — kilobytes, never on the timed path.  One process per GPU, torch.distributed
(NCCL on the GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np

from .api import CHECKSUM_DTYPE

MASK64 = (1 << 64) - 1


def frame_shard(n_frames: int, rank: int, world: int) -> list[int]:
    """Global frame indices owned by `rank` (round robin, f mod world == rank)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return list(range(rank, n_frames, world))


def row_shard(grid_len: int, rank: int, world: int) -> tuple[int, int]:
    """[row_begin, row_end) owned by `rank`; contiguous, covers [0, grid_len) exactly, sizes differ by <= 1."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(grid_len, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def combine_checksums(records: np.ndarray) -> dict:
    """Folds digests of disjoint vertex sets (row shards, or frames) into one: sums mod 2^64, min/max."""
    r = np.asarray(records, dtype=CHECKSUM_DTYPE).ravel()

    def s(name):
        return int(np.sum(r[name], dtype=np.uint64)) & MASK64
    return {"position_bits": s("position_bits"), "normal_bits": s("normal_bits"), "shade_bits": s("shade_bits"),
            "min_y": float(r["min_y"].min()), "max_y": float(r["max_y"].max()), "vertices": s("vertices")}


def all_gather_checksums(local: np.ndarray, device=None) -> np.ndarray:
    """all_gather of equally sized per-rank digest arrays -> (world, n) CHECKSUM_DTYPE.  Uses the
    default process group (NCCL needs `device` = this rank's cuda device; gloo works on CPU)."""
    import torch
    import torch.distributed as dist
    local = np.ascontiguousarray(local, dtype=CHECKSUM_DTYPE).ravel()
    t = torch.from_numpy(local.view(np.uint8).copy())
    if device is not None:
        t = t.to(device)
    world = dist.get_world_size()
    out = torch.empty(world * t.numel(), dtype=torch.uint8, device=t.device)   # flat: gloo wants rank-major 1-D
    dist.all_gather_into_tensor(out, t)
    return out.cpu().numpy().view(CHECKSUM_DTYPE).reshape(world, -1)


def max_over_ranks(value: float, device=None) -> float:
    """MAX all_reduce of a scalar (the timing rule: a multi-GPU time is the max over ranks)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
