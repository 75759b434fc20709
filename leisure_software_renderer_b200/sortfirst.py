"""Sort-first multi-GPU plumbing: one process per GPU, scene replicated, screen bands
or cameras partitioned across ranks, final frame(s) assembled on a root rank with one gather.

There is no data-path collective inside a frame: screen tiles are independent once triangles are binned.  The
only exchange is the frame assembly (disjoint pixels -> pure concatenation, no reduction), done with
torch.distributed (NCCL over NVLink on GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

TILE = 16


def cameras_of_rank(n_cameras: int, world: int, rank: int) -> List[int]:
    """Camera batch mode: rank r renders cameras {c : c mod world == r}."""
    return [c for c in range(n_cameras) if c % world == rank]


def band_of_rank(height: int, world: int, rank: int, tile: int = TILE) -> Tuple[int, int]:
    """Screen-band mode: contiguous band of tile rows [row0, row1) in TOP-anchored tile rows, balanced to +-1 row.
    Returns the pixel rows [y0, y1) counted from the top of the frame."""
    tile_rows = (height + tile - 1) // tile
    base, extra = divmod(tile_rows, world)
    r0 = rank * base + min(rank, extra)
    r1 = r0 + base + (1 if rank < extra else 0)
    return min(r0 * tile, height), min(r1 * tile, height)


def band_sizes(height: int, width: int, world: int, bytes_per_pixel: int, tile: int = TILE) -> List[int]:
    return [(band_of_rank(height, world, r, tile)[1] - band_of_rank(height, world, r, tile)[0]) * width * bytes_per_pixel for r in range(world)]


def gather_frames(local, world: int, rank: int, root: int = 0):
    """Gathers equally-sized per-rank frame tensors on `root` (camera batch).  Returns a list on root, None elsewhere."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return [local]
    bufs = [torch.empty_like(local) for _ in range(world)] if rank == root else None
    dist.gather(local, bufs, dst=root)
    return bufs


def gather_bands(local_band, sizes: Sequence[int], world: int, rank: int, root: int = 0):
    """Assembles a frame from per-rank bands of different sizes (flat uint8 tensors): root receives every band
    in rank order with point-to-point sends (a gather of ragged messages) and concatenates them top to bottom."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return local_band
    if rank == root:
        parts = []
        for r in range(world):
            if r == root:
                parts.append(local_band)
            else:
                buf = torch.empty(sizes[r], dtype=local_band.dtype, device=local_band.device)
                dist.recv(buf, src=r)
                parts.append(buf)
        return torch.cat(parts)
    dist.send(local_band, dst=root)
    return None
