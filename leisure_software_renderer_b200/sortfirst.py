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


def owned_rows(tile_rows: int, first: int, count: int, stride: int) -> List[bool]:
    """ShsbFrameParams::own_row_* as a mask over tile rows (row 0 = top): owned iff ty >= first and (ty - first) % stride < count;
    count <= 0 owns every row."""
    return [count <= 0 or (ty >= first and (ty - first) % stride < count) for ty in range(tile_rows)]


def stripes_of_rank(world: int, rank: int, rows_per_stripe: int = 2) -> Tuple[int, int, int]:
    """Interleaved stripes: (own_row_first, own_row_count, own_row_stride) of rank r -- balanced by construction, no draw can be dropped."""
    return rank * rows_per_stripe, rows_per_stripe, world * rows_per_stripe


def balanced_band_cuts(row_cost: Sequence[float], world: int) -> List[int]:
    """Cost-balanced contiguous bands: cuts[r] .. cuts[r + 1] are rank r's tile rows, every rank gets at least one row and about
    1 / world of sum(row_cost) (e.g. last frame's shaded pixels + triangles per tile row).  Deterministic, so every rank computes
    the same partition from the same costs."""
    n = len(row_cost)
    if world <= 0 or n < world:
        raise ValueError("need at least one tile row per rank")
    cum = [0.0]
    for c in row_cost:
        cum.append(cum[-1] + max(float(c), 0.0))
    total = cum[-1]
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        k = cuts[-1] + 1
        while k < n - (world - r) and cum[k] < target:   # leave a row for every later rank
            k += 1
        cuts.append(k)
    cuts.append(n)
    return cuts


def band_framebuffer_rows(height: int, cuts: Sequence[int], rank: int, tile: int = TILE) -> Tuple[int, int]:
    """Framebuffer (bottom-origin) pixel rows [a, b) of rank r's band: a band of top-anchored tile rows is one contiguous range of
    render-target memory, so it can be sent / received in place."""
    return max(0, height - cuts[rank + 1] * tile), height - cuts[rank] * tile


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
