"""world_size-2 gloo tests of the sort-first partition + frame assembly logic (CPU, no GPU needed)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from leisure_software_renderer_b200 import sortfirst


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, H, W, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # a synthetic "frame": pixel value encodes its position, rows counted from the top
        full = (np.arange(H * W * 4, dtype=np.uint32) % 251).astype(np.uint8).reshape(H, W, 4)
        y0, y1 = sortfirst.band_of_rank(H, world, rank)
        band = torch.from_numpy(full[y0:y1].copy().reshape(-1))
        sizes = sortfirst.band_sizes(H, W, world, 4)
        assert sizes[rank] == band.numel()
        frame = sortfirst.gather_bands(band, sizes, world, rank)
        cams = sortfirst.cameras_of_rank(5, world, rank)
        local = torch.full((8,), float(rank))
        frames = sortfirst.gather_frames(local, world, rank)
        if rank == 0:
            ok = np.array_equal(frame.numpy().reshape(H, W, 4), full) and all(float(f[0]) == i for i, f in enumerate(frames))
            np.save(os.path.join(out_dir, "ok.npy"), np.array([ok, len(cams)]))
        else:
            assert frame is None and frames is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("H", [1080, 37])
def test_band_gather_world2(tmp_path, H):
    mp.spawn(_worker, args=(2, _free_port(), H, 64, str(tmp_path)), nprocs=2, join=True)
    ok = np.load(tmp_path / "ok.npy")
    assert ok[0] == 1 and ok[1] == 3  # cameras 0, 2, 4 on rank 0


def test_partitions_cover_frame_exactly():
    for H in (1, 15, 16, 17, 1080, 4320):
        for world in (1, 2, 3, 4, 8):
            bands = [sortfirst.band_of_rank(H, world, r) for r in range(world)]
            assert bands[0][0] == 0 and bands[-1][1] == H
            for a, b in zip(bands, bands[1:]):
                assert a[1] == b[0]
            assert all(y0 % 16 == 0 or y0 == H for y0, _ in bands)
    for n in (1, 7, 64):
        for world in (1, 2, 8):
            got = sorted(c for r in range(world) for c in sortfirst.cameras_of_rank(n, world, r))
            assert got == list(range(n))


def test_ownership_layouts_cover_every_row_once():
    for tile_rows in (1, 7, 13, 68, 270):
        for world in (1, 2, 3, 4, 8):
            if tile_rows < world:
                continue
            # interleaved stripes
            for rows_per_stripe in (1, 2, 3):
                count = [0] * tile_rows
                for r in range(world):
                    for ty, o in enumerate(sortfirst.owned_rows(tile_rows, *sortfirst.stripes_of_rank(world, r, rows_per_stripe))):
                        count[ty] += int(o)
                assert count == [1] * tile_rows
            # cost-balanced bands
            rng = np.random.default_rng(tile_rows * 31 + world)
            for cost in (np.ones(tile_rows), rng.random(tile_rows) ** 4, np.concatenate([np.zeros(tile_rows - 1), [5.0]]), np.zeros(tile_rows)):
                cuts = sortfirst.balanced_band_cuts(cost, world)
                assert cuts[0] == 0 and cuts[-1] == tile_rows and len(cuts) == world + 1
                assert all(b > a for a, b in zip(cuts, cuts[1:])), "every rank owns at least one tile row"
                count = [0] * tile_rows
                for r in range(world):
                    for ty, o in enumerate(sortfirst.owned_rows(tile_rows, cuts[r], cuts[r + 1] - cuts[r], 1 << 24)):
                        count[ty] += int(o)
                assert count == [1] * tile_rows
    # balance: with a smooth cost no band exceeds its fair share by more than one row's cost
    cost = np.linspace(1.0, 3.0, 270)
    cuts = sortfirst.balanced_band_cuts(cost, 8)
    shares = [cost[a:b].sum() for a, b in zip(cuts, cuts[1:])]
    assert max(shares) <= cost.sum() / 8 + cost.max()
    # framebuffer rows of the bands tile the frame bottom-up without gaps (ragged last tile row included)
    H = 4320 + 5
    cuts = sortfirst.balanced_band_cuts(np.ones((H + 15) // 16), 8)
    spans = [sortfirst.band_framebuffer_rows(H, cuts, r) for r in range(8)]
    assert spans[0][1] == H and spans[-1][0] == 0
    assert all(spans[r][0] == spans[r + 1][1] for r in range(7))
