"""N > 1 host logic on CPU: world_size 2 over gloo. Each rank takes its slice of a ray batch / its sample range of a frame
(rtp_b200.dist), computes it with the CPU oracle standing in for the device, and the pieces are combined the way
bench.py combines them (gather of slices, all-reduce of raw sums). The result must equal the single-rank result."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from rtp_b200 import _abi as A  # noqa: E402
from rtp_b200 import dist as rdist  # noqa: E402


def test_slices_partition_exactly():
    for n in (0, 1, 7, 16, 2073600, (1 << 24) + 3):
        for world in (1, 2, 3, 4, 8):
            parts = [rdist.ray_slice(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[r][1] == parts[r + 1][0] for r in range(world - 1))
            sizes = [e - b for b, e in parts]
            assert max(sizes) - min(sizes) <= 1
    assert [rdist.sample_range(16, r, 8) for r in range(8)] == [(2 * r, 2 * r + 2) for r in range(8)]
    assert [rdist.sample_range(3, r, 2) for r in range(2)] == [(0, 2), (2, 3)]
    for h in (1, 2, 7, 1080):
        for world in (1, 2, 3, 8):
            assert sum(rdist.rows_of(h, r, world) for r in range(world)) == h
            assert all(rdist.rows_of(h, r, world) == len(range(h)[r::world]) for r in range(world))


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from rtp_b200 import api, scenes

        sc = scenes.bunny_lambert()
        o = oracle.Scene(sc)
        # --- ray batch: contiguous slices, no exchange on the data path; rank 0 gathers only to compare ---------------
        cam = api.Camera(16 / 9, sc.camera.fov, sc.camera.focal_dist, 0.0, sc.camera.transformation)
        rays = np.concatenate([oracle.camera_rays(cam, 96, 54), scenes.incoherent_rays(3001)])
        b, e = rdist.ray_slice(len(rays), rank, world)
        mine = o.hit(rays[b:e])
        parts = [None] * world
        dist.all_gather_object(parts, mine.tobytes())
        # --- frame: sample ranges into raw sums, one all-reduce, divide ------------------------------------------------
        w, h, spp = 48, 27, 5
        sb, se = rdist.sample_range(spp, rank, world)
        rgb, fg, _ = o.render(w, h, spp, seed=9, sample_range=(sb, se), flags=A.RENDER_RAW_SUMS)
        acc = torch.from_numpy(np.concatenate([rgb.reshape(-1), fg.reshape(-1)]))
        rdist.reduce_frame(acc)
        # --- frame: rows dealt out round-robin, every rank renders all samples of its rows, rank 0 gathers: no sum at all ---
        ro, rs = rdist.row_split(rank, world)
        rrgb, rfg, rst = o.render(w, h, spp, seed=9, rows=(ro, rs))
        assert rst.paths == rdist.rows_of(h, rank, world) * w * spp
        rows = [None] * world
        dist.all_gather_object(rows, (rrgb[ro::rs].tobytes(), rfg[ro::rs].tobytes()))
        if rank == 0:
            full_hits = o.hit(rays)
            assert b"".join(parts) == full_hits.tobytes()
            ref, ref_fg, _ = o.render(w, h, spp, seed=9)
            img = rdist.finish_frame(acc[: w * h * 3].numpy().reshape(h, w, 3), spp)
            # the per-pixel sum is re-associated across ranks: a few ulp at most
            assert np.allclose(img, ref, rtol=1e-14, atol=1e-15)
            assert np.array_equal(acc[w * h * 3:].numpy().reshape(h, w) / spp, ref_fg)
            by_rows = np.zeros_like(ref)
            by_rows_fg = np.zeros_like(ref_fg)
            for r in range(world):
                by_rows[r::world] = np.frombuffer(rows[r][0], dtype=np.float64).reshape(-1, w, 3)
                by_rows_fg[r::world] = np.frombuffer(rows[r][1], dtype=np.float64).reshape(-1, w)
            assert by_rows.tobytes() == ref.tobytes() and by_rows_fg.tobytes() == ref_fg.tobytes()  # bit-identical
            open(os.path.join(tmp, "ok"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world_size_2_gloo(tmp_path):
    import socket

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").exists()
