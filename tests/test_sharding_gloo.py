"""CPU, world_size 2 over gloo: the sort-first partition + gather logic (swegl_b200/sharding.py).  Each rank
renders its band of rows with the CPU oracle (the band scissor is part of the viewport descriptor), bands are
gathered to rank 0 and must reassemble to the single-process frame."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle.binding import Oracle
    from swegl_b200 import configs, sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    scene, vps, screen, cfg = configs.build("box_640_close")
    vp = vps[0]
    y0, y1 = sharding.band_rows(vp.h, world, rank)
    vp.band = (y0, y1)
    o = Oracle().render(scene, vp, screen_wh=screen)
    band = torch.from_numpy(o["pixels"][y0:y1].view(np.int32).copy())
    frame = sharding.gather_bands(band, vp.h, vp.w, dist, dst=0)
    # the copy-free variant: every rank holds a full-size frame with only its band rendered
    mine = torch.zeros((vp.h, vp.w), dtype=torch.int32)
    mine[y0:y1] = band
    inplace = sharding.gather_bands_inplace(mine, vp.h, dist, dst=0)
    if rank == 0:
        assert torch.equal(inplace, frame)
    frames = sharding.frames_for_rank(7, world, rank)
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([len(frames)], dtype=torch.int64))
    if rank == 0:
        assert sum(int(c) for c in counts) == 7
        np.save(out_path, frame.numpy().view(np.uint32))
    dist.destroy_process_group()


def test_band_partition_covers_every_row_once():
    from swegl_b200 import sharding
    for h in (1, 7, 480, 2160, 4321):
        for world in (1, 2, 3, 4, 8):
            rows = []
            for r in range(world):
                y0, y1 = sharding.band_rows(h, world, r)
                rows += list(range(y0, y1))
            assert rows == list(range(h))
            inter = sorted(y for r in range(world) for (a, b) in sharding.interleaved_bands(h, world, r) for y in range(a, b))
            assert inter == list(range(h))


def test_two_rank_band_render_and_gather(tmp_path, oracle):
    import torch.multiprocessing as mp
    from swegl_b200 import configs
    out = str(tmp_path / "frame.npy")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    scene, vps, screen, cfg = configs.build("box_640_close")
    full = oracle.render(scene, vps[0], screen_wh=screen)
    got = np.load(out)
    assert (got == full["pixels"]).all()
