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
    # cost-balanced bands from the gathered frame's coverage (rank 0 decides, everyone follows)
    cuts = torch.zeros(world + 1, dtype=torch.int64)
    if rank == 0:
        cov = ((frame >> 24) != 0).sum(dim=1).to(torch.float64) + 0.05 * vp.w
        b = sharding.balanced_bands(cov.tolist(), world)
        cuts = torch.tensor([b[0][0]] + [e for _, e in b], dtype=torch.int64)
    dist.broadcast(cuts, src=0)
    bands = [(int(cuts[k]), int(cuts[k + 1])) for k in range(world)]
    vp.band = bands[rank]
    o2 = Oracle().render(scene, vp, screen_wh=screen)
    mine2 = torch.zeros((vp.h, vp.w), dtype=torch.int32)
    mine2[bands[rank][0]:bands[rank][1]] = torch.from_numpy(o2["pixels"][bands[rank][0]:bands[rank][1]].view(np.int32).copy())
    balanced = sharding.gather_bands_inplace(mine2, vp.h, dist, dst=0, bands=bands)
    if rank == 0:
        assert torch.equal(balanced, frame) and bands[0][1] != sharding.band_rows(vp.h, world, 0)[1]
    # peer-memory output handshake (CUDA IPC in production): the handle reaches every rank, dst keeps its own screen
    class FakeRenderer:
        target = None

        def export_screen(self):
            return bytes(range(64))

        def import_screen(self, handle):
            assert handle == bytes(range(64))
            return 0xABC000

        def set_color_target(self, ptr):
            self.target = ptr
    fr = FakeRenderer()
    got = sharding.share_screen(fr, dist, dst=0)
    assert (got, fr.target) == ((None, None) if rank == 0 else (0xABC000, 0xABC000))
    tok = sharding.frame_barrier(dist, torch.ones(1))
    assert float(tok) == world
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


def test_balanced_bands_split_cost_evenly():
    from swegl_b200 import sharding
    h = 2160
    cost = [max(0.0, 1000.0 - abs(y - 1400) * 2.0) + 10.0 for y in range(h)]         # an off-centre blob
    for world in (2, 3, 4, 8):
        b = sharding.balanced_bands(cost, world)
        assert b[0][0] == 0 and b[-1][1] == h and all(b[k][1] == b[k + 1][0] for k in range(world - 1))
        assert all(y1 - y0 >= 8 for y0, y1 in b)
        sums = [sum(cost[y0:y1]) for y0, y1 in b]
        assert max(sums) <= 1.05 * sum(cost) / world + max(cost)
    assert sharding.balanced_bands([0.0] * 100, 4) == [sharding.band_rows(100, 4, r) for r in range(4)]
    assert sharding.balanced_bands([1.0] * 10, 4) == [sharding.band_rows(10, 4, r) for r in range(4)]   # too few rows


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


def test_rebalance_bands_equalises_cost():
    """rebalance_bands: bands whose rows cost different amounts converge to equal per-band time in a few steps"""
    from swegl_b200 import sharding
    height, world = 4320, 8
    # true cost per row: triangle-dense poles + coverage hump in the middle
    row_cost = [3.0 if (y < 300 or y > 4020) else 1.0 + 2.0 * (1 - abs(y - 2160) / 2160.0) for y in range(height)]
    bands = [sharding.band_rows(height, world, r) for r in range(world)]
    def times(b):
        return [sum(row_cost[y0:y1]) for y0, y1 in b]
    t0 = times(bands)
    for _ in range(4):
        bands = sharding.rebalance_bands(bands, times(bands))
        assert bands[0][0] == 0 and bands[-1][1] == height and all(a[1] == b[0] for a, b in zip(bands, bands[1:]))
        assert all(y1 - y0 >= 8 for y0, y1 in bands)
    t = times(bands)
    assert max(t) / (sum(t) / world) < 1.03 < max(t0) / (sum(t0) / world)
    # damping moves part of the way; world 1 is the identity
    half = sharding.rebalance_bands([sharding.band_rows(height, world, r) for r in range(world)], t0, damping=0.5)
    assert half[0][0] == 0 and half[-1][1] == height
    assert sharding.rebalance_bands([(0, height)], [1.0]) == [(0, height)]


def _viewport_worker(rank, world, port, out_path):
    """config 4: one viewport_t per rank (viewport v -> rank v mod world), rectangles gathered to rank 0"""
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle.binding import Oracle
    from swegl_b200 import configs, sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    scene, vps, screen, cfg = configs.build("multiview_1080")
    mine = sharding.viewports_for_rank(len(vps), world, rank)
    px = np.zeros((screen[1], screen[0]), dtype=np.uint32)
    o = Oracle()
    for v in mine:
        o.render(scene, vps[v], screen_wh=screen, pixels=px)
    frame = torch.from_numpy(px.view(np.int32))
    rects = [(vp.x, vp.y, vp.w, vp.h) for vp in vps]
    got = sharding.gather_rects_inplace(frame, rects, dist, dst=0)
    if rank == 0:
        np.save(out_path, got.numpy().view(np.uint32))
    else:
        assert got is None
    dist.destroy_process_group()


def test_viewports_for_rank_cover_every_viewport_once():
    from swegl_b200 import sharding
    for n in (1, 2, 4, 5):
        for world in (1, 2, 3, 4, 8):
            got = sorted(v for r in range(world) for v in sharding.viewports_for_rank(n, world, r))
            assert got == list(range(n))
    assert sharding.viewports_for_rank(4, 8, 5) == []           # more GPUs than viewports: the rest idle


def test_two_rank_viewport_per_rank_and_rect_gather(tmp_path, oracle):
    import torch.multiprocessing as mp
    from swegl_b200 import configs
    out = str(tmp_path / "frame.npy")
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_viewport_worker, args=(2, port, out), nprocs=2, join=True)
    scene, vps, screen, cfg = configs.build("multiview_1080")
    want = np.zeros((screen[1], screen[0]), dtype=np.uint32)
    for vp in vps:
        oracle.render(scene, vp, screen_wh=screen, pixels=want)
    assert (np.load(out) == want).all()


def _protocol_worker(rank, world, port, log_dir):
    """arm_frame_sync: rank 0 resets its flag block BEFORE any other rank arms (their first frame raises done flags in
    rank 0's memory); a rank other than 0 cannot be the assembling one"""
    sys.path.insert(0, ROOT)
    import time
    import torch.distributed as dist
    from swegl_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    class FakeRenderer:
        def set_frame_sync(self, r, w=0):
            with open(os.path.join(log_dir, f"armed_{r}"), "w") as f:
                f.write(f"{time.monotonic_ns()} {w}")
    fr = FakeRenderer()
    if rank == 1:
        time.sleep(0.2)                                  # a late rank must not let the others run ahead of rank 0's reset
    sharding.arm_frame_sync(fr, dist, dst=0)
    try:
        sharding.arm_frame_sync(fr, dist, dst=1)
        raise AssertionError("dst != 0 accepted")
    except ValueError:
        pass
    dist.destroy_process_group()


def test_frame_protocol_arms_rank0_first(tmp_path):
    import torch.multiprocessing as mp
    port = 33500 + (os.getpid() % 2000)
    mp.spawn(_protocol_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    t0, w0 = (int(v) for v in open(tmp_path / "armed_0").read().split())
    t1, w1 = (int(v) for v in open(tmp_path / "armed_1").read().split())
    assert t0 < t1 and w0 == 2 and w1 == 2
