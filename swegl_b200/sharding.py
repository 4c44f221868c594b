"""Sort-first multi-GPU partitioning: scene replicated on every rank, each rank renders a band of
rows of the FULL viewport (a scissor, not a smaller viewport: a sub-viewport would change the projection,
viewport.cpp:30-35), finished bands are gathered to rank 0 for single-frame output.  Batches of independent
frames need no collective at all (frame i -> rank i mod world).  A multi-viewport frame (config 4) can also be split
one viewport_t per rank (viewports_for_rank / gather_rects_inplace).

Pure host logic over torch.distributed: works with CUDA tensors over NCCL (NVLink) and with CPU tensors over
gloo (tests/test_sharding_gloo.py).
"""


def band_rows(height, world, rank):
    """rows [y0, y1) of a `height`-row viewport owned by `rank` (contiguous, sizes differ by at most 1)."""
    base, rem = divmod(height, world)
    y0 = rank * base + min(rank, rem)
    return y0, y0 + base + (1 if rank < rem else 0)


def balanced_bands(row_cost, world, min_rows=8):
    """Contiguous bands [(y0, y1)] * world whose summed `row_cost` (a sequence, one non-negative number per row --
    e.g. the covered pixels of each row in the previous frame plus a constant for the per-row fixed work) is as equal
    as a prefix-sum split allows.  Every band gets at least `min_rows` rows.  Same result on every rank."""
    cost = [float(c) for c in row_cost]
    height = len(cost)
    if world <= 1:
        return [(0, height)]
    if height < world * min_rows:
        return [band_rows(height, world, r) for r in range(world)]
    total = sum(cost)
    if total <= 0:
        return [band_rows(height, world, r) for r in range(world)]
    cuts, acc, y = [0], 0.0, 0
    for k in range(1, world):
        target = total * k / world
        while y < height and acc + cost[y] <= target:
            acc += cost[y]
            y += 1
        lo = cuts[-1] + min_rows                         # leave room for this band and for the bands after it
        hi = height - (world - k) * min_rows
        cuts.append(min(max(y, lo), hi))
        if cuts[-1] != y:                                # clamped: keep the running sum consistent with the cut
            y = cuts[-1]
            acc = sum(cost[:y])
    cuts.append(height)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def rebalance_bands(bands, times, min_rows=8, damping=1.0):
    """Feedback step of the band balance: `bands` = the contiguous [(y0, y1)] the ranks just rendered, `times` = what each
    took (any unit).  Every band's cost is taken as uniform over its rows (time / rows); the new cuts split that
    piecewise-constant cost profile into equal shares.  Unlike the coverage model this sees everything a rank really
    pays: triangle-dense rows (the poles of a sphere), the halo of the DoF pass, rank 0's extra duty as the assembling
    GPU.  `damping` < 1 moves only part of the way (for noisy timings).  Same result on every rank."""
    world = len(bands)
    if world <= 1:
        return list(bands)
    height = bands[-1][1]
    cost = []
    for (y0, y1), t in zip(bands, times):
        cost += [max(float(t), 1e-9) / max(1, y1 - y0)] * (y1 - y0)
    new = balanced_bands(cost, world, min_rows)
    if damping >= 1.0:
        return new
    cuts = [0]
    for k in range(1, world):
        c = int(round(bands[k][0] + damping * (new[k][0] - bands[k][0])))
        cuts.append(min(max(c, cuts[-1] + min_rows), height - (world - k) * min_rows))
    cuts.append(height)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def interleaved_bands(height, world, rank, band=64):
    """list of [y0, y1) bands for interleaved assignment (load balance for centred objects)."""
    out = []
    for i, y0 in enumerate(range(0, height, band)):
        if i % world == rank:
            out.append((y0, min(height, y0 + band)))
    return out


def frames_for_rank(n_frames, world, rank):
    """frame-parallel batches: frame i -> rank i mod world, no collective."""
    return list(range(rank, n_frames, world))


def viewports_for_rank(n_views, world, rank):
    """config 4 (renderer.hpp:20-34): one viewport_t per GPU.  swegl::render(scene, vp1, vp2, ...) runs
    original_to_world once and then every viewport independently, so viewport v -> rank v mod world needs no exchange
    before the output; ranks beyond the viewport count idle."""
    return list(range(rank, n_views, world))


def gather_rects_inplace(frame, rects, dist, dst=0):
    """Output step of the viewport partition: `frame` is every rank's full (H, W) screen tensor, rank
    v mod world has rendered rectangle rects[v] = (x, y, w, h) of it.  On return rank `dst` holds every
    rectangle.  A rectangle is a strided block, so it travels through one contiguous staging tensor per side
    (send/recv pairs batched in one group).  Peer-memory output (share_screen) needs none of this: the rectangles are
    stored into dst's screen by the kernels that produce them."""
    world, rank = dist.get_world_size(), dist.get_rank()
    if world == 1:
        return frame
    ops, landing = [], []
    for v, (x, y, w, h) in enumerate(rects):
        owner = v % world
        if owner == dst:
            continue
        if rank == dst:
            buf = frame.new_empty((h, w))
            landing.append((buf, (x, y, w, h)))
            ops.append(dist.P2POp(dist.irecv, buf, owner, tag=v))
        elif rank == owner:
            ops.append(dist.P2POp(dist.isend, frame[y:y + h, x:x + w].contiguous(), dst, tag=v))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    for buf, (x, y, w, h) in landing:
        frame[y:y + h, x:x + w] = buf
    return frame if rank == dst else None


def gather_bands_inplace(frame, height, dist, dst=0, bands=None):
    """Copy-free gather for the common case: `frame` is every rank's full (height, width) screen tensor (contiguous),
    rank r has rendered rows bands[r] of it (default: band_rows(height, world, r)).  On return rank `dst`'s tensor
    holds the whole frame: each peer's band is received straight into its place (one NCCL send/recv pair per peer,
    batched)."""
    world, rank = dist.get_world_size(), dist.get_rank()
    if world == 1:
        return frame
    if bands is None:
        bands = [band_rows(height, world, r) for r in range(world)]
    ops = []
    if rank == dst:
        for r in range(world):
            if r != dst:
                y0, y1 = bands[r]
                ops.append(dist.P2POp(dist.irecv, frame[y0:y1], r))
    else:
        y0, y1 = bands[rank]
        ops.append(dist.P2POp(dist.isend, frame[y0:y1], dst))
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    return frame if rank == dst else None


def share_screen(renderer, dist, dst=0, device=None):
    """Peer-memory output: rank `dst` exports its device screen (CUDA IPC handle, broadcast as 64 bytes), every other
    rank imports it and makes it its colour target, so that its band is written straight into dst's frame by the
    last kernel of its frame.  Afterwards a frame is complete on dst once every rank's stream has passed a barrier.
    `renderer` needs export_screen() / import_screen(handle) / set_color_target(ptr) (swegl_b200.Renderer).
    Returns the pointer the rank now renders into (None on dst: its own screen)."""
    import torch
    rank = dist.get_rank()
    h = torch.zeros(64, dtype=torch.uint8, device=device)
    if rank == dst:
        h = torch.tensor(list(renderer.export_screen()), dtype=torch.uint8, device=device)
    dist.broadcast(h, src=dst)
    if rank == dst:
        return None
    ptr = renderer.import_screen(bytes(h.cpu().tolist()))
    renderer.set_color_target(ptr)
    return ptr


def arm_frame_sync(renderer, dist, dst=0):
    """Switch the frame protocol on (after share_screen): from now on a banded view submitted with
    render_device(..., stats=False) is complete in dst's screen when dst's stream has run past it -- flags over peer
    memory instead of a collective (include/swegl_b200.h: swegl_b200_set_frame_sync).  dst's flags are reset first."""
    world, rank = dist.get_world_size(), dist.get_rank()
    if dst != 0:
        raise ValueError("the frame protocol assembles on rank 0")
    if rank == dst:
        renderer.set_frame_sync(0, world)
    dist.barrier()
    if rank != dst:
        renderer.set_frame_sync(rank, world)
    dist.barrier()


def frame_barrier(dist, token):
    """What is left of the gather with share_screen(): every rank's frame kernels (ordered before this call on the
    current stream) have finished -- and with them their stores into dst's screen -- when the all-reduce of the
    one-element `token` tensor completes."""
    dist.all_reduce(token)
    return token


def gather_bands(local_rows, height, width, dist, dst=0):
    """Gather every rank's band (a (rows, width) uint32/int32 tensor for band_rows(height, world, rank)) to `dst`.
    Returns the assembled (height, width) frame on dst, None elsewhere.  Bands may differ by one row, so the
    payload is padded to the largest band and trimmed on arrival.  (General form; gather_bands_inplace avoids
    the staging copies when every rank holds a full-size frame buffer.)"""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    max_rows = -(-height // world)
    send = torch.zeros((max_rows, width), dtype=local_rows.dtype, device=local_rows.device)
    send[: local_rows.shape[0]] = local_rows
    if rank == dst:
        recv = [torch.empty_like(send) for _ in range(world)]
        dist.gather(send, recv, dst=dst)
        frame = torch.empty((height, width), dtype=local_rows.dtype, device=local_rows.device)
        for r in range(world):
            y0, y1 = band_rows(height, world, r)
            frame[y0:y1] = recv[r][: y1 - y0]
        return frame
    dist.gather(send, None, dst=dst)
    return None
