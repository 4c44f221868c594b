"""Sort-first multi-GPU partitioning: scene replicated on every rank, each rank renders a band of
rows of the FULL viewport (a scissor, not a smaller viewport: a sub-viewport would change the projection,
viewport.cpp:30-35), finished bands are gathered to rank 0 for single-frame output.  Batches of independent
frames need no collective at all (frame i -> rank i mod world).

Pure host logic over torch.distributed: works with CUDA tensors over NCCL (NVLink) and with CPU tensors over
gloo (tests/test_sharding_gloo.py).
"""


def band_rows(height, world, rank):
    """rows [y0, y1) of a `height`-row viewport owned by `rank` (contiguous, sizes differ by at most 1)."""
    base, rem = divmod(height, world)
    y0 = rank * base + min(rank, rem)
    return y0, y0 + base + (1 if rank < rem else 0)


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


def gather_bands_inplace(frame, height, dist, dst=0):
    """Copy-free gather for the common case: `frame` is every rank's full (height, width) screen tensor (contiguous),
    rank r has rendered rows band_rows(height, world, r) of it.  On return rank `dst`'s tensor holds the whole frame:
    each peer's band is received straight into its place (one NCCL send/recv pair per peer, batched)."""
    world, rank = dist.get_world_size(), dist.get_rank()
    if world == 1:
        return frame
    ops = []
    if rank == dst:
        for r in range(world):
            if r != dst:
                y0, y1 = band_rows(height, world, r)
                ops.append(dist.P2POp(dist.irecv, frame[y0:y1], r))
    else:
        y0, y1 = band_rows(height, world, rank)
        ops.append(dist.P2POp(dist.isend, frame[y0:y1], dst))
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    return frame if rank == dst else None


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
