"""Per-rank cost of a band-sharded frame, emulated on ONE GPU: for G in --ranks, split the frame into G
coverage-balanced row bands (swegl_b200.sharding.balanced_bands, as bench.py's sharded_frame does) and time every
band's kernel chain on its own (CUDA events around the captured graph, L2 flushed between frames).  max over the
bands = what the slowest rank of a G-GPU run spends before the gather.  With --cull both (default) it is done with
and without the sort-first band culling, and the culled frame is compared with the full frame.

    python tools/band_probe.py --workload sphere1000_8k --ranks 1,2,4,8
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="sphere1000_8k")
    ap.add_argument("--ranks", default="1,2,4,8")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--stages", action="store_true", help="also per-stage times of every band (timing mode, direct launches)")
    args = ap.parse_args()
    import torch
    from swegl_b200 import Renderer, configs, sharding
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream()            # not the legacy stream (handle 0 = "the context's own stream" for the ABI)
    torch.cuda.set_stream(stream)
    scene, vps, screen, cfg = configs.build(args.workload)
    vp = vps[0]
    nodes = scene.node_matrices()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    out = {"workload": args.workload, "ranks": {}}
    rs = {}
    for pol in (0, 1):
        r = Renderer(0, stream=stream.cuda_stream)
        r.set_band_culling(pol)
        r.upload_scene(scene)
        r.set_screen(*screen)
        rs[pol] = r
    r = rs[0]
    full = np.zeros((screen[1], screen[0]), np.uint32)
    r.begin_frame(scene, nodes)
    r.render(vp, full)
    cov = ((full >> 24) != 0).sum(axis=1).astype(np.float64) + 0.05 * screen[0]

    def time_band(r, band):
        vp.band = band
        d = vp.desc()
        r.begin_frame(scene, nodes)
        r.render_device(d, stats=True)                  # sizes the pools
        for _ in range(3):
            r.begin_frame(scene, nodes)
            r.render_device(d, stats=False)
        torch.cuda.synchronize()
        ms = []
        for _ in range(args.steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r.begin_frame(scene, nodes)
            r.render_device(d, stats=False)
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        stages = None
        if args.stages:
            r.set_timing(True)
            acc = {}
            for _ in range(5):
                r.begin_frame(scene, nodes)
                st = r.render_device(d, stats=True)
                for k in ("ms_vertex", "ms_setup", "ms_raster", "ms_fragment", "ms_post"):
                    acc[k] = acc.get(k, 0.0) + getattr(st, k) / 5
            r.set_timing(False)
            stages = {k: round(v, 4) for k, v in acc.items()}
        vp.band = (0, 0)
        return float(np.median(ms)), stages

    for G in [int(x) for x in args.ranks.split(",")]:
        bands = [(0, vp.h)] if G == 1 else sharding.balanced_bands(cov.tolist(), G)
        row = {"bands": bands}
        for pol, key in ((0, "no_cull"), (1, "cull")):
            if G == 1 and pol == 1:
                continue
            r = rs[pol]
            per, stg, counts = [], [], []
            for b in bands:
                ms, st = time_band(r, (0, 0) if G == 1 else b)
                per.append(round(ms, 4)); stg.append(st)
                if pol == 1:
                    vp.band = b
                    r.begin_frame(scene, nodes); r.render_device(vp.desc(), stats=True)
                    counts.append(r.cull_counts())
                    vp.band = (0, 0)
            row[key] = {"ms_per_band": per, "ms_max": max(per)}
            if args.stages:
                row[key]["stages"] = stg
            if counts:
                row[key]["marked_frac"] = [round(c["marked"] / max(1, c["clusters"]), 3) for c in counts]
                row[key]["live_frac"] = [round(c["live"] / max(1, c["clusters"]), 3) for c in counts]
                row[key]["vertex_blocks_frac"] = [round(c["vertex_blocks_needed"] / max(1, c["vertex_blocks"]), 3) for c in counts]
        if G > 1:
            r = rs[1]
            got = np.zeros_like(full)
            for b in bands:
                vp.band = b
                r.begin_frame(scene, nodes)
                r.render(vp, got)
            vp.band = (0, 0)
            row["culled_bands_equal_full_frame"] = bool((got == full).all())
        out["ranks"][str(G)] = row
    base = out["ranks"].get("1", {}).get("no_cull", {}).get("ms_max")
    if base:
        for G, row in out["ranks"].items():
            for key in ("no_cull", "cull"):
                if key in row:
                    row[key]["speedup_vs_1"] = round(base / row[key]["ms_max"], 3)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
