"""Per-stage device times of a few workloads for ONE build of the library (select it with SWEGL_B200_LIB=...), exact and
fast shading side by side; one JSON line.  Used to A/B kernel variants on the GPU box in a single gpurun call:

    for v in "" minb6 ...; do SWEGL_B200_LIB=swegl_b200/libswegl_b200_$v.so python tools/ab_probe.py --tag $v; done

Stage times come from the context's timing mode (CUDA events between the kernels of a frame, direct launches);
`graph_ms` is the whole captured chain (render_device without stats) between two events with a 256 MiB L2 flush
between frames; `pipe_fps` is swegl_b200.FramePipeline (4 contexts, frames round robin).
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", default="")
    ap.add_argument("--workloads", default="truck_4k_dof,brainstem_4k_dof,truck_1080,sphere1000_8k")
    ap.add_argument("--frames", type=int, default=30)
    ap.add_argument("--pipe", action="store_true")
    ap.add_argument("--modes", default="exact,fast")
    args = ap.parse_args()
    import torch
    from swegl_b200 import Renderer, configs, _abi
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    out = {"tag": args.tag, "lib": os.environ.get("SWEGL_B200_LIB", "default"), "workloads": {}}
    for name in args.workloads.split(","):
        scene, vps, screen, cfg = configs.build(name)
        r = Renderer(0, stream=stream.cuda_stream)
        r.upload_scene(scene)
        r.set_screen(*screen)
        nodes = scene.node_matrices()
        row = {}
        for mode, key in ((_abi.SHADING_EXACT, "exact"), (_abi.SHADING_FAST, "fast")):
            if key not in args.modes.split(","):
                continue
            r.set_shading(mode)
            r.begin_frame(scene, nodes)
            for vp in vps:
                r.render_device(vp, stats=True)
            r.set_timing(True)
            acc = {}
            for i in range(args.frames + 2):
                r.begin_frame(scene, nodes)
                for vp in vps:
                    st = r.render_device(vp, stats=True)
                    if i >= 2:
                        for k in ("ms_vertex", "ms_setup", "ms_raster", "ms_fragment", "ms_post", "ms_total"):
                            acc[k] = acc.get(k, 0.0) + getattr(st, k) / args.frames
            r.set_timing(False)
            descs = [vp.desc() for vp in vps]
            for _ in range(3):
                r.begin_frame(scene, nodes)
                for d in descs:
                    r.render_device(d, stats=False)
            torch.cuda.synchronize()
            ms = []
            for _ in range(args.frames):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                r.begin_frame(scene, nodes)
                for d in descs:
                    r.render_device(d, stats=False)
                e1.record()
                torch.cuda.synchronize()
                ms.append(e0.elapsed_time(e1))
            r.synchronize()
            acc = {k: round(v, 4) for k, v in acc.items()}
            acc["graph_ms"] = round(float(np.median(ms)), 4)
            acc["covered"] = int(st.n_covered)
            row[key] = acc
        r.close()
        if args.pipe and not name.startswith("sphere"):
            from swegl_b200.pipeline import FramePipeline
            pipe = FramePipeline(0, 4)
            try:
                pipe.upload_scene(scene)
                pipe.set_screen(*screen)
                n = 400
                row["pipe_fps"] = round(n / (pipe.measure(scene, vps, n, warmup=12) / 1e3), 1)
            finally:
                pipe.close()
        out["workloads"][name] = row
    print(json.dumps(out))


if __name__ == "__main__":
    main()
