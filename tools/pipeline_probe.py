"""Throughput of independent frames when R contexts (each with its own stream, pools and screen) render them round
robin on ONE GPU: the latency-bound head of frame i+1 (vertex / mark / set-up / spans) overlaps the fragment and DoF
kernels of frame i.  Device-resident, CUDA events on a main stream that forks to / joins the context streams.

    python tools/pipeline_probe.py --workload truck_4k_dof --depths 1,2,3,4
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="truck_4k_dof")
    ap.add_argument("--depths", default="1,2,3,4")
    ap.add_argument("--steps", type=int, default=240)
    args = ap.parse_args()
    import torch
    from swegl_b200 import configs
    from swegl_b200.pipeline import FramePipeline
    torch.cuda.set_device(0)
    main_stream = torch.cuda.Stream()
    torch.cuda.set_stream(main_stream)
    scene, vps, screen, cfg = configs.build(args.workload)
    out = {"workload": args.workload, "steps": args.steps, "fps": {}, "ms_per_frame": {}, "host_submit_ms_per_frame": {}}
    for depth in [int(x) for x in args.depths.split(",")]:
        pipe = FramePipeline(0, depth)
        pipe.upload_scene(scene)
        pipe.set_screen(*screen)
        ms = pipe.measure(scene, vps, args.steps, warmup=3 * depth)
        out["fps"][str(depth)] = round(1e3 * args.steps / ms, 1)
        out["ms_per_frame"][str(depth)] = round(ms / args.steps, 5)
        out["host_submit_ms_per_frame"][str(depth)] = round(pipe.host_submit_ms / args.steps, 5)
        pipe.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
