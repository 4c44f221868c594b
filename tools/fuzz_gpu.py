"""Long-run parity fuzz on the GPU box: seeds [a, b) of swegl_b200.configs.fuzz_case through the CUDA path and the C
oracle (depth / coverage / alpha identical, colour within 1 LSB), in three flavours: the small soup of the test suite,
a heavy one (30 primitives x 120 vertices at 1280x720: deep bin lists, pool growth) and the transparency-layer one.
Writes a JSON report; exit status 1 if anything differs.
    python tools/fuzz_gpu.py 0 400 gpurun_out/fuzz.json
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    a, b = int(sys.argv[1]), int(sys.argv[2])
    out_path = sys.argv[3] if len(sys.argv) > 3 else None
    from oracle.binding import Oracle
    from swegl_b200 import Renderer, _abi, configs
    from swegl_b200.scene import Viewport
    r, o = Renderer(0), Oracle()
    bad, n = [], 0
    t0 = time.time()
    worst = 0
    exact = 0

    def check(tag, seed, scene, vp, screen):
        nonlocal n, worst, exact
        print(f"[fuzz] {tag} seed {seed} light {vp.light_mode} tex {vp.tex_mode} post {vp.post_mode} layers {vp.transparency_layers}", file=sys.stderr, flush=True)
        r.upload_scene(scene); r.set_screen(*screen); r.begin_frame(scene)
        px = np.zeros((screen[1], screen[0]), np.uint32)
        z = np.empty((vp.h, vp.w), np.float32)
        st = r.render(vp, px, z)
        ref = o.render(scene, vp, screen_wh=screen)
        dz = int((z.view(np.uint32) != ref["z"].view(np.uint32)).sum())
        d = np.abs(px.view(np.uint8).astype(np.int16) - ref["pixels"].view(np.uint8).astype(np.int16)).reshape(screen[1], screen[0], 4)
        da, dc = int(d[..., 3].max()), int(d.max())
        n += 1
        worst = max(worst, dc)
        exact += int(dc == 0)
        if dz or da or dc > 1:
            bad.append({"flavour": tag, "seed": seed, "depth_words": dz, "alpha_max": da, "channel_max": dc,
                        "pixels_off_by_more_than_1": int((d.max(axis=2) > 1).sum())})

    for seed in range(a, b):
        scene, vp, screen, pose = configs.fuzz_case(seed)
        check("small", seed, scene, vp, screen)
        if seed % 4 == 0:
            scene, vp, screen, pose = configs.fuzz_layers_case(seed)
            check("layers", seed, scene, vp, screen)
        if seed % 8 == 0:
            _, vp0, _, pose = configs.fuzz_case(seed)
            heavy = configs.fuzz_scene(seed, n_prims=30, verts_per_prim=120)
            vp = Viewport(0, 0, 1280, 720, light_mode=vp0.light_mode, tex_mode=vp0.tex_mode, post_mode=vp0.post_mode,
                          focal_distance=vp0.focal_distance, focal_depth=vp0.focal_depth)
            vp.camera.apply(pose)
            check("heavy", seed, heavy, vp, (1280, 720))
    rep = {"seeds": [a, b], "frames_checked": n, "frames_bit_identical_in_colour": exact, "worst_channel_difference": worst,
           "mismatches": bad, "seconds": round(time.time() - t0, 1)}
    s = json.dumps(rep)
    print(s)
    if out_path:
        with open(out_path, "w") as f:
            f.write(s + "\n")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
