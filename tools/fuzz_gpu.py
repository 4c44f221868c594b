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

    rc = Renderer(0)
    rc.set_band_culling(1)                              # culling on every banded view

    def check_paths(seed, scene, vp, screen):
        """the other ways into the same frame: culled row bands, the captured-graph path (render_device) and the pipelined
        read-back (render_async/wait) must give the blocking call's frame bit for bit; then a second frame of the same
        uploaded scene with moved nodes, other lights (more than the light block was laid out for) against the oracle"""
        nonlocal n
        rc.upload_scene(scene); rc.set_screen(*screen); rc.begin_frame(scene)
        full = np.zeros((screen[1], screen[0]), np.uint32)
        fz = np.empty((vp.h, vp.w), np.float32)
        rc.render(vp, full, fz)
        nb = 2 + seed % 5
        cuts = sorted({0, vp.h} | {int(vp.h * (k / nb) ** 1.3) for k in range(1, nb)})
        px = np.zeros_like(full); z = np.empty_like(fz)
        for b0, b1 in zip(cuts, cuts[1:]):
            vp.band = (b0, b1)
            rc.render(vp, px, z)
        vp.band = (0, 0)
        if not ((px == full).all() and (z.view(np.uint32) == fz.view(np.uint32)).all()):
            bad.append({"flavour": "culled_bands", "seed": seed, "bands": nb})
        rc.begin_frame(scene)
        rc.render_device(vp, stats=False); rc.synchronize()
        if not (rc.read_screen() == full).all():
            bad.append({"flavour": "render_device", "seed": seed})
        host = [rc.alloc_host((screen[1], screen[0]), np.uint32) for _ in range(2)]
        for h in host:
            h[:] = 0
        tickets = []
        for k in range(2):
            rc.begin_frame(scene)
            tickets.append(rc.render_async(vp.desc(), host[k]))
        for t in tickets:
            rc.wait(t)
        if not ((host[0] == full).all() and (host[1] == full).all()):
            bad.append({"flavour": "render_async", "seed": seed})
        # frame 2: the app moved things (scene.animate / key handlers, test_1.cpp:288-303,378)
        rng = np.random.default_rng(5000 + seed)
        scene.node_translation = (scene.node_translation + rng.uniform(-0.3, 0.3, scene.node_translation.shape)).astype(np.float32)
        scene.node_scale = (scene.node_scale * rng.uniform(0.8, 1.2, scene.node_scale.shape)).astype(np.float32)
        lights = [(float(rng.uniform(-3, 3)), float(rng.uniform(-1, 4)), float(rng.uniform(-4, 2)), float(rng.choice([0.5, 5.0, 80.0])))
                  for _ in range(int(rng.integers(0, 14)))]
        scene.set_lights(float(rng.uniform(0.05, 0.4)), (float(rng.uniform(-1, 1)), -1.0, float(rng.uniform(-1, 1))), 0.6, lights)
        rc.begin_frame(scene)
        px2 = np.zeros_like(full); z2 = np.empty_like(fz)
        rc.render(vp, px2, z2)
        ref = o.render(scene, vp, screen_wh=screen)
        d = np.abs(px2.view(np.uint8).astype(np.int16) - ref["pixels"].view(np.uint8).astype(np.int16))
        if (z2.view(np.uint32) != ref["z"].view(np.uint32)).any() or d.max() > 1:
            bad.append({"flavour": "second_frame_moved_nodes_and_lights", "seed": seed, "channel_max": int(d.max()),
                        "depth_words": int((z2.view(np.uint32) != ref["z"].view(np.uint32)).sum()), "lights": len(lights)})
        n += 5

    for seed in range(a, b):
        scene, vp, screen, pose = configs.fuzz_case(seed)
        check("small", seed, scene, vp, screen)
        if seed % 2 == 1:
            check_paths(seed, *configs.fuzz_case(seed)[:3])
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
