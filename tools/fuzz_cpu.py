"""Long-run pinning of the oracle: seeds [a, b) of swegl_b200.configs.fuzz_case through the UNMODIFIED reference
(oracle/_ref/libswegl_ref.so) and through the C restatement, bit for bit (colour, depth, vertex state, `yes` marks), in
the flavours of tools/fuzz_gpu.py: small soup, transparency layers (every 4th seed), heavy overdraw (every 8th).
Runs in the build container only (needs the reference build).
    python tools/fuzz_cpu.py 0 2000 profiles/r01_fuzz_cpu.json
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
    from oracle.binding import Oracle, Ref
    from swegl_b200 import _abi, configs
    from swegl_b200.scene import Viewport
    ref, orc = Ref(), Oracle()
    bad, ub, n, t0 = [], [], 0, time.time()
    fragments = 0

    def check(tag, seed, scene, vp, screen, pose):
        nonlocal n, fragments
        h = ref.import_scene(scene); scr = ref.lib.ref_screen_new(*screen); rv = ref.make_viewport(scr, vp, pose)
        rpx, rz = ref.render(h, rv, scr, screen[0], screen[1], vp.w, vp.h)
        rvs = ref.vertex_state(h, scene.n_vertices)
        ref.lib.ref_viewport_free(rv); ref.lib.ref_screen_free(scr); ref.lib.ref_scene_free(h)
        o = orc.render(scene, vp, screen_wh=screen, want_vertices=True)
        ok = (rpx == o["pixels"]).all() and (rz.view(np.uint32) == o["z"].view(np.uint32)).all() and (rvs["yes"] == o["yes"]).all() \
            and all((rvs[k].view(np.uint32) == o[k].view(np.uint32)).all() for k in ("v_world", "v_viewport", "normal_world"))
        n += 1
        fragments += int(o["n_fragments"])
        if not ok:
            rec = {"flavour": tag, "seed": seed, "pixels": int((rpx != o["pixels"]).sum()),
                   "depth_words": int((rz.view(np.uint32) != o["z"].view(np.uint32)).sum()),
                   "oracle_texel_guard_fragments": int(o["n_texel_guard"])}
            # the reference indexes the bitmap with a negative row / column on this frame (pixel_shaders.cpp:354-377 with a
            # texture coordinate extrapolated below zero by a near-plane sliver): it reads outside the bitmap, the oracle
            # wraps.  Only colour may differ then, and only where the oracle counted such fragments.
            if o["n_texel_guard"] > 0 and rec["depth_words"] == 0:
                ub.append(rec)
            else:
                bad.append(rec)

    for seed in range(a, b):
        scene, vp, screen, pose = configs.fuzz_case(seed)
        vp.post_mode = _abi.POST_NULL
        check("small", seed, scene, vp, screen, pose)
        if seed % 4 == 0:
            scene, vp, screen, pose = configs.fuzz_layers_case(seed)
            check("layers", seed, scene, vp, screen, pose)
        if seed % 8 == 0:
            _, vp0, _, pose = configs.fuzz_case(seed)
            heavy = configs.fuzz_scene(seed, n_prims=30, verts_per_prim=120)
            vp = Viewport(0, 0, 640, 360, light_mode=vp0.light_mode, tex_mode=vp0.tex_mode)
            vp.camera.apply(pose)
            check("heavy", seed, heavy, vp, (640, 360), pose)
    rep = {"seeds": [a, b], "frames_checked": n, "fragments_shaded": fragments, "mismatches": bad,
           "frames_where_the_reference_reads_outside_a_bitmap": ub, "seconds": round(time.time() - t0, 1),
           "what": "unmodified reference vs oracle/swegl_oracle.c: colour, depth, v_world, v_viewport, normal_world, yes -- bit for bit"}
    s = json.dumps(rep)
    print(s)
    if out_path:
        with open(out_path, "w") as f:
            f.write(s + "\n")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
