"""Generate tests/golden/ from the UNMODIFIED reference (oracle/_ref) — run in the build container.

For every workload in swegl_b200.configs.CONFIGS:
  1. render it with the reference itself (swegl::render through oracle/ref_driver.cpp),
  2. render it with the C restatement (oracle/swegl_oracle.c),
  3. require bit-identical frames, depth buffers and per-vertex state,
  4. record FNV-1a-64 hashes + counters in tests/golden/MANIFEST.json.
DoF workloads: the reference's DoF is broken at HEAD, so the pre-DoF frame is pinned against the
unmodified reference and the DoF-R frame against the reference built with oracle/dof_r.patch
(oracle/_ref/libswegl_ref_dofr.so, `make -C oracle refdofr`): both must be bit-identical to the oracle's.
Small full-frame fixtures (compressed) are stored for the 640x480 config so the golden check does not
depend on hashes only.

    python tools/make_golden.py
"""
import copy
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.binding import Oracle, Ref, REF_DOFR_LIB  # noqa: E402
from swegl_b200 import _abi, configs  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def ref_render_config(ref, scene, vps, screen, with_dof=False):
    h = ref.import_scene(scene)
    scr = ref.lib.ref_screen_new(*screen)
    rvs = [ref.make_viewport(scr, vp, vp.pose, with_dof=with_dof) for vp in vps]
    if len(rvs) == 1:
        ref.lib.ref_render(h, rvs[0])
    elif len(rvs) == 4:
        ref.lib.ref_render4(h, *rvs)
    else:
        raise ValueError("1 or 4 viewports")
    import ctypes as C
    px = np.ctypeslib.as_array(C.cast(ref.lib.ref_screen_pixels(scr), C.POINTER(C.c_uint32)), shape=(screen[1], screen[0])).copy()
    zs = [np.ctypeslib.as_array(C.cast(ref.lib.ref_viewport_zbuffer(rv), C.POINTER(C.c_float)), shape=(vp.h, vp.w)).copy()
          for rv, vp in zip(rvs, vps)]
    vstate = ref.vertex_state(h, scene.n_vertices)
    for rv in rvs:
        ref.lib.ref_viewport_free(rv)
    ref.lib.ref_screen_free(scr)
    ref.lib.ref_scene_free(h)
    return px, zs, vstate


def main():
    os.makedirs(OUT, exist_ok=True)
    orc, ref, ref_dofr = Oracle(), Ref(), Ref(REF_DOFR_LIB)
    manifest = {}
    for name in configs.CONFIGS:
        scene, vps, screen, cfg = configs.build(name)
        has_dof = any(vp.post_mode == _abi.POST_DOF for vp in vps)
        vps_nodof = [copy.copy(vp) for vp in vps]
        for vp in vps_nodof:
            vp.post_mode = _abi.POST_NULL
        rpx, rzs, rv = ref_render_config(ref, scene, vps_nodof, screen)
        opx = np.zeros((screen[1], screen[0]), np.uint32)
        outs = [orc.render(scene, vp, screen_wh=screen, pixels=opx, want_vertices=True) for vp in vps_nodof]
        assert (rpx == opx).all(), f"{name}: oracle frame differs from the reference"
        for rz, o in zip(rzs, outs):
            assert (rz.view(np.uint32) == o["z"].view(np.uint32)).all(), f"{name}: oracle depth differs from the reference"
        for k in ("v_world", "v_viewport", "normal_world"):
            assert (rv[k].view(np.uint32) == outs[-1][k].view(np.uint32)).all(), f"{name}: {k} differs from the reference"
        assert (rv["yes"] == outs[-1]["yes"]).all(), f"{name}: yes flags differ from the reference"
        entry = {"description": cfg["desc"], "screen": list(screen), "pinned_against_reference": True,
                 "frame_fnv1a64": "%016x" % orc.fnv(opx),
                 "depth_fnv1a64": ["%016x" % orc.fnv(o["z"].view(np.uint32)) for o in outs],
                 "covered": [int(o["n_covered"]) for o in outs], "fragments": [int(o["n_fragments"]) for o in outs],
                 "spans": [int(o["n_spans"]) for o in outs], "setup_triangles": [int(o["n_setup_triangles"]) for o in outs],
                 "vertex_fnv1a64": {k: "%016x" % orc.fnv(outs[-1][k].view(np.uint32)) for k in ("v_world", "v_viewport", "normal_world")},
                 "yes_count": int(outs[-1]["yes"].sum())}
        if has_dof:
            dpx = np.zeros((screen[1], screen[0]), np.uint32)
            for vp in vps:
                orc.render(scene, vp, screen_wh=screen, pixels=dpx)
            rdpx, _, _ = ref_render_config(ref_dofr, scene, vps, screen, with_dof=True)
            assert (rdpx == dpx).all(), f"{name}: the oracle's DoF-R frame differs from the patched reference"
            entry["pre_dof_frame_fnv1a64"] = entry.pop("frame_fnv1a64")
            entry["frame_fnv1a64"] = "%016x" % orc.fnv(dpx)
            entry["dof_r_pinned_against_patched_reference"] = True
        manifest[name] = entry
        print(name, entry["frame_fnv1a64"], entry["covered"])
        if name == "box_640":
            np.savez_compressed(os.path.join(OUT, "box_640.npz"), pixels=opx, z=outs[0]["z"], v_world=outs[0]["v_world"],
                                v_viewport=outs[0]["v_viewport"], normal_world=outs[0]["normal_world"], yes=outs[0]["yes"])
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    print("wrote", os.path.join(OUT, "MANIFEST.json"))


if __name__ == "__main__":
    main()
