"""Per-CTA timeline of k_fragments from a probe build (-DFRAG_PROBE_TIMELINE; tools only, not the product):
    SWEGL_B200_LIB=swegl_b200/libswegl_b200_tl.so python tools/timeline_probe.py [workload]
Prints when the CTAs start, how long the dependency wait is, the duration of busy-tile and streaming items, and when the
CTAs end, all relative to the first CTA's entry."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "truck_4k_dof"
    import torch
    from swegl_b200 import Renderer, configs, _abi
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
    scene, vps, screen, cfg = configs.build(name)
    r = Renderer(0, stream=stream.cuda_stream)
    r.upload_scene(scene); r.set_screen(*screen)
    r.set_shading(_abi.SHADING_FAST)
    nodes = scene.node_matrices()
    lib = r.lib
    lib.swegl_b200_probe_timeline.argtypes = [C.c_void_p, C.c_int]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for i in range(4):
        r.begin_frame(scene, nodes)
        r.render_device(vps[0], stats=True)
    flush.zero_(); torch.cuda.synchronize()
    lib.swegl_b200_probe_timeline(None, 1)
    r.begin_frame(scene, nodes)
    st = r.render_device(vps[0], stats=True)
    tl = np.zeros((2048, 16), np.uint64)
    lib.swegl_b200_probe_timeline(tl.ctypes.data, 0)
    used = tl[:, 0] != 0
    tl = tl[used]
    t0 = tl[:, 0].min()
    start = (tl[:, 0] - t0).astype(np.float64) / 1e3
    waited = (tl[:, 1] - t0).astype(np.float64) / 1e3
    print(f"{name}: {len(tl)} CTAs, covered {st.n_covered}")
    q = lambda a: " ".join(f"{np.percentile(a, p):7.2f}" for p in (0, 10, 50, 90, 100))
    print("                     min     p10     p50     p90     max  (us)")
    print(f"CTA entry        {q(start)}")
    print(f"after dep. wait  {q(waited)}")
    busy_d, grp_d, end = [], [], []
    for row in tl:
        prev = row[1]
        last = prev
        for k in range(2, 8):
            if row[k] == 0:
                break
            t = row[k] >> np.uint64(1)
            (grp_d if (row[k] & np.uint64(1)) else busy_d).append((float(t) - float(prev)) / 1e3)
            prev = t; last = t
        end.append((float(last) - float(t0)) / 1e3)
    busy_d, grp_d, end = np.array(busy_d), np.array(grp_d), np.array(end)
    print(f"busy item   n={len(busy_d):5d} {q(busy_d)}   sum/CTA-slots {busy_d.sum() / len(tl):.2f}")
    print(f"group item  n={len(grp_d):5d} {q(grp_d)}   sum/CTA-slots {grp_d.sum() / len(tl):.2f}")
    print(f"CTA end          {q(end)}")
    # first-round busy items vs later ones
    first = np.array([(float(row[2] >> np.uint64(1)) - float(row[1])) / 1e3 for row in tl if row[2] != 0 and not (row[2] & np.uint64(1))])
    print(f"first item (busy) n={len(first)} {q(first)}")


    # phases of warp 0's row of the CTA's first busy tile (slots 8..15)
    ph = tl[(tl[:, 8] != 0) & (tl[:, 11] != 0)]
    names = ["busy_list load", "bin_cnt load + reset", "records staged", "bin A resolved (last even bin)", "bin A shaded+stored", "bin B resolved (last odd bin)", "bin B shaded+stored"]
    cols = [(8, 9), (9, 10), (10, 11), (11, 12), (12, 13), (11, 14), (14, 15)]
    print(f"phases of warp 0, first busy item ({len(ph)} CTAs):")
    for nm, (a, b) in zip(names, cols):
        ok = (ph[:, a] != 0) & (ph[:, b] != 0) & (ph[:, b] >= ph[:, a])
        d = (ph[ok, b] - ph[ok, a]).astype(np.float64) / 1e3
        if len(d):
            print(f"  {nm:34s} n={len(d):4d} {q(d)}")


if __name__ == "__main__":
    main()
