"""k_dof cost split: the same 4K DoF frame with the scene in view and with the camera turned away (every DoF tile is
background), per-stage CUDA-event times of the context's timing mode."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from swegl_b200 import Renderer, configs
    r = Renderer(0)
    out = {}
    for name in sys.argv[1:] or ("truck_4k_dof", "brainstem_4k_dof"):
        scene, vps, screen, cfg = configs.build(name)
        r.upload_scene(scene); r.set_screen(*screen)
        for tag, extra in (("in_view", []), ("turned_away", [("rotate_y", 3.1)])):
            vp = configs.build(name)[1][0]
            vp.camera.apply(extra)
            r.set_timing(True)
            acc, n = {}, 0
            for i in range(30):
                r.begin_frame(scene)
                st = r.render_device(vp, stats=True)
                if i >= 5:
                    for k in ("ms_fragment", "ms_post", "ms_total"):
                        acc[k] = acc.get(k, 0.0) + getattr(st, k)
                    n += 1
            r.set_timing(False)
            out[f"{name}/{tag}"] = {k: round(v / n, 5) for k, v in acc.items()} | {"covered": int(st.n_covered)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
