import sys, os, time
sys.path.insert(0, "/root/repo")
import numpy as np
from swegl_b200 import Renderer, configs
name = sys.argv[1] if len(sys.argv) > 1 else "sphere100_1080"
scene, vps, screen, cfg = configs.build(name)
vp = vps[0]
a, b, ref = Renderer(0), Renderer(0), Renderer(0)
for r in (a, b, ref):
    r.set_band_culling(1)
    r.upload_scene(scene); r.set_screen(*screen)
cut = int(vp.h * 0.45)
for r, band in ((a, (0, cut)), (b, (cut, vp.h))):
    vp.band = band
    r.begin_frame(scene); r.render_device(vp, stats=True)
vp.band = (0, 0)
b.set_color_target(a.device_buffers()[0])
a.set_frame_sync(0, 2); b.set_frame_sync(1, 2)
for step in range(4):
    vp.camera.apply([("rotate_y", 0.2), ("translate", 0.15, 0.05, 0)])
    want = np.zeros((screen[1], screen[0]), np.uint32)
    ref.begin_frame(scene); ref.render(vp, want)
    order = ((b, (cut, vp.h)), (a, (0, cut))) if step % 2 else ((a, (0, cut)), (b, (cut, vp.h)))
    t0 = time.time()
    for r, band in order:
        vp.band = band
        r.begin_frame(scene); r.render_device(vp, stats=False)
    vp.band = (0, 0)
    t1 = time.time()
    a.synchronize(); t2 = time.time()
    got = a.read_screen()
    b.synchronize(); t3 = time.time()
    print(f"step {step}: submit {t1-t0:.3f}s a.sync {t2-t1:.3f}s b.sync {t3-t2:.3f}s a.status {a.frame_sync_status()} b.status {b.frame_sync_status()} equal {(got == want).all()} diff_rows {np.unique(np.nonzero(got != want)[0])[:6]}", flush=True)
