import sys, time
sys.path.insert(0, '/root/repo')
import torch, numpy as np
from swegl_b200 import Renderer, configs
name = sys.argv[1] if len(sys.argv) > 1 else 'truck_4k_dof'
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
r = Renderer(0, stream=st.cuda_stream)
scene, vps, screen, cfg = configs.build(name)
r.upload_scene(scene); r.set_screen(*screen)
nodes = scene.node_matrices(); d = vps[0].desc()
r.begin_frame(scene, nodes); st = r.render_device(d, stats=True)
print('stats', st.n_setup_triangles, st.n_spans, st.n_chunks, st.n_covered, st.n_launches, st.pool_grows)
for _ in range(20): r.begin_frame(scene, nodes); r.render_device(d, stats=False)
torch.cuda.synchronize()
K = 300
t0 = time.perf_counter()
for _ in range(K): r.begin_frame(scene, nodes); r.render_device(d, stats=False)
t_issue = time.perf_counter() - t0
torch.cuda.synchronize()
t1 = time.perf_counter() - t0
print(f'{name}: wall {t1/K*1e3:.4f} ms/frame (host issue {t_issue/K*1e3:.4f} ms/frame)')
# events around the whole batch
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(K): r.begin_frame(scene, nodes); r.render_device(d, stats=False)
e1.record(); torch.cuda.synchronize()
print(f'events over batch: {e0.elapsed_time(e1)/K:.4f} ms/frame')
# render only (no begin_frame)
e0.record()
for _ in range(K): r.render_device(d, stats=False)
e1.record(); torch.cuda.synchronize()
print(f'render_device only: {e0.elapsed_time(e1)/K:.4f} ms/frame')
r.set_timing(True)
acc = None
for i in range(20):
    r.begin_frame(scene, nodes); st = r.render_device(d, stats=True)
    v = [st.ms_vertex, st.ms_setup, st.ms_raster, st.ms_fragment, st.ms_post, st.ms_total]
    acc = v if acc is None else [a + b for a, b in zip(acc, v)]
print('timing mode avg ms:', [round(a / 20, 4) for a in acc])
