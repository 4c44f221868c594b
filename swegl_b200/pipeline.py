"""Frame-parallel rendering on ONE GPU: `depth` contexts (swegl_b200_ctx: own stream, pools, screen), frames round robin.

swegl's frame loop (src/test_1.cpp:366-385) renders one frame at a time; a batch of independent frames (an offline
sequence, several clients) has no such dependency.  One frame is a chain of kernels of which only the last two
(k_fragments, k_dof) fill the GPU -- the first four (k_vertex, k_mark, k_setup, k_spans) are latency bound on a
glTF-sized scene and leave most SMs idle.  With two or three contexts the head of frame i+1 runs under the tail of
frame i.  Every context holds its own copy of the scene (HBM is not the constraint: 180 GB), nothing is shared, so
the frames are exactly the frames a single context renders.

PyTorch is used for streams and events only.
"""

from .renderer import Renderer


class FramePipeline:
    def __init__(self, device=0, depth=2):
        import torch
        self.torch = torch
        self.device = int(device)
        self.depth = int(depth)
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(self.depth)]
        self.renderers = [Renderer(self.device, stream=s.cuda_stream) for s in self.streams]
        if self.depth > 1:
            for r in self.renderers:
                r.set_shared_gpu(True)          # the contexts' frames overlap on the GPU
        self.next = 0

    def close(self):
        for r in self.renderers:
            r.close()
        self.renderers = []

    def upload_scene(self, scene):
        for r in self.renderers:
            r.upload_scene(scene)

    def set_screen(self, w, h):
        for r in self.renderers:
            r.set_screen(w, h)

    def submit(self, scene, viewport_descs, node_mats=None):
        """begin_frame + render_device of every viewport on the next context; returns that context's index.  The frame
        stays in that context's device screen until the context is used again (`depth` submits later)."""
        k = self.next
        r = self.renderers[k]
        r.begin_frame(scene, node_mats)
        for d in viewport_descs:
            r.render_device(d, stats=False)
        self.next = (k + 1) % self.depth
        return k

    def synchronize(self):
        for r in self.renderers:
            r.synchronize()                 # raises if a frame overflowed a pool (the frame has to be submitted again)

    def size_pools(self, scene, viewports):
        """one synchronous frame per context: sizes the span / chunk / fragment pools for this workload"""
        for r in self.renderers:
            r.begin_frame(scene)
            for vp in viewports:
                r.render_device(vp, stats=True)

    def measure(self, scene, viewports, steps, warmup=6):
        """-> milliseconds for `steps` frames (CUDA events on the current torch stream, which forks to the context
        streams before the first frame and joins them after the last)"""
        torch = self.torch
        main = torch.cuda.current_stream(self.device)
        descs = [vp.desc() for vp in viewports]
        nodes = scene.node_matrices()
        self.size_pools(scene, viewports)
        for _ in range(warmup):
            self.submit(scene, descs, nodes)
        self.synchronize()
        torch.cuda.synchronize(self.device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        import time
        e0.record(main)
        for s in self.streams:
            s.wait_event(e0)
        t0 = time.perf_counter()
        for _ in range(steps):
            self.submit(scene, descs, nodes)
        self.host_submit_ms = 1e3 * (time.perf_counter() - t0)      # host side of the same frames (enqueue only)
        for s in self.streams:
            done = torch.cuda.Event()
            done.record(s)
            main.wait_event(done)
        e1.record(main)
        self.synchronize()
        torch.cuda.synchronize(self.device)
        return e0.elapsed_time(e1)

    def measure_batches(self, scene, viewports, batches, frames_per_batch, warmup=6):
        """-> [milliseconds of each of `batches` batches of `frames_per_batch` frames]: every batch is bracketed by its own
        pair of CUDA events on the current torch stream (fork to the context streams before its first frame, join after
        its last), the batches follow each other without a host synchronisation"""
        torch = self.torch
        main = torch.cuda.current_stream(self.device)
        descs = [vp.desc() for vp in viewports]
        nodes = scene.node_matrices()
        self.size_pools(scene, viewports)
        for _ in range(warmup):
            self.submit(scene, descs, nodes)
        self.synchronize()
        torch.cuda.synchronize(self.device)
        ev = []
        for _ in range(batches):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(main)
            for s in self.streams:
                s.wait_event(e0)
            for _ in range(frames_per_batch):
                self.submit(scene, descs, nodes)
            for s in self.streams:
                done = torch.cuda.Event()
                done.record(s)
                main.wait_event(done)
            e1.record(main)
            ev.append((e0, e1))
        self.synchronize()
        torch.cuda.synchronize(self.device)
        return [a.elapsed_time(b) for a, b in ev]

    def read_screen(self, k):
        return self.renderers[k].read_screen()
