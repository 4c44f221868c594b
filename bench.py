#!/usr/bin/env python
"""bench.py — fps / shaded Mpixels/s of the swegl hot path on B200, beside swegl's own CPU renderer.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

A step = one BATCH of FRAMES_PER_STEP independent frames of the workload (a frame = swegl::render(scene, viewport) =
vertex stage + cull/mark + setup + scan rasterisation with z test + Phong/bilinear shading + DoF-R; one frame takes
~0.08 ms, far too short a timed region on its own).  Default workload = the north-star
target of BASELINE.json (CesiumMilkTruck 3840x2160, Phong + bilinear, sun + 2 point lights, DoF-R);
the 1080p configs[1] frame is measured in the same run and reported under "also".

  value     : frames/s over K batches of independent frames with the scene resident in HBM, rendered round robin by
              --pipeline-depth contexts per GPU (swegl_b200.FramePipeline: the latency-bound head of frame i+1 runs
              under the fragment/DoF kernels of frame i); one pair of CUDA events around the batch; max over ranks.
              one_frame_at_a_time: the same frames on one context, CUDA events around each frame on the launching
              stream, a 256 MiB L2 flush between frames (outside the events).
  e2e       : the same frames through the public host API (Renderer.begin_frame + render_async/wait, three frames in
              flight): node matrices/lights H2D and the finished frame D2H into pinned memory every step;
              e2e.blocking_call_fps is the same through the blocking Renderer.render.
  roofline  : the dominant kernel's algorithmic bytes / its CUDA-event duration.
  cpu_baseline : the unmodified reference (oracle/_ref) timed on this box's host cores (rank 0, N=1).
N>1 (torchrun): frame-parallel, scene replicated, frame i on GPU i mod N, no collective ("weak").
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "fps"
UNIT = "frames/s"
DEFAULT_WORKLOAD = "truck_4k_dof"
ALSO_WORKLOAD = "truck_1080"
FRAMES_PER_STEP = 256           # `value`: a step is a batch of this many frames (K = 50 -> ~1 s of timed device work at 4K)
E2E_FRAMES_PER_STEP = 32        # the end-to-end loops are PCIe / host bound: smaller batches keep the default run within minutes
E2E_IN_FLIGHT = 3               # frames submitted before the oldest is waited for (the library keeps three staging images)
SERIAL_FRAMES_MAX = 1000        # one_frame_at_a_time: a 256 MiB L2 flush sits between the frames
SINGLE_FRAME_STEPS = 50         # timed frames of the single-frame (sharded / multiview) measurements


def golden_fnv(name):
    """FNV-1a-64 of the workload's frame as the unmodified reference renders it (tests/golden/MANIFEST.json, pinned by
    tools/make_golden.py against oracle/_ref; DoF workloads: against the reference built with oracle/dof_r.patch)"""
    try:
        return int(json.load(open(os.path.join(ROOT, "tests", "golden", "MANIFEST.json")))[name]["frame_fnv1a64"], 16)
    except (OSError, KeyError, ValueError):
        return None


def workload_config(name, cfg, scene, screen, n_gpus=1, depth=4):
    """`config` of the JSON line: the workload and how the swegl_b200 arm measures `value` on it.  BOTH arms print exactly
    this dict for the same command line (the driver compares them); what is specific to an arm's own run -- the reference
    arm's processes, the host's NUMA binding -- lives in top-level keys of its line."""
    return {"workload": name, "description": cfg["desc"], "resolution": f"{screen[0]}x{screen[1]}",
            "triangles": scene.n_triangles(), "vertices": scene.n_vertices,
            "shader": "phong+bilinear", "data": "bundled glTF scene pack" if not name.startswith("sphere") else "synthetic make_sphere + seeded LCG texture",
            "step": f"one batch of {FRAMES_PER_STEP} independent frames", "frames_per_step": FRAMES_PER_STEP,
            "parallelism": "single GPU" if n_gpus == 1 else
                           f"frame-parallel x{n_gpus}: scene replicated, frame i on GPU i mod N, no collective",
            "frames_in_flight_per_gpu": depth,
            "l2": (f"value: {depth} contexts per GPU render the batches of independent frames round robin; "
                   "their frame buffers and pools (about 140 MB each at 4K) rotate, so the working set exceeds "
                   "the 126 MB L2 and no flush is inserted; one_frame_at_a_time: 256 MiB memset between timed "
                   "frames, outside the CUDA events") if depth > 1 else
                  "256 MiB memset between timed frames, outside the CUDA events"}


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# the CPU reference arm
# ------------------------------------------------------------------------------------------------
_W = {}


def _ref_worker_init(name):
    """one process = one copy of the single-threaded reference renderer (scene imported once)"""
    os.environ["SWEGL_B200_IMAGE_DECODER"] = "pil"      # (spawned workers inherit it anyway)
    from oracle.binding import Ref, Oracle, REF_LIB
    from swegl_b200 import configs
    scene, vps, screen, cfg = configs.build(name)
    _W.update(scene=scene, vp=vps[0], screen=screen, orc=Oracle(), use_ref=os.path.exists(REF_LIB))
    if _W["use_ref"]:
        ref = Ref()
        h = ref.import_scene(scene)
        scr = ref.lib.ref_screen_new(*screen)
        _W.update(ref=ref, h=h, scr=scr, rv=ref.make_viewport(scr, vps[0], vps[0].pose))


def _ref_worker_frames(frames):
    """render `frames` frames; returns (seconds per frame list, kind)"""
    from swegl_b200 import _abi
    vp, screen, orc = _W["vp"], _W["screen"], _W["orc"]
    times = []
    for _ in range(frames):
        t0 = time.perf_counter()
        if _W["use_ref"]:
            px, z = _W["ref"].render(_W["h"], _W["rv"], _W["scr"], screen[0], screen[1], vp.w, vp.h)
            if vp.post_mode == _abi.POST_DOF:       # the reference's DoF is broken at HEAD: DoF-R from the C oracle
                orc.dof_r(px, z, vp.focal_distance, vp.focal_depth)
        else:
            orc.render(_W["scene"], vp, screen_wh=screen)
        times.append(time.perf_counter() - t0)
    return times, ("reference" if _W["use_ref"] else "port")


def cpu_reference_single(name, frames, warmup):
    """the reference as shipped: one process, one raster thread. -> (frames/s, kind, per-frame seconds)"""
    _ref_worker_init(name)
    _ref_worker_frames(warmup)
    times, kind = _ref_worker_frames(frames)
    return frames / sum(times), kind, times


def run_reference_arm(args, rank, world):
    """bench.py --impl reference: every host core renders its own frames with the (single-threaded)
    reference; a step = one frame per process; value = aggregate frames/s (median over steps)."""
    if rank != 0:
        return
    # the arm runs the unmodified reference and nothing of the product: the scene packs' embedded PNG / JPEG textures are
    # decoded with PIL here instead of libswegl_b200.so's decoder (same texels, guarded by the packs' digests)
    os.environ["SWEGL_B200_IMAGE_DECODER"] = "pil"
    import multiprocessing as mp
    from swegl_b200 import configs
    name = args.workload
    scene, vps, screen, cfg = configs.build(name)
    procs = max(1, min(os.cpu_count() or 1, 64))
    vals, kind = [], "reference"
    t_start = time.perf_counter()
    with mp.get_context("spawn").Pool(procs, initializer=_ref_worker_init, initargs=(name,)) as pool:
        for _ in range(max(args.warmup, 0)):
            pool.map(_ref_worker_frames, [1] * procs, chunksize=1)
            if time.perf_counter() - t_start > 60:
                break
        for _ in range(args.steps):
            t0 = time.perf_counter()
            res = pool.map(_ref_worker_frames, [1] * procs, chunksize=1)
            wall = time.perf_counter() - t0
            kind = res[0][1]
            vals.append(procs / wall)
            if time.perf_counter() - t_start > 200:
                break
    value = statistics.median(vals)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(vals),
            "warmup": args.warmup, "ms_per_step": 1e3 * procs / value, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic" if name.startswith("sphere") else "bundled scene",
            "config": workload_config(name, cfg, scene, screen, n_gpus=max(1, args.gpus), depth=args.pipeline_depth),
            "arm": "the unmodified reference on the host cores (config.step / parallelism / l2 describe the swegl_b200 arm of the same command; "
                   "here a step is one frame per renderer process, no GPU involved)",
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": kind,
                             "sample": f"{procs} independent single-threaded renderer processes, 1 frame each per step, "
                                       f"{len(vals)} steps (swegl has one raster thread, so frame-parallel processes are all the "
                                       f"host threads it can use); DoF-R by the C oracle because the reference's DoF is broken at HEAD"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# the GPU arm
# ------------------------------------------------------------------------------------------------
def algorithmic_bytes(scene, vp, covered):
    """B_frag = 8*W*H + T_unique ; B_dof = 12*W*H"""
    wh = vp.w * vp.h
    tex_bytes = sum(int(t.size) * 4 for t in scene.textures)
    t_unique = min(tex_bytes, 16 * covered)
    return {"fragment": 8 * wh + t_unique, "dof": 12 * wh}


def stage_fracs(scene, vp, covered, ms_per_stage):
    """k_fragments / k_dof of a workload against the measured HBM bandwidth: algorithmic bytes / stage time / peak"""
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = float(json.load(open(peaks_path))["hbm_gbs"]) if os.path.exists(peaks_path) else 6650.0
    ab = algorithmic_bytes(scene, vp, covered)
    out = {"k_fragments": ab["fragment"] / (ms_per_stage["ms_fragment"] * 1e-3) / 1e9 / peak if ms_per_stage["ms_fragment"] > 0 else None}
    if ms_per_stage.get("ms_post", 0) > 0.005:              # (a null post pass is a 2.7 us copy-less stage, not k_dof)
        out["k_dof"] = ab["dof"] / (ms_per_stage["ms_post"] * 1e-3) / 1e9 / peak
    return out


def measure_gpu(r, torch, scene, vps, screen, steps, warmup, flush):
    """-> (seconds for `steps` frames measured with per-frame CUDA events, stats of the last frame)"""
    for w in range(warmup):
        r.begin_frame(scene)
        for vp in vps:
            r.render_device(vp, stats=(w == 0))        # the first frame sizes the span/chunk/fragment pools
    torch.cuda.synchronize()
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    descs = [vp.desc() for vp in vps]
    fd_nodes = scene.node_matrices()
    for i in range(steps):
        flush.zero_()
        ev0[i].record()
        r.begin_frame(scene, fd_nodes)
        for d in descs:
            r.render_device(d, stats=False)
        ev1[i].record()
    r.synchronize()                                     # raises if any timed frame overflowed a pool (incomplete work)
    torch.cuda.synchronize()
    ms = [a.elapsed_time(b) for a, b in zip(ev0, ev1)]
    return sum(ms) / 1e3, ms


def measure_e2e(r, torch, scene, vps, screen, steps, warmup, pixels):
    fd_nodes = scene.node_matrices()
    descs = [vp.desc() for vp in vps]
    for _ in range(max(1, warmup)):
        r.begin_frame(scene, fd_nodes)
        for d in descs:
            r.render(d, pixels, stats=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        r.begin_frame(scene, fd_nodes)
        for d in descs:
            r.render(d, pixels, stats=False)
    torch.cuda.synchronize()
    return time.perf_counter() - t0


def measure_e2e_pipelined(r, torch, scene, vps, screen, steps, warmup, images):
    """the same per-frame host work (begin_frame: node matrices + lights H2D; finished frame D2H into pinned memory)
    through render_async/wait: frames i+1 and i+2 are submitted before frame i is waited for (the library keeps three
    staging images), three host images take turns"""
    fd_nodes = scene.node_matrices()
    descs = [vp.desc() for vp in vps]

    def run(n):
        pending = []
        for i in range(n):
            r.begin_frame(scene, fd_nodes)
            pending.append([r.render_async(d, images[i % len(images)]) for d in descs])
            if len(pending) > len(images) - 1:
                for t in pending.pop(0):
                    r.wait(t)
        for fr in pending:
            for t in fr:
                r.wait(t)
    run(max(2, warmup))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run(steps)
    torch.cuda.synchronize()
    return time.perf_counter() - t0


def measure_kernels(r, scene, vps, steps):
    """per-stage CUDA-event times (context timing mode), averaged over `steps` frames"""
    r.set_timing(True)
    acc = {"ms_vertex": 0.0, "ms_setup": 0.0, "ms_raster": 0.0, "ms_fragment": 0.0, "ms_post": 0.0, "ms_total": 0.0}
    last = None
    n = 0
    for i in range(steps + 2):
        r.begin_frame(scene)
        for vp in vps:
            st = r.render_device(vp, stats=True)
            if i >= 2:
                for k in acc:
                    acc[k] += getattr(st, k)
                n += 1
            last = st
    r.set_timing(False)
    return {k: v / max(n, 1) for k, v in acc.items()}, last


def measure_pipelined(torch, local_rank, scene, vps, screen, steps, warmup, depth, dist=None):
    """`steps` batches of FRAMES_PER_STEP independent frames through swegl_b200.FramePipeline: `depth` contexts on this GPU,
    frames round robin, device-resident; one pair of CUDA events around every batch (fork/join over the context streams).
    The contexts' frame buffers and pools rotate, so at 4K the working set (depth x ~140 MB) never fits the 126 MB L2.
    -> [seconds per batch]"""
    from swegl_b200.pipeline import FramePipeline
    pipe = FramePipeline(local_rank, depth)
    try:
        pipe.upload_scene(scene)
        pipe.set_screen(*screen)
        pipe.measure(scene, vps, 2 * depth, warmup=max(warmup, 3) * depth)         # warm-up: pools sized, graphs captured
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        ms = pipe.measure_batches(scene, vps, steps, FRAMES_PER_STEP, warmup=depth)
    finally:
        pipe.close()
    return [m / 1e3 for m in ms]


def gpu_workload(r, torch, name, steps, warmup, flush, world, dist, do_e2e=True, depth=4, local_rank=0):
    """steps = batches.  -> dict with the per-batch seconds of the pipelined (`value`) measurement, the serial one-frame-at-
    a-time measurement and the two end-to-end loops, each reduced with MAX over the ranks"""
    from swegl_b200 import configs
    scene, vps, screen, cfg = configs.build(name)
    r.upload_scene(scene)
    r.set_screen(*screen)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    serial_frames = min(steps * FRAMES_PER_STEP, SERIAL_FRAMES_MAX)
    secs, ms = measure_gpu(r, torch, scene, vps, screen, serial_frames, warmup, flush)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    batch_secs = measure_pipelined(torch, local_rank, scene, vps, screen, steps, warmup, depth, dist if world > 1 else None) if depth > 1 else None
    if batch_secs is None:                              # --pipeline-depth 1: the serial measurement in batches
        batch_secs = []
        for _ in range(steps):
            bs, _ = measure_gpu(r, torch, scene, vps, screen, FRAMES_PER_STEP, 0, flush)
            batch_secs.append(bs)
    if world > 1:
        t = torch.tensor([secs] + batch_secs, device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)        # per batch: the slowest rank
        secs, batch_secs = float(t[0].item()), [float(x) for x in t[1:].tolist()]
        dist.barrier()
    out = {"scene": scene, "vps": vps, "screen": screen, "cfg": cfg, "secs": secs, "serial_frames": serial_frames, "ms": ms,
           "batch_secs": batch_secs, "pipe_secs": sum(batch_secs), "depth": depth}
    if do_e2e:
        images = [r.alloc_host((screen[1], screen[0]), np.uint32) for _ in range(E2E_IN_FLIGHT)]
        n_sync, n_async = max(steps * E2E_FRAMES_PER_STEP // 4, 8), steps * E2E_FRAMES_PER_STEP
        sync_secs = measure_e2e(r, torch, scene, vps, screen, n_sync, warmup, images[0])
        r.readback_stats(reset=True)
        e2e_secs = measure_e2e_pipelined(r, torch, scene, vps, screen, n_async, warmup, images)
        rb_bytes, rb_frames = r.readback_stats()
        if world > 1:
            t = torch.tensor([e2e_secs, sync_secs], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_secs, sync_secs = float(t[0].item()), float(t[1].item())
        out.update(e2e_secs=e2e_secs, e2e_frames=n_async, e2e_sync_secs=sync_secs, e2e_sync_frames=n_sync)
        nodes = scene.n_nodes
        out["h2d"] = nodes * (64 + 36) + 16 * len(scene.point_lights)
        out["d2h_full"] = sum(vp.w * vp.h * 4 for vp in vps)
        out["d2h"] = rb_bytes / max(rb_frames, 1) + 64            # what really crossed PCIe per frame (+ the frame's counters)
        # the partial read-back must leave the same host image as a full copy: check it against the golden frame hash
        from swegl_b200 import _abi
        from swegl_b200.renderer import frame_hash
        want = golden_fnv(name)
        if want is not None and len(vps) == 1:
            r.set_shading(_abi.SHADING_EXACT)           # the manifest pins the bit-exact frame; the timed frames use the +-1 LSB shading
            fd_nodes = scene.node_matrices()
            for i in range(2 * len(images)):            # through the pipelined, partial path: every host image (and staging image) twice
                r.begin_frame(scene, fd_nodes)
                r.wait(r.render_async(vps[0].desc(), images[i % len(images)]))
            out["frame_fnv_ok"] = all(frame_hash(im) == want for im in images)
            r.set_shading(_abi.SHADING_FAST)
    return out


class _DevArray:
    """torch.as_tensor() view of a raw device pointer (the context's screen buffer) for the NCCL gather"""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<i4", "data": (int(ptr), False), "version": 2}


def nvlink_kib(local_rank):
    """(tx, rx) payload KiB this GPU has moved over all its NVLinks so far (NVML field values
    NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX / _RX, summed over the links), or None where NVML does not report them"""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        out = []
        for fid in (pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX, pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX):
            total, seen = 0, False
            vals = pynvml.nvmlDeviceGetFieldValues(h, [(fid, 0xFFFFFFFF)])        # scope UINT_MAX: all links
            v = vals[0]
            if v.nvmlReturn == pynvml.NVML_SUCCESS:
                total, seen = int(v.value.ullVal), True
            else:
                for link in range(18):
                    v = pynvml.nvmlDeviceGetFieldValues(h, [(fid, link)])[0]
                    if v.nvmlReturn == pynvml.NVML_SUCCESS:
                        total += int(v.value.ullVal); seen = True
            if not seen:
                print(f"[bench] NVLink counters: field {fid} not reported (nvmlReturn {vals[0].nvmlReturn})", file=sys.stderr)
                return None
            out.append(total)
        return tuple(out)
    except Exception as e:
        print(f"[bench] NVLink counters unavailable: {type(e).__name__}: {e}", file=sys.stderr)
        return None


def nvlink_delta(torch, dist, world, local_rank, before, frames):
    """per-frame NVLink payload of every rank between `before` (nvlink_kib) and now: {"tx_kib_per_frame": [...], "rx_...": [...]}"""
    after = nvlink_kib(local_rank)
    ok = before is not None and after is not None
    t = torch.tensor([(after[0] - before[0]) / frames if ok else -1.0, (after[1] - before[1]) / frames if ok else -1.0], device="cuda", dtype=torch.float64)
    allt = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allt, t)
    if any(float(a[0]) < 0 for a in allt):
        return None
    return {"tx_kib_per_frame": [round(float(a[0]), 1) for a in allt], "rx_kib_per_frame": [round(float(a[1]), 1) for a in allt],
            "source": "NVML NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX/RX, all links, sampled around warm-up + timed frames"}


def screen_fnv_ok(r, torch, dist, world, rank, name, render_one):
    """The assembled frame on rank 0 against the golden FNV-1a-64 of the reference's frame (tests/golden/MANIFEST.json).
    The manifest pins the bit-exact frame, the timed frames use the +-1 LSB shading: every rank renders the frame once
    more with exact shading through the very same path (`render_one`), rank 0 reads its device screen back and hashes it.
    -> True / False on rank 0 (None elsewhere, or when the manifest has no entry)"""
    from swegl_b200 import _abi
    from swegl_b200.renderer import frame_hash
    want = golden_fnv(name)
    if want is None:
        return None
    r.set_shading(_abi.SHADING_EXACT)
    for _ in range(2):
        render_one()
    r.synchronize(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ok = None
    if rank == 0:
        ok = frame_hash(r.read_screen()) == want
    r.set_shading(_abi.SHADING_FAST)
    if world > 1:
        dist.barrier()
    return ok


def multiview_frame(r, torch, dist, name, steps, warmup, flush, world, rank):
    """config 4: ONE split-screen frame, one viewport_t per GPU (viewport v -> rank v mod world, SURVEY §8e / renderer.hpp:
    20-34), scene replicated.  Output on rank 0 either by an NCCL rectangle gather after rendering or by peer-memory
    stores of the producing kernels (share_screen) + a one-element all-reduce.  Returns ms per frame, max over ranks."""
    from swegl_b200 import configs, sharding
    scene, vps, screen, cfg = configs.build(name)
    r.upload_scene(scene)
    r.set_screen(*screen)
    nodes = scene.node_matrices()
    mine = sharding.viewports_for_rank(len(vps), world, rank)
    descs = [vps[v].desc() for v in mine]
    rects = [(vp.x, vp.y, vp.w, vp.h) for vp in vps]
    screen_ptr, _ = r.device_buffers()
    full = torch.as_tensor(_DevArray(screen_ptr, (screen[1], screen[0])), device="cuda")
    token = torch.zeros(1, device="cuda")
    start_token = torch.zeros(1, device="cuda")
    state = {"peer": False}

    def one():
        r.begin_frame(scene, nodes)
        for d in descs:
            r.render_device(d, stats=False)
        if world > 1:
            if state["peer"]:
                sharding.frame_barrier(dist, token)
            else:
                sharding.gather_rects_inplace(full, rects, dist, dst=0)
    r.begin_frame(scene, nodes)
    for v in mine:
        r.render_device(vps[v], stats=True)             # sizes the pools

    def timed():
        for _ in range(max(warmup, 3)):
            one()
        r.synchronize(); torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = []
        for _ in range(steps):
            flush.zero_()
            if world > 1:
                dist.all_reduce(start_token)            # stream-ordered, outside the events: the ranks start together
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); one(); e1.record()
            r.synchronize(); torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        t = torch.tensor([sum(ms) / len(ms)], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    ms_gather = timed()
    if rank == 0:
        full.zero_()
    gather_ok = screen_fnv_ok(r, torch, dist, world, rank, name, one)
    ms_peer, peer_ok = None, None
    if world > 1:
        sharding.share_screen(r, dist, dst=0, device="cuda")
        state["peer"] = True
        if rank == 0:
            full.zero_()
        ms_peer = timed()
        if rank == 0:
            full.zero_()
        peer_ok = screen_fnv_ok(r, torch, dist, world, rank, name, one)
        r.set_color_target(None)
        dist.barrier()
    best = min(m for m in (ms_gather, ms_peer) if m is not None)
    checks = [c for c in (gather_ok, peer_ok) if c is not None]
    return {"workload": name, "description": cfg["desc"], "ms_per_frame": best, "fps": 1e3 / best, "timed_frames": steps,
            "ms_per_frame_nccl_rect_gather": ms_gather if world > 1 else None, "ms_per_frame_peer_write": ms_peer,
            "frame_fnv_ok": (all(checks) if checks else None), "frame_fnv_ok_gather": gather_ok, "frame_fnv_ok_peer_write": peer_ok,
            "n_gpus": world, "active_gpus": min(world, len(vps)),
            "viewports_per_rank": [len(sharding.viewports_for_rank(len(vps), world, k)) for k in range(world)],
            "partition": "one viewport_t per GPU (viewport v on rank v mod N), scene replicated" if world > 1 else "single GPU, 4 viewports one after the other",
            "scaling": "strong"}


def sharded_frame(r, torch, dist, name, steps, warmup, flush, world, rank):
    """sort-first: ONE frame split in row bands over the ranks (scene replicated), bands gathered to rank 0 with
    NCCL over NVLink.  Strong scaling of a single frame; returns ms per frame (max over ranks)."""
    from swegl_b200 import configs, sharding
    scene, vps, screen, cfg = configs.build(name)
    vp = vps[0]
    r.upload_scene(scene)
    r.set_screen(*screen)
    bands = [sharding.band_rows(vp.h, world, k) for k in range(world)]
    nodes = scene.node_matrices()
    screen_ptr, _ = r.device_buffers()
    full = torch.as_tensor(_DevArray(screen_ptr, (screen[1], screen[0])), device="cuda")
    state = {}

    def set_bands(b):
        state["bands"] = b
        vp.band = b[rank] if world > 1 else (0, 0)
        state["desc"] = vp.desc()

    token = torch.zeros(1, device="cuda")

    def one():
        r.begin_frame(scene, nodes)
        r.render_device(state["desc"], stats=False)
        if world > 1:
            if state.get("peer"):
                sharding.frame_barrier(dist, token)     # the bands were written into rank 0's screen by the kernels
            else:
                sharding.gather_bands_inplace(full, vp.h, dist, dst=0, bands=state["bands"])   # NCCL send/recv into place
        return full
    set_bands(bands)
    r.begin_frame(scene, nodes)
    r.render_device(state["desc"], stats=True)          # sizes the pools for this workload
    one()
    if world > 1:
        # temporal coherence: rebalance the bands on the coverage of the frame just gathered (covered pixels per row
        # + a share for the per-row fixed work), computed on rank 0 and broadcast
        cuts = torch.zeros(world + 1, dtype=torch.int64, device="cuda")
        if rank == 0:
            r.synchronize(); torch.cuda.synchronize()
            cov = ((full >> 24) != 0).sum(dim=1).to(torch.float64) + 0.05 * screen[0]
            b = sharding.balanced_bands(cov.cpu().tolist(), world)
            cuts = torch.tensor([b[0][0]] + [y1 for _, y1 in b], dtype=torch.int64, device="cuda")
        dist.broadcast(cuts, src=0)
        c = [int(v) for v in cuts.cpu().tolist()]
        set_bands([(c[k], c[k + 1]) for k in range(world)])
        r.begin_frame(scene, nodes)
        r.render_device(state["desc"], stats=True)      # pools for the new band
    def timed():
        for _ in range(max(warmup, 3)):
            one()
        r.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = []
        for _ in range(steps):
            flush.zero_()
            if world > 1:
                dist.all_reduce(start_token)            # stream-ordered, no host wait: the ranks' frames start together on the GPUs
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            one()
            e1.record()
            r.synchronize()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        t = torch.tensor([sum(ms) / len(ms)], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    start_token = torch.zeros(1, device="cuda")
    # one GPU, whole frame: the denominator of the strong-scaling figure, measured by every rank on its own GPU in this very
    # run (the ranks do not share anything here); rank 0's number is reported
    ms_1gpu = None
    if world > 1:
        saved = (state["bands"], vp.band)
        vp.band = (0, 0)
        full_desc = vp.desc()
        r.begin_frame(scene, nodes)
        r.render_device(full_desc, stats=True)          # pools for the whole frame
        for _ in range(max(warmup, 3)):
            r.begin_frame(scene, nodes); r.render_device(full_desc, stats=False)
        r.synchronize(); torch.cuda.synchronize()
        ms1 = []
        for _ in range(steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); r.begin_frame(scene, nodes); r.render_device(full_desc, stats=False); e1.record()
            r.synchronize(); torch.cuda.synchronize()
            ms1.append(e0.elapsed_time(e1))
        ms_1gpu = sum(ms1) / len(ms1)
        set_bands(saved[0])
        r.begin_frame(scene, nodes)
        r.render_device(state["desc"], stats=True)
        dist.barrier()
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_timed = max(warmup, 3) + steps
    nv0 = nvlink_kib(local_rank) if world > 1 else None
    ms_gather = timed()                                 # bands sent to rank 0 after rendering (NCCL send/recv)
    nvl_gather = nvlink_delta(torch, dist, world, local_rank, nv0, n_timed) if world > 1 else None
    if rank == 0:
        full.zero_()
    gather_ok = screen_fnv_ok(r, torch, dist, world, rank, name, one)
    ms_peer, peer_ok, ms_sync, sync_ok, sync_info = None, None, None, None, None
    nvl_peer, nvl_sync = None, None
    if world > 1:
        # the fused form: every rank's last kernel stores its band into rank 0's screen over NVLink (CUDA IPC mapping)
        sharding.share_screen(r, dist, dst=0, device="cuda")
        state["peer"] = True
        if rank == 0:
            full.zero_()
        nv0 = nvlink_kib(local_rank)
        ms_peer = timed()
        nvl_peer = nvlink_delta(torch, dist, world, local_rank, nv0, n_timed)
        if rank == 0:
            full.zero_()
        peer_ok = screen_fnv_ok(r, torch, dist, world, rank, name, one)
        # the frame protocol (swegl_b200_set_frame_sync): no collective at all -- rank 0 clears the others' rows locally,
        # they store only the tiles they drew into and raise a flag in rank 0's memory; bands rebalanced on measured time
        state["peer"] = False
        state["sync"] = True
        sharding.arm_frame_sync(r, dist, dst=0)

        def one_sync():
            r.begin_frame(scene, nodes)
            r.render_device(state["desc"], stats=False)
            return full
        nonlocal_one = one_sync

        def own_time_ms():
            """this rank's own share of one frame: CUDA events around its chain (rank 0: the device-side stamps, which
            exclude its wait for the others)"""
            for _ in range(2):
                nonlocal_one()
            r.synchronize(); torch.cuda.synchronize(); dist.barrier()
            acc = 0.0
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); nonlocal_one(); e1.record()
                r.synchronize(); torch.cuda.synchronize()
                acc += (r.frame_sync_status()[1] if rank == 0 else e0.elapsed_time(e1)) / 5
                dist.barrier()
            return acc
        history = []
        for it in range(4):
            t_own = torch.tensor([own_time_ms()], device="cuda", dtype=torch.float64)
            ts = [torch.zeros_like(t_own) for _ in range(world)]
            dist.all_gather(ts, t_own)
            ts = [float(t.item()) for t in ts]
            history.append({"bands": state["bands"], "own_ms": [round(t, 4) for t in ts]})
            if it == 3:
                break
            nb = sharding.rebalance_bands(state["bands"], ts, damping=0.8)
            set_bands(nb)
            r.set_frame_sync(-1)                        # pools for the new band are sized outside the protocol
            r.begin_frame(scene, nodes)
            r.render_device(state["desc"], stats=True)
            sharding.arm_frame_sync(r, dist, dst=0)
        one_saved = one

        def timed_sync():
            for _ in range(max(warmup, 3)):
                nonlocal_one()
            r.synchronize(); torch.cuda.synchronize(); dist.barrier()
            ms = []
            for _ in range(steps):
                flush.zero_()
                dist.all_reduce(start_token)            # stream-ordered, outside the events: the ranks' frames start together
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); nonlocal_one(); e1.record()
                r.synchronize(); torch.cuda.synchronize()
                ms.append(e0.elapsed_time(e1))
            t = torch.tensor([sum(ms) / len(ms)], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        if rank == 0:
            full.zero_()
        nv0 = nvlink_kib(local_rank)
        ms_sync = timed_sync()
        nvl_sync = nvlink_delta(torch, dist, world, local_rank, nv0, n_timed)
        errs = torch.tensor([r.frame_sync_errors()], device="cuda", dtype=torch.int64)
        dist.all_reduce(errs)
        # the reference's frame?  (the bands moved, the frame must not; rank 0 does not clear its own band: zero it first)
        if rank == 0:
            r.synchronize(); torch.cuda.synchronize()
            full.zero_()
        dist.barrier()
        sync_ok = screen_fnv_ok(r, torch, dist, world, rank, name, nonlocal_one)
        if rank == 0:
            sync_ok = bool(sync_ok) and int(errs.item()) == 0
        sync_info = {"balance_history": history, "timed_out_waits": int(errs.item())}
        r.set_frame_sync(-1)
        r.set_color_target(None)
        state["sync"] = False
        dist.barrier()
    st = r.render_device(state["desc"], stats=True)
    # what the peer-memory variants put on NVLink by construction (rank 0's own band stays local): peer_write stores every
    # row of a band, background included; under the protocol only the colour of the tiles something was drawn into travels
    peer_bytes = None
    if world > 1:
        t = torch.tensor([int(st.n_busy_tiles)], device="cuda", dtype=torch.int64)
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        busy = [int(a.item()) for a in allt]
        bands_now = state["bands"]
        peer_bytes = {"busy_tiles_per_rank": busy,
                      "peer_write_into_rank0": sum((b[1] - b[0]) * vp.w * 4 for b in bands_now[1:]),
                      "peer_protocol_into_rank0": sum(n * 8 * 128 * 4 for n in busy[1:]),
                      "how": "algorithmic: rows x width x 4 B of the other ranks' bands / 4 KiB of colour per busy tile of theirs"}
    stage_ms, stage_frac = None, None
    if world == 1:                                      # the whole frame on one GPU: its stages against the roofline as well
        stage_ms, last_ = measure_kernels(r, scene, [vp], 5)
        stage_frac = stage_fracs(scene, vp, int(last_.n_covered), stage_ms)
    cands = [m for m in (ms_gather, ms_peer, ms_sync) if m is not None]
    best = min(cands)
    checks = [c for c in (gather_ok, peer_ok, sync_ok) if c is not None]
    if ms_1gpu is None:
        ms_1gpu = best
    return {"workload": name, "description": cfg["desc"], "ms_per_frame": best, "fps": 1e3 / best, "timed_frames": steps,
            "ms_per_frame_1gpu_same_run": ms_1gpu, "speedup_vs_1gpu": ms_1gpu / best,
            "ms_per_frame_nccl_gather": ms_gather if world > 1 else None,
            "ms_per_frame_peer_write": ms_peer, "ms_per_frame_peer_protocol": ms_sync, "peer_protocol": sync_info,
            "frame_fnv_ok": (all(checks) if checks else None), "frame_fnv_ok_gather": gather_ok, "frame_fnv_ok_peer_write": peer_ok,
            "frame_fnv_ok_peer_protocol": sync_ok,
            "nvlink": ({"nccl_gather": nvl_gather, "peer_write": nvl_peer, "peer_protocol": nvl_sync,
                        "note": "hardware counters (NVML NVLINK_THROUGHPUT_DATA_TX/RX, nvidia-smi nvlink -gt d); None = not exposed on this box"} if world > 1 else None),
            "peer_bytes_per_frame": peer_bytes, "ms_per_stage": stage_ms, "roofline_frac": stage_frac,
            "n_gpus": world,
            "partition": "row bands, sort-first, scene replicated, band culling (DESIGN.md 6: nccl_gather / peer_write / peer_protocol)" if world > 1 else "single GPU",
            "bands": state["bands"], "rank0_band_covered_pixels": int(st.n_covered), "scaling": "strong"}


def bind_to_gpu_numa_node(local_rank):
    """Pin this rank's host threads to the CPUs of its GPU's NUMA node BEFORE any page-locked memory is allocated, so
    that the pinned frame buffers of the end-to-end path land in the memory next to the GPU's PCIe root (with several
    ranks on one box the default placement sends most D2H copies across the socket link).  Returns the node or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(local_rank)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dev = "/sys/bus/pci/devices/" + bus[-12:].lower()                    # 00000000:1B:00.0 -> 0000:1b:00.0
        node = int(open(dev + "/numa_node").read())
        cpus = set()
        for part in open(dev + "/local_cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus |= set(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if node < 0 or not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true")
    ap.add_argument("--pipeline-depth", type=int, default=4,
                    help="contexts per GPU that render the independent frames of `value` round robin (1 = one frame at a time)")
    ap.add_argument("--multiview", default="multiview_4k", help="workload of the viewport-per-GPU split-screen measurement (config 4; '' = skip)")
    ap.add_argument("--sharded", default="sphere1000_8k", help="workload of the band-sharded single-frame measurement ('' = skip)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — swegl_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    numa_node = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    torch.cuda.set_device(local_rank)
    # stdout carries exactly one JSON line: anything a library prints there (NCCL's version banner) goes to stderr.
    # The original stdout is kept aside for that line.
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from swegl_b200 import Renderer
    # an explicit stream: torch's default stream is the legacy stream (handle 0), which the ABI reads as "use the
    # context's own stream" -- events recorded by torch would then not bracket the kernels
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    r = Renderer(local_rank, stream=stream.cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    main_res = gpu_workload(r, torch, args.workload, args.steps, args.warmup, flush, world, dist, depth=args.pipeline_depth, local_rank=local_rank)
    clocks = sampler.stop() if rank == 0 else None

    scene, vps, screen, cfg = main_res["scene"], main_res["vps"], main_res["screen"], main_res["cfg"]
    kern, last = measure_kernels(r, scene, vps, min(args.steps, 20))
    covered = int(last.n_covered)
    frames = args.steps * FRAMES_PER_STEP
    fps = world * frames / main_res["pipe_secs"]
    batch_fps = sorted(world * FRAMES_PER_STEP / b for b in main_res["batch_secs"])
    serial_fps = world * main_res["serial_frames"] / main_res["secs"]
    e2e_fps = world * main_res["e2e_frames"] / main_res["e2e_secs"]

    also = None
    if not args.no_also and args.workload == DEFAULT_WORKLOAD and world == 1:
        a = gpu_workload(r, torch, ALSO_WORKLOAD, args.steps, args.warmup, flush, world, dist, depth=args.pipeline_depth, local_rank=local_rank)
        ak, al = measure_kernels(r, a["scene"], a["vps"], min(args.steps, 20))
        also = {"workload": ALSO_WORKLOAD, "description": a["cfg"]["desc"], "fps": frames / a["pipe_secs"],
                "one_frame_at_a_time_fps": a["serial_frames"] / a["secs"],
                "e2e_fps": a["e2e_frames"] / a["e2e_secs"], "e2e_blocking_call_fps": a["e2e_sync_frames"] / a["e2e_sync_secs"],
                "e2e_d2h_bytes_per_frame": a["d2h"], "frame_fnv_ok": a.get("frame_fnv_ok"),
                "shaded_mpix_per_s": al.n_covered * a["serial_frames"] / a["secs"] / 1e6, "ms_per_stage": ak,
                "roofline_frac": stage_fracs(a["scene"], a["vps"][0], int(al.n_covered), ak)}

    # the remaining BASELINE.json configs that fit one GPU (config 1 and config 3), device-resident, one frame at a time
    others = None
    if not args.no_also and args.workload == DEFAULT_WORKLOAD and world == 1:
        others = {}
        for oname in ("box_640", "brainstem_4k_dof"):
            from swegl_b200 import configs as _cfg
            osc, ovps, oscreen, ocfg = _cfg.build(oname)
            r.upload_scene(osc)
            r.set_screen(*oscreen)
            n_ = 100
            osecs, _ = measure_gpu(r, torch, osc, ovps, oscreen, n_, args.warmup, flush)
            ok_, ol_ = measure_kernels(r, osc, ovps, 10)
            others[oname] = {"description": ocfg["desc"], "one_frame_at_a_time_fps": n_ / osecs, "ms_per_frame": 1e3 * osecs / n_,
                             "covered_pixels": int(ol_.n_covered), "shaded_mpix_per_s": ol_.n_covered * n_ / osecs / 1e6,
                             "ms_per_stage": ok_, "roofline_frac": stage_fracs(osc, ovps[0], int(ol_.n_covered), ok_)}

    sharded = None
    if args.sharded and not args.no_also:
        sharded = sharded_frame(r, torch, dist, args.sharded, SINGLE_FRAME_STEPS, args.warmup, flush, world, rank)
    multiview = None
    if args.multiview and not args.no_also:
        rm = Renderer(local_rank, stream=stream.cuda_stream)      # its own context: own screen, own peer mappings
        try:
            multiview = multiview_frame(rm, torch, dist, args.multiview, SINGLE_FRAME_STEPS, args.warmup, flush, world, rank)
        finally:
            rm.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ----
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"]); peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak = 6650.0; peak_src = "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"
    ab = algorithmic_bytes(scene, vps[0], covered)
    per_kernel = {"k_fragments": (ab["fragment"], kern["ms_fragment"])}
    if kern["ms_post"] > 0:
        per_kernel["k_dof"] = (ab["dof"], kern["ms_post"])
    dom = max(per_kernel, key=lambda k: per_kernel[k][1])
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(args.workload, {}).get(dom)

    def roof(k):
        b, ms_ = per_kernel[k]
        ach = b / (ms_ * 1e-3) / 1e9 if ms_ > 0 else 0.0
        return {"kernel": k, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "algorithmic_bytes": b, "ms": ms_, "peak_source": peak_src}
    roofline = roof(dom)
    roofline["traffic"] = traffic
    roofline["all_kernels"] = [roof(k) for k in per_kernel]

    # ---- CPU baseline (bounded sample of the same workload) ----
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, kind, times = cpu_reference_single(args.workload, 3, 1)
        from oracle.binding import Oracle                           # the checker, here only to COUNT the CPU path's fragments
        n_frag_cpu = sum(int(Oracle().render(scene, vp_, screen_wh=screen)["n_fragments"]) for vp_ in vps)
        cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": kind, "fragments_shaded_per_s": n_frag_cpu * v,
               "fragments_per_frame": n_frag_cpu,
               "sample": f"3 frames of {args.workload} after 1 warm-up, swegl::render as shipped (1 raster thread) "
                         f"via oracle/_ref + DoF-R by the C oracle; {1e3 / v:.1f} ms/frame"}

    checks = [c for c in (main_res.get("frame_fnv_ok"), (sharded or {}).get("frame_fnv_ok"), (multiview or {}).get("frame_fnv_ok")) if c is not None]
    line = {"metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * main_res["pipe_secs"] / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic" if args.workload.startswith("sphere") else "bundled scene",
            "config": workload_config(args.workload, cfg, scene, screen, n_gpus=world, depth=main_res["depth"]),
            "host_numa_binding": (f"rank 0 on node {numa_node}, every rank bound to its GPU's node" if numa_node is not None else "none"),
            "frames_timed": frames * world, "timed_region_s": main_res["pipe_secs"],
            "batches": {"n": len(batch_fps), "frames_each": FRAMES_PER_STEP, "fps_median": statistics.median(batch_fps),
                        "fps_min": batch_fps[0], "fps_max": batch_fps[-1]},
            "ms_per_frame": 1e3 * main_res["pipe_secs"] / frames,
            "shaded_mpix_per_s": covered * fps / 1e6, "viewport_mpix_per_s": sum(v.w * v.h for v in vps) * fps / 1e6,
            "covered_pixels": covered,
            "one_frame_at_a_time": {"fps": serial_fps, "ms_per_frame": 1e3 * main_res["secs"] / main_res["serial_frames"], "frames": main_res["serial_frames"],
                                    "how": "one context, frame i+1 starts when frame i is done: per-frame CUDA events on the "
                                           "launching stream, 256 MiB L2 flush between frames outside the events"},
            "frame_stats": {"setup_triangles": int(last.n_setup_triangles), "spans": int(last.n_spans), "chunks": int(last.n_chunks)},
            "e2e": {"value": e2e_fps, "unit": UNIT, "frames_per_step": E2E_FRAMES_PER_STEP,
                    "h2d_bytes_per_step": main_res["h2d"] * E2E_FRAMES_PER_STEP, "d2h_bytes_per_step": int(main_res["d2h"] * E2E_FRAMES_PER_STEP),
                    "h2d_bytes_per_frame": main_res["h2d"], "d2h_bytes_per_frame": int(main_res["d2h"]), "full_frame_bytes": main_res["d2h_full"],
                    "timed_region_s": main_res["e2e_secs"],
                    "api": "Renderer.begin_frame + render_async/wait (swegl_b200_render_viewport_async): 3 frames in flight, "
                           "every frame's node matrices/lights go H2D; D2H into pinned memory of the rectangle that differs from "
                           "what the host image already holds (the frame's bounding box united with the previous frame's; "
                           "d2h_bytes_* are the bytes the library counted, swegl_b200_readback_stats)",
                    "blocking_call_fps": world * main_res["e2e_sync_frames"] / main_res["e2e_sync_secs"],
                    "blocking_call_api": "Renderer.begin_frame + render (swegl_b200_render_viewport, what swegl::render maps to): "
                                         "returns with the whole frame in host memory (full copy)"},
            "gpu_launches": (int(last.n_launches) * len(vps) + 1) * frames,
            "frame_fnv_ok": (all(checks) if checks else None), "frame_fnv_ok_main": main_res.get("frame_fnv_ok"),
            "ms_per_stage": kern, "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks}
    if also:
        line["also"] = also
    if others:
        line["other_workloads"] = others
    if sharded:
        line["sharded_frame"] = sharded
    if multiview:
        line["multiview_frame"] = multiview
    # last key, short: what survives in a truncated tail of the line
    if sharded:
        line["strong"] = {"workload": sharded["workload"], "n_gpus": world, "ms_1gpu": round(sharded["ms_per_frame_1gpu_same_run"], 4),
                          "ms": round(sharded["ms_per_frame"], 4), "speedup": round(sharded["speedup_vs_1gpu"], 3),
                          "frames": sharded["timed_frames"], "fnv_ok": sharded["frame_fnv_ok"]}
    json_out.write(json.dumps(line) + "\n")
    json_out.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
