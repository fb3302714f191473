#!/usr/bin/env python
"""bench.py -- rendered views/sec (forward + backward) of the MVRenderer hot path on N B200s.

Workload (config.workload): BASELINE.json configs[1] -- mesh rendering of synthetic ~10k-face meshes,
batch 32 x 12 views per GPU, 224x224, Phong shading, forward + backward (gradients to azim/elev/dist).
`--workload points` switches to configs[2] (2048-pt clouds, alpha compositing) for exploration.

One JSON line on rank 0 (contract in the task statement):
  value      whole-job views/s (forward + backward) with inputs resident in HBM (device-timed, max over ranks)
  forward_only  the same for the forward pass alone
  e2e        same metric through MVRenderer.forward/backward from HOST buffers (H2D + D2H inside)
  roofline   dominant kernel: algorithmic bytes per launch / CUDA-event time vs measured HBM peak
  cpu_baseline  the CPU oracle timed on this box's host cores on a bounded sample (rank 0, N=1)
`--impl reference` times the reference's CPU implementation of the path: PyTorch3D cannot be installed
here (no network, not vendored), so this arm runs the oracle port with all host threads.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "rendered_views_per_sec_fwd_bwd"
UNIT = "views/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="mesh", choices=["mesh", "points"])
    ap.add_argument("--batch", type=int, default=32, help="objects per GPU")
    ap.add_argument("--views", type=int, default=12)
    ap.add_argument("--image-size", type=int, default=224)
    ap.add_argument("--faces", type=int, default=10000)
    ap.add_argument("--points", type=int, default=2048)
    ap.add_argument("--points-per-pixel", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cuda-graph", action="store_true",
                    help="replay the device-resident step from two captured CUDA graphs (mvtn_b200.graphs)")
    ap.add_argument("--cpu-sample-objects", type=int, default=0, help="0 = size the sample for ~10-20 s")
    return ap.parse_args()


def workload_name(a):
    if a.workload == "mesh":
        return (f"mesh fwd+bwd: {a.batch} objects/GPU x {a.views} views, ~{a.faces}-face synthetic meshes, "
                f"{a.image_size}x{a.image_size}, Phong, faces_per_pixel=1 (BASELINE configs[1])")
    return (f"points fwd+bwd: {a.batch} clouds/GPU x {a.views} learned_spherical views, {a.points} pts, "
            f"{a.image_size}x{a.image_size}, alpha compositing K={a.points_per_pixel} (BASELINE configs[2])")


# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ts, ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            inside = t0 - 0.05 <= ts <= t1 + 0.15
            try:
                if inside:
                    sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            if inside:
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------------
def make_inputs(a, rank):
    from mvtn_b200 import synth
    seed = 1236 + 1000 * rank
    if a.workload == "mesh":
        meshes = synth.make_meshes(a.batch, a.faces, seed)
        azim, elev, dist = synth.circular_views(a.batch, a.views)   # config.yaml:23-24 canonical 30 deg / 2.2
        return {"meshes": meshes, "views": (azim, elev, dist)}
    pts = synth.make_clouds(a.batch, a.points, seed + 1)
    return {"points": pts, "views": synth.learned_spherical_views(a.batch, a.views, seed + 2)}


def algorithmic_bytes_per_view(a, inp, which):
    """compulsory traffic, each tensor once, geometry counted once per view."""
    hw = a.image_size * a.image_size
    if a.workload == "mesh":
        V = sum(v.shape[0] for v, _ in inp["meshes"]) / len(inp["meshes"])
        F = sum(f.shape[0] for _, f in inp["meshes"]) / len(inp["meshes"])
        geo = 12 * V + 12 * F
        return geo + hw * (12 + 4)            # fwd: RGB + pix_to_face ; bwd: grad RGB + pix_to_face
    K = a.points_per_pixel
    return 12 * a.points + hw * (12 + 4 * K)


def run_ours(a):
    from mvtn_b200 import MVRenderer, Meshes, ops, parallel
    from mvtn_b200 import _lib as L
    rank, local_rank, world = parallel.init_distributed()
    if world != a.gpus and world > 1:
        raise SystemExit(f"--gpus {a.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    lib = L.load()
    lib.mvr_host_set_threads(max(1, (os.cpu_count() or 1) // max(world, 1)))     # ranks share the host cores
    inp = make_inputs(a, rank)
    B, M, S = a.batch, a.views, a.image_size
    N = B * M
    azim_h, elev_h, dist_h = (t.contiguous().pin_memory() for t in inp["views"])
    cot = torch.randn(N, 3, S, S, device=dev, generator=torch.Generator(device=dev).manual_seed(7 + rank)) / (3 * S * S)
    bg = torch.tensor([0.99999] * 3, device=dev)
    bg_black = torch.zeros(3, device=dev)
    obj = torch.tensor([0.99999] * 3, device=dev)
    light = torch.tensor([[0.0, 1.0, 0.0]], device=dev)

    if a.workload == "mesh":
        nv = [v.shape[0] for v, _ in inp["meshes"]]
        nf = [f.shape[0] for _, f in inp["meshes"]]
        verts_d = torch.cat([v for v, _ in inp["meshes"]]).to(dev)
        faces_d = torch.cat([f for _, f in inp["meshes"]]).to(dev)
        mesh_list = [Meshes([v], [f]) for v, f in inp["meshes"]]          # what run_mvtn.py's loader hands over (CPU)
        from mvtn_b200 import collate_meshes
        mesh_host = collate_meshes(mesh_list)     # the loader's collate_fn: one packed, pinned host batch
        renderer = MVRenderer(M, image_size=S, pc_rendering=False, light_direction="fixed").to(dev)
        renderer_pipe = MVRenderer(M, image_size=S, pc_rendering=False, light_direction="fixed", copy_stream=True).to(dev).train()
        kernels = ["mesh_scatter_kernel", "mesh_shade_kernel", "mesh_backward_kernel"]
    else:
        pts_d = inp["points"].to(dev)
        pts_h = inp["points"].pin_memory()
        renderer = MVRenderer(M, image_size=S, pc_rendering=True, points_per_pixel=a.points_per_pixel,
                              background_color="black", compositor="alpha", cuda_graph=a.cuda_graph).to(dev)
        tiled = a.points_per_pixel in (1, 2, 4, 8) and os.environ.get("MVR_POINTS_TILED", "1") != "0"
        kernels = (["points_bin_kernel", "points_tile_kernel", "points_backward_kernel"] if tiled
                   else ["points_scatter_kernel", "points_resolve_kernel", "points_backward_kernel"])
    renderer.train()
    azim_d, elev_d, dist_d = (t.to(dev) for t in (azim_h, elev_h, dist_h))

    graphed = None
    if a.cuda_graph:
        from mvtn_b200 import graphs
        sample = (azim_d, elev_d, dist_d)
        if a.workload == "mesh":
            geom_static = ops.PackedMeshes.from_packed(verts_d, faces_d, nv, nf)
            graphed = graphs.graphed_mesh_render(geom_static, M, light, obj, bg, S, sample)      # prepare is in the graph
        else:
            graphed = graphs.graphed_points_render(pts_d, obj, M, renderer.points_radius, bg * 0, S, sample,
                                                   points_per_pixel=a.points_per_pixel, compositor="alpha")

    def step_resident():
        """Hot path with inputs already in HBM: prepare + look_at + forward + backward + look_at backward."""
        az = azim_d.detach().requires_grad_(); el = elev_d.detach().requires_grad_(); di = dist_d.detach().requires_grad_()
        if graphed is not None:
            img = graphed(az, el, di)
            img.backward(cot)
            return az.grad, el.grad, di.grad
        if a.workload == "mesh":
            R, T, C, _bad = ops._LookAt.apply(az.reshape(-1), el.reshape(-1), di.reshape(-1))
            geom = ops.PackedMeshes.from_packed(verts_d, faces_d, nv, nf)
            img, _ = ops.render_meshes(geom, M, R, T, C, light, obj, bg, S)
        else:
            img, _cams, _ = ops.render_points_from_angles(pts_d, obj, M, az, el, di, renderer.points_radius, bg_black, S,
                                                          points_per_pixel=a.points_per_pixel, compositor="alpha")
        img.backward(cot)
        return az.grad, el.grad, di.grad

    def step_forward_only():
        """Forward pass alone (inference: render_and_save, evaluation loops), inputs resident, no autograd graph."""
        with torch.no_grad():
            R, T, C, _bad = ops._LookAt.apply(azim_d, elev_d, dist_d)
            if a.workload == "mesh":
                geom = ops.PackedMeshes.from_packed(verts_d, faces_d, nv, nf)
                img, _ = ops.render_meshes(geom, M, R, T, C, light, obj, bg, S)
            else:
                img, _ = ops.render_points(pts_d, obj, M, R, T, None, renderer.points_radius, bg_black, S,
                                           points_per_pixel=a.points_per_pixel, compositor="alpha", dist=dist_d)
        return img

    g_host = torch.empty(3, B, M, pin_memory=True)

    def step_e2e(list_api=False):
        """Through the public API from HOST buffers: H2D of the step's inputs (pinned host memory: the collated mesh
        batch / the point tensor + the view tensors), render, backward, D2H of the result (gradients w.r.t.
        azim/elev/dist).  list_api=True hands MVRenderer the reference's python list of per-object CPU meshes instead,
        so the multi-threaded gather into pinned memory is inside the timed region too."""
        az = azim_h.to(dev, non_blocking=True).requires_grad_()
        el = elev_h.to(dev, non_blocking=True).requires_grad_()
        di = dist_h.to(dev, non_blocking=True).requires_grad_()
        if a.workload == "mesh":
            img, _ = renderer(mesh_list if list_api else mesh_host, None, az, el, di)
        else:
            img, _ = renderer(None, pts_h, az, el, di)
        img.backward(cot.view_as(img))
        g_host[0].copy_(az.grad, non_blocking=True)
        g_host[1].copy_(el.grad, non_blocking=True)
        g_host[2].copy_(di.grad, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return g_host

    g_ring = [torch.empty(3, B, M, pin_memory=True) for _ in range(2)]
    ev_ring = [torch.cuda.Event() for _ in range(2)]
    ring = {"i": 0}

    def step_e2e_pipelined():
        """step_e2e without the per-step device sync: the gradients of step i travel to a pinned ring buffer behind an
        event and are read on the host while step i+1 is in flight (what a training loop that logs with one step of lag
        does).  Every step still pays its own H2D and D2H inside the timed region; only the wait moves."""
        i = ring["i"]; ring["i"] = i + 1
        az = azim_h.to(dev, non_blocking=True).requires_grad_()
        el = elev_h.to(dev, non_blocking=True).requires_grad_()
        di = dist_h.to(dev, non_blocking=True).requires_grad_()
        if a.workload == "mesh":
            img, _ = renderer_pipe(mesh_host, None, az, el, di)
        else:
            img, _ = renderer(None, pts_h, az, el, di)
        img.backward(cot.view_as(img))
        g = g_ring[i & 1]
        g[0].copy_(az.grad, non_blocking=True)
        g[1].copy_(el.grad, non_blocking=True)
        g[2].copy_(di.grad, non_blocking=True)
        ev_ring[i & 1].record()
        if i > 0:
            ev_ring[(i - 1) & 1].synchronize()
        return g_ring[(i - 1) & 1]

    if a.workload == "mesh":
        # verts fp32 + faces narrowed to int32 by the multi-threaded host gather + the three (B, M) view tensors
        h2d = sum(v.numel() * 4 + f.numel() * 4 for v, f in inp["meshes"]) + 3 * B * M * 4
    else:
        h2d = pts_h.numel() * 4 + 3 * B * M * 4
    d2h = 3 * B * M * 4 + 4    # gradients + the rotation-validity flag

    def timed(fn, steps, profile=None):
        parallel.barrier(); torch.cuda.synchronize()
        if profile:
            lib.mvr_profile_enable(profile.encode())
        l0 = lib.mvr_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.time()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        w1 = time.time()
        parallel.barrier()
        ms = e0.elapsed_time(e1)
        prof = None
        if profile:
            import ctypes
            tot, n = ctypes.c_double(0), ctypes.c_int(0)
            L.check(lib.mvr_profile_collect(ctypes.byref(tot), ctypes.byref(n)), "mvr_profile_collect")
            prof = (tot.value, n.value)
        return ms, lib.mvr_launch_count() - l0, prof, (w0, w1)

    # warm-up (also sizes workspaces / staging buffers)
    for _ in range(max(a.warmup, 3)):
        step_resident()
        step_forward_only()
    for _ in range(max(a.warmup, 3)):
        step_e2e()
        step_e2e_pipelined()
        if a.workload == "mesh":
            step_e2e(list_api=True)
    torch.cuda.synchronize()
    # which kernel dominates?  one profiled step per candidate
    shares = {}
    for k in kernels:
        _, _, prof, _ = timed(step_resident, 2, profile=k)
        shares[k] = prof[0] / max(prof[1], 1)
    top = max(shares, key=shares.get)

    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    ms, launches, prof, (w0, w1) = timed(step_resident, a.steps, profile=top)
    ms_fwd, _, _, _ = timed(step_forward_only, a.steps)
    ms_e2e, _, _, (w2, w3) = timed(step_e2e, a.steps)
    ms_e2e_pipe, _, _, (_, w3) = timed(step_e2e_pipelined, a.steps)
    ms_e2e_list = None
    if a.workload == "mesh":
        ms_e2e_list, _, _, (_, w3) = timed(lambda: step_e2e(list_api=True), a.steps)
    clocks = sampler.stop(w0, w3)

    ms_max = parallel.max_over_ranks(ms, dev)
    ms_fwd_max = parallel.max_over_ranks(ms_fwd, dev)
    ms_e2e_max = parallel.max_over_ranks(ms_e2e, dev)
    ms_e2e_pipe_max = parallel.max_over_ranks(ms_e2e_pipe, dev)
    ms_e2e_list_max = parallel.max_over_ranks(ms_e2e_list, dev) if ms_e2e_list is not None else None
    total_views = parallel.sum_over_ranks(N * a.steps, dev)
    value = total_views / (ms_max / 1e3)
    e2e_value = total_views / (ms_e2e_max / 1e3)

    peak, peak_src = measured_peaks()
    per_view = algorithmic_bytes_per_view(a, inp, top)
    k_ms = prof[0] / max(prof[1], 1)
    achieved = (per_view * N) / (k_ms / 1e3) / 1e9 if k_ms > 0 else 0.0
    traffic = None
    try:   # DRAM bytes per launch of this kernel from the committed ncu capture (only valid for the default workload)
        default_mesh = a.workload == "mesh" and (B, M, S, a.faces) == (32, 12, 224, 10000)
        default_points = a.workload == "points" and (B, M, S, a.points, a.points_per_pixel) == (32, 12, 224, 2048, 4)
        if default_mesh or default_points:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[top]["bytes"]
    except Exception:
        traffic = None
    roofline = {"bound": "hbm", "kernel": top, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": int(per_view * N), "kernel_ms": round(k_ms, 4),
                "kernel_ms_all": {k: round(v, 4) for k, v in shares.items()}}

    out = {"metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": a.steps,
           "warmup": max(a.warmup, 3), "ms_per_step": round(ms_max / a.steps, 4), "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": workload_name(a), "objects_per_gpu": B, "views": M, "image_size": S,
                      "cuda_graph": bool(a.cuda_graph),
                      "l2": "inputs_exceed_l2 (images + cotangent + pix_to_face > 126 MB per step)"},
           "e2e": {"value": round(e2e_value, 1), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                   "ms_per_step": round(ms_e2e_max / a.steps, 4),
                   "input": ("collated pinned host batch (mvtn_b200.collate_meshes) + pinned view tensors" if a.workload == "mesh"
                             else "pinned host point tensor + pinned view tensors")},
           "forward_only": {"value": round(total_views / (ms_fwd_max / 1e3), 1), "unit": UNIT, "ms_per_step": round(ms_fwd_max / a.steps, 4)},
           "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline}

    out["e2e"]["pipelined"] = {"value": round(total_views / (ms_e2e_pipe_max / 1e3), 1), "ms_per_step": round(ms_e2e_pipe_max / a.steps, 4),
                               "note": "same per-step H2D / D2H, but the result of step i is awaited on the host during step i+1 (no per-step device sync); mesh batches go through MVRenderer(copy_stream=True)"}
    if ms_e2e_list_max is not None:
        out["e2e"]["list_api"] = {"value": round(total_views / (ms_e2e_list_max / 1e3), 1), "ms_per_step": round(ms_e2e_list_max / a.steps, 4),
                                  "input": "python list of per-object CPU meshes (the reference's loader output); gather into pinned memory inside the timed region"}
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        out["cpu_baseline"], out["parity"] = cpu_baseline(a, inp)
    if rank == 0:
        emit(out)
    if world > 1:
        torch.distributed.destroy_process_group()


# ---------------------------------------------------------------------------------------------------
def oracle_step(a, inp, n_obj, keep=None):
    """One forward+backward of the CPU oracle on the first n_obj objects of the workload.  Returns seconds; `keep` (a dict)
    receives the cameras, fragment indices, images and gradients for the parity report."""
    import numpy as np
    from oracle import oracle as orc
    from mvtn_b200 import ops
    M, S = a.views, a.image_size
    az, el, di = (t[:n_obj].reshape(-1).numpy() for t in inp["views"])
    t0 = time.time()
    R, T, C = orc.look_at(az, el, di)
    if a.workload == "mesh":
        ms = inp["meshes"][:n_obj]
        vp = np.concatenate([v.numpy() for v, _ in ms]); fp = np.concatenate([f.numpy() for _, f in ms]).astype(np.int32)
        voff = np.cumsum([0] + [v.shape[0] for v, _ in ms]).astype(np.int32)
        foff = np.cumsum([0] + [f.shape[0] for _, f in ms]).astype(np.int32)
        nrm = orc.packed_vertex_normals(vp, fp, voff, foff)
        k00, k11 = ops.fov_projection_scale()
        rgb = np.full(3, 0.99999, np.float32); light = np.array([[0, 1.0, 0]], np.float32)
        o = orc.mesh_forward(vp, fp, voff, foff, nrm, rgb, M, R, T, C, light, rgb, k00, k11, 0.5, S, S, 1,
                             orc.PERSPECTIVE_CORRECT, fragments=False)
        g = np.full((n_obj * M, 3, S, S), 1.0 / (3 * S * S), np.float32)
        b = orc.mesh_backward(vp, fp, voff, foff, nrm, rgb, M, R, T, C, light, k00, k11, S, S, 1,
                              orc.PERSPECTIVE_CORRECT, o["pix_to_face"], g)
        gv = orc.look_at_backward(az, el, di, b["gR"], b["gT"], b["gC"])
        if keep is not None:
            keep.update(R=R, T=T, C=C, index=o["pix_to_face"], images=o["images"], gR=b["gR"], gT=b["gT"], gC=b["gC"], g_views=gv)
    else:
        pts = inp["points"][:n_obj].numpy()
        rgb = np.full(3, 0.99999, np.float32)
        inv = (1.0 / di).astype(np.float32)
        K = a.points_per_pixel
        o = orc.points_forward(pts, rgb, M, R, T, inv, 0.006, np.zeros(3, np.float32), S, S, K, orc.COMPOSITE_ALPHA,
                               fragments=False)
        g = np.full((n_obj * M, 3, S, S), 1.0 / (3 * S * S), np.float32)
        b = orc.points_backward(pts, rgb, M, R, T, inv, 0.006, S, S, K, orc.COMPOSITE_ALPHA, o["idx"], g)
        gv = orc.look_at_backward(az, el, di, b["gR"], b["gT"], np.zeros_like(b["gT"]))      # the orthographic path does not use C
        if keep is not None:
            keep.update(R=R, T=T, C=C, index=o["idx"], images=o["images"], gR=b["gR"], gT=b["gT"], g_scale=b.get("g_inv_dist"), g_views=gv)
    return time.time() - t0


def cpu_baseline(a, inp):
    from oracle import oracle as orc
    cores = os.cpu_count() or 1
    orc.set_num_threads(cores)
    n = a.cpu_sample_objects
    if n <= 0:
        t1 = oracle_step(a, inp, 1)                     # probe: one object
        n = max(1, min(a.batch, int(12.0 / max(t1, 1e-3))))
    keep = {}
    t = oracle_step(a, inp, n, keep)
    base = {"value": round(n * a.views / t, 2), "unit": UNIT, "cores": orc.num_threads(), "kind": "port",
            "sample": f"{n} object(s) x {a.views} views of the same workload, fwd+bwd, {t:.2f} s of CPU work "
                      f"(oracle/mvr_oracle.c, OpenMP over image rows)"}
    return base, parity_report(a, inp, n, keep)


def parity_report(a, inp, n_obj, ref):
    """the parity gates that go with every throughput number.  The CUDA path renders the objects the CPU
    baseline has just rendered (same constant cotangent) and is compared with the oracle's outputs in the two stages of
    the test protocol: (A) rasterizer / shader / compositor and their backward from the SAME cameras -- fragment indices
    bit-exact, images and camera gradients within tolerance; (B) the camera kernels on their own -- look_at forward
    against the oracle's R, T, C and look_at backward fed with the oracle's camera gradients.  (An end-to-end comparison
    through DIFFERENT cameras would measure how many edge pixels a 1e-7 change of R flips, not the kernels.)
    Exact depth ties are counted from the K-list (meshes: a K = 2 render)."""
    import numpy as np
    import torch
    from mvtn_b200 import ops
    dev = torch.device("cuda", torch.cuda.current_device())
    M, S = a.views, a.image_size
    N = n_obj * M
    cot = torch.full((N, 3, S, S), 1.0 / (3 * S * S), device=dev)
    col = torch.tensor([0.99999] * 3, device=dev)
    Rd, Td, Cd = (torch.from_numpy(ref[k]).to(dev).requires_grad_() for k in ("R", "T", "C"))
    views = [t[:n_obj].reshape(-1).to(dev).requires_grad_() for t in inp["views"]]

    def rel(x, y):
        x = x.detach().cpu().numpy().reshape(-1); y = np.asarray(y).reshape(-1)
        return float(np.abs(x - y).max() / max(np.abs(y).max(), 1e-30))

    out = {"sample": f"the cpu_baseline sample ({n_obj} object(s) x {M} views)", "pixels": int(N * S * S)}
    if a.workload == "mesh":
        ms = inp["meshes"][:n_obj]
        geom = ops.PackedMeshes([v for v, _ in ms], [f for _, f in ms], dev)
        light = torch.tensor([[0.0, 1.0, 0.0]], device=dev)
        img, fr = ops.render_meshes(geom, M, Rd, Td, Cd, light, col, col, S)
        img.backward(cot)
        index = fr["pix_to_face"]
        _, fr2 = ops.render_meshes(geom, M, Rd.detach(), Td.detach(), Cd.detach(), light, col, col, S, faces_per_pixel=2, fragments=True)
        zb = fr2["zbuf"]
        ties = int(((zb[..., 0] == zb[..., 1]) & (fr2["pix_to_face"][..., 1] >= 0)).sum())
        g_cam = max(rel(Rd.grad, ref["gR"]), rel(Td.grad, ref["gT"]), rel(Cd.grad, ref["gC"]))
        tol = {"index": "bit-exact", "images_abs": 1e-5, "gradients_rel": 1e-4, "look_at_abs": 2e-6}
    else:
        pts = inp["points"][:n_obj].to(dev)
        inv = (1.0 / views[2].detach()).requires_grad_()
        img, fr = ops.render_points(pts, col, M, Rd, Td, inv, 0.006, col * 0, S, points_per_pixel=a.points_per_pixel,
                                    compositor="alpha", fragments=True)
        img.backward(cot)
        index = fr["idx"]
        zb = fr["zbuf"]
        ties = int(((zb[..., 1:] == zb[..., :-1]) & (index[..., 1:] >= 0)).sum()) if zb.shape[-1] > 1 else 0
        g_cam = max(rel(Rd.grad, ref["gR"]), rel(Td.grad, ref["gT"]), rel(inv.grad, ref["g_scale"]))
        tol = {"index": "bit-exact", "images_abs": 1e-5, "gradients_rel": 1e-5, "look_at_abs": 2e-6}
    out["index_mismatches"] = int((index.cpu().numpy() != ref["index"]).sum())
    out["exact_depth_ties"] = ties
    out["image_max_abs_err"] = round(float(np.abs(img.detach().cpu().numpy() - ref["images"]).max()), 9)
    out["grad_camera_max_rel_err"] = round(g_cam, 9)
    # stage B: the camera kernels
    R2, T2, C2, _ = ops._LookAt.apply(*views)
    out["look_at_max_abs_err"] = round(max(float((R2.detach().cpu() - torch.from_numpy(ref["R"])).abs().max()),
                                           float((T2.detach().cpu() - torch.from_numpy(ref["T"])).abs().max()),
                                           float((C2.detach().cpu() - torch.from_numpy(ref["C"])).abs().max())), 9)
    gC_ref = ref.get("gC")
    loss = (R2 * torch.from_numpy(ref["gR"]).to(dev)).sum() + (T2 * torch.from_numpy(ref["gT"]).to(dev)).sum()
    if gC_ref is not None:
        loss = loss + (C2 * torch.from_numpy(gC_ref).to(dev)).sum()
    loss.backward()
    # one scale for the three view gradients: with orthographic cameras d/d dist through the cameras is identically ~0
    out["look_at_backward_max_rel_err"] = round(rel(torch.cat([v.grad for v in views]), np.concatenate([np.asarray(g) for g in ref["g_views"]])), 9)
    out["tolerance"] = tol
    return out


def run_reference(a):
    """Reference arm: the reference's own CPU implementation of the path.  PyTorch3D is not vendored and
    cannot be installed offline, so this is the oracle port with all host threads (kind "port")."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    from oracle import oracle as orc
    cores = os.cpu_count() or 1
    orc.set_num_threads(cores)
    inp = make_inputs(a, 0)
    t1 = oracle_step(a, inp, 1)
    n = max(1, min(a.batch, int(3.0 / max(t1, 1e-3))))      # ~3 s of CPU work per step
    for _ in range(min(a.warmup, 1)):
        oracle_step(a, inp, n)
    steps = max(1, min(a.steps, 5))
    t0 = time.time()
    for _ in range(steps):
        oracle_step(a, inp, n)
    dt = time.time() - t0
    v = round(n * a.views * steps / dt, 2)
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": steps,
           "warmup": min(a.warmup, 1), "ms_per_step": round(dt / steps * 1e3, 2), "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": workload_name(a), "objects_per_gpu": a.batch, "views": a.views, "image_size": a.image_size},
           "cpu_baseline": {"value": v, "unit": UNIT, "cores": orc.num_threads(), "kind": "port",
                            "sample": f"{n} object(s) x {a.views} views per step, fwd+bwd (PyTorch3D is not installable "
                                      f"offline: oracle port of its CPU path)"},
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    emit(out)


def emit(obj):
    """The ONE JSON line of the contract goes to the process's original stdout; everything libraries print meanwhile
    (NCCL's version banner, torchrun notices) has been diverted to stderr."""
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


_REAL_STDOUT = 1

if __name__ == "__main__":
    args = parse()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
