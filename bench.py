#!/usr/bin/env python
"""bench.py -- frames/s and Gfragments/s of the raster path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[4], the configuration the metric is quoted on): synthetic
icosphere, class-I frequency 708 -> 10,025,280 triangles, rendered at 3840x2160 in image mode
along the reference's 64-frame turntable angle sequence (main.rs:55-58,92-96).  A "step" is one
frame: update + clear + draw_mesh over the whole scene (main.rs:78-83).

* value     device-timed frames/s, scene resident in HBM (uploaded once, like the reference loads
            its meshes once), results left on the device (sloth_render_device_batch: geometry of
            frame k+1 overlaps the resolve of frame k); CUDA events on the library's stream;
            inputs (401 MB of triangles) exceed L2, so no flush is needed between steps.
* e2e       the same frames through the public C-ABI call with HOST buffers
            (sloth_render_batch: per frame the rotation goes host->device, the cell buffer comes
            back device->host into pinned memory), wall clock, max over ranks.
* roofline  dominant kernel (k_tri on the indexed path, k_geom3 on the soup path; it visits every triangle
            of the scene): algorithmic 40 B/triangle / its average duration (CUDA events recorded inside
            the library around that kernel, one frame per sample, same frames as the timed region)
            against the measured HBM peak; `frame_frac` is the whole frame (40 N + 4 W H bytes over the
            batch time per frame) against the same peak.  `roofline_resolve`: the write-out kernel on 4 bytes per
            cell, with the fraction its measured DRAM traffic reaches beside it.
* post_check  after the timed batch the last frame it left on the device is compared with a single
            sloth_render_device of the same rotation (and hashed): the timed path is verified at size.
* cpu_baseline / --impl reference: the CPU oracle (oracle/sloth_oracle.c, a port: the Rust reference
            cannot be built in this image), timed on a bounded sample (every S-th triangle) and
            scaled to frames/s.  The reference is single-threaded, so `value` is the ONE-thread figure;
            a frames-across-cores run is reported beside it (`all_cores`).  That arm touches nothing of
            the product library: rotations and angles come from the oracle.
Multi-GPU: frames are independent -> rank r renders frames r, r+N, ... ("weak": K frames per rank).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FREQ, WIDTH, HEIGHT, TURNTABLE_FRAMES = 708, 3840, 2160, 64
METRIC = "frames/s at 3840x2160, 10M-triangle icosphere (Gfragments/s alongside)"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def make_scene(freq):
    return icosphere_soup(freq)


def rotations(n):
    """main.rs:55-58,76-77,92-96 through the product's host helpers (the b200 arm)."""
    import rust_sloth_b200 as rs
    pitches = rs.turntable_pitches(0.0, n)
    return np.stack([rs.rotation_from_euler(0.0, p, 0.0) for p in pitches])


def rotations_oracle(n):
    """The same sequence from the oracle (the reference arm must not load the product library)."""
    import oracle
    return np.stack([oracle.rotation(0.0, p, 0.0) for p in oracle.turntable(0.0, n)])


def workload_config(n_tri, freq, world):
    """`config` of the JSON line -- one function for both arms, so the two dicts are identical."""
    return {"workload": f"icosphere f={freq} ({n_tri} triangles) at {WIDTH}x{HEIGHT}, image mode, "
                        f"{TURNTABLE_FRAMES}-frame turntable angles",
            "l2": "inputs exceed L2 (10 M triangles streamed per frame), no flush",
            "sharding": "independent frames per GPU, no collective" if world > 1 else "single GPU"}


def icosphere_soup(freq):
    """meshes.icosphere without importing the product package's __init__ (reference arm)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_sloth_meshes", os.path.join(ROOT, "rust-sloth_b200", "meshes.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    t = time.time()
    xyz, rgb, s0 = mod.icosphere(freq)
    log(f"[bench] icosphere f={freq}: {len(xyz)} triangles in {time.time() - t:.1f}s")
    return xyz, rgb, s0


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def __enter__(self):
        if os.environ.get("BENCH_NO_CLOCKS"):   # diagnostic only
            return self
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=3)
            except Exception:
                self.proc.kill()

    def samples_since(self, t0):
        return sum(1 for t, line in list(self.lines) if t >= t0 and line.count(",") >= 8)

    def summary(self, t0, t1):
        sm, mx, reasons = [], [], set()
        for t, line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                if t0 - 0.15 <= t <= t1 + 0.15:
                    sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            if t0 - 0.15 <= t <= t1 + 0.15:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        if not sm:
            sm = [float(x.split(",")[1]) for _, x in self.lines[-3:] if len(x.split(",")) > 2] or [0.0]
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx) if mx else 0.0),
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_sample(xyz, rgb, s0, rot, target_s=12.0):
    """Oracle (mode 0 = the reference's scan domain) on every S-th triangle, S chosen for ~target_s."""
    import oracle
    t = time.perf_counter()
    oracle.render(xyz[:0], rgb[:0], s0, WIDTH, HEIGHT, rot, mode=0)
    t_clear = time.perf_counter() - t
    stride = 1024
    t = time.perf_counter()
    oracle.render(xyz, rgb, s0, WIDTH, HEIGHT, rot, mode=0, tri_first=0, tri_step=stride)
    t_probe = max(time.perf_counter() - t - t_clear, 1e-4)
    while stride > 1 and t_probe * (1024 / (stride // 2)) < target_s:
        stride //= 2
    return stride, t_clear


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU algorithm (oracle port).  Each timed step is one frame sample
    (every S-th triangle over the reference's full scan domain) on ONE thread -- the reference is single-threaded;
    afterwards one frames-across-cores pass (a sample per host thread) gives the all-cores figure."""
    if rank != 0:
        return
    import oracle
    from concurrent.futures import ThreadPoolExecutor
    xyz, rgb, s0 = icosphere_soup(args.freq)
    rots = rotations_oracle(TURNTABLE_FRAMES)
    threads = max(1, len(os.sched_getaffinity(0)))
    # size a step so that warmup+steps (one thread) plus the all-cores pass stay within ~2.5 minutes
    step_target = min(args.ref_step_seconds, 120.0 / max(1, args.steps + args.warmup + 2))
    stride, t_clear = cpu_sample(xyz, rgb, s0, rots[0], target_s=step_target)
    log(f"[bench/reference] every {stride}-th triangle per frame sample, clear {t_clear * 1e3:.0f} ms, {threads} host threads")

    def one(i):
        _, _, cnt = oracle.render(xyz, rgb, s0, WIDTH, HEIGHT, rots[i % len(rots)], mode=0,
                                  tri_first=i % stride, tri_step=stride)
        return cnt["covered"]

    k = 0
    for _ in range(args.warmup):
        one(k)
        k += 1
    t0 = time.perf_counter()
    covered = 0
    for _ in range(args.steps):
        covered += one(k)
        k += 1
    t_step = (time.perf_counter() - t0) / args.steps
    t_frame = t_clear + stride * max(t_step - t_clear, 1e-9)     # one full frame on one thread
    fps = 1.0 / t_frame
    frags_per_frame = covered * stride / args.steps
    # all host cores: one frame sample per thread, concurrently (frames are independent)
    pool = ThreadPoolExecutor(threads)
    list(pool.map(one, range(k, k + threads)))
    k += threads
    t0 = time.perf_counter()
    list(pool.map(one, range(k, k + threads)))
    t_par = time.perf_counter() - t0
    fps_all = threads / (t_clear + stride * max(t_par - t_clear, 1e-9))
    sample = (f"each step = one frame sample on one thread, a sample = every {stride}-th triangle of the "
              f"{len(xyz)}-triangle frame over the reference's full scan domain; "
              f"frames/s = 1 / (t_clear + {stride} * (t_step - t_clear))")
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_frame * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "gfragments_per_s": frags_per_frame * fps / 1e9,
        "config": workload_config(len(xyz), args.freq, world),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": 1, "kind": "port", "sample": sample},
        "all_cores": {"value": fps_all, "unit": "frames/s", "cores": threads,
                      "sample": f"{threads} concurrent frame samples, one per host thread, same sampling"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_b200(args, rank, local_rank, world):
    import torch
    import rust_sloth_b200 as rs

    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)

    xyz, rgb, s0 = make_scene(args.freq)
    n_tri = len(xyz)
    rots = rotations(TURNTABLE_FRAMES)
    ctx = rs.Context.blank(True, device=local_rank)
    ctx.set_scene(xyz, rgb, s0)
    ctx.resize(WIDTH, HEIGHT)
    cpf = ctx.cells_per_frame()
    d_cells = torch.empty(cpf + 2, dtype=torch.int32, device=f"cuda:{local_rank}")
    stream = torch.cuda.ExternalStream(ctx.stream_ptr(), device=f"cuda:{local_rank}")
    K, Wm = args.steps, args.warmup
    my_frames = [(rank + world * i) % len(rots) for i in range(K + Wm)]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # fragments per frame (untimed; the counter adds atomics)
    ctx.stats_enable(count_fragments=True)
    frags = []
    for f in my_frames[Wm:Wm + min(K, 4)]:
        ctx.render_device(rots[f], d_cells.data_ptr())
        frags.append(ctx.stats()["fragments"])
    frags_per_frame = float(np.mean(frags))
    ctx.stats_enable()

    # ---- value: device-timed, results stay on the device ----------------------------------
    # One sloth_render_device_batch call for the K timed frames: inside it the geometry of frame k+1
    # (issue-bound) overlaps the resolve of frame k (memory-bound) on a second stream.
    warm_rots = np.stack([rots[f] for f in my_frames[:Wm]])
    timed_rots = np.stack([rots[f] for f in my_frames[Wm:]])
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:   # started before the warm-up: nvidia-smi needs a moment to come up
        ctx.render_device_batch(warm_rots, d_cells.data_ptr(), 0)
        ctx.sync()
        barrier()
        launches0 = ctx.stats()["kernel_launches"]
        t_wall0 = time.time()
        ev0.record(stream)
        ctx.render_device_batch(timed_rots, d_cells.data_ptr(), 0)
        ev1.record(stream)
        ctx.sync()
        launches_timed = ctx.stats()["kernel_launches"] - launches0   # kernels of this library, counted at launch
        d_timed = d_cells[:cpf].clone()   # what the timed batch left behind (the load loop below overwrites d_cells)
        # keep the GPU in the same state a little longer so the 100 ms sampler sees the load (until it has
        # delivered a few samples: with 8 ranks starting nvidia-smi at once the first line can take a second)
        t_busy = time.time()
        while time.time() - t_busy < 0.5 or (clk.proc and clk.samples_since(t_wall0) < 3 and time.time() - t_busy < 6.0):
            ctx.render_device_batch(timed_rots[:8], d_cells.data_ptr(), 0)
            ctx.sync()
        t_wall1 = time.time()
    barrier()
    dev_ms = ev0.elapsed_time(ev1)

    # ---- post-check: the frame the timed batch left on the device == a single render of that rotation ----
    import hashlib
    d_check = torch.empty(cpf + 2, dtype=torch.int32, device=f"cuda:{local_rank}")
    ctx.render_device(timed_rots[-1], d_check.data_ptr())
    ctx.sync()
    same = bool(torch.equal(d_timed, d_check[:cpf]))
    post_check = {"last_timed_frame_equals_single_render": same,
                  "cells_sha256": hashlib.sha256(d_timed.cpu().numpy().tobytes()).hexdigest(),
                  "frame": int(my_frames[-1])}
    del d_check, d_timed

    # ---- roofline: per-kernel events, one frame per sample --------------------------------
    ctx.stats_enable(kernel_timing=True)
    geom_ms, xform_ms, walk_ms, resolve_ms, frame_ms = [], [], [], [], []
    for f in my_frames[Wm:]:
        ctx.render_device(rots[f], d_cells.data_ptr())
        st = ctx.stats()
        geom_ms.append(st["geom_ms"]); walk_ms.append(st["walk_ms"]); xform_ms.append(st["xform_ms"])
        resolve_ms.append(st["resolve_ms"]); frame_ms.append(st["last_frame_ms"])
    last = ctx.stats()
    ctx.stats_enable()
    indexed = last["geom_path"] == rs.PATH_INDEXED
    kernel_name = "k_tri" if indexed else "k_geom3"

    # ---- bare device->host ceiling: the same bytes per frame, nothing else running on this GPU (all ranks at once)
    d2h_pin = torch.empty(cpf, dtype=torch.int32, pin_memory=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(2):
        d2h_pin.copy_(d_cells[:cpf], non_blocking=True)
    barrier()
    e0.record()
    for _ in range(16):
        d2h_pin.copy_(d_cells[:cpf], non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    d2h_ms = e0.elapsed_time(e1) / 16
    del d2h_pin

    # ---- e2e: public API, host buffers ------------------------------------------------------
    # Two passes over the same frames through sloth_render_batch: the plain wire (4-byte cells over PCIe) and the span
    # wire (run lists over PCIe, the cells rebuilt by host threads inside the library); both leave the same bytes in
    # the caller's buffer, which is checked on the last frame.
    ring = max(1, min(K, 64))   # page-locked frames (2.1 GB at 4K); a larger K reuses them, 64 frames per call
    pinned = rs.PinnedBuffer(ring * cpf)
    my_rots = np.stack([rots[f] for f in my_frames[Wm:]])
    last_slot = (K - 1) % ring

    def e2e_pass(wire):
        ctx.set_wire(wire)
        ctx.render_batch(my_rots[:min(K, 3)], pinned.array)
        ctx.set_wire(wire)        # resets the wire statistics after the warm-up call
        barrier()
        t = time.perf_counter()
        for i in range(0, K, ring):
            ctx.render_batch(my_rots[i:i + ring], pinned.array)
        dt = time.perf_counter() - t
        barrier()
        digest = hashlib.sha256(pinned.array[last_slot * cpf:(last_slot + 1) * cpf].tobytes()).hexdigest()
        return dt, digest, ctx.wire_stats()

    e2e_cells_s, cells_digest, _ = e2e_pass(rs.WIRE_CELLS)
    e2e_s, spans_digest, wire_stats = e2e_pass(rs.WIRE_SPANS)
    ctx.set_wire(rs.WIRE_CELLS)
    post_check["e2e_last_frame_same_on_both_wires"] = cells_digest == spans_digest
    post_check["e2e_last_frame_equals_device_frame"] = cells_digest == post_check["cells_sha256"]

    times = torch.tensor([dev_ms, e2e_s * 1e3, d2h_ms, e2e_cells_s * 1e3], dtype=torch.float64, device=f"cuda:{local_rank}")
    if dist is not None:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms_max, e2e_ms_max, d2h_ms_max, e2e_cells_ms_max = float(times[0]), float(times[1]), float(times[2]), float(times[3])

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            hbm_peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            hbm_peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        g_ms = float(np.mean(geom_ms))
        achieved = 40.0 * n_tri / (g_ms * 1e-3) / 1e9
        # DRAM traffic of the dominant kernel: not measurable outside ncu, so it is the figure of the committed
        # `ncu --set full` capture, labelled with the commit it was taken at (profiles/traffic.json)
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            traffic = tj.get(f"{kernel_name}_dram_bytes_per_launch")
            traffic_src = f"{tj.get('capture', 'ncu --set full')} at commit {tj.get('commit', '?')}"
        r_ms = float(np.mean(resolve_ms))
        r_traffic = json.load(open(tpath)).get("k_resolve_even_dram_bytes_per_launch") if os.path.exists(tpath) else None
        fps = K * world / (dev_ms_max * 1e-3)
        e2e_fps = K * world / (e2e_ms_max * 1e-3)
        e2e_cells_fps = K * world / (e2e_cells_ms_max * 1e-3)
        d2h_ceiling_fps = world / (d2h_ms_max * 1e-3)
        line = {
            "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": dev_ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "gfragments_per_s": frags_per_frame * fps / 1e9,
            "fragments_per_frame": frags_per_frame,
            "config": workload_config(n_tri, args.freq, world),
            # headline: the span wire (SLOTH_WIRE_SPANS, one sloth_ctx_set_wire call); d2h bytes are what the library
            # counted for rank 0's frames (run lists + their lengths, plain cells for frames without runs)
            "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": 64,
                    "d2h_bytes_per_step": wire_stats["d2h_bytes"] / max(wire_stats["frames"], 1),
                    "wire": "spans (run lists over PCIe, 4-byte cells rebuilt into the caller's buffer by "
                            f"{wire_stats['threads']} host threads per rank)",
                    "frames_sent_as_plain_cells": wire_stats["plain_frames"],
                    "host_bytes_written_per_step": cpf * 4,
                    "gfragments_per_s": frags_per_frame * e2e_fps / 1e9},
            # the same call over the plain wire, against its own roofline: the measured device->host rate of those bytes
            "e2e_plain_cells": {"value": e2e_cells_fps, "unit": "frames/s", "h2d_bytes_per_step": 64, "d2h_bytes_per_step": cpf * 4,
                                "d2h_gbs_measured": cpf * 4 * world / (d2h_ms_max * 1e-3) / 1e9,
                                "d2h_ceiling_frames_per_s": d2h_ceiling_fps,
                                "frac_of_d2h_ceiling": e2e_cells_fps / d2h_ceiling_fps},
            "roofline": {"bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": 40 * n_tri, "avg_launch_ms": g_ms,
                         "frame_bytes": 40 * n_tri + 4 * WIDTH * HEIGHT,
                         "frame_frac": (40.0 * n_tri + 4.0 * WIDTH * HEIGHT) / (dev_ms_max / K * 1e-3) / 1e9 / hbm_peak,
                         "kernel_ms": {kernel_name: g_ms, "k_xform": float(np.mean(xform_ms)),
                                       "k_tail": float(np.mean(walk_ms)),
                                       "k_resolve": float(np.mean(resolve_ms)),
                                       "frame_unoverlapped": float(np.mean(frame_ms))},
                         "geometry_path": "indexed" if indexed else "soup", "unique_vertices": int(last["n_vert"])},
            # the write-out kernel against the same peak: 4 bytes per cell are its algorithmic bytes; what it actually
            # moves is the 8-byte key per slot in (and back out where a fragment landed) plus the cells
            "roofline_resolve": {"bound": "hbm", "kernel": "k_resolve_even", "achieved": 4.0 * WIDTH * HEIGHT / (r_ms * 1e-3) / 1e9,
                                 "peak": hbm_peak, "unit": "GB/s", "frac": 4.0 * WIDTH * HEIGHT / (r_ms * 1e-3) / 1e9 / hbm_peak,
                                 "algorithmic_bytes_per_launch": 4 * WIDTH * HEIGHT, "avg_launch_ms": r_ms,
                                 "traffic": r_traffic, "traffic_frac": (r_traffic / (r_ms * 1e-3) / 1e9 / hbm_peak) if r_traffic else None,
                                 "traffic_source": traffic_src},
            "post_check": post_check,
            "gpu_launches": int(launches_timed),
            "clocks": clk.summary(t_wall0, t_wall1),
            "last_frame_stats": {k: last[k] for k in ("walk_tris", "walk_items", "irregular_tris", "stamp_fixups")},
        }
        if args.cpu_baseline and world == 1:
            import oracle
            stride, t_clear = cpu_sample(xyz, rgb, s0, rots[0], target_s=args.cpu_seconds)
            t = time.perf_counter()
            _, _, cnt = oracle.render(xyz, rgb, s0, WIDTH, HEIGHT, rots[my_frames[Wm]], mode=0, tri_first=0, tri_step=stride)
            t_s = time.perf_counter() - t
            t_frame = t_clear + stride * max(t_s - t_clear, 1e-9)
            line["cpu_baseline"] = {
                "value": 1.0 / t_frame, "unit": "frames/s", "cores": 1, "kind": "port",
                "sample": f"every {stride}-th triangle of one frame ({t_s:.1f}s measured) over the reference's full scan "
                          f"domain, scaled: t_frame = t_clear + {stride}*(t_sample - t_clear) = {t_frame:.1f}s; "
                          f"host has {len(os.sched_getaffinity(0))} cores, the reference is single-threaded",
                "gfragments_per_s": cnt["covered"] * stride / t_frame / 1e9}
        print(json.dumps(line), flush=True)
    pinned.free()
    ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    # Only the final JSON line may reach stdout: libraries (NCCL prints its version banner there when
    # NCCL_DEBUG is set in the environment) write to fd 1 behind Python's back, so fd 1 is pointed at
    # stderr for the whole run and the saved descriptor is used for the one line that matters.
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    global print
    _print = print

    def print(*a, **k):  # noqa: A001 - shadows the builtin on purpose for the JSON line
        if k.get("file") is None:
            os.write(saved_stdout, (" ".join(str(x) for x in a) + "\n").encode())
        else:
            _print(*a, **k)

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--freq", type=int, default=FREQ)
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--ref-step-seconds", type=float, default=4.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
