#!/usr/bin/env python
"""bench.py — gorender hot path on B200: Mtriangles/s and FPS on the 200k-triangle
sphere at 1280x720 (BASELINE.json configs[2], "C3"), beside the CPU reference arm.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of the reference

A "step" renders `--frames` consecutive frames of the demo spin (main.go:229-233:
Rotation.Y += 0.01 per frame) of the C3 scene, issued as batched Draws of `--batch`
frames, every frame of a batch into its own device framebuffer.  N > 1 is frame-parallel
(SURVEY.md §8e): each rank renders its own `--frames` frames per step, no data-path
collective, weak scaling.  One JSON line is printed by rank 0.

  value        whole-job Mtriangles/s (submitted scene faces x frames / time), scene and
               framebuffers resident in HBM, CUDA-event timed, max over ranks
  e2e          the same metric through the C-ABI calls with HOST buffers: per step the pinned-host ->
               device copy of the per-frame matrices and every frame's pixels + z-buffer kept exact in pinned
               host memory by host mirrors (only the tiles that changed cross PCIe), overlapped with rendering
  roofline     per SURVEY.md §8(d): the dominant kernel's share of the compulsory bytes per launch / its
               CUDA-event time, vs the measured HBM copy bandwidth (MEASURED_PEAKS.json); both kernels listed,
               with their ncu DRAM traffic; the whole path's fraction under path_frac
  latency      one frame per call through the literal drop-in call (grb_draw_present: Draw + host framebuffer)
  other_configs  C1, C2 pose B and C4 (single GPU) with their own kernel times, fractions and CPU baselines
  strips       (N > 1) BASELINE.json configs[3]: the 2M-triangle 3840x2160 frame as sort-first strips, every
               rank's raster kernel writing its rows into rank 0's framebuffer over NVLink (parallel.StripGroup)
  cpu_baseline the oracle's threaded restatement of the reference timed on this host
"""
from __future__ import annotations

import argparse
import glob
import json
import os
import shutil
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOAD = ("C3: 200k-triangle geodesic sphere (n=100), 1280x720, flat shading, untextured, default camera, "
            "demo spin (BASELINE.json configs[2])")
WIDTH, HEIGHT = 1280, 720
SPHERE_N = 100            # 20*n^2 = 200 000 faces, 10*n^2+2 = 100 002 vertices
METRIC = "Mtriangles/s (submitted scene triangles x FPS), 200k-tri mesh @1280x720"
UNIT = "Mtri/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=8192, help="frames per step (8192 x 20 steps: a timed region of ~1.4 s)")
    ap.add_argument("--batch", type=int, default=64, help="frames per batched Draw call")
    ap.add_argument("--cpu-sample-frames", type=int, default=2000, help="frames of the CPU baseline sample (~14 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the latency / other_configs / strips / PCIe-ceiling legs")
    ap.add_argument("--mode", default="frames", choices=["frames", "strips"],
                    help="frames: frame-parallel C3 batches (the headline metric, with the strips leg inside for N > 1); "
                         "strips: only the sort-first strips of the 2M-triangle 3840x2160 C4 frame (strong scaling)")
    ap.add_argument("--strips-exchange", default="peer", choices=["peer", "nccl", "none"],
                    help="peer: raster kernels store into rank 0's framebuffer over NVLink (StripGroup); nccl: round 1's "
                         "grouped send/recv gather")
    ap.add_argument("--strips-balance", default="busy", choices=["busy", "equal"])
    ap.add_argument("--strips-frames-per-call", type=int, default=8,
                    help="consecutive frames of the animation one strip draw call renders (every frame split across all ranks)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.proc = None
        self.lines = []
        if shutil.which("nvidia-smi") is None:
            return
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        lines = self.lines
        if t0 is not None:
            inside = [(ts, ln) for ts, ln in lines if t0 <= ts <= t1 + 0.05]
            # a very short timed region may fall between two samples: take the nearest ones
            lines = inside or sorted(lines, key=lambda x: abs(x[0] - 0.5 * (t0 + t1)))[:2]
        for ts, ln in lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def ncu_summary(pattern: str):
    """The newest committed per-kernel ncu summary matching profiles/<pattern> (written from an `ncu --set full`
    capture by scripts/ncu_summary.py), or None."""
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", pattern)))
    if not files:
        return None
    d = json.load(open(files[-1]))
    d["file"] = os.path.relpath(files[-1], ROOT)
    return d


def build_scene():
    from gorender_b200 import workloads

    objs, cam = workloads.config_c3(SPHERE_N)
    return objs, cam


def spin_frames(first: int, count: int):
    from gorender_b200 import geometry

    return geometry.spin_rotations(count, start=first)


def bench_config(args, world: int):
    """`config` of the JSON line — the same dictionary for both arms (the driver compares them)."""
    B = max(1, min(args.batch, args.frames))
    F = (args.frames // B) * B
    return {
        "workload": WORKLOAD, "frames_per_step": F, "frames_per_draw_call": B, "parallelism": f"frame-parallel x{world}",
        "streams_per_gpu": 2,
        "l2": f"no flush needed: every batched draw writes {B} x 7.4 MB of framebuffers and ~{B * 12} MB of "
              "intermediates, far more than the 126 MB L2",
        "published_reference": "README.md:10-13: ~10 Mtps HUD metric / ~100 FPS on an Intel MacBook Pro",
    }


# ---------------------------------------------------------------- reference arm

def run_reference(args, rank: int, world: int):
    """The reference's own CPU implementation of the path.  The Go toolchain does not exist in
    this image, so this is the oracle's restatement in the reference's threaded structure (one
    projection task per object, 16 tile raster tasks on a pool: renderer.go:145-156,452-465),
    g++ -O3 with the SSE transform of asm_amd64.s.  Each step is a bounded sample of the step's
    workload: 64 consecutive frames of the same spin."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_binding import Oracle
    import scene_defs

    orc = Oracle()
    objs, cam = build_scene()
    r = scene_defs.SceneDef(WIDTH, HEIGHT, objs, cam).renderer(None)
    threads = 16  # numTiles workers (renderer.go:151-155)
    frames = max(1, min(args.frames, 64))  # bounded sample per step: 64 consecutive frames (~0.45 s of CPU)
    nfaces = sum(len(o.Mesh.Faces) for o in objs)
    nsteps = args.warmup + args.steps
    timer = orc.sequence_timer(r, objs, [cam] * (nsteps * frames), spin_frames(0, nsteps * frames), threads=threads)
    total = 0.0
    tpf = 0
    for s in range(nsteps):
        sec, tpf = timer.run(s * frames, frames)
        if s >= args.warmup:
            total += sec
    timer.close()
    fps = args.steps * frames / total
    value = fps * nfaces / 1e6
    sample = (f"{frames} consecutive demo-spin frames of the step's workload per step x {args.steps} steps, C++ restatement of "
              "the reference (oracle/) in its threaded structure; Go is not installed in this image")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "fps": fps, "mtps_hud": fps * tpf / 1e6,
        "config": bench_config(args, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": os.cpu_count(), "threads": threads, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------- helpers of the b200 arm

def cpu_sample(orc, renderer, objs, cams, rots, threads=16, warm=3):
    """Seconds per frame of the oracle (threaded reference structure) over len(cams) frames."""
    timer = orc.sequence_timer(renderer, objs, cams, rots, threads=threads)
    timer.run(0, min(len(cams), warm))
    sec, _ = timer.run(0, len(cams))
    timer.close()
    return sec / len(cams), sec


def kernel_roofline(ktimes_ms, alg_bytes, frames_per_launch, hbm_peak, ncu=None):
    """Per-kernel roofline entries: SURVEY §8(d) compulsory bytes of the kernel x frames per launch / its time."""
    out = {}
    for k, ms in ktimes_ms.items():
        if k not in alg_bytes or ms <= 0:
            continue
        b = alg_bytes[k] * frames_per_launch
        e = {"algorithmic_bytes_per_launch": b, "ms_per_launch": ms, "achieved": b / (ms * 1e-3) / 1e9,
             "frac": b / (ms * 1e-3) / 1e9 / hbm_peak}
        nk = (ncu or {}).get("kernels", {}).get(k) if ncu else None
        if nk and nk.get("frames_per_launch") == frames_per_launch:
            e["traffic"] = nk["dram_bytes_total"]
            e["traffic_over_algorithmic"] = nk["dram_bytes_total"] / b
            e["ncu"] = {kk: nk.get(kk) for kk in ("issue_active_pct", "dram_throughput_pct", "sm_throughput_pct", "warps_active_pct",
                                                   "warp_instructions", "fp32_pipe_fma_pct", "l1_hit_pct", "registers")}
        out[k] = e
    return out


def timed_batches(g, dev, renderer, packed_list, reps):
    """CUDA-event time (ms per draw call) and per-kernel times of `reps` passes over the batched draws in packed_list."""
    import torch

    for p in packed_list:
        renderer.draw_packed(p, 0, sync=False)
    dev.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st = torch.cuda.current_stream()
    e0.record(st)
    for _ in range(reps):
        for p in packed_list:
            renderer.draw_packed(p, 0, sync=False)
    e1.record(st)
    dev.synchronize()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (reps * len(packed_list))
    dev.set_kernel_timing(True)
    dev.kernel_times()
    for p in packed_list:
        renderer.draw_packed(p, 0, sync=False)
    dev.synchronize()
    kt, _ = dev.kernel_times()
    dev.set_kernel_timing(False)
    kt = {k: v / len(packed_list) for k, v in kt.items() if k in ("setup", "raster") and v > 0}   # (the other slots are empty event pairs)
    return ms, kt


def other_configs(args, g, dev, stream, hbm_peak, orc):
    """C1, C2 pose B, C4 on one GPU: device-resident batched draws (CUDA events), per-kernel times, §8(d) fractions,
    and the CPU restatement on a bounded sample of the same frames."""
    import torch
    from gorender_b200 import workloads

    out = {}
    specs = [
        ("c1", "C1: models/suzanne.obj (967 faces), 1280x720, default options, demo spin (BASELINE.json configs[0])",
         lambda: workloads.config_c1(), 1280, 720, 64, 300),
        ("c2b", "C2 pose B: models/cube.obj + textures-16.png, camera (0.6,0.3,1.7), clipped, 804k textured pixels per frame "
                "(BASELINE.json configs[1]); static pose", lambda: workloads.config_c2("B"), 1280, 720, 64, 60),
        ("c4", "C4: 10 x textured Gouraud 200k-triangle spheres (2.0 M faces), 3840x2160, one GPU (BASELINE.json configs[3])",
         lambda: workloads.config_c4(SPHERE_N), 3840, 2160, 8, 6),
    ]
    for key, desc, build, w, h, B, cpu_frames in specs:
        objs, cam = build()
        nfaces = sum(len(o.Mesh.Faces) for o in objs)
        meshes = {id(o.Mesh): o.Mesh for o in objs}.values()
        inst = len(objs) // max(len(meshes), 1)
        nverts = sum(len(m.Vertices) for m in meshes)
        nvn = sum(len(m.VertexNormals) for m in meshes)
        mfaces = sum(len(m.Faces) for m in meshes)
        textured = any(len(m.Faces.Textures) for m in meshes)
        rots = None
        with torch.cuda.stream(stream):
            fb = g.FrameBuffer(w, h, B, dev)
            r = g.Renderer(fb)
            if key == "c2b":
                packed = [np.ascontiguousarray(r.pack_objects(objs, [cam] * B))]
            elif key == "c4":
                base = [o.Rotation.copy() for o in objs]
                rot = spin_frames(0, B)
                rows = []
                for f in range(B):
                    for o, b0 in zip(objs, base):
                        o.Rotation = np.array([b0[0], np.float32(b0[1] + rot[f]), b0[2]], dtype=np.float32)
                    rows.append(r.pack_objects(objs, [cam]))
                for o, b0 in zip(objs, base):
                    o.Rotation = b0
                packed = [np.ascontiguousarray(np.concatenate(rows, axis=0))]
            else:
                rots = spin_frames(0, max(4 * B, cpu_frames))
                packed = [np.ascontiguousarray(r.pack_objects(objs, [cam] * B, rots[i * B:(i + 1) * B])) for i in range(4)]
            # enough repetitions for ~0.3 s of GPU time
            ms0, _ = timed_batches(g, dev, r, packed, 1)
            reps = int(max(2, min(400, 300.0 / max(ms0 * len(packed), 1e-3))))
            ms, kt = timed_batches(g, dev, r, packed, reps)
            stats = np.zeros(B, dtype=g._cabi.STATS_DTYPE)
            dev.check(dev.lib.grb_frame_stats_read(dev.h, B, stats.ctypes.data))
            fb.close()
        fps = B / (ms * 1e-3)
        # B = 16 Nv + 12 Nf + 16 Nf [+ 16 Nvn + 12 Nf if Gouraud] [+ 24 Nf + 4 Nf if textured] + 8 W H, per instance
        per_inst = 16.0 * nverts + 28.0 * mfaces
        if nvn:
            per_inst += 16.0 * nvn + 12.0 * mfaces
        if textured:
            per_inst += 28.0 * mfaces
        scene_bytes = per_inst * inst
        fb_bytes = 8.0 * w * h
        alg = {"setup": scene_bytes, "raster": fb_bytes}
        ncu = ncu_summary(f"r*_ncu_{key}_*.json")
        rec = {
            "workload": desc, "value": fps * nfaces / 1e6, "unit": UNIT, "fps": fps, "us_per_frame": ms * 1e3 / B,
            "frames_per_draw_call": B, "timed_draw_calls": reps * len(packed),
            "kernel_ms_per_launch": kt,
            "roofline": {"bound": "hbm", "unit": "GB/s", "peak": hbm_peak,
                         "path_bytes_per_frame": scene_bytes + fb_bytes,
                         "path_frac": (scene_bytes + fb_bytes) * fps / 1e9 / hbm_peak,
                         "kernels": kernel_roofline(kt, alg, B, hbm_peak, ncu), "ncu_file": ncu["file"] if ncu else None},
            "frame_stats": {"triangles_rasterised": float(stats["triangles"].mean()), "tpf": float(stats["tpf"].mean()),
                            "out_of_domain": int(stats["out_of_domain"].sum()), "list_fallbacks": int(stats["list_fallbacks"].sum())},
        }
        if orc is not None:
            import scene_defs

            rr = scene_defs.SceneDef(w, h, objs, cam).renderer(None)
            spf, sec = cpu_sample(orc, rr, objs, [cam] * cpu_frames, rots[:cpu_frames] if rots is not None else None,
                                  warm=1 if key == "c4" else 3)
            rec["cpu_baseline"] = {"value": nfaces / spf / 1e6, "unit": UNIT, "fps": 1.0 / spf, "cores": os.cpu_count(), "threads": 16,
                                   "kind": "port", "sample": f"{cpu_frames} frames of the same scene ({sec:.1f} s)"}
        out[key] = rec
    return out


def latency_leg(g, dev, stream, objs, cam, frames=3000):
    """One frame per call through grb_draw_present (Draw + host framebuffer + stats, one synchronisation per frame,
    CUDA-graph replay): the reference's contract is one Draw per frame (main.go:201-208)."""
    import ctypes as C
    import torch

    with torch.cuda.stream(stream):
        fb = g.FrameBuffer(WIDTH, HEIGHT, 1, dev)
        r = g.Renderer(fb)
        n = 512
        packed = np.ascontiguousarray(r.pack_objects(objs, [cam] * n, spin_frames(0, n)))
        p = r.draw_params(None)
        stats = np.zeros(1, dtype=g._cabi.STATS_DTYPE)
        nfaces = sum(len(o.Mesh.Faces) for o in objs)
        res = {}
        for label, zb in (("pixels_and_z", fb.mirror("ZBuffer")), ("pixels_only", None)):
            color = fb.mirror("Pixels")
            lib, h, stride, nobj = dev.lib, dev.h, packed.strides[0], packed.shape[1]

            def one(i):
                rc = lib.grb_draw_present(h, fb.handle, 0, 1, C.c_void_p(packed.ctypes.data + (i % n) * stride), nobj, C.byref(p),
                                          color.h, 0, zb.h if zb is not None else None, 0, C.c_void_p(stats.ctypes.data))
                if rc:
                    dev.check(rc)

            for i in range(30):
                one(i)
            w0 = color.stats()
            r0 = dev.graph_replays()
            t0 = time.perf_counter()
            for i in range(frames):
                one(i)
            sec = time.perf_counter() - t0
            w1 = color.stats()
            res[label] = {"fps": frames / sec, "us_per_frame": sec / frames * 1e6, "value": frames / sec * nfaces / 1e6, "unit": UNIT,
                          "frames": frames, "graph_replays": dev.graph_replays() - r0,
                          "tiles_written_frac": (w1[0] - w0[0]) / max(w1[1] - w0[1], 1)}
        res["checksum"] = int(fb.Pixels.sum())
        res["call"] = ("grb_draw_present: matrices H2D + setup + raster + host-mirror update + stats D2H as one CUDA graph, one "
                       "stream synchronisation per frame; C3 spin, 1280x720")
        fb.close()
    return res


def pcie_ceiling(torch, dist, world, seconds=0.4):
    """Platform ceiling of the end-to-end leg: every rank DMA-copies device frames into pinned host memory at the same
    time, nothing else running.  Returns aggregate GB/s (bytes of all ranks / max time)."""
    nbytes = 64 * WIDTH * HEIGHT * 4
    src = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    dst = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(2, int(seconds * 50e9 / nbytes))
    e0.record()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    return world * reps * nbytes / (ms * 1e-3) / 1e9


# ---------------------------------------------------------------- sort-first strips (C4)

def strips_leg(args, g, torch, dist, rank, world, local_rank, frames_timed=256, warm=16, fpc=None):
    """BASELINE.json configs[3]: 10 textured Gouraud spheres (2.0 M faces) at 3840x2160 as sort-first strips.  Total work is
    fixed as N grows (strong scaling).  `peer`: rank 0's two framebuffers are shared over CUDA IPC, every rank's raster
    kernel stores its rows into them over NVLink, device-side flags hand each frame over (parallel.StripGroup), strips
    balanced by the busy tiles of a probe frame.  `nccl`: round 1's grouped send/recv gather of torch-owned framebuffers."""
    from gorender_b200 import parallel, workloads

    W4, H4 = 3840, 2160
    objs, cam = workloads.config_c4(SPHERE_N)
    nfaces = sum(len(o.Mesh.Faces) for o in objs)
    stream = torch.cuda.Stream()
    dev = g.Device(local_rank, stream.cuda_stream)
    FR = 4                                  # distinct frames of the spin, cycled
    fpc = fpc or args.strips_frames_per_call  # consecutive frames one call renders (each split across all ranks)
    exchange = args.strips_exchange if world > 1 else "peer"
    with torch.cuda.stream(stream):
        probe = g.FrameBuffer(W4, H4, 1, dev)
        pr = g.Renderer(probe)
        rots = spin_frames(0, FR)
        packed = []
        base_rot = [o.Rotation.copy() for o in objs]
        for f in range(FR):
            for o, b in zip(objs, base_rot):  # every instance spins from its own start angle
                o.Rotation = np.array([b[0], np.float32(b[1] + rots[f]), b[2]], dtype=np.float32)
            packed.append(np.ascontiguousarray(pr.pack_objects(objs, [cam])))
        for o, b in zip(objs, base_rot):
            o.Rotation = b
        pr.draw_packed(packed[0], 0)
        # one call = fpc consecutive frames
        calls = [np.ascontiguousarray(np.concatenate([packed[(c * fpc + i) % FR] for i in range(fpc)], axis=0)) for c in range(FR)]
        flags = probe.tile_flags(0)
        rows = (parallel.balanced_strip_rows(flags.sum(axis=1), world, H4) if args.strips_balance == "busy"
                else [parallel.strip_rows(H4, world, r) for r in range(world)])
        busy_tiles = int(flags.sum())
        probe.close()
        # the same calls on ONE GPU (rank 0 alone, whole frames), measured in this run: what the strips are compared with
        single = None
        if world > 1:
            if rank == 0:
                sfb = g.FrameBuffer(W4, H4, fpc, dev)
                sr = g.Renderer(sfb)
                nc = max(2, min(frames_timed, 64) // fpc)
                for c in range(3):
                    sr.draw_packed(calls[c % FR], 0, sync=False)
                dev.synchronize()
                s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s0.record(stream)
                for c in range(nc):
                    sr.draw_packed(calls[c % FR], 0, sync=False)
                s1.record(stream)
                dev.synchronize()
                torch.cuda.synchronize()
                sms = s0.elapsed_time(s1) / (nc * fpc)
                spx, sz = sfb.read(fpc - 1, 1)
                single = {"ms_per_frame": sms, "fps": 1e3 / sms, "value": 1e3 / sms * nfaces / 1e6,
                          "checksum_last_frame": int(spx[0].astype(np.uint64).sum()), "last_call": int((nc - 1) % FR)}
                sfb.close()
            dist.barrier()

        if exchange == "none":      # diagnostic: every rank draws its strip into its own framebuffer, nothing is exchanged
            lfbs = [g.FrameBuffer(W4, H4, fpc, dev) for _ in range(2)]
            lrs = [g.Renderer(fb) for fb in lfbs]

            def one_frame(n):
                y0, y1 = rows[rank]
                if y1 > y0:
                    lrs[n & 1].draw_packed(calls[n % FR], 0, rows=(y0, y1) if world > 1 else None, sync=False)

            def finish():
                pass
        elif exchange == "peer":
            # if any rank cannot map rank 0's framebuffers (CUDA IPC refused by the platform), every rank leaves the leg together
            grp, err = None, ""
            try:
                grp = parallel.StripGroup(dev, W4, H4, nbuf=2, rows=rows, frames=fpc)
            except Exception as e:  # noqa: BLE001
                err = f"{type(e).__name__}: {e}"
            if dist is not None:
                okt = torch.tensor([0 if grp is None else 1], device="cuda", dtype=torch.int32)
                dist.all_reduce(okt, op=dist.ReduceOp.MIN)
                if int(okt[0]) == 0:
                    if grp is not None:
                        grp.close_local()
                    dev.close()
                    return {"error": "strip group could not be set up on every rank" + (": " + err if err else "")} if rank == 0 else None
            elif grp is None:
                raise RuntimeError(err)

            def one_frame(n):
                k = n & 1
                grp.draw(k, calls[n % FR])
                grp.release(k)     # rank 0 consumes nothing here: the buffer is free as soon as the frame is complete

            def finish():
                pass
        else:
            comm = torch.cuda.Stream()
            tfbs = [parallel.TorchFrameBuffer(W4, H4, fpc, dev, torch.device("cuda", local_rank)) for _ in range(2)]
            rs = [g.Renderer(t.fb) for t in tfbs]
            gathered = [None, None]

            def one_frame(n):
                k = n & 1
                if gathered[k] is not None:
                    stream.wait_event(gathered[k])        # do not overwrite a strip still being sent
                y0, y1 = rows[rank]
                if y1 > y0:
                    rs[k].draw_packed(calls[n % FR], 0, rows=(y0, y1), sync=False)
                comm.wait_stream(stream)
                with torch.cuda.stream(comm):
                    for f in range(fpc):
                        parallel.gather_strips_to_rank0(tfbs[k].color[f], tfbs[k].depth[f], H4, rows=rows)
                    gathered[k] = torch.cuda.Event()
                    gathered[k].record(comm)

            def finish():
                stream.wait_stream(comm)

        def barrier():
            dev.synchronize()
            torch.cuda.synchronize()
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize()

        n = 0
        for _ in range(max(2, warm // fpc)):
            one_frame(n)
            n += 1
        finish()
        barrier()
        l0 = dev.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ncalls = max(1, frames_timed // fpc)
        frames_timed = ncalls * fpc
        for _ in range(ncalls):
            one_frame(n)
            n += 1
        finish()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        launches = dev.launch_count() - l0
        timeouts = dev.signal_timeouts()
        if dist is not None:
            t = torch.tensor([ms, float(timeouts)], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, timeouts = float(t[0]), int(t[1])
        res = None
        if rank == 0:
            if exchange == "none":
                px, z = lfbs[(n - 1) & 1].read(fpc - 1, 1)
            elif exchange == "peer":
                px, z = grp.fbs[(n - 1) & 1].read(fpc - 1, 1)
            else:
                px, z = tfbs[(n - 1) & 1].color[fpc - 1:].cpu().numpy(), tfbs[(n - 1) & 1].depth[fpc - 1:].cpu().numpy()
            rows_other = sum(max(0, y1 - y0) for i, (y0, y1) in enumerate(rows) if i != 0)
            # the other ranks send rank 0 only the tiles that are busy (or were, in that buffer): background stays put
            busy_other = int(sum(flags[y0 // 32:(y1 + 31) // 32].sum() for i, (y0, y1) in enumerate(rows) if i != 0))
            fps = frames_timed / (ms * 1e-3)
            res = {
                "metric": "Mtriangles/s (submitted scene triangles x FPS), 2M-tri scene @3840x2160, sort-first strips",
                "value": fps * nfaces / 1e6, "unit": UNIT, "n_gpus": world, "fps": fps, "ms_per_frame": ms / frames_timed,
                "frames_timed": frames_timed, "frames_per_call": fpc, "scaling": "strong", "gpu_launches": int(launches),
                "covered_pixels": int((z[0] > -1).sum()), "checksum": int(px[0].astype(np.uint64).sum()),
                "signal_timeouts": timeouts,
                "rows_per_rank": [[int(a), int(b)] for a, b in rows], "busy_tiles_in_probe_frame": busy_tiles,
                "nvlink_bytes_per_frame": int(busy_other * 8192) if (world > 1 and exchange == "peer") else int(rows_other * W4 * 8) if world > 1 else 0,
                "nvlink_bytes_per_frame_if_whole_strips": int(rows_other * W4 * 8) if world > 1 else 0,
                "workload": "C4: 10 x textured Gouraud 200k-triangle spheres, 3840x2160 (BASELINE.json configs[3])",
                "exchange": ("every rank's raster kernel stores its rows into rank 0's framebuffer over NVLink (CUDA IPC peer memory), "
                             "device-side flags per rank, two framebuffers in flight" if exchange == "peer" else
                             "grouped NCCL send/recv of torch-owned framebuffers to rank 0, overlapped with the next frame"),
                "balance": args.strips_balance, "last_call": int((n - 1) % FR),
            }
            if single is not None:
                res["single_gpu_same_calls"] = single
                res["speedup_vs_single_gpu"] = single["ms_per_frame"] / res["ms_per_frame"]
        if exchange == "peer":
            grp.close()
        dev.close()
    return res


def run_strips(args, rank: int, world: int, local_rank: int):
    import torch
    import gorender_b200 as g

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist  # noqa: F811

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    res = strips_leg(args, g, torch, dist, rank, world, local_rank, frames_timed=max(args.steps, 1) * 8, warm=max(args.warmup, 1) * 4)
    if rank == 0 and "error" in res:
        print(json.dumps(res), flush=True)
    elif rank == 0:
        res.update({"steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_frame"] * 8, "higher_is_better": True,
                    "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                    "config": {"workload": res["workload"], "frames_per_step": 8, "parallelism": f"sort-first strips x{world}"}})
        print(json.dumps(res), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


# ---------------------------------------------------------------- this repo's arm

def run_b200(args, rank: int, world: int, local_rank: int):
    import torch
    import gorender_b200 as g
    from gorender_b200.renderer import Mirror

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the gorender_b200 hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist  # noqa: F811

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    B = max(1, min(args.batch, args.frames))
    F = (args.frames // B) * B            # frames per step, a whole number of batches
    NB = F // B
    K, W = args.steps, args.warmup
    objs, cam = build_scene()
    nfaces = sum(len(o.Mesh.Faces) for o in objs)
    nverts = sum(len(o.Mesh.Vertices) for o in objs)

    # Two contexts, each with its own CUDA stream, workspace and framebuffer; batches alternate
    # between them, so the tail of one batch's raster kernel overlaps the next batch's setup kernel.
    streams = [torch.cuda.Stream() for _ in range(2)]
    devs = [g.Device(local_rank, st.cuda_stream) for st in streams]
    stream, dev = streams[0], devs[0]
    fbs = [g.FrameBuffer(WIDTH, HEIGHT, B, devs[k]) for k in range(2)]
    rends = [g.Renderer(fb) for fb in fbs]

    # per-frame matrices, computed on the host like the Go caller would (renderer.go:255-262);
    # each rank renders its own frames of the spin.  A window of distinct batches is prepared and cycled.
    nprep = min(NB, 48) & ~1 or 1
    first = rank * nprep * B
    rot = spin_frames(first, nprep * B)
    packed = [np.ascontiguousarray(rends[b & 1].pack_objects(objs, [cam] * B, rot[b * B:(b + 1) * B])) for b in range(nprep)]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        for d in devs:
            d.synchronize()

    def step_device(s, only=None, nb=NB):
        for b in range(nb):
            k = (b & 1) if only is None else only
            rends[k].draw_packed(packed[(s * NB + b) % nprep], 0, sync=False)

    # ---- leg 1: device-resident throughput (the `value`)
    with torch.cuda.stream(stream):
        for s in range(W):
            step_device(s)
        barrier()
        sampler = ClockSampler(local_rank) if rank == 0 else None
        if sampler:
            time.sleep(0.1)
        launches0 = sum(d.launch_count() for d in devs)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(streams[0])
        streams[1].wait_event(e0)              # both streams start after e0 ...
        for s in range(W, W + K):
            step_device(s)
        tail = torch.cuda.Event()
        tail.record(streams[1])
        streams[0].wait_event(tail)            # ... and e1 is recorded after both have finished
        e1.record(streams[0])
        barrier()
        t1 = time.perf_counter()
        ms = e0.elapsed_time(e1)
        launches = sum(d.launch_count() for d in devs) - launches0
        clocks = sampler.stop(t0, t1) if sampler else None
    stats = np.zeros(B, dtype=g._cabi.STATS_DTYPE)
    dev.check(dev.lib.grb_frame_stats_read(dev.h, B, stats.ctypes.data))

    # ---- leg 2: end to end through the C ABI with host buffers: host mirrors (tile-sparse write-back into pinned
    # host memory) of every frame's pixels and z-buffer; `full` = whole-frame DMA copies instead (round 1's form)
    # four framebuffers (two per context) and four pairs of host planes in flight, so that a mirror update — the PCIe-bound
    # part — always has a finished batch to move while the next ones render
    NFB = 4
    fbs_e = fbs + [g.FrameBuffer(WIDTH, HEIGHT, B, devs[k & 1]) for k in range(2, NFB)]
    rends_e = rends + [g.Renderer(fb) for fb in fbs_e[2:]]
    mir_c = [Mirror(devs[k & 1], WIDTH, HEIGHT, B, g._cabi.GRB_PLANE_COLOR) for k in range(NFB)]
    mir_z = [Mirror(devs[k & 1], WIDTH, HEIGHT, B, g._cabi.GRB_PLANE_DEPTH) for k in range(NFB)]
    host_px = [m.array for m in mir_c]
    host_z = [m.array for m in mir_z]

    def step_e2e(s, with_depth=True, full=False):
        for b in range(NB):
            k = b % NFB
            rends_e[k].draw_packed(packed[(s * NB + b) % nprep], 0, sync=False)   # H2D of the matrices happens inside
            if full:
                fbs_e[k].read_async(0, B, host_px[k], host_z[k] if with_depth else None)  # overlaps the next draws
            else:
                fbs_e[k].update_mirrors_async(0, B, mir_c[k], mir_z[k] if with_depth else None)

    def time_e2e(steps, count_tiles=False, **kw):
        step_e2e(0, **kw)
        barrier()
        w0 = [m.stats() for m in mir_c + mir_z] if count_tiles else None     # (synchronises; outside the timed region)
        t0 = time.perf_counter()
        for s in range(W, W + steps):
            step_e2e(s, **kw)
        for d in devs:
            d.synchronize()
        sec = time.perf_counter() - t0
        barrier()
        if count_tiles:
            w1 = [m.stats() for m in mir_c + mir_z]
            return sec / steps, sum(b[0] - a[0] for a, b in zip(w0, w1)), sum(b[1] - a[1] for a, b in zip(w0, w1))
        return sec / steps

    with torch.cuda.stream(stream):
        e2e_full_sec = time_e2e(max(1, min(K, 2)), full=True)
        for m in mir_c + mir_z:
            m.invalidate()       # the DMA copies wrote the planes behind the mirrors' backs
        time_e2e(1)
        e2e_sec, tiles_w, tiles_f = time_e2e(K, count_tiles=True)     # tiles written during the K timed steps only
        # colour only (what the reference's presenter consumes: Pixels2, main.go:297)
        e2e_px_sec = time_e2e(max(1, K // 2), with_depth=False)
    checksum = int(host_px[(NB - 1) % NFB][B - 1].sum())  # the read-back is real

    # ---- leg 3: per-kernel CUDA-event times (roofline of the dominant kernel)
    dev.set_kernel_timing(True)     # one context only: kernels timed back to back, no overlap
    nbt = min(NB, 16)
    with torch.cuda.stream(stream):
        step_device(0, only=0, nb=nbt)
        dev.synchronize()
        dev.kernel_times()
        nt = 2
        for s in range(nt):
            step_device(s, only=0, nb=nbt)
        dev.synchronize()
    ktimes, _ = dev.kernel_times()
    dev.set_kernel_timing(False)
    ktimes = {k: v / (nt * nbt) for k, v in ktimes.items() if k in ("setup", "raster") and v > 0}  # ms per launch (one launch of each per batch)

    # ---- max over ranks
    if dist is not None:
        t = torch.tensor([ms, e2e_sec, e2e_px_sec, e2e_full_sec], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_sec, e2e_px_sec, e2e_full_sec = float(t[0]), float(t[1]), float(t[2]), float(t[3])

    total_frames = world * F * K
    fps = total_frames / (ms * 1e-3)
    value = fps * nfaces / 1e6
    e2e_fps = world * F / e2e_sec
    e2e_value = e2e_fps * nfaces / 1e6

    # ---- roofline, SURVEY.md §8(d): B = 16 Nv + 12 Nf + 16 Nf + 8 W H compulsory bytes per C3 frame (scene read once,
    # framebuffer written once, clear generated in-kernel).  The setup kernel's share is the scene, the raster kernel's
    # the framebuffer.  `design_bytes` adds the intermediates this design moves between its two kernels.
    tris = float(stats["triangles"].mean())
    hbm_peak, peak_src = peaks()
    ntiles = ((WIDTH + 31) // 32) * ((HEIGHT + 31) // 32)
    alg = {"setup": 16.0 * nverts + 12.0 * nfaces + 16.0 * nfaces, "raster": 8.0 * WIDTH * HEIGHT}
    design = {"setup": 48.0 * nfaces + 16.0 * tris + 48.0 * tris + 8.0 * tris / 6.0,
              "raster": 8.0 * WIDTH * HEIGHT + 48.0 * tris + 8.0 * tris / 6.0 + 4.0 * ntiles}
    ncu = ncu_summary("r*_ncu_c3_batch*.json")
    kroof = kernel_roofline(ktimes, alg, B, hbm_peak, ncu)
    for k, e in kroof.items():
        e["design_bytes_per_launch"] = design[k] * B
        e["design_frac"] = design[k] * B / (e["ms_per_launch"] * 1e-3) / 1e9 / hbm_peak
    dom = max(kroof, key=lambda k: kroof[k]["ms_per_launch"])
    path_bytes = alg["setup"] + alg["raster"]
    roofline = {
        "bound": "hbm", "kernel": dom + "_kernel", "achieved": kroof[dom]["achieved"], "peak": hbm_peak, "unit": "GB/s",
        "frac": kroof[dom]["frac"], "traffic": kroof[dom].get("traffic"), "peak_source": peak_src,
        "definition": "SURVEY.md section 8(d) compulsory bytes of the kernel (raster: 8 W H framebuffer out; setup: 16 Nv + 28 Nf scene in) "
                      "x frames per launch / CUDA-event time of the kernel / measured HBM copy bandwidth",
        "algorithmic_bytes_per_launch": kroof[dom]["algorithmic_bytes_per_launch"], "ms_per_launch": kroof[dom]["ms_per_launch"],
        "frames_per_launch": B, "kernels": kroof,
        "kernel_share": {k: e["ms_per_launch"] / max(sum(x["ms_per_launch"] for x in kroof.values()), 1e-12) for k, e in kroof.items()},
        "ncu_file": ncu["file"] if ncu else None,
        "note": "the path is instruction-issue bound, not HBM bound: see kernels.*.ncu (issue slots active vs DRAM % of peak)",
        "path_bytes_per_frame": path_bytes, "path_achieved_gbs": path_bytes * fps / world / 1e9,
        "path_frac": path_bytes * fps / world / 1e9 / hbm_peak,
    }

    # ---- extras: the platform's D2H ceiling, latency (one frame per call), the other configurations, strips
    extras = {}
    orc = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from oracle_binding import Oracle

        orc = Oracle()
    if not args.no_extras:
        extras["d2h_ceiling_gbs"] = pcie_ceiling(torch, dist, world)
        if rank == 0:
            extras["latency"] = latency_leg(g, dev, stream, objs, cam)
        if world == 1:
            extras["other_configs"] = other_configs(args, g, dev, stream, hbm_peak, orc)
    h2d_step = int(NB * packed[0].nbytes)
    d2h_full_step = int(NB * (host_px[0].nbytes + host_z[0].nbytes))
    for m in mir_c + mir_z:
        m.close()
    host_px = host_z = None
    if not args.no_extras and world > 1:
        for fb in fbs_e:
            fb.close()
        for d in devs:
            d.trim()
        extras["strips"] = strips_leg(args, g, torch, dist, rank, world, local_rank)

    cpu = None
    if orc is not None:
        import scene_defs

        n = args.cpu_sample_frames
        rr = scene_defs.SceneDef(WIDTH, HEIGHT, objs, cam).renderer(None)
        timer = orc.sequence_timer(rr, objs, [cam] * n, spin_frames(0, n), threads=16)
        timer.run(0, min(n, 5))
        sec, _ = timer.run(0, n)
        timer.close()
        cpu = {"value": n / sec * nfaces / 1e6, "unit": UNIT, "fps": n / sec, "cores": os.cpu_count(), "threads": 16,
               "kind": "port", "sample": f"{n} consecutive demo-spin frames of the same C3 scene ({sec:.1f} s), oracle "
               "in the reference's threaded structure (1 projection task per object + 16 tile tasks)"}

    if rank == 0:
        d2h_step = tiles_w * 4096.0 / max(K, 1)     # bytes this rank's mirrors received per step
        e2e = {"value": e2e_value, "unit": UNIT, "fps": e2e_fps,
               "h2d_bytes_per_step": h2d_step, "d2h_bytes_per_step": int(d2h_step), "checksum": checksum,
               "reads_back": "pixels (RGBA8) and z-buffer (f32) of every frame in pinned host memory, kept exact by host "
                             "mirrors: only tiles that are busy now or were busy in the host copy cross PCIe",
               "tiles_written_frac": tiles_w / max(tiles_f, 1),
               "d2h_gbs": world * d2h_step / e2e_sec / 1e9,
               "full_frame_copies": {"value": world * F / e2e_full_sec * nfaces / 1e6, "fps": world * F / e2e_full_sec,
                                     "d2h_bytes_per_step": d2h_full_step, "d2h_gbs": world * d2h_full_step / e2e_full_sec / 1e9},
               "pixels_only": {"value": world * F / e2e_px_sec * nfaces / 1e6, "fps": world * F / e2e_px_sec}}
        if "d2h_ceiling_gbs" in extras:
            e2e["d2h_ceiling_gbs"] = extras["d2h_ceiling_gbs"]
            e2e["frac_of_d2h_ceiling"] = e2e["d2h_gbs"] / extras["d2h_ceiling_gbs"]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "fps": fps,
            "mtps_hud": fps * float(stats["tpf"].mean()) / 1e6,
            "config": bench_config(args, world),
            "e2e": e2e,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "frame_stats": {"triangles_rasterised": tris, "tpf": float(stats["tpf"].mean()),
                            "out_of_domain": int(stats["out_of_domain"].sum()), "list_fallbacks": int(stats["list_fallbacks"].sum())},
        }
        for k in ("latency", "other_configs", "strips"):
            if extras.get(k) is not None:
                line[k] = extras[k]
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    # stdout carries exactly ONE line, the JSON: native libraries that write to file descriptor 1 on their own
    # (NCCL prints its version banner there) are sent to stderr for the whole run, and print() keeps the real stdout
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    if args.mode == "strips":
        run_strips(args, rank, world, local_rank)
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
