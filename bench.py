#!/usr/bin/env python
"""bench.py - Mrays/s and ms/frame of the kd_tree_simd_accel hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2] [--spp S] [--mode exact|ordered|fast|fast+ordered]
    python bench.py --impl reference ...      the reference's own CPU path (oracle/_ref, else the oracle port) on the host cores

A step = one frame of the workload through the whole wavefront (ray generation, primary / secondary / shadow closest-hit
queries, shading, resolve).  Default workload = BASELINE.json configs[1]: scenes/hw09/scene5.crtscene, 1920x1080, 1 spp,
max_ray_depth 5 (2,073,600 primary + 652,885 shadow/reflection queries per frame).  --workload cfg4 renders configs[3] as stated
(128 spp), --workload cfg5 [--spp S] configs[4] (S samples of the 4K GI frame per step; multi-pass frames are queued without a
host round trip per pass).

value  : all closest-hit queries of the K timed frames / device time (CUDA events on the launching stream, L2 flushed
         between frames outside the events); scene resident in HBM, frame left in HBM.
e2e    : the same frames through the C ABI as a frame SEQUENCE (include/rt_b200.h): parameters in, frame out to pinned HOST
         memory every step, wall clock; frame i's download overlaps frame i+1's render.  e2e.value is the sequence that delivers
         what the reference's main.cpp consumes - the 8-bit frame of io/image/ppm.hpp:17-19, quantised on the device
         (rt_render_frame_rgb8_begin); e2e.float_sequence is the float frame (rt_render_frame_begin), e2e.one_call_per_frame the
         synchronous rt_render_frame.
roofline: the dominant trace kernel against the bound that holds for the workload - instruction issue for the cache-resident
         scenes (configs 1-4), HBM for config 5 - from the committed ncu counts of the same kernels (profiles/kernel_counts.json)
         and the kernel time measured live; the reference algorithm's bytes / flop are reported apart (reference_equivalent).
N > 1  : one process per GPU, scene replicated, each rank renders its own sample of every pixel (weak scaling: the N-GPU
         job is the same frame at spp = N); the framebuffers are combined by ONE fused kernel per rank over NVLink peer memory
         (rt_peer_*: wait + reduce + resolve; --combine nccl keeps the ncclReduce baseline); time = max over ranks.
north_star_scaling (every run): the north-star frame - config 5, 3840x2160, GI 1, a fixed TOTAL sample count split over the
         N GPUs (strong scaling), by sample slice and by row band (render/tile/bucket.hpp:7-21) - ms/frame, efficiency against
         the same job on one GPU measured in the same run, combined frame verified against the single-GPU evaluation.
"""
from __future__ import annotations

import argparse
import gzip
import importlib
import json
import os
import sys
import tempfile
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

WORKLOADS = {
    "cfg1": "cfg1_hw15_scene2", "cfg2": "cfg2_hw09_scene5", "cfg3": "cfg3_hw11_scene8_d10", "cfg3d5": "cfg3_hw11_scene8_d5",
    "cfg4": "cfg4_hw12_scene4", "cfg5": "cfg5_synthetic",
}
# BASELINE.json configs[4]: the synthetic random-mesh scene (SURVEY.md section 8d "Config 5"): N triangles in a diffuse box,
# 3840x2160, GI 1, max_ray_depth 5, kd<24,64>; one sample of every pixel per step and per GPU (the 512-spp frame is 512 such steps)
SYNTHETIC = {"cfg5": dict(n_tris=10_000_000, width=3840, height=2160, seed=1234, spp=1, max_ray_depth=5, gi_rays=1, kd=(24, 64))}
# samples per pixel and GI rays as BASELINE.json states them, where that differs from the 1-spp pass workloads.json describes
AS_STATED = {"cfg4": dict(spp=128, gi_rays=1)}
MODES = {"exact": 0, "fast": 2, "ordered": 4, "fast+ordered": 6}


def load_workload(name: str, n_tris: int | None = None, cpu_only: bool = False, spp: int | None = None) -> dict:
    w = _load_workload(name, n_tris, cpu_only)
    w["spp_pass"] = w["spp"]                      # what workloads.json's per-frame counts refer to
    for k, v in AS_STATED.get(name, {}).items():
        w[k] = v
    if spp:
        w["spp"] = spp
    return w


def _load_workload(name: str, n_tris: int | None = None, cpu_only: bool = False) -> dict:
    with open(os.path.join(REPO, "tests", "golden", "workloads.json")) as fh:
        table = json.load(fh)
    if name in SYNTHETIC:
        from tests.helpers import crtscene                       # scene synthesis (numpy), not the oracle
        cfg = dict(SYNTHETIC[name])
        if n_tris:
            cfg["n_tris"] = n_tris
        w = dict(table.get(f"{WORKLOADS[name]}_{cfg['n_tris']}", {}))
        w.update(key=f"{WORKLOADS[name]}_{cfg['n_tris']}", scene=f"synthetic_{cfg['n_tris']}_triangles_seed{cfg['seed']}", width=cfg["width"],
                 height=cfg["height"], spp=cfg["spp"], max_ray_depth=cfg["max_ray_depth"], gi_rays=cfg["gi_rays"], kd=list(cfg["kd"]),
                 n_triangles=cfg["n_tris"], synthetic=True)
        # the CPU arms render a 1/64-pixel version of the same camera (every 8th pixel in x and y): a bounded sample
        wd, ht = (cfg["width"] // 8, cfg["height"] // 8) if cpu_only else (cfg["width"], cfg["height"])
        w["rtsc"] = crtscene.to_rtsc_bytes(crtscene.synthetic_scene(n_tris=cfg["n_tris"], seed=cfg["seed"], width=wd, height=ht))
        w["cpu_sample_div"] = 64
        return w
    w = table[WORKLOADS[name]]
    w["key"] = WORKLOADS[name]
    with gzip.open(os.path.join(REPO, "tests", "golden", "scenes", w["scene"] + ".rtsc.gz"), "rb") as fh:
        w["rtsc"] = fh.read()
    return w


def measured_peaks() -> tuple[dict, str]:
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as fh:
            return json.load(fh), "measured (MEASURED_PEAKS.json)"
    except OSError:
        return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock + throttle reasons during the timed region (NVML; falls back to nvidia-smi)."""

    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._th = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                     "hw_power_brake_slowdown": 0x80, "sync_boost": 0x10, "display_clock_setting": 0x100}
            while not self._stop.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.reasons |= {k for k, bit in names.items() if r & bit}
                time.sleep(0.001)
        except Exception:  # noqa: BLE001
            import subprocess
            while not self._stop.is_set():
                try:
                    out = subprocess.check_output(["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm,"
                                                   "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                                                   "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                                                   "--format=csv,noheader,nounits"], text=True, timeout=5).strip().split(", ")
                    self.samples.append(int(out[0]))
                    self.max_mhz = int(out[1])
                    for k, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), out[2:]):
                        if v.strip() == "Active":
                            self.reasons.add(k)
                except Exception:  # noqa: BLE001
                    pass
                time.sleep(0.1)

    def __enter__(self):
        self._th.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._th.join(timeout=10)

    def summary(self) -> dict:
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own CPU implementation on the host cores
# ------------------------------------------------------------------------------------------------------------------------
def cpu_reference_frames(w: dict, frames: int, warmup: int) -> dict:
    """Times `frames` frames of the workload with the UNMODIFIED reference compiled into oracle/_ref (speed build:
    -O3, FMA contraction on, widest ISA the host has; all host threads, BUCKET_TILES), falling back to the oracle port.
    This is the one place bench.py executes oracle/ code, as the baseline - never as the thing measured.
    Synthetic workloads (cfg5) render the 1/64-pixel version of the frame held in w["rtsc"] (same camera, every 8th pixel)."""
    from tests.helpers import refimpl
    kd = tuple(w.get("kd", (8, 64)))
    base = dict(spp=w["spp"], depth=w["max_ray_depth"], gi=w["gi_rays"], kd_depth=kd[0], kd_leaf=kd[1])
    variant = None
    for fp, isa in (("speed", "v4"), ("speed", "v3"), ("canon", "v3")):
        if refimpl.available(fp=fp, isa=isa, **base):
            variant = dict(fp=fp, isa=isa, **base)
            break
    times = []
    what = (f"frames of {w['key']}" if not w.get("synthetic") else
            f"frames of {w['key']} at 1/{w['cpu_sample_div']} of the pixels (every 8th pixel in x and y, same camera)")
    if variant:
        with tempfile.NamedTemporaryFile(suffix=".rtsc") as tf:
            tf.write(w["rtsc"])
            tf.flush()
            ref = refimpl.RefImpl(tf.name, **variant)
            if w.get("synthetic") or "rays" not in w or w["spp"] != w.get("spp_pass", w["spp"]):
                c = ref.count()                                  # counting pass of the harness, outside the timed frames
                rays = c["cull"] + c["nocull"]
            else:
                rays = w["rays"]
            for i in range(warmup + frames):
                _, sec = ref.render(want_image=False)
                if i >= warmup:
                    times.append(sec)
            kind, cores = "reference", ref.threads
            sample = (f"{frames} {what} by render_frame(BUCKET_TILES) of the reference headers "
                      f"({os.path.basename(ref.path)}, W={ref.W}), wall clock around render_frame as src/main.cpp:16-20")
            ref.close()
    else:
        from tests.helpers import oracle
        o = oracle.Oracle(w["rtsc"], kd[0], kd[1])
        p = oracle.default_params(spp=w["spp"], max_ray_depth=w["max_ray_depth"], gi_rays=w["gi_rays"])
        rays = None
        for i in range(warmup + frames):
            t0 = time.perf_counter()
            _, c = o.render(p)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
            rays = int(c[0]) + int(c[2]) + int(c[4])
        kind, cores = "port", os.cpu_count()
        sample = f"{frames} {what} by oracle/rt_oracle.c (C port, all host threads); oracle/_ref has no build for this configuration"
    mean = sum(times) / len(times)
    return {"value": rays / mean / 1e6, "unit": "Mrays/s", "cores": int(cores), "kind": kind, "sample": sample,
            "ms_per_frame": 1e3 * mean, "ms_per_frame_best": 1e3 * min(times), "rays_per_frame": int(rays)}


def run_reference_arm(args, w: dict) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base = cpu_reference_frames(w, max(args.steps, 1), max(args.warmup, 1))
    line = {"impl": "reference", "metric": "Mrays/s", "value": base["value"], "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": base["ms_per_frame"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": data_desc(w),
            "config": workload_config(w, args, int(os.environ.get("WORLD_SIZE", "1"))),
            "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": base["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def data_desc(w: dict) -> str:
    if w.get("synthetic"):
        return "synthetic random-triangle mesh in a diffuse box (numpy PCG64, seed 1234) + camera rays generated on the device"
    return "synthetic camera rays over the reference's own scene file (fixture copy)"


def workload_config(w: dict, args, world: int, accel: str | None = None) -> dict:
    kd = w.get("kd", [8, 64])
    src = (f"{w['scene']} (tests/helpers/crtscene.synthetic_scene)" if w.get("synthetic") else
           f"scenes/{w['scene'].replace('_', '/', 1)}.crtscene")
    return {"workload": f"{w['key']}: {src} {w['width']}x{w['height']}, "
                        f"spp {w['spp']} per GPU, max_ray_depth {w['max_ray_depth']}, gi_rays {w['gi_rays']}, kd<{kd[0]},{kd[1]}>",
            "rays_per_frame_per_gpu": w.get("rays"), "mode": args.mode, "accel": accel, "sharding": "replicated scene, 1 sample slice per GPU" if world > 1 else "none",
            "l2": "flushed between timed frames (256 MiB write, outside the CUDA events)"}



# ------------------------------------------------------------------------------------------------------------------------
# north-star scaling record: config 5's 4K GI frame, fixed total sample count, split over the GPUs of the run (strong scaling)
# ------------------------------------------------------------------------------------------------------------------------
def north_star_scaling(rt, args, rank: int, world: int, local: int, flags: int):
    """BASELINE.json north_star: "near-linear tile-parallel scaling to 8xB200 on a 4K multi-spp GI frame".  Every rank holds the
    replicated config-5 scene (args.ns_tris triangles) and the frame - 3840x2160, GI 1, max_ray_depth 5, args.ns_spp samples per
    pixel in TOTAL - is split two ways: by sample slice (every rank renders its samples of every pixel, RT_FLAG_RAW_SUM) and by
    row band (the reference's bucket decomposition, render/tile/bucket.hpp:7-21, dealt round-robin; every rank renders all
    samples of its bands).  Both are combined by the one fused peer-memory kernel per rank (rt_peer_combine) into rank 0's frame.
    The same job on ONE GPU is timed in the same run (rank 0 alone), efficiency = t(1) / (N x t(N)); the combined frames are
    verified against the single-GPU evaluation.  Device time (CUDA events around render + combine), max over ranks."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from tests.helpers import crtscene
    cfg = SYNTHETIC["cfg5"]
    T = int(args.ns_spp)
    t0 = time.perf_counter()
    data = crtscene.to_rtsc_bytes(crtscene.synthetic_scene(n_tris=args.ns_tris, seed=cfg["seed"], width=cfg["width"], height=cfg["height"]))
    scene = rt.Scene.from_rtsc(data, kd_max_depth=cfg["kd"][0], kd_max_leaf_size=cfg["kd"][1], device=local, accel_width=args.accel_width,
                               accel_build=rt.ACCEL_BUILD_DEVICE if args.accel_build == "device" else rt.ACCEL_BUILD_HOST)
    del data
    t_build = time.perf_counter() - t0
    H, W = scene.height, scene.width
    stream = torch.cuda.current_stream()
    kw = dict(max_ray_depth=cfg["max_ray_depth"], diffuse_reflection_ray_count=cfg["gi_rays"])
    fb = torch.zeros((H, W, 3), dtype=torch.float32, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(frame, frames=2, warm=1):
        for _ in range(warm):
            frame()
        barrier()
        ms = 0.0
        for _ in range(frames):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            a.record(stream)
            frame()
            b.record(stream)
            b.synchronize()
            t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms += float(t[0])
        return ms / frames

    # ---- the job on one GPU (every rank renders it, independently; rank 0's time is the N = 1 reference) ----
    full = rt.default_params(samples_per_pixel=T, flags=flags, **kw)
    def one_gpu():
        scene.render_frame_device(full, fb.data_ptr(), stream=stream.cuda_stream)
    for _ in range(1):
        one_gpu()
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream); one_gpu(); b.record(stream); b.synchronize()
    t1 = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.broadcast(t1, src=0)
    ms1 = float(t1[0])
    c = scene.counters()
    rays = int(c.primary + c.shadow + c.secondary)
    out = {"workload": f"config 5: synthetic {args.ns_tris:,}-triangle mesh in a diffuse box, {W}x{H}, GI 1, max_ray_depth {cfg['max_ray_depth']}, "
                       f"{T} spp in total, kd<{cfg['kd'][0]},{cfg['kd'][1]}>, accel bvh{int(scene.info.accel_width)}",
           "scaling": "strong", "n_gpus": world, "total_spp": T, "rays_per_frame": rays, "passes_per_frame_n1": int(c.passes),
           "n1": {"ms_per_frame": ms1, "mrays_s": rays / ms1 / 1e3}, "host_build_s_per_rank": round(t_build, 2)}
    if world > 1:
        frame1 = fb.clone()                                         # the one-GPU frame (sequential sample order)
        def group():
            g = rt.PeerGroup(world, rank, local, W, H)
            handles = [None] * world
            dist.all_gather_object(handles, g.handle)
            g.connect(handles)
            dist.barrier()
            return g
        # ---- sample slices ----
        first, count = rt.spp_slice(T, rank, world)
        sl = rt.default_params(samples_per_pixel=max(count, 1), sample_offset=first, spp_total=T, flags=flags | rt.FLAG_RAW_SUM, **kw)
        g = group()
        def by_samples():
            scene.render_frame_device(sl, g.framebuffer, stream=stream.cuda_stream)
            g.combine(T, rt.PEER_OUT_RGB, stream=stream.cuda_stream)
        ms = timed(by_samples)
        rec = {"ms_per_frame": ms, "mrays_s": rays / ms / 1e3, "efficiency_vs_n1": ms1 / (world * ms), "samples_per_gpu": count}
        if rank == 0:
            got, _ = g.read_result(stream.cuda_stream, want_rgb8=False)
            # the single-GPU evaluation of the same partition: every slice's raw sum, added in rank order, divided once
            acc = None
            tmp = torch.zeros_like(fb)
            for r in range(world):
                f, n = rt.spp_slice(T, r, world)
                scene.render_frame_device(rt.default_params(samples_per_pixel=max(n, 1), sample_offset=f, spp_total=T, flags=flags | rt.FLAG_RAW_SUM, **kw),
                                          tmp.data_ptr(), stream=stream.cuda_stream)
                stream.synchronize()
                acc = tmp.clone() if acc is None else acc.add_(tmp)
            scene.resolve_sum_device(acc.data_ptr(), T, d_rgb=acc.data_ptr(), stream=stream.cuda_stream)
            stream.synchronize()
            want = acc.cpu().numpy()
            rec["combined_frame_bit_identical_to_single_gpu_evaluation_of_the_partition"] = bool(np.array_equal(got.view(np.uint32), want.view(np.uint32)))
            rec["max_abs_diff_vs_sequential_single_gpu_frame"] = float(np.abs(got - frame1.cpu().numpy()).max())
        barrier()
        g.close()
        out["sample_slices"] = rec
        # ---- row bands (bucket rows dealt round-robin): all samples of the rank's bands; the other rows of its frame stay zero ----
        g = group()
        BAND = 16                                                   # rows per band: 135 bands at 4K, so 16 or 17 per rank at N = 8
        bands = rt.row_bands(H, BAND, rank, world)
        # all of the rank's bands in ONE call (rt_params.band_rows / band_period / band_phase): one wavefront over 1/N of the frame
        band_params = rt.default_params(samples_per_pixel=T, band_rows=BAND, band_period=world, band_phase=rank, flags=flags, **kw)
        def by_bands():
            scene.render_frame_device(band_params, g.framebuffer, stream=stream.cuda_stream)
            g.combine(1, rt.PEER_OUT_RGB, stream=stream.cuda_stream)      # sum of disjoint bands; dividing by 1 is exact
        # both frame slots of the group must hold this rank's bands and zeros elsewhere: three frames touch both slots
        ms = timed(by_bands, frames=2, warm=2)
        rec = {"ms_per_frame": ms, "mrays_s": rays / ms / 1e3, "efficiency_vs_n1": ms1 / (world * ms), "bands_per_gpu": len(bands), "band_rows": BAND}
        if rank == 0:
            got, _ = g.read_result(stream.cuda_stream, want_rgb8=False)
            rec["combined_frame_equals_single_gpu_frame"] = bool(np.array_equal(got, frame1.cpu().numpy()))
        barrier()
        g.close()
        out["row_bands"] = rec
    scene.close()
    return out

# ------------------------------------------------------------------------------------------------------------------------
def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--mode", default="ordered", choices=sorted(MODES),
                    help="ordered = the accelerated query (the backend's own bounding-volume hierarchy, near child first; frames "
                         "bit-identical to exact); exact = the reference's kd-tree in the reference's visit order")
    ap.add_argument("--tris", type=int, default=None, help="triangle count of the synthetic workload (cfg5; default 10,000,000)")
    ap.add_argument("--spp", type=int, default=None, help="samples per pixel per GPU and step (default: the workload's own)")
    ap.add_argument("--accel-width", type=int, default=0, choices=[0, 2, 4], help="hierarchy width of the ordered modes (0 = default)")
    ap.add_argument("--accel-build", default="host", choices=["host", "device"],
                    help="where the backend's hierarchy is built: binned SAH on the host threads, or a linear BVH by CUDA kernels (csrc/rt_lbvh.cuh)")
    ap.add_argument("--ns-tris", type=int, default=1_000_000, help="north_star_scaling: triangles of the config-5 scene (0 = skip)")
    ap.add_argument("--ns-spp", type=int, default=64, help="north_star_scaling: total samples per pixel of the 4K GI frame")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--combine", default="peer", choices=["peer", "nccl"],
                    help="N > 1: peer = one fused wait+reduce+resolve kernel per rank over NVLink peer memory (rt_peer_combine); "
                         "nccl = ncclReduce to rank 0 + resolve kernel (the baseline it replaces)")
    args = ap.parse_args()
    w = load_workload(args.workload, args.tris, cpu_only=(args.impl == "reference"), spp=args.spp)
    if args.impl == "reference":
        run_reference_arm(args, w)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the backend has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rt = importlib.import_module("simd-raytracer_b200")
    kd = w.get("kd", [8, 64])
    t_build = time.perf_counter()
    scene = rt.Scene.from_rtsc(w["rtsc"], kd_max_depth=kd[0], kd_max_leaf_size=kd[1], device=local, accel_width=args.accel_width,
                               accel_build=rt.ACCEL_BUILD_DEVICE if args.accel_build == "device" else rt.ACCEL_BUILD_HOST)
    t_build = time.perf_counter() - t_build
    H, W = scene.height, scene.width
    flags = MODES[args.mode]
    spp_total = w["spp"] * world
    first, count = rt.spp_slice(spp_total, rank, world)
    params = rt.default_params(samples_per_pixel=count, sample_offset=first, spp_total=spp_total, max_ray_depth=w["max_ray_depth"],
                               diffuse_reflection_ray_count=w["gi_rays"], flags=flags | (rt.FLAG_RAW_SUM if world > 1 else 0))
    stream = torch.cuda.current_stream()
    fb = torch.zeros((H, W, 3), dtype=torch.float32, device="cuda")
    rgb8 = torch.zeros((H, W, 3), dtype=torch.uint8, device="cuda")
    host = torch.zeros((H, W, 3), dtype=torch.float32).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    peer = None
    if world > 1 and args.combine == "peer":
        # cudaIpc handles of every rank's framebuffer block, all-gathered once (plumbing); the per-frame combine is device-only
        peer = rt.PeerGroup(world, rank, local, W, H)
        handles = [None] * world
        dist.all_gather_object(handles, peer.handle)
        peer.connect(handles)
        dist.barrier()
        # the end-to-end frames land in ONE host frame that every rank's process maps and pins (POSIX shared memory): every
        # rank stores its slice of the combined frame over its own PCIe link, nobody downloads it (rt_peer_host_result_attach)
        # (falls back to downloading the frame from rank 0 when the shared-memory object cannot be created, e.g. a small /dev/shm)
        host_frames = None
        shm = [f"/rt_b200_bench_{os.getpid()}" if rank == 0 else None]
        if rank == 0:
            try:
                host_frames = peer.attach_host_result(shm[0], create=True)
            except rt.RtError as e:
                print(f"bench: no shared host frame ({e}); rank 0 downloads the combined frame instead", file=sys.stderr)
                shm[0] = None
        dist.broadcast_object_list(shm, src=0)
        if rank != 0 and shm[0]:
            try:
                host_frames = peer.attach_host_result(shm[0], create=False)
            except rt.RtError as e:
                print(f"bench: rank {rank} cannot attach the shared host frame ({e})", file=sys.stderr)
        ok = torch.tensor([1 if host_frames is not None else 0], dtype=torch.int32, device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)             # all ranks or none
        all_attached = int(ok.item()) == 1                    # .item() waits for the collective: every rank has tried by now
        if rank == 0 and shm[0]:
            os.unlink("/dev/shm" + shm[0])                    # the mappings keep the object alive
        if not all_attached:
            host_frames = None
    shared_host = peer is not None and host_frames is not None
    # N > 1, peer combine: one step = the render of frame i (main stream) and, concurrently on a second stream, the fused
    # wait+reduce+resolve of frame i-1 (rt_peer_* keeps two frame slots per rank).  The combine is released by the step's
    # start event, so every timed step [a, b] contains exactly one render and one combine; nothing runs during the L2 flush.
    # BENCH_SIDE_PRIORITY=-1 (high priority for the combine) was measured slower: 0.470 vs 0.448 ms per step at N = 2
    side = torch.cuda.Stream(priority=int(os.environ.get("BENCH_SIDE_PRIORITY", "0"))) if peer else None
    ev_done = torch.cuda.Event() if peer else None
    ev_rendered = torch.cuda.Event() if peer else None
    outputs = rt.PEER_OUT_RGB                      # N = 1 leaves a float frame in HBM; so does the combine (rank 0)
    pending = [False]

    def frame_device():
        """serial frame: render, then combine, on the launching stream (warm-up, verification, NCCL baseline)"""
        if peer:
            scene.render_frame_device(params, peer.framebuffer, stream=stream.cuda_stream)
            peer.combine(spp_total, rt.PEER_OUT_RGB | rt.PEER_OUT_RGB8, stream=stream.cuda_stream)
            return
        scene.render_frame_device(params, fb.data_ptr(), stream=stream.cuda_stream)
        if world > 1:
            dist.reduce(fb, dst=0, op=dist.ReduceOp.SUM)
            if rank == 0:
                scene.resolve_sum_device(fb.data_ptr(), spp_total, d_rgb=fb.data_ptr(), d_rgb8=rgb8.data_ptr(), stream=stream.cuda_stream)

    trace = [] if os.environ.get("BENCH_TRACE_COMBINE") else None      # developer probe: where a pipelined step spends its time

    def combine_pending(start_event, out=None):
        """the combine of the frame signalled last, on the side stream, not before `start_event`"""
        side.wait_event(start_event)
        peer.reduce_resolve(spp_total, out or outputs, stream=side.cuda_stream)
        if trace is not None:
            e = torch.cuda.Event(enable_timing=True); e.record(side); trace[-1]["reduce_end"] = e
        peer.wait_done(stream=side.cuda_stream)
        if trace is not None:
            e = torch.cuda.Event(enable_timing=True); e.record(side); trace[-1]["wait_done_end"] = e
        ev_done.record(side)

    def step_pipelined(a, b):
        a.record(stream)
        if trace is not None:
            trace.append({"a": a, "b": b})
        if pending[0]:
            combine_pending(a)
        t = scene.render_frame_device_begin(params, peer.framebuffer, stream=stream.cuda_stream)     # queued, no host sync
        if trace is not None:
            e = torch.cuda.Event(enable_timing=True); e.record(stream); trace[-1]["render_end"] = e
        # "frame i is rendered" is published from the SIDE stream, behind an event of the render stream: the system-scope release
        # towards the peers (~10 us) then runs beside the next frame's first kernels instead of in front of them; the combine of
        # this frame follows it in side-stream order
        ev_rendered.record(stream)
        side.wait_event(ev_rendered)
        peer.signal_ready(side.cuda_stream)                                                          # flips the frame slot
        if pending[0]:
            stream.wait_event(ev_done)             # frame i+1 renders into the slot the combined frame used
        pending[0] = True
        b.record(stream)
        return t

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        frame_device()
    barrier()
    if world == 1:
        # the timed frames are queued frames: warm that path up as well (its streams, events and pinned words are created on first use)
        for _ in range(max(args.warmup, 3)):
            scene.frame_wait(scene.render_frame_device_begin(params, fb.data_ptr(), stream=stream.cuda_stream))
        barrier()
        frame_device()                              # the counters below are those of a serial frame
        barrier()
    c0 = scene.counters()
    rays_frame = c0.primary + c0.shadow + c0.secondary

    # ---- timed: device ------------------------------------------------------------------------------------------------
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    acc = {k: 0.0 for k in ("ms_primary", "ms_secondary", "ms_shadow", "ms_shade", "ms_resolve", "ms_total")}
    launches = 0
    rerendered = 0
    if peer:
        # prime the pipeline (untimed): frame -1 rendered and signalled, its combine is the first timed step's
        e0 = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        scene.frame_wait(step_pipelined(*e0))
    barrier()
    with ClockSampler(local) as clocks:
        t_wall0 = time.perf_counter()
        for a, b in ev:
            flush.fill_(1)
            if peer:
                rerendered += bool(scene.frame_wait(step_pipelined(a, b)))
                b.synchronize()
                c = scene.counters()        # of the queued frame: device time of the render; no per-kernel events inside it
                launches += c.kernel_launches + 3
            elif world == 1:
                # the frame is QUEUED (rt_render_frame_device_begin): no per-kernel timing events between its kernels, so the
                # kernels of the pass chain through programmatic dependent launch exactly as they do in a frame sequence
                a.record(stream)
                t = scene.render_frame_device_begin(params, fb.data_ptr(), stream=stream.cuda_stream)
                b.record(stream)
                rerendered += bool(scene.frame_wait(t))
                b.synchronize()
                c = scene.counters()
                launches += c.kernel_launches
            else:
                a.record(stream)
                frame_device()
                b.record(stream)
                c = scene.counters()        # blocks until the frame is done; reads the per-class CUDA events of this frame
                launches += c.kernel_launches + (1 if world > 1 and rank == 0 else 0)
            for k in acc:
                acc[k] += getattr(c, k)
        barrier()
        t_wall = time.perf_counter() - t_wall0
    ms_dev = sum(a.elapsed_time(b) for a, b in ev)
    if trace:
        import statistics
        rows = [t for t in trace if "reduce_end" in t and t["a"] in [x for x, _ in ev]]
        for k in ("reduce_end", "wait_done_end", "render_end", "b"):
            print(f"rank {rank}: a -> {k:14s} median {statistics.median(t['a'].elapsed_time(t[k]) for t in rows) * 1e3:8.1f} us", file=sys.stderr)
    if peer:
        # drain (untimed): the combine of the last timed frame
        e1 = torch.cuda.Event()
        e1.record(stream)
        combine_pending(e1)
        stream.wait_event(ev_done)
        pending[0] = False
        barrier()
    if peer or world == 1:
        if rerendered:
            raise SystemExit("a timed frame had to be rendered again (pool overflow after warm-up): the measurement is void")
        # per-kernel-class times: the queued frames carry no per-kernel events, so the split comes from serial frames
        for k in acc:
            acc[k] = 0.0
        for _ in range(3):
            frame_device()
            c = scene.counters()
            for k in acc:
                acc[k] += getattr(c, k) * args.steps / 3.0
        barrier()
        if acc["ms_primary"] == 0.0 and c0.passes > 1 and world == 1:
            # a frame of very many passes carries no per-launch events: split one serial frame of ONE pass's samples and scale
            per_pass = -(-count // int(c0.passes))
            pp = rt.default_params(samples_per_pixel=per_pass, sample_offset=first, spp_total=spp_total, max_ray_depth=w["max_ray_depth"],
                                   diffuse_reflection_ray_count=w["gi_rays"], flags=flags)
            scene.render_frame_device(pp, fb.data_ptr(), stream=stream.cuda_stream)
            c = scene.counters()
            for k in acc:
                acc[k] = getattr(c, k) * (count / per_pass) * args.steps
            barrier()

    # ---- timed: end to end through the C ABI with a host framebuffer ---------------------------------------------------------
    e2e_params = rt.default_params(samples_per_pixel=count, sample_offset=first, spp_total=spp_total, max_ray_depth=w["max_ray_depth"],
                                   diffuse_reflection_ray_count=w["gi_rays"], flags=flags)
    host_np = host.numpy()
    host2_np = torch.zeros((H, W, 3), dtype=torch.float32).pin_memory().numpy()
    host8 = [torch.zeros((H, W, 3), dtype=torch.uint8).pin_memory().numpy() for _ in range(2)]

    def sequence_rgb8(k):
        """N = 1: the frame sequence that delivers the 8-bit frame (what write_ppm consumes, io/image/ppm.hpp:17-19): quantised on
        the device, 3 bytes per pixel over PCIe, frame i's download behind frame i+1's render"""
        prev = None
        for i in range(k):
            t = scene.render_frame_rgb8_begin(e2e_params, host8[i & 1])
            if prev is not None:
                scene.frame_wait(prev)
            prev = t
        scene.frame_wait(prev)

    def frame_e2e():
        """one call per frame, synchronous"""
        if world == 1:
            scene.render_frame(e2e_params, out=host_np)          # params in, float frame out to pinned host memory
            return
        # N > 1: every rank renders its slice, the combine runs, rank 0 downloads the float frame
        frame_device()
        if rank == 0:
            if peer:
                peer.read_result(stream.cuda_stream, rgb=host_np, want_rgb8=False)
            else:
                host.copy_(fb, non_blocking=True)
                stream.synchronize()
        else:
            stream.synchronize()

    def sequence(k):
        """the same k frames as a frame SEQUENCE: every frame still takes its parameters from the host and lands as a float
        frame in pinned host memory inside the timed region; frame i's download (and, N > 1, its combine over NVLink) overlaps
        frame i+1's render, nothing waits for the host in between"""
        if world == 1:
            prev = None
            for i in range(k):
                t = scene.render_frame_begin(e2e_params, host2_np if i & 1 else host_np)
                if prev is not None:
                    scene.frame_wait(prev)
                prev = t
            scene.frame_wait(prev)
            return 0
        redo = 0
        ev_dl = torch.cuda.Event()
        for i in range(k + 1):
            e = torch.cuda.Event()
            e.record(stream)
            if i > 0 and shared_host:
                # frame i-1: every rank reduces its slice and copies it into the shared HOST frame (side stream); complete when
                # this rank has seen every rank's flags (rt_peer_wait_done inside combine_pending)
                combine_pending(e, rt.PEER_OUT_HOST_RGB)
            elif i > 0:
                combine_pending(e)                                # frame i-1: reduce + resolve into rank 0, side stream
                if rank == 0:
                    peer.download_result(host2_np if (i - 1) & 1 else host_np, stream=side.cuda_stream)
                    ev_dl.record(side)
            t = None
            if i < k:
                t = scene.render_frame_device_begin(params, peer.framebuffer, stream=stream.cuda_stream)
                peer.signal_ready(stream.cuda_stream)
                if i > 0:
                    stream.wait_event(ev_done)
            if i > 0:
                (ev_dl if (rank == 0 and not shared_host) else ev_done).synchronize()   # frame i-1 is in host memory
            if t is not None:
                redo += bool(scene.frame_wait(t))
        return redo

    for _ in range(3):
        frame_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        frame_e2e()
    barrier()
    t_e2e_sync = time.perf_counter() - t0
    t_e2e = t_e2e_sync
    if world == 1 or peer:
        sequence(3)
        barrier()
        t0 = time.perf_counter()
        redo = sequence(args.steps)
        barrier()
        t_e2e = time.perf_counter() - t0
        if redo:
            raise SystemExit("a timed frame of the sequence had to be rendered again after warm-up: the measurement is void")
    t_e2e_rgb8 = None
    if world == 1:
        sequence_rgb8(3)
        barrier()
        t0 = time.perf_counter()
        sequence_rgb8(args.steps)
        barrier()
        t_e2e_rgb8 = time.perf_counter() - t0
        want8 = scene.render_frame_rgb8(e2e_params)
        if not (np.array_equal(host8[0], want8) and np.array_equal(host8[1], want8)):
            raise SystemExit("the 8-bit frames of the sequence differ from rt_render_frame_rgb8: the measurement is void")

    # ---- N > 1: the combined frame must be the single-GPU frame at spp = N (bit for bit with one sample per rank) ----------
    verify = None
    if world > 1:
        frame_device()
        barrier()
        if rank == 0:
            if peer:
                got, got8 = peer.read_result(stream.cuda_stream)
            else:
                got, got8 = fb.cpu().numpy(), rgb8.cpu().numpy()
            if H * W <= 1920 * 1920:
                want = scene.render_frame(rt.default_params(samples_per_pixel=spp_total, max_ray_depth=w["max_ray_depth"],
                                                            diffuse_reflection_ray_count=w["gi_rays"], flags=flags))
                same = bool(np.array_equal(got.view(np.uint32), want.view(np.uint32)))
                verify = {"combined_frame_equals_single_gpu_spp_N_frame": same,
                          "max_abs_diff": float(np.abs(got - want).max()), "rgb8_nonzero": int((got8 != 0).sum())}
            else:
                verify = {"combined_frame_finite": bool(np.isfinite(got).all()), "rgb8_nonzero": int((got8 != 0).sum())}
        barrier()

    if shared_host and rank == 0 and verify is not None and "combined_frame_equals_single_gpu_spp_N_frame" in verify:
        # both host slots hold the frame of the last two sequence steps (same parameters every step)
        verify["host_frames_equal_single_gpu_spp_N_frame"] = bool(np.array_equal(host_frames[0].view(np.uint32), want.view(np.uint32)) and
                                                                  np.array_equal(host_frames[1].view(np.uint32), want.view(np.uint32)))
    t = torch.tensor([ms_dev, t_e2e, float(rays_frame), t_e2e_sync], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms_dev, t_e2e, rays_all, t_e2e_sync = float(tmax[0]), float(tmax[1]), float(tsum[2]), float(tmax[3])
    else:
        rays_all = float(rays_frame)

    ns = None
    if args.ns_tris > 0:
        try:
            ns = north_star_scaling(rt, args, rank, world, local, flags)
        except Exception as e:  # noqa: BLE001 - the headline line must not be lost to the secondary record
            if world > 1:
                raise
            ns = {"error": repr(e)}
    if rank == 0:
        peaks, peak_src = measured_peaks()
        K = args.steps
        ms_step = ms_dev / K
        kinds = w.get("kinds") or {k: {"alg_bytes": 0, "alg_flop": 0} for k in ("primary", "secondary", "shadow")}
        spp_scale = w["spp"] / max(w.get("spp_pass", 1), 1)          # workloads.json counts one sample per pixel
        cls = {"primary": acc["ms_primary"] / K, "secondary": acc["ms_secondary"] / K, "shadow": acc["ms_shadow"] / K}
        dom = max(cls, key=cls.get)
        accel = args.mode.endswith("ordered")
        # the structure the timed trace kernels walk: the reference's own kd<depth,leaf> tree in reference order (exact / fast), or
        # the backend's bounding-volume hierarchy, two- or four-wide (ordered modes; rt_build_opts.accel_width)
        built_on_device = int(scene.info.accel_build) == rt.ACCEL_BUILD_DEVICE
        accel_name = f"{'lbvh' if built_on_device else 'bvh'}{int(scene.info.accel_width)}" if accel else f"reference kd<{kd[0]},{kd[1]}>"
        class_kernels = ({"primary": ["k_tile_cull", "k_stream_primary_sparse", "k_stream_primary"], "secondary": ["k_stream_level"],
                          "shadow": ["k_stream_shadow"]} if accel else
                         {"primary": ["k_primary"], "secondary": ["k_trace_level"], "shadow": ["k_shadow"]})
        dom_kernel = class_kernels[dom][-1] if dom != "primary" else ("k_stream_primary_sparse" if accel else "k_primary")
        # what the kernels of every class executed and moved, per frame: ncu counts of the SAME code, workload, mode and hierarchy,
        # committed under profiles/ (scripts/make_kernel_counts.py); the time they are divided by is measured live, above
        counts, counts_src = {}, None
        try:
            with open(os.path.join(REPO, "profiles", "kernel_counts.json")) as fh:
                entry = json.load(fh).get(w["key"], {}).get(f"{args.mode}/{accel_name}/spp{w['spp']}", {})
                counts, counts_src = entry.get("kernels", {}), entry.get("_source")
        except OSError:
            pass

        def class_counts(kind):
            got = [counts[k] for k in class_kernels[kind] if k in counts]
            if not got:
                return None
            return {f: sum(g[f] for g in got) for f in ("launches", "warp_inst", "thread_inst", "dram_bytes", "ncu_us")}

        sm_hz = peaks.get("sm_max_mhz", 1965.0) * 1e6
        issue_peak = 148 * 4 * sm_hz / 1e9                                              # G warp-instructions / s: 4 schedulers per SM
        fp32_peak = 148 * 128 * sm_hz / 1e12                                            # T FP32 instr/s

        def roofline_of(kind):
            """issue bound (scene resident in L1/L2: configs 1-4) or HBM bound (config 5), from what the kernels really did"""
            cc, ms = class_counts(kind), cls[kind]
            if not cc or ms <= 0:
                return None
            ginst = cc["warp_inst"] / (ms * 1e-3) / 1e9
            gbs = cc["dram_bytes"] / (ms * 1e-3) / 1e9
            r = {"kernels": [k for k in class_kernels[kind] if k in counts], "launches_per_frame": cc["launches"], "ms_per_frame_live": ms,
                 "issue": {"achieved": ginst, "peak": issue_peak, "unit": "G warp-inst/s", "frac": ginst / issue_peak},
                 "lanes_per_inst": cc["thread_inst"] / max(cc["warp_inst"], 1),
                 "thread_issue_frac": ginst / issue_peak * cc["thread_inst"] / max(cc["warp_inst"], 1) / 32.0,
                 "hbm": {"achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"]},
                 "dram_bytes_per_frame": cc["dram_bytes"], "warp_inst_per_frame": cc["warp_inst"],
                 "ncu_share_check": {"ncu_us_per_frame": cc["ncu_us"], "live_us_per_frame": ms * 1e3}}
            return r

        per_kind = {k: roofline_of(k) for k in cls}
        hbm_bound = bool(w.get("synthetic"))
        rd = per_kind.get(dom)
        if rd:
            top = rd["hbm"] if hbm_bound else rd["issue"]
            roofline = {"bound": "hbm" if hbm_bound else "issue", "kernel": dom_kernel, "achieved": top["achieved"], "peak": top["peak"],
                        "unit": top["unit"], "frac": top["frac"], "traffic": rd["dram_bytes_per_frame"] / max(rd["launches_per_frame"], 1),
                        "lanes_per_inst": rd["lanes_per_inst"], "thread_issue_frac": rd["thread_issue_frac"],
                        "launches_per_frame": rd["launches_per_frame"], "peak_source": peak_src, "counts_source": counts_src,
                        "other_bound": rd["issue"] if hbm_bound else rd["hbm"],
                        "note": ("dominant kernel class = %s (%.0f %% of the frame's device time).  achieved = warp instructions the kernel "
                                 "executed (ncu smsp__inst_executed.sum of this code on this workload, profiles/kernel_counts.json) / the "
                                 "kernel time measured live in this run; peak = 148 SMs x 4 schedulers x the SM clock.  lanes_per_inst = "
                                 "active threads per executed warp instruction (32 = no divergence); thread_issue_frac weights the issue "
                                 "fraction by it.  traffic = measured DRAM bytes per launch.  The scene is L1/L2 resident, so HBM is not the "
                                 "bound (other_bound)." % (dom, 100.0 * cls[dom] / max(ms_step, 1e-9)) if not hbm_bound else
                                 "dominant kernel class = %s.  The scene (GBs) exceeds L2: achieved = measured DRAM bytes of the kernel "
                                 "(ncu dram__bytes_read+write.sum, profiles/kernel_counts.json) / the kernel time measured live; the issue "
                                 "view is under other_bound." % dom)}
        else:
            roofline = {"bound": "hbm" if hbm_bound else "issue", "kernel": dom_kernel, "achieved": None, "peak": peaks["hbm_gbs"] if hbm_bound else issue_peak,
                        "unit": "GB/s" if hbm_bound else "G warp-inst/s", "frac": None, "traffic": None, "peak_source": peak_src,
                        "note": f"no ncu counts committed for {w['key']} {args.mode}/{accel_name}/spp{w['spp']} (profiles/kernel_counts.json)"}
        own = (kinds[dom].get("own") or {}) if accel else {}
        e2e_main_t, e2e_main_bytes, e2e_main_api = t_e2e, int(host.numel() * 4), None
        if t_e2e_rgb8 is not None:
            e2e_main_t, e2e_main_bytes = t_e2e_rgb8, int(host.numel())
            e2e_main_api = ("rt_render_frame_rgb8_begin + rt_frame_wait (frame sequence: params in, the 8-bit frame of io/image/ppm.hpp:17-19 - "
                            "what the reference's main.cpp writes - out to pinned host memory every step, quantised on the device; two frames "
                            "in flight: they render on two streams into two pool sets, so frame i's thinning launches and its download run "
                            "under frame i+1's render - which is why a frame of the sequence can cost less than ms_per_step, the time of "
                            "one frame alone on the GPU; the frames were verified against rt_render_frame_rgb8)")
        float_api = ("rt_render_frame_begin + rt_frame_wait (frame sequence: params in, float frame out to pinned host memory "
                     "every step; two frames in flight on two streams and pool sets, frame i's download overlaps frame i+1's render)" if world == 1 else
                     ("rt_render_frame_device_begin per rank, rt_peer_* combine on a second stream (overlaps frame i+1's render) "
                      "with RT_PEER_OUT_HOST_RGB: every rank copies its slice of the combined float frame into the shared pinned "
                      "host frame over its own PCIe link (rt_peer_host_result_attach); d2h_bytes_per_step is the sum over ranks"
                      if shared_host else
                      "rt_render_frame_device_begin per rank, rt_peer_* combine + rt_peer_download_result on a second stream "
                      "(frame i's combine and download overlap frame i+1's render), float frame in pinned host memory on rank 0 "
                      "every step" if peer else
                      "rt_render_frame_device per rank + ncclReduce + resolve + float frame to pinned host memory on rank 0"))
        line = {
            "metric": "Mrays/s", "value": rays_all * K / (ms_dev * 1e-3) / 1e6, "unit": "Mrays/s", "n_gpus": world, "steps": K,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": data_desc(w),
            "config": workload_config(w, args, world, accel_name),
            "rays": {"primary_per_frame": int(c0.primary), "shadow_per_frame": int(c0.shadow), "secondary_per_frame": int(c0.secondary),
                     "primary_mrays_s": c0.primary / cls["primary"] / 1e3 if cls["primary"] else None,
                     "shadow_mrays_s": c0.shadow / cls["shadow"] / 1e3 if cls["shadow"] else None,
                     "secondary_mrays_s": c0.secondary / cls["secondary"] / 1e3 if cls["secondary"] else None,
                     "passes_per_frame": int(c0.passes),
                     "ms": {k: v / K for k, v in acc.items()},
                     "ms_source": ("per-class times from serial frames with a CUDA event pair around every launch (which also "
                                   "switches programmatic dependent launch off between them); ms_per_step is the queued frame"
                                   if (world == 1 or peer) else "CUDA event pair around every launch of the timed frames")},
            "e2e": {"value": rays_all * K / e2e_main_t / 1e6, "unit": "Mrays/s", "ms_per_frame": 1e3 * e2e_main_t / K,
                    "h2d_bytes_per_step": int(np.dtype(np.uint8).itemsize * __import__("ctypes").sizeof(rt.Params)),
                    "d2h_bytes_per_step": e2e_main_bytes,
                    "api": e2e_main_api or float_api,
                    "vs_device_time": (1e3 * e2e_main_t / K) / ms_step,
                    "float_sequence": {"value": rays_all * K / t_e2e / 1e6, "ms_per_frame": 1e3 * t_e2e / K, "d2h_bytes_per_step": int(host.numel() * 4),
                                       "api": float_api},
                    "one_call_per_frame": ({"value": rays_all * K / t_e2e_sync / 1e6, "ms_per_frame": 1e3 * t_e2e_sync / K,
                                            "d2h_bytes_per_step": int(host.numel() * 4),
                                            "api": ("rt_render_frame (synchronous: render, download the float frame, return)" if world == 1 else
                                                    "serial: render, combine, download, per frame")} if (world == 1 or peer) else None)},
            "gpu_launches": int(launches),
            "roofline": roofline,
            # every ray kind against both bounds, from the same measured counts (north star: "each number as an absolute value and
            # as a fraction of its roofline ... achieved L2/HBM GB/s for traversal fetches ... issue utilisation for intersection")
            "roofline_by_kind": {k: (dict(per_kind[k], mrays_s=(n / cls[k] / 1e3 if cls[k] else None)) if per_kind[k] else
                                     {"mrays_s": (n / cls[k] / 1e3 if cls[k] else None)})
                                 for k, n in (("primary", int(c0.primary)), ("secondary", int(c0.secondary)), ("shadow", int(c0.shadow)))},
            # the REFERENCE algorithm's work for the same rays (8 B/node + 36 B/triangle test + ray/hit I/O,
            # 24 flop/node + 45 flop/triangle, counted by the oracle on the reference's tree in the reference's visit order) per
            # second of this backend's kernel time.  NOT a roofline fraction: the accelerated query does a small part of that work
            # (e.g. 11 instead of 237 triangle tests per config-2 shadow ray), so these rates can exceed the hardware's peaks.
            "reference_equivalent": {k: {"alg_bytes_per_frame": kinds[k]["alg_bytes"] * spp_scale, "alg_flop_per_frame": kinds[k]["alg_flop"] * spp_scale,
                                         "gbs": (kinds[k]["alg_bytes"] * spp_scale / (cls[k] * 1e-3) / 1e9 if cls[k] else None),
                                         "tflops": (kinds[k]["alg_flop"] * spp_scale / (cls[k] * 1e-3) / 1e12 if cls[k] else None)}
                                     for k in cls},
            "own_traversal_bytes": ({"kernel": dom_kernel, "structure": own.get("structure"), "own_bytes_per_frame": own.get("own_bytes"),
                                     "gbs": own["own_bytes"] * spp_scale / (cls[dom] * 1e-3) / 1e9,
                                     "note": "bytes the two-wide traversal asks for (64 B per node visit + 48 B per triangle record + ray/hit "
                                             "I/O), counted by running its source on the CPU over the same queries; served by L1/L2"}
                                    if own.get("own_bytes") and cls[dom] > 0 else None),
            "clocks": clocks.summary(),
            "wall_s_timed_region": t_wall,
            "combine": (None if world == 1 else ("rt_peer_*: fused wait+reduce+resolve over NVLink peer memory, frame i-1's combine on a "
                                                 "second stream inside frame i's timed step (two frame slots per rank)" if peer else
                                                 "ncclReduce(sum) to rank 0 + rt_resolve_sum_device")),
            "verify": verify,
            "north_star_scaling": ns,
            "scene": {"triangles": int(scene.info.n_triangles), "kd_nodes": int(scene.info.n_nodes), "packets": int(scene.info.n_packets),
                      "bvh_nodes": int(scene.info.bvh_n_nodes), "bvh_depth": int(scene.info.bvh_depth),
                      "device_bytes": int(scene.info.device_bytes), "host_build_s": round(t_build, 3),
                      "accel_build": "device" if int(scene.info.accel_build) == rt.ACCEL_BUILD_DEVICE else "host",
                      "accel_build_s": round(float(scene.info.accel_build_seconds), 4),
                      "kd_build_s": round(float(scene.info.build_seconds), 3), "bvh4_stack_need": int(scene.info.bvh4_stack_need),
                      "bvh_leaf_size": int(scene.info.bvh_leaf_size)},
        }
        if world == 1 and not args.no_cpu_baseline:
            if w.get("synthetic"):
                w = load_workload(args.workload, args.tris, cpu_only=True)
            base = cpu_reference_frames(w, frames=3 if w.get("synthetic") else 20, warmup=1 if w.get("synthetic") else 2)
            line["cpu_baseline"] = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}
            line["cpu_baseline"]["ms_per_frame"] = base["ms_per_frame"]
        print(json.dumps(line), flush=True)
    if peer:
        barrier()
        host_frames = None                       # a view of the mapping peer.close() unmaps
        peer.close()
    scene.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
