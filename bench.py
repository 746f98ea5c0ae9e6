#!/usr/bin/env python3
"""bench.py — Mrays/s of the per-pixel trace loop (BASELINE.json metric) on 1..8 B200.

A "step" is one frame of the workload rendered by the hot path: ray generation, wavefront
trace/shade/shadow loop over all bounce levels, device-side 8-bit resolve and (N > 1) the
NCCL gather of the interleaved tiles to rank 0.  A "ray" is one reference castRay call
(primary + shadow + secondary; reference src/scene.cpp:65,91,127,134).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

N > 1 is launched by the driver with torch.distributed.run (one rank per GPU, NCCL).
`--impl reference` times the reference's own CPU implementation of the path
(oracle/_ref/libref.so = the unmodified reference sources behind a C shim) on the host
cores, each step a bounded pixel sample of the same workload.

Only the cpu_baseline leg and `--impl reference` execute anything under oracle/.
"""
import argparse
import ctypes as C
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
PKG_DIR = ROOT / "cs184-raytracer_b200"
GOLDEN = ROOT / "tests" / "golden"

WORKLOADS = {
    # name: (scene spec, width, height, depth, description)
    "synthetic": (("synthetic", 708, 1000, 184), 7680, 4320, 5,
                  "BASELINE configs[4]: synthetic 1,002,528-triangle height field + 1000 spheres, 8 shadow lights, depth 5, 7680x4320"),
    "synthetic4k": (("synthetic", 708, 1000, 184), 3840, 2160, 5,
                    "configs[4] scene at 3840x2160 (the >=1 Grays/s target is stated at 4K)"),
    "teapot": (("file", "inputs/input-02.rti"), 1920, 1080, 10, "BASELINE configs[1]: teapot.obj scene at 1920x1080, depth 10"),
    "refraction3": (("file", "excess_inputs/refraction3.rti"), 3840, 2160, 10, "BASELINE configs[2]: refraction3.rti at 3840x2160"),
    "bunny": (("file", "excess_inputs/bunny4.rti"), 3840, 2160, 10, "BASELINE configs[3] stand-in: bunny.obj, 4 lights, 3840x2160"),
    "input01": (("file", "inputs/input-01.rti"), 1000, 1000, 10, "BASELINE configs[0]: input-01.rti at 1000x1000"),
}


def load_package():
    name = "cs184_raytracer_b200"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, PKG_DIR / "__init__.py", submodule_search_locations=[str(PKG_DIR)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def build_scene(pkg, spec):
    if spec[0] == "synthetic":
        return pkg.HostScene.synthetic(spec[1], spec[2], spec[3])
    return pkg.HostScene.load(GOLDEN / spec[1])


# ---------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the unmodified reference hot path on the host cores
# ---------------------------------------------------------------------------------------
class RefLib:
    def __init__(self, counting):
        path = ROOT / "oracle" / "_ref" / ("libref_count.so" if counting else "libref.so")
        if not path.exists():
            raise FileNotFoundError(str(path))
        self.lib = C.CDLL(str(path))
        self.lib.ref_scene_load.restype = C.c_void_p
        self.lib.ref_scene_load.argtypes = [C.POINTER(C.c_char_p), C.c_int, C.c_char_p, C.c_int]
        self.lib.ref_scene_from_flat.restype = C.c_void_p
        self.lib.ref_scene_from_flat.argtypes = [C.c_void_p, C.c_char_p, C.c_int]

    def scene(self, pkg, spec, host_scene):
        err = C.create_string_buffer(512)
        if spec[0] == "synthetic":
            h = self.lib.ref_scene_from_flat(C.cast(host_scene.flat, C.c_void_p), err, 512)
        else:
            arr = (C.c_char_p * 1)(str(GOLDEN / spec[1]).encode())
            h = self.lib.ref_scene_load(arr, 1, err, 512)
        if not h:
            raise RuntimeError(err.value.decode())
        return C.c_void_p(h)

    def render(self, h, w, hh, depth, threads):
        rgb = np.zeros((hh, w, 3))
        sec, calls = C.c_double(), C.c_uint64()
        self.lib.ref_render(h, w, hh, depth, 0, threads, rgb.ctypes.data_as(C.c_void_p), None, C.byref(sec), C.byref(calls))
        return sec.value, int(calls.value)


def sample_size(width, height, px_budget):
    """Same camera, same aspect, fewer pixels (the ray mix is resolution independent)."""
    if width * height <= px_budget:
        return width, height
    s = (px_budget / (width * height)) ** 0.5
    return max(16, int(width * s)), max(9, int(height * s))


def reference_runs(pkg, wl_name, steps, warmup, target_s):
    """Times the reference on a bounded sample.  Returns dict(value Mrays/s, ms_per_step, ...)."""
    spec, width, height, depth, _ = WORKLOADS[wl_name]
    threads = os.cpu_count() or 1
    # silence the parsers' per-line .obj warnings
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(2)
    os.dup2(devnull, 2)
    try:
        host_scene = build_scene(pkg, spec)
        timed, counted = RefLib(False), RefLib(True)
        h_t = timed.scene(pkg, spec, host_scene)
        h_c = counted.scene(pkg, spec, host_scene)
    finally:
        os.dup2(saved, 2)
        os.close(devnull)
        os.close(saved)
    # calibrate on a tiny frame, then size the sample for ~target_s per step; the counting
    # build (castRay counter behind a linker --wrap) renders the sample once: that run gives
    # the exact ray count and doubles as the first warm-up
    cw, ch = sample_size(width, height, 12 * 7)
    t_cal, _ = timed.render(h_t, cw, ch, depth, threads)
    per_px = max(t_cal / (cw * ch), 1e-9)
    sw, sh = sample_size(width, height, max(cw * ch, int(target_s / per_px)))
    t_cnt, rays = counted.render(h_c, sw, sh, depth, threads)
    if t_cnt < 0.4 * target_s and (sw, sh) != (width, height):      # calibration was pessimistic: resize once
        sw, sh = sample_size(width, height, int(sw * sh * 0.8 * target_s / max(t_cnt, 1e-6)))
        t_cnt, rays = counted.render(h_c, sw, sh, depth, threads)
    for _ in range(max(warmup - 1, 0)):
        timed.render(h_t, sw, sh, depth, threads)
    times = [timed.render(h_t, sw, sh, depth, threads)[0] for _ in range(steps)]
    total = sum(times)
    return {
        "value": rays * steps / total / 1e6, "ms_per_step": 1e3 * total / steps, "cores": threads,
        "sample": f"{sw}x{sh} pixel subsample of the {width}x{height} frame (same camera, depth {depth}), "
                  f"{rays} rays per step, all {threads} host threads, unmodified reference traceRay/castRay",
        "rays_per_step": rays, "sample_w": sw, "sample_h": sh,
    }


# ---------------------------------------------------------------------------------------
# clocks sampler
# ---------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons during the timed region: NVML polled every 10 ms from a thread (the same
    counters nvidia-smi prints; the timed region of a small workload is shorter than one nvidia-smi period),
    falling back to `nvidia-smi -lms 100` when the NVML binding is unavailable."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw"
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.index, self.rows, self.proc, self.thread = index, [], None, None
        self.stop_flag, self.nvml, self.source = threading.Event(), None, None

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [x.strip() for x in vis.split(",") if x.strip()]
            if self.index < len(ids) and ids[self.index].isdigit():
                return int(ids[self.index])
        return self.index

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            bits = [pynvml.nvmlClocksThrottleReasonHwSlowdown, pynvml.nvmlClocksThrottleReasonHwThermalSlowdown,
                    pynvml.nvmlClocksThrottleReasonSwThermalSlowdown, pynvml.nvmlClocksThrottleReasonSwPowerCap]

            def poll():
                while not self.stop_flag.is_set():
                    try:
                        sm = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                        r = int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                        self.rows.append([sm, mx] + ["Active" if r & b else "Not Active" for b in bits] + [0.0])
                    except Exception:
                        pass
                    self.stop_flag.wait(0.01)
            self.nvml, self.source = pynvml, "nvml, 10 ms"
            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self._physical_index()}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return
        self.source = "nvidia-smi -lms 100"

        def pump():
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self):
        if self.nvml is not None:
            self.stop_flag.set()
            if self.thread:
                self.thread.join(timeout=2)
        elif self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
            if self.thread:
                self.thread.join(timeout=2)
        else:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except ValueError:
                continue
            for k, nm in enumerate(self.NAMES):
                if str(r[2 + k]).lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": self.source}


# ---------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="synthetic", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus != 1:
        print(f"bench.py: --gpus {args.gpus} needs torch.distributed.run (WORLD_SIZE={world})", file=sys.stderr)
        sys.exit(2)
    spec, width, height, depth, desc = WORKLOADS[args.workload]
    pkg = load_package()
    config = {"workload": f"{args.workload}: {desc}", "width": width, "height": height, "bounce_depth": depth,
              "partition": f"interleaved 32x32 tiles over {world} rank(s), scene replicated",
              "l2": "no explicit flush: every frame streams its ray and hit queues (80 B / 116 B per record, 10^6..10^8 records per bounce "
                    "level) through the 126 MB L2 between two uses of any scene data"}

    # -------------------------------------------------------------- reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        try:
            r = reference_runs(pkg, args.workload, args.steps, max(args.warmup, 0), target_s=6.0)
        except FileNotFoundError as e:
            print(json.dumps({"impl": "reference", "unavailable": f"reference shim not built: {e}"}))
            return
        line = {"impl": "reference", "metric": "Mrays/s (primary+shadow+secondary)", "value": r["value"], "unit": "Mrays/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": r["value"], "unit": "Mrays/s", "cores": r["cores"], "kind": "reference",
                                 "sample": r["sample"]},
                "e2e": {"value": r["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # -------------------------------------------------------------- B200 arm
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    host_scene = build_scene(pkg, spec)
    ren = pkg.Renderer(local_rank)
    ren.upload(host_scene)
    up_stats = ren.stats()
    stream = torch.cuda.current_stream().cuda_stream
    base_flags = pkg.RT_FLAG_TIME_KERNELS
    p = pkg.make_params(width, height, depth, tile_rank=rank, tile_world=world, flags=base_flags)
    emulate = os.environ.get("RT_BENCH_EMULATE_RANK")      # "r/n": time rank r's share of an n-rank job on one GPU (development aid)
    if emulate and world == 1:
        er, en = (int(x) for x in emulate.split("/"))
        p = pkg.make_params(width, height, depth, tile_rank=er, tile_world=en, flags=base_flags)
        config["partition"] = f"EMULATED rank {er} of {en} (its tiles only, no gather)"
    own_tiles, max_tiles, total_tiles = pkg.tile_counts(p)
    if world == 1 and emulate:
        out = torch.zeros(max_tiles * pkg.RT_TILE_PIXELS * 3, dtype=torch.uint8, device=dev)
        gathered = frame = None
        args.no_e2e = True
    elif world == 1:
        out = torch.empty(height * width * 3, dtype=torch.uint8, device=dev)
        gathered = frame = None
    else:
        out = torch.zeros(max_tiles * pkg.RT_TILE_PIXELS * 3, dtype=torch.uint8, device=dev)
        # rank r's tiles land in slice r of one buffer: the gather writes them in place (no concatenation pass)
        packed_all = torch.empty(world * out.numel(), dtype=torch.uint8, device=dev) if rank == 0 else None
        gathered = list(packed_all.chunk(world)) if rank == 0 else None
        frame = torch.empty(height * width * 3, dtype=torch.uint8, device=dev) if rank == 0 else None

    def step():
        ren.render_device(p, out.data_ptr(), rgb8=True, stream=stream)
        if world > 1:
            dist.gather(out, gathered, dst=0)
            if rank == 0:
                ren.unpack_tiles(p, packed_all.data_ptr(), frame.data_ptr(), rgb8=True, stream=stream)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 0)):
        step()
    sync_all()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    agg = {"ms_kernel": np.zeros(4), "launches_kernel": np.zeros(4), "launches": 0}
    e0.record()
    for _ in range(args.steps):
        step()
        st = ren.stats()
        agg["ms_kernel"] += np.array(st["ms_kernel"]); agg["launches_kernel"] += np.array(st["launches_kernel"])
        agg["launches"] += st["kernel_launches"] + (1 if (world > 1 and rank == 0) else 0)
    e1.record()
    sync_all()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    rays = torch.tensor([st["rays_primary"], st["rays_shadow"], st["rays_secondary"], agg["launches"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
        dist.all_reduce(rays, op=dist.ReduceOp.SUM)
    ms_total = float(ms_total.item())
    rays_cls = [int(x) for x in rays[:3].tolist()]
    rays_frame = sum(rays_cls)
    launches_total = int(rays[3].item())
    value = rays_frame * args.steps / (ms_total * 1e-3) / 1e6

    # ---------------- end-to-end: scene upload (H2D) + LBVH build + render + gather + D2H, host buffers
    e2e = None
    if not args.no_e2e:
        flat = host_scene.flat.contents
        cudart = torch.cuda.cudart()
        nf = flat.num_faces
        pinned = []
        for ptr in (flat.face_points, flat.face_normals):
            addr = C.cast(ptr, C.c_void_p).value
            if nf and addr and cudart.cudaHostRegister(addr, nf * 72, 0) in (0, cudart.cudaError.success):
                pinned.append(addr)
        host_frame = torch.empty(height * width * 3, dtype=torch.uint8, pin_memory=True)
        pe = pkg.make_params(width, height, depth, tile_rank=rank, tile_world=world)

        def e2e_step():
            ren.upload(host_scene)
            ren.render_device(pe, out.data_ptr(), rgb8=True, stream=stream)
            if world > 1:
                dist.gather(out, gathered, dst=0)
                if rank == 0:
                    ren.unpack_tiles(pe, packed_all.data_ptr(), frame.data_ptr(), rgb8=True, stream=stream)
                    host_frame.copy_(frame, non_blocking=True)
            else:
                host_frame.copy_(out, non_blocking=True)
            torch.cuda.synchronize()

        e2e_step()
        sync_all()
        n_e2e = max(1, min(args.steps, 3))
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            e2e_step()
        sync_all()
        t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        st_e = ren.stats()
        e2e = {"value": rays_frame * n_e2e / float(t_e2e.item()) / 1e6, "unit": "Mrays/s",
               "h2d_bytes_per_step": int(st_e["scene_bytes_h2d"]) * world, "d2h_bytes_per_step": height * width * 3,
               "ms_per_step": 1e3 * float(t_e2e.item()) / n_e2e,
               "ms_scene_upload": st_e["ms_upload"], "ms_lbvh_build": st_e["ms_build"], "steps": n_e2e,
               "path": "rt_scene_upload (host arrays, pinned) + rt_render_device_rgb8 + NCCL gather + D2H into pinned host frame"}
        for addr in pinned:
            cudart.cudaHostUnregister(addr)

    # ---------------- roofline of the dominant kernel (rank 0)
    roofline = None
    if rank == 0:
        # one frame with the kernels strictly serialised: durations per kernel class that add up to the frame
        # (in the timed region k_shadow of level l runs beside k_trace / k_shade of level l+1)
        ps = pkg.make_params(width, height, depth, tile_rank=p.tile_rank, tile_world=p.tile_world,
                             flags=pkg.RT_FLAG_TIME_KERNELS | pkg.RT_FLAG_SERIAL)
        ren.render_device(ps, out.data_ptr(), rgb8=True, stream=stream)
        serial = ren.stats()
        pc = pkg.make_params(width, height, depth, tile_rank=p.tile_rank, tile_world=p.tile_world, flags=pkg.RT_FLAG_COUNT_WORK)
        ren.render_device(pc, out.data_ptr(), rgb8=True, stream=stream)
        cs = ren.stats()
        names = ["k_trace", "k_shade", "k_shadow", "other"]
        # dominance from the serialised frame (the overlapped k_trace bracket also counts its wait for SMs)
        dom = int(np.argmax(np.array(serial["ms_kernel"])[:3]))
        nsl = cs["rays_shadow"] // max(cs["hits"], 1)
        # algorithmic bytes: 32 B per BVH child box tested, 80 B per exact
        # face test, 128 B per sphere test, + the kernel's queue records
        if dom == 2:
            k = 1
            queue = cs["hits"] * 116 + cs["rays_shadow"] * 24
        elif dom == 0:
            k = 0
            queue = (cs["rays_primary"] + cs["rays_secondary"]) * 80 + cs["hits"] * 116
        else:
            k = None
            queue = cs["hits"] * (116 + 24) + cs["rays_secondary"] * 80
        bytes_frame = queue
        if k is not None:
            bytes_frame += 32 * cs["nodes_fetched"][k] + 80 * cs["tris_tested"][k] + 128 * cs["spheres_tested"][k]
        n_launch = max(agg["launches_kernel"][dom] / args.steps, 1)
        if dom == 0:
            # k_trace shares the GPU with the previous level's k_shadow in the timed region: take its
            # duration from the serialised frame instead
            ms_launch = serial["ms_kernel"][0] / max(serial["launches_kernel"][0], 1)
            dur_src = "RT_FLAG_SERIAL frame after the timed region (CUDA events on the launching stream)"
        else:
            ms_launch = agg["ms_kernel"][dom] / max(agg["launches_kernel"][dom], 1)
            dur_src = "timed region (CUDA events on the stream the kernel is launched on)"
        achieved = bytes_frame / n_launch / (ms_launch * 1e-3) / 1e9
        peak, peak_src = 6650.0, "fallback"
        try:
            mp = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
            peak, peak_src = float(mp["hbm_gbs"]), "measured"
        except Exception:
            pass
        node_bytes, face_bytes = ren.device_bytes()
        gather = ren.microbench_gather(max(node_bytes, 1 << 20), 64)
        traffic = None
        tf = ROOT / "profiles" / "traffic.json"
        if tf.exists():
            try:
                traffic = json.loads(tf.read_text()).get(args.workload, {}).get(names[dom])
            except Exception:
                traffic = None
        roofline = {"bound": "hbm", "kernel": names[dom], "achieved": achieved, "peak": peak, "peak_source": peak_src,
                    "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                    "bytes_per_launch": bytes_frame / n_launch, "ms_per_launch": ms_launch, "launches_per_step": n_launch,
                    "duration_source": dur_src,
                    "kernel_share_of_step": float(serial["ms_kernel"][dom] / max(serial["ms_trace"], 1e-9)),
                    "ms_kernel_per_step": {n: float(v / args.steps) for n, v in zip(names, agg["ms_kernel"])},
                    "ms_kernel_serial_frame": {n: float(v) for n, v in zip(names, serial["ms_kernel"])},
                    "ms_serial_frame": float(serial["ms_trace"]),
                    "note": "ms_kernel_per_step: event-bracketed durations inside the timed region, where k_trace/k_shade of "
                            "bounce level l+1 run beside k_shadow of level l (k_trace's bracket includes waiting for SMs); "
                            "ms_kernel_serial_frame: one extra frame with RT_FLAG_SERIAL, classes add up to ms_serial_frame",
                    "node_gather_gbs": gather, "frac_of_node_gather": achieved / gather if gather else None,
                    "node_array_bytes": node_bytes, "face_record_bytes": face_bytes,
                    "work_per_step": {"nodes": cs["nodes_fetched"], "tris": cs["tris_tested"], "spheres": cs["spheres_tested"],
                                      "hits": cs["hits"], "shadow_lights": int(nsl)}}

    # ---------------- cpu baseline (rank 0, N = 1)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            r = reference_runs(pkg, args.workload, 1, 0, target_s=12.0)
            cpu = {"value": r["value"], "unit": "Mrays/s", "cores": r["cores"], "kind": "reference", "sample": r["sample"]}
        except FileNotFoundError as e:
            cpu = {"value": None, "unit": "Mrays/s", "cores": 0, "kind": "reference", "sample": f"unavailable: {e}"}

    if rank == 0:
        line = {"metric": "Mrays/s (primary+shadow+secondary)", "value": value, "unit": "Mrays/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "rays_per_step": {"primary": rays_cls[0], "shadow": rays_cls[1], "secondary": rays_cls[2]},
                "frame_ms": ms_total / args.steps, "ms_scene_upload": up_stats["ms_upload"], "ms_lbvh_build": up_stats["ms_build"],
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches_total, "roofline": roofline, "cpu_baseline": cpu}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
