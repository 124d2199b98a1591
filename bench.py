#!/usr/bin/env python
"""bench.py -- FPN multi-level RoIAlign fwd+bwd throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A "step" is one pass of the hot path over one synthetic batch: plan (level
assignment + binning) + fused forward + backward (zero-fill + scatter).  The
workload is BASELINE.json configs[1] (mask head 14x14, 2 images at 1333x800,
2048 RoIs/image, P2-P5, 256 channels fp32) per GPU; with N > 1 every rank runs
that workload on its own images (sharded by image, no data-path collective,
weak scaling) and NCCL only reduces the timings.

Prints ONE JSON line on rank 0 (see DESIGN.md "Measurement").
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
sys.path.insert(0, os.path.join(ROOT, "tests"))

import synth  # noqa: E402

METRIC = "roialign_fwd_bwd_throughput"
UNIT = "RoIs/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=1, help="index into BASELINE.json configs (0..3)")
    ap.add_argument("--sampling-ratio", type=int, default=2)
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="skip the CUDA-graph replay timing")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true")
    ap.add_argument("--gpu-baseline-rois", type=int, default=512)
    ap.add_argument("--cpu-sample-rois", type=int, default=0, help="0 = sized automatically")
    ap.add_argument("--deterministic", action="store_true",
                    help="time the deterministic (segmented reduction) backward instead of the atomic one")
    ap.add_argument("--tune", default="", help="comma list key=value for rpool_set_tuning")
    ap.add_argument("--shard", action="store_true",
                    help="strong scaling: ONE instance of the config, its images dealt round-robin to the "
                         "ranks (BASELINE.json configs[3]); default is the config on every rank (weak)")
    return ap.parse_args()


def workload(cfg_id, rank):
    cfg = dict(synth.CONFIGS[cfg_id])
    rng = np.random.RandomState(cfg_id + 1000 * rank)
    L = cfg["n_levels"]
    shapes = synth.pyramid_shapes(cfg["n_images"], cfg["channels"], cfg["height"], cfg["width"], L)
    rois = synth.make_rois(rng, cfg["n_images"], cfg["rois_per_image"], cfg["height"], cfg["width"],
                           aspect_range=cfg["aspect"])
    scales = [1.0 / s for s in synth.STRIDES[:L]]
    return cfg, rng, shapes, rois, scales


def algorithmic_bytes(cfg, shapes, rois, levels, scales, S):
    """fwd = O + U*C*4 + 20R; bwd = O + F + 20R."""
    C = cfg["channels"]
    R = rois.shape[0]
    O = sum(R * C * P * P * 4 for P in cfg["out_sizes"])
    F = sum(int(np.prod(s)) * 4 for s in shapes)
    U = synth.window_cells_touched(rois, levels, shapes, scales, max(cfg["out_sizes"]) * max(S, 1))
    return dict(O=O, F=F, U_bytes=U * C * 4, fwd=O + U * C * 4 + 20 * R, bwd=O + F + 20 * R)


def ncu_traffic(cfg_id, S, kernels):
    """DRAM bytes per launch of the named kernels from the committed ncu --set full
    capture (profiles/ncu_traffic.json); None when that workload was not captured."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f)["cfg%d_S%d" % (cfg_id, S)]
        return float(sum(t[k]["read"] + t[k]["write"] for k in kernels))
    except Exception:  # noqa: BLE001
        return None


def ncu_view(cfg_id, S):
    """What the committed ncu capture says about the pooling kernels of this workload:
    DRAM GB/s (dram bytes / gpu__time_duration) and L2 hit rate per kernel, against the
    nominal 8 TB/s of the part (north_star) -- None when the workload was not captured."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f)["cfg%d_S%d" % (cfg_id, S)]
        out = {"source": "profiles/ncu_traffic.json (ncu --set full --clock-control none, one launch each)"}
        for name, v in t.items():
            if "us" not in v:
                continue
            gbs = (v["read"] + v["write"]) / (v["us"] * 1e-6) / 1e9
            out[name] = {"dram_GBps": gbs, "frac_of_8TBps": gbs / 8000.0, "l2_hit_pct": v.get("l2_hit_pct"),
                         "us": v["us"]}
        return out
    except Exception:  # noqa: BLE001
        return None


def reference_numpy_numbers(cfg_name):
    """The reference's own NumPy path cannot travel to the GPU box; what it does on this
    workload was measured in the development container by tools/reference_numpy_timing.py
    and is quoted from the committed result."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_reference_numpy_cpu.json")) as f:
            d = json.load(f)
        c = d["configs"][cfg_name]
        return {"rois_per_s_1core": c["one_core"]["rois_per_s"],
                "rois_per_s_all_cores": c["all_cores"]["rois_per_s"], "cores": c["all_cores"]["processes"],
                "sample_rois": c["sample_rois"], "where": d["where"],
                "source": "profiles/r01_reference_numpy_cpu.json (tools/reference_numpy_timing.py)"}
    except Exception:  # noqa: BLE001
        return None


def bind_near_device(local_index):
    """Restricts this process to the CPUs NVML lists as the device's ideal affinity, so that
    the pinned host buffers of the end-to-end leg are first-touched on the GPU's NUMA node
    (a box with the buffers on the far socket copied 30 % slower).  Returns (old_mask,
    n_cpus) or (None, 0) when NVML or the mask is unavailable; restore with
    os.sched_setaffinity(0, old_mask)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = [v.strip() for v in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if v.strip()]
        idx = int(vis[local_index]) if vis and local_index < len(vis) and vis[local_index].isdigit() \
            else local_index
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        old = os.sched_getaffinity(0)
        cpus &= old
        if not cpus:
            return None, 0
        os.sched_setaffinity(0, cpus)
        return old, len(cpus)
    except Exception:  # noqa: BLE001
        return None, 0


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(object):
    """SM clock and clock-event (throttle) reasons sampled DURING the timed region.

    NVML is polled from a thread every 2 ms (the timed region of the default run is
    tens of milliseconds: `nvidia-smi -lms` cannot sample faster than 100 ms); the
    `nvidia-smi` loop is the fallback when the NVML binding is missing."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []          # nvidia-smi fallback
        self.sm, self.reasons, self.power = [], set(), []
        self.sm_max = None
        self.proc = None
        self.gpu = gpu_index
        self.nvml = None
        self.stop_flag = threading.Event()
        self.t = None

    def _nvml_index(self):
        # CUDA_VISIBLE_DEVICES remaps torch's device order; NVML enumerates the physical parts
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        ids = [v.strip() for v in vis.split(",") if v.strip()]
        if ids and self.gpu < len(ids) and ids[self.gpu].isdigit():
            return int(ids[self.gpu])
        return self.gpu

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._nvml_index())
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:  # noqa: BLE001
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _poll(self):
        nv = self.nvml
        names = [("hw_slowdown", "nvmlClocksEventReasonHwSlowdown"),
                 ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown"),
                 ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown"),
                 ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap"),
                 ("hw_power_brake_slowdown", "nvmlClocksEventReasonHwPowerBrakeSlowdown")]
        bits = [(n, getattr(nv, a)) for n, a in names if hasattr(nv, a)]
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons", None)
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
                if get_reasons is not None:
                    r = int(get_reasons(self.handle))
                    for n, b in bits:
                        if r & b:
                            self.reasons.add(n)
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.handle) / 1000.0)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.002)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nvml is not None:
            self.stop_flag.set()
            self.t.join(timeout=2)
            out = {"sm_mhz": float(np.median(self.sm)) if self.sm else None,
                   "sm_min_mhz": float(min(self.sm)) if self.sm else None,
                   "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
                   "samples": len(self.sm), "power_w_max": max(self.power) if self.power else None,
                   "how": "NVML polled every 2 ms inside the timed region"}
            try:
                self.nvml.nvmlShutdown()
            except Exception:  # noqa: BLE001
                pass
            return out
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:  # noqa: BLE001
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "how": "nvidia-smi -lms 100"}


# ---------------------------------------------------------------------------
# CPU arm: the oracle port (and the reference's own C++ forward) on the host cores
# ---------------------------------------------------------------------------
_CPU_DATA = {}


def cpu_arm(cfg_id, S, sample_rois, repeats=1):
    """Times the CPU restatement of the reference path (oracle/, all host threads)
    on a bounded sample of the same workload.  Returns (rois_per_s, info)."""
    import oracle
    if cfg_id not in _CPU_DATA:
        cfg, rng, shapes, rois, scales = workload(cfg_id, 0)
        feats = [rng.standard_normal(s).astype(np.float32) for s in shapes]
        _CPU_DATA[cfg_id] = (cfg, rng, shapes, rois, scales, feats)
    cfg, rng, shapes, rois, scales, feats = _CPU_DATA[cfg_id]
    threads = oracle.max_threads()
    if sample_rois <= 0:
        sample_rois = min(rois.shape[0], 4096)
    sample_rois = min(sample_rois, rois.shape[0])
    sel = np.sort(np.random.RandomState(99).choice(rois.shape[0], sample_rois, replace=False))
    sub = rois[sel]
    levels = oracle.levels_for_pyramid(sub[:, 1:], cfg["n_levels"])
    mode = "chainer" if S == 1 else "caffe2"
    gys = [synth.make_gy(np.random.RandomState(7), sub.shape[0], cfg["channels"], P)
           for P in cfg["out_sizes"]]
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        for P in cfg["out_sizes"]:
            oracle.fpn_forward(feats, sub, levels, scales, P, mode, S, threads=threads)
        t1 = time.perf_counter()
        for g in gys:
            oracle.fpn_backward(g, shapes, sub, levels, scales, mode, S, threads=threads)
        t2 = time.perf_counter()
        if best is None or (t2 - t0) < best[0]:
            best = (t2 - t0, t1 - t0, t2 - t1)
    info = {
        "value": sub.shape[0] / best[0], "unit": UNIT, "cores": threads, "kind": "port",
        "sample": "%d of %d RoIs of %s (drawn without replacement, seed 99); C port of the "
                  "reference path (oracle/roialign_oracle.c, %s semantics, sampling_ratio %d), "
                  "OpenMP over RoIs (fwd) / channels (bwd); fwd %.3f s + bwd %.3f s"
                  % (sub.shape[0], rois.shape[0], cfg["name"], mode, S, best[1], best[2]),
        "fwd_s": best[1], "bwd_s": best[2],
    }
    if oracle.have_ref() and len(cfg["out_sizes"]) == 1:
        # the reference's own compiled C++ forward, single-threaded as shipped
        n = min(sub.shape[0], 128)
        rois_xy = oracle.roi_yx_to_xy(sub[:n])
        t0 = time.perf_counter()
        P = cfg["out_sizes"][0]
        for l in range(cfg["n_levels"]):
            m = np.nonzero(levels[:n] == l)[0]
            if m.size:
                oracle.ref_caffe2_forward(feats[l], rois_xy[m], P, P, scales[l], max(S, 1))
        info["reference_cpp_forward_rois_per_s_1thread"] = n / (time.perf_counter() - t0)
    info["reference_numpy_path"] = reference_numpy_numbers(cfg["name"])
    return info["value"], info


def gpu_baseline_arm(cfg_id, sample_rois, device, repeats=3):
    """Same-box GPU baseline: the reference's CuPy kernels restated in CUDA
    (baseline/refgpu_baseline.cu) and dispatched per RoI like the reference's FPN
    heads, on a bounded sample of the workload (its cost grows with the level
    map, not with the RoI: every backward call touches a whole dense gradient)."""
    import torch
    from baseline import refgpu
    from chainer_maskrcnn_b200 import _engine
    cfg, rng, shapes, rois, scales = workload(cfg_id, 0)
    sample_rois = min(sample_rois, rois.shape[0])
    sel = np.sort(np.random.RandomState(99).choice(rois.shape[0], sample_rois, replace=False))
    sub = rois[sel]
    # levels by the product's own device mapper (bit-exact image of the reference rule)
    levels = _engine.assign_levels(torch.from_numpy(sub).to(device), as_int=True,
                                   k_cap=cfg["n_levels"] - 1).cpu().numpy()
    feats = [torch.randn(s, device=device, dtype=torch.float32) for s in shapes]
    state = refgpu.FpnState(feats, scales)
    rois_xy = torch.from_numpy(np.ascontiguousarray(sub[:, [0, 2, 1, 4, 3]])).to(device)
    total_ms, ops = 0.0, 0
    for P in cfg["out_sizes"]:
        top = torch.empty((sample_rois, cfg["channels"], P, P), device=device)
        gy = torch.rand_like(top) * 2 - 1
        refgpu.fpn_step(state, rois_xy, levels, P, top, gy)       # warm-up
        torch.cuda.synchronize()
        best = None
        for _ in range(repeats):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            n_ops = refgpu.fpn_step(state, rois_xy, levels, P, top, gy)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        total_ms += best
        ops += n_ops
    return {"value": sample_rois / (total_ms * 1e-3), "unit": UNIT, "kind": "reference CuPy kernels restated "
            "(baseline/refgpu_baseline.cu), per-RoI dispatch of fpn_roi_mask_head.py:57-63, sampling_ratio 1, "
            "NCHW", "sample": "%d of %d RoIs of %s (seed 99)" % (sample_rois, rois.shape[0], cfg["name"]),
            "ms": total_ms, "device_ops": ops}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    oracle.build()
    # bounded sample per step, sized so that the whole --steps K run stays within a few minutes
    # on the host cores (the port does ~5 k RoIs/s on 16 cores: about 60 s of work in total)
    sample = args.cpu_sample_rois if args.cpu_sample_rois > 0 else \
        min(1024, max(64, 300000 // max(args.steps, 1)))
    for _ in range(max(args.warmup, 0) and 1):
        cpu_arm(args.config, args.sampling_ratio, min(sample, 64))
    vals, info = [], None
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, info = cpu_arm(args.config, args.sampling_ratio, sample)
        vals.append(v)
    dt = time.perf_counter() - t0
    cfg = synth.CONFIGS[args.config]
    value = float(np.mean(vals))
    info["value"] = value
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": cfg["name"], "sampling_ratio": args.sampling_ratio,
                   "note": "CPU arm: bounded sample per step on the host cores"},
        "cpu_baseline": info,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    import chainer_maskrcnn_b200 as pkg
    from chainer_maskrcnn_b200 import _engine, _lib, _sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        # stdout carries exactly one JSON line: keep NCCL's version banner off it
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=device)
    for kv in filter(None, args.tune.split(",")):
        k, v = kv.split("=")
        _lib.set_tuning(**{k: int(v)})

    S = args.sampling_ratio
    cfg, rng, shapes, rois_np, scales = workload(args.config, 0 if args.shard else rank)
    if args.shard and world > 1:
        # image n -> rank n mod world; a rank holds only its own images' pyramids, RoIs and gradients
        if cfg["n_images"] < world:
            raise SystemExit("--shard: %s has %d images, fewer than %d ranks" % (cfg["name"], cfg["n_images"], world))
        rois_np, _ = _sharding.shard_rois(rois_np, cfg["n_images"], world, rank)
        cfg["n_images"] = len(_sharding.images_of_rank(cfg["n_images"], world, rank))
        shapes = synth.pyramid_shapes(cfg["n_images"], cfg["channels"], cfg["height"], cfg["width"],
                                      cfg["n_levels"])
    C, R = cfg["channels"], rois_np.shape[0]
    sizes = cfg["out_sizes"]
    # features live in HBM channels-last (the layout B200 convolutions produce)
    feats = [torch.randn(s, device=device, dtype=torch.float32,
                         generator=torch.Generator(device=device).manual_seed(17 + l + 100 * rank))
             .contiguous(memory_format=torch.channels_last) for l, s in enumerate(shapes)]
    rois = torch.from_numpy(rois_np).to(device)
    gys = [(torch.rand((R, C, P, P), device=device, dtype=torch.float32) * 2 - 1)
           .contiguous(memory_format=torch.channels_last) for P in sizes]
    grads = [torch.empty(s, device=device, dtype=torch.float32,
                         memory_format=torch.channels_last) for s in shapes]

    def step(ev=None):
        if ev:
            ev[0].record()
        outs, plan = _engine.forward(feats, rois, None, scales, sizes, sampling_ratio=S,
                                     roi_format=_lib.ROI_YX)
        if ev:
            ev[1].record()
        _engine.backward(plan, gys, out=grads, deterministic=args.deterministic)
        if ev:
            ev[2].record()
        return outs

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()

    K = args.steps
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = _lib.launch_count()
    e0.record()
    for k in range(K):
        step(evs[k])
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    total_ms = _sharding.max_over_ranks(e0.elapsed_time(e1), device)
    fwd_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in evs]))
    bwd_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in evs]))
    fwd_ms_max = _sharding.max_over_ranks(fwd_ms, device)
    bwd_ms_max = _sharding.max_over_ranks(bwd_ms, device)
    total_rois = _sharding.sum_over_ranks(R, device)
    value = total_rois * K / (total_ms * 1e-3)

    # ---- the same step replayed from a CUDA graph (launch-bound small workloads) ----
    graph_info = None
    if not args.no_graph and not args.deterministic:
        try:
            side = torch.cuda.Stream(device=device)
            side.wait_stream(torch.cuda.current_stream(device))
            with torch.cuda.stream(side):
                step()                                    # allocator warm-up on the capture stream
            torch.cuda.current_stream(device).wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                step()
            for _ in range(3):
                g.replay()
            torch.cuda.synchronize()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            for _ in range(K):
                g.replay()
            g1.record()
            torch.cuda.synchronize()
            gms = _sharding.max_over_ranks(g0.elapsed_time(g1), device) / K
            graph_info = {"ms_per_step": gms, "value": total_rois / (gms * 1e-3), "unit": UNIT,
                          "note": "plan + forward + zero fill + backward captured once with torch.cuda.graph "
                                  "and replayed; RoIs are read from the same device buffer at replay time"}
            del g
        except Exception as e:  # noqa: BLE001
            graph_info = {"ms_per_step": None, "error": repr(e)}

    # ---- end to end through the public host-array API ---------------------
    e2e = None
    if not args.no_e2e:
        old_mask, near_cpus = bind_near_device(local)
        pin = lambda t: t.cpu().contiguous().pin_memory()
        feats_h = [pin(f).numpy() for f in feats]           # NCHW host arrays, as the reference holds them
        rois_h = pin(rois).numpy()
        gys_h = [pin(g).numpy() for g in gys]
        h2d = sum(a.nbytes for a in feats_h) + rois_h.nbytes + sum(a.nbytes for a in gys_h)
        d2h = sum(a.nbytes for a in gys_h) + sum(a.nbytes for a in feats_h)
        for _ in range(3):      # warm-up: the pinned result buffers come from torch's caching allocator
            pooled, g = pkg.fpn_roi_align_host(feats_h, rois_h, None, scales, sizes, S, gys=gys_h)
            del pooled, g
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        check = 0.0
        for _ in range(args.e2e_steps):
            pooled, g = pkg.fpn_roi_align_host(feats_h, rois_h, None, scales, sizes, S, gys=gys_h)
            check += float(pooled[0][0, 0, 0, 0]) + float(g[0][0, 0, 0, 0])   # the result is on the host
            del pooled, g     # consumed: the pinned buffers go back to the caching allocator
        torch.cuda.synchronize()
        dt = _sharding.max_over_ranks(time.perf_counter() - t0, device)
        e2e = {"value": total_rois * args.e2e_steps / dt, "unit": UNIT,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": 1e3 * dt / args.e2e_steps,
               "api": "chainer_maskrcnn_b200.fpn_roi_align_host (pinned NumPy in, NumPy out; uploads, "
                      "kernels and downloads on three streams; NCHW->NHWC conversion on the device "
                      "inside the timed region)"}
        e2e["host_cpus"] = ("%d CPUs of the device's NVML affinity mask" % near_cpus) if near_cpus else "unbound"
        del feats_h, gys_h
        if old_mask is not None:
            os.sched_setaffinity(0, old_mask)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    levels_np = _engine.read_plan(_engine.make_plan(shapes, rois, None, scales, sizes, S))[0]
    ab = algorithmic_bytes(cfg, shapes, rois_np, levels_np, scales, S)
    peak, peak_src = measured_peak()
    dom = "backward" if bwd_ms >= fwd_ms else "forward"
    dom_ms = bwd_ms if dom == "backward" else fwd_ms
    dom_bytes = ab["bwd"] if dom == "backward" else ab["fwd"]
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": ncu_traffic(args.config, S, ["rpool_zero_kernel", "rpool_backward_kernel"]
                               if dom == "backward" else ["rpool_forward_kernel"]),
        "traffic_source": "profiles/ncu_traffic.json (dram__bytes_read.sum + dram__bytes_write.sum, "
                          "ncu --set full, per launch)",
        "peak_source": peak_src,
        "ncu": ncu_view(args.config, S),
        "kernel": ("rpool_zero_kernel + rpool_backward_kernel" if dom == "backward"
                   else "rpool_plan_kernel + rpool_forward_kernel"),
        "algorithmic_bytes_per_launch": int(dom_bytes), "ms_per_launch": dom_ms,
        "forward": {"ms": fwd_ms, "bytes": int(ab["fwd"]),
                    "GBps": ab["fwd"] / (fwd_ms * 1e-3) / 1e9,
                    "frac": ab["fwd"] / (fwd_ms * 1e-3) / 1e9 / peak},
        "backward": {"ms": bwd_ms, "bytes": int(ab["bwd"]),
                     "GBps": ab["bwd"] / (bwd_ms * 1e-3) / 1e9,
                     "frac": ab["bwd"] / (bwd_ms * 1e-3) / 1e9 / peak},
        "fwd_plus_bwd": {"bytes": int(ab["fwd"] + ab["bwd"]),
                         "GBps": (ab["fwd"] + ab["bwd"]) / ((fwd_ms + bwd_ms) * 1e-3) / 1e9,
                         "frac": (ab["fwd"] + ab["bwd"]) / ((fwd_ms + bwd_ms) * 1e-3) / 1e9 / peak},
    }
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            import oracle
            oracle.build()
            _, cpu = cpu_arm(args.config, S, args.cpu_sample_rois, repeats=3)
        except Exception as e:  # noqa: BLE001
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "failed: %r" % (e,)}
    gpu_base = None
    if world == 1 and not args.no_gpu_baseline:
        try:
            gpu_base = gpu_baseline_arm(args.config, args.gpu_baseline_rois, device)
        except Exception as e:  # noqa: BLE001
            gpu_base = {"value": None, "unit": UNIT, "kind": "failed: %r" % (e,)}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K,
        "warmup": max(args.warmup, 3), "ms_per_step": total_ms / K, "higher_is_better": True,
        "scaling": "strong" if args.shard else "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {
            "workload": cfg["name"] + (" sharded by image over %d GPUs" % world if args.shard else " per GPU"),
            "baseline_config_index": args.config,
            "rois_per_gpu": R, "channels": C, "out_sizes": sizes, "sampling_ratio": S,
            "levels": "P2-P%d, assigned on device by the reference rule" % (cfg["n_levels"] + 1),
            "layout": "channels-last features/pooled/gradients resident in HBM",
            "step": ("rpool_plan + rpool_forward + rpool_backward (deterministic: rectangles, scan, private "
                     "windows, ordered gather; its scratch-size query synchronises once per step)"
                     if args.deterministic else
                     "rpool_plan + rpool_forward + rpool_backward (zero-fill included)"),
            "l2": "no flush: one step touches %.0f MB >> 126 MB L2"
                  % ((ab["O"] * 2 + ab["F"] * 2) / 1e6),
            "sharding": "by image, one process per GPU, no data-path collective",
            "tuning": {k: _lib.get_tuning(k) for k in ("prefetch", "threads", "order", "force_path", "split_heads")},
        },
        "fwd_ms": fwd_ms_max, "bwd_ms": bwd_ms_max,
        "roofline": roofline, "cpu_baseline": cpu, "gpu_baseline": gpu_base, "e2e": e2e,
        "cuda_graph": graph_info,
        "gpu_launches": int(launches), "clocks": clocks,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
