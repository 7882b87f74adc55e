#!/usr/bin/env python
"""bench.py -- throughput of the KLT matching hot path on synthetic Sentinel-2
shaped scene pairs (BASELINE.json: matches/s and scene-pairs/s, HBM roofline).

    python bench.py --gpus N --steps K --warmup W          # CUDA arm
    python bench.py --impl reference --steps K --warmup W   # the unmodified reference on the host CPUs

One step = one 10980 x 10980 uint16 scene pair through the whole path (auto mask,
min/max, uint8 + Laplacian k7 of both rasters, Shi-Tomasi corners, pyramidal LK
forward + backward, back-check, (x0,y0) sort, ZNCC of the rows with score >= 0.4),
default KARIOS config.  Both arms run this same workload (same `config`).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

S2 = 10980
WORKLOAD = "s2_b04_10980x10980_pair_klt_zncc_default_config"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--size", type=int, default=S2, help="scene side (default: the S2 10 m band)")
    ap.add_argument("--scenes", type=int, default=2, help="distinct synthetic scenes per rank")
    ap.add_argument("--depth", type=int, default=6,
                    help="scene pairs in flight per GPU (independent contexts + streams)")
    ap.add_argument("--batches", type=int, default=5,
                    help="the K-step timed batch is repeated this many times; the median batch is reported")
    ap.add_argument("--ref-budget", type=float, default=float(os.environ.get("KR_REF_BUDGET_S", "150")),
                    help="--impl reference: stop timing further full-scene steps after this many seconds")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-numa", action="store_true", help="do not bind the rank to its GPU's NUMA node")
    ap.add_argument("--quick", action="store_true",
                    help="tuning runs: resident throughput and stage times only (no e2e, CPU or next-row legs)")
    a = ap.parse_args()
    if a.quick:
        a.no_e2e = a.no_cpu_baseline = True
    return a


def default_conf(cls, **kw):
    """karios/configuration/processing_configuration.json:8-19 (CLI default)."""
    base = dict(minDistance=10, blocksize=15, maxCorners=20000, matching_winsize=25,
                qualityLevel=0.1, xStart=0, tile_size=20000, laplacian_kernel_size=7,
                outliers_filtering=False, laplacian_invert_polarity=False)
    base.update(kw)
    return cls(**base)


def bench_config(size, world):
    """The workload both arms run (the driver compares the two `config` objects)."""
    return {"workload": WORKLOAD if size == S2 else f"synthetic_{size}x{size}_pair_klt_zncc",
            "scene": f"{size}x{size} uint16 pair, synthetic texture (karios_b200/synth.py, seed 1234 + scene index), "
                     "monitored shifted by (+0.30,-0.20) px",
            "klt": "maxCorners 20000, minDistance 10, blocksize 15, winsize 25, q 0.1, k7, tile_size 20000 (1 tile)",
            "scoring": "ZNCC of the rows with KLT score >= 0.4",
            "parallelism": f"scene pairs sharded over {world} GPU(s), no data-path collective",
            "l2": "inputs (482 MB per pair) exceed the 126 MB L2; no flush needed"}


# ------------------------------------------------------------------ CPU legs
def scene_host(size, seed):
    """Scene `seed` as host uint16 arrays (generated on the GPU when there is one: the
    torch generator is bit-identical on both, and minutes faster there)."""
    import torch
    from karios_b200 import synth
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    ref, mon = synth.make_pair(size, size, seed=seed, device=dev)
    to_np = lambda t: t.cpu().view(torch.int16).numpy().view(np.uint16)  # noqa: E731
    return to_np(ref), to_np(mon)


def reference_pass(ref, mon):
    """One pass of the UNMODIFIED reference over (ref, mon): KLT.match + compute_zncc from
    oracle/_ref (placed by build(); see oracle/vendor_ref.py, oracle/ref_run.py)."""
    import logging
    from oracle import ref_run, refimport
    refimport.use_vendored()
    logging.getLogger("karios").setLevel(logging.ERROR)      # one warning per border row otherwise
    df, secs, how = ref_run.run_pair(mon, ref, None, threshold=0.4)
    return len(df), secs, how


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    size = args.size
    ref, mon = scene_host(size, 1234)
    how = ""
    t_all = time.perf_counter()
    # the CPU path has no compilation or caches to warm beyond the first pass: one warm-up
    # pass of the full scene, whatever W is (stated in the line)
    warm = min(1, args.warmup)
    for _ in range(warm):
        _, _, how = reference_pass(ref, mon)
    matches, secs, done = 0, 0.0, 0
    per_step = []
    for _ in range(args.steps):
        m, s, how = reference_pass(ref, mon)
        matches += m
        secs += s
        per_step.append(round(s, 3))
        done += 1
        if time.perf_counter() - t_all > args.ref_budget and done >= 3:
            break
    value = matches / secs if secs > 0 else 0.0
    cores = os.cpu_count()
    sample = (f"{done} full {size}x{size} scene pairs timed (of {args.steps} requested: full-scene passes stop "
              f"after {args.ref_budget:.0f} s), {warm} warm-up pass; {how}; cv2.setNumThreads({cores})")
    line = {
        "impl": "reference", "metric": "matches_per_sec", "value": value, "unit": "matches/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "steps_timed": done, "warmup_run": warm,
        "ms_per_step": 1e3 * secs / max(1, done), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": bench_config(size, world),
        "scene_pairs_per_sec": done / secs if secs > 0 else 0.0,
        "matches_per_scene": matches / max(1, done),
        "seconds_per_step": per_step,
        "cpu_baseline": {"value": value, "unit": "matches/s", "cores": cores, "kind": "reference",
                         "sample": sample},
        "e2e": {"value": value, "unit": "matches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------- clock log
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons = index, threading.Event(), [], set()
        self.active = threading.Event()
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self.stop_flag.is_set():
            if self.active.is_set():            # only while a timed batch is running
                try:
                    self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    for k, bit in names.items():
                        if r & bit:
                            self.reasons.add(k)
                except Exception:  # noqa: BLE001
                    pass
            time.sleep(0.003)

    def summary(self):
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ----------------------------------------------------------------- CUDA arm
def stage_bytes(P, C, n_corners, n_z):
    """Algorithmic (compulsory) bytes of each stage of one scene pair: inputs read
    once, outputs written once (adapted to the launch split used
    here -- the auto mask is written by the min/max pass)."""
    win_bytes = 28 * 28 + 26 * 26
    return {
        "minmax_mask": 2 * 2 * P + P,
        "laplacian_mon": 2 * P + P,
        "laplacian_ref": 2 * P + P,
        "corner_response": P + P + 8 * C,
        "select": 8 * C + 8 * C,
        "nms": 0, "corner_sort": 0,
        "pyramids": 2 * P + 2 * (P // 4),
        "lk_roundtrip": n_corners * 2 * 2 * win_bytes,
        "rows": n_corners * 20,
        "zncc": n_z * 2 * 43 * 43 * 2,
        "mutual_info": 0,          # not part of the headline path (timed under next_rows)
    }


def _max_over_ranks(dist, dev, world, *vals):
    import torch
    if world == 1:
        return vals
    t = torch.tensor(list(vals), device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return tuple(float(v) for v in t)


def _sum_over_ranks(dist, dev, world, *vals):
    import torch
    if world == 1:
        return vals
    t = torch.tensor(list(vals), device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return tuple(float(v) for v in t)


def cuda_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    from karios_b200 import sharding
    numa = {"bound": False, "skipped": True}
    if not args.no_numa:
        numa = sharding.bind_to_gpu_numa(local)       # before any pinned allocation

    import torch
    import torch.distributed as dist
    from karios_b200 import _native as N
    from karios_b200 import synth
    from karios_b200.api import SceneMatcher, ScenePipeline
    from karios_b200.core.configuration import KLTConfiguration

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    size = args.size
    conf = default_conf(KLTConfiguration)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    lib = N.load_library()

    # scenes resident in HBM: scene i of this rank has seed 1234 + rank + i * world
    scenes = []
    for i in range(args.scenes):
        ref, mon = synth.make_pair(size, size, seed=1234 + rank + i * world, device=dev)
        scenes.append((mon, ref))
    torch.cuda.synchronize()
    sm = SceneMatcher(size, size, conf, 0.4, device=dev, depth=args.depth)
    sm.trace_units = bool(os.environ.get("KR_TRACE_UNITS"))
    cap = sm.rows.capacity
    side = torch.cuda.Stream(device=dev)

    def batch(n_pairs):
        """n_pairs scene pairs of this rank (+ the exchange when sharded) -> (rows of this rank, gathered)"""
        seq = [scenes[i % len(scenes)] for i in range(n_pairs)]
        if world > 1:
            g, tot = sharding.match_many_exchange(sm, seq, None, side)
            m = sharding.batch_moments(g)          # statistics of the whole batch, on the device
            return tot, (g, m)
        _, tot = sm.match_many(seq)
        return tot, None

    sampler = ClockSampler(local)          # NVML is initialised before the timed region
    sampler.start()
    # warm-up: W steps, then one untimed batch of the timed size (the caching allocator then holds
    # the arena / gather blocks of that size: no cudaMalloc inside a timed batch; NCCL channels up)
    batch(max(args.warmup, 1))
    batch(args.steps)
    torch.cuda.synchronize()

    batches = []
    launches = None
    diag = None
    for b in range(max(1, args.batches)):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = int(lib.kr_launch_count())
        sampler.active.set()
        e0.record()
        # K steps = K scene pairs, `depth` of them in flight (each on its own context and stream);
        # sharded runs end with the one exchange step of the path (all_gather of the fixed-size
        # unit records, issued on a side stream when the last unit is enqueued)
        tot, gathered = batch(args.steps)
        e1.record()
        torch.cuda.synchronize()
        sampler.active.clear()
        launches = int(lib.kr_launch_count()) - l0
        ms_b = e0.elapsed_time(e1)
        if world > 1:
            dist.barrier()
        (ms_b,) = _max_over_ranks(dist, dev, world, ms_b)
        (tot_all,) = _sum_over_ranks(dist, dev, world, float(tot))
        done_t = list(getattr(sm, "unit_done_t", []))
        gaps = [y - x for x, y in zip(done_t, done_t[1:])]
        batches.append({"ms": ms_b, "matches": int(tot_all), "redo_units": int(getattr(sm, "n_redo", 0)),
                        "max_unit_gap_ms": round(1e3 * max(gaps), 3) if gaps else None})
        if world > 1 and b == 0:
            assert gathered[0].shape[0] == world and gathered[0].shape[1] == args.steps
            diag = sharding.moments_dict(gathered[1])
    sampler.stop_flag.set()
    order = sorted(range(len(batches)), key=lambda i: batches[i]["ms"])
    med = batches[order[len(order) // 2]]
    ms, matches = med["ms"], med["matches"]
    secs = ms / 1e3
    value = matches / secs
    pairs_per_sec = world * args.steps / secs

    # exchange alone (sharded runs): the collective on an idle device, for the record
    exchange_ms = 0.0
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g = sharding.exchange(sm.last_arena)
        sharding.batch_moments(g)
        b_.record()
        torch.cuda.synchronize()
        (exchange_ms,) = _max_over_ranks(dist, dev, world, a.elapsed_time(b_))

    # ---- per-stage device times (CUDA events on the launching stream) --------
    sm.ctx.set_profiling(True)
    acc = {}
    reps = 3
    for i in range(reps):
        mon, ref = scenes[i % len(scenes)]
        st = sm.ctx.match_tile(mon, ref, None, sm.windows[0], sm.kconf, sm.rows)
        torch.cuda.synchronize()
        for k, v in sm.ctx.stage_ms().items():
            acc[k] = acc.get(k, 0.0) + v / reps
    sm.ctx.set_profiling(False)
    # one pair alone, no other pair in flight (latency of the path)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(5):
        sm.ctx.match_tile(*scenes[i % len(scenes)], None, sm.windows[0], sm.kconf, sm.rows)
    torch.cuda.synchronize()
    single_pair_ms = 1e3 * (time.perf_counter() - t0) / 5
    P = size * size
    n_z = int((sm.rows.f32[4, : st.n_kept] >= 0.4).sum().item())
    sb = stage_bytes(P, st.n_candidates, st.n_corners, n_z)
    stages = {k: {"ms": round(acc[k], 4), "alg_bytes": sb[k],
                  "gbs": round(sb[k] / (acc[k] * 1e6), 1) if acc[k] > 0 else None,
                  "frac_of_hbm_peak": round(sb[k] / (acc[k] * 1e6) / hbm_peak, 4) if acc[k] > 0 and sb[k] else None}
              for k in acc}
    dom = max(acc, key=lambda k: acc[k])
    achieved = sb[dom] / (acc[dom] * 1e6) if acc[dom] > 0 else 0.0
    traffic, traffic_src = None, None
    try:        # dram__bytes_read + write of the stage's kernels, one ncu --set full capture (profiles/)
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        if size == S2 and dom in tj["stages"]:
            traffic, traffic_src = tj["stages"][dom]["dram_bytes"], tj["source"]
    except Exception:  # noqa: BLE001
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": round(achieved, 1), "peak": hbm_peak,
                "unit": "GB/s", "frac": round(achieved / hbm_peak, 4), "traffic": traffic,
                "traffic_source": traffic_src,
                "peak_source": peak_src, "kernel_ms": round(acc[dom], 4),
                "alg_bytes_per_launch": sb[dom]}
    total_alg = sum(sb.values())
    total_ms = sum(v for v in acc.values() if v > 0)

    # ---- SURVEY 8(f) rows built so far, timed on the rows of the last pair -----
    next_rows = {}
    try:
        if args.quick:
            raise RuntimeError("skipped (--quick)")
        mon, ref = scenes[0]
        cols = [sm.rows.f32[i, : st.n_kept].contiguous() for i in range(4)]
        for _ in range(2):
            sm.ctx.mutual_info(ref, mon, *cols)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        reps_mi = 5
        for _ in range(reps_mi):
            mi_out = sm.ctx.mutual_info(ref, mon, *cols)
        b.record()
        torch.cuda.synchronize()
        mi_ms = a.elapsed_time(b) / reps_mi
        n_mi = int((~torch.isnan(mi_out[0])).sum().item())
        mi_bytes = n_mi * 2 * 57 * 57 * 2
        next_rows["mutual_info"] = {
            "what": "MutualInfoService.compute_mutual_info + ZNCCService.compute_mi (one launch, both scores)",
            "rows": int(st.n_kept), "rows_scored": n_mi, "ms": round(mi_ms, 4),
            "rows_per_sec": round(int(st.n_kept) / (mi_ms / 1e3), 1),
            "alg_bytes": mi_bytes, "gbs": round(mi_bytes / (mi_ms * 1e6), 1),
            "bound": "shared-memory atomics + L1/L2 gathers (chips are re-read from cache)"}
    except Exception as e:  # noqa: BLE001
        next_rows["mutual_info"] = {"error": repr(e)}
    try:
        if args.quick:
            raise RuntimeError("skipped (--quick)")
        from karios_b200 import api as kapi
        from karios_b200.core.image import DeviceRaster
        from karios_b200.matcher.large_offset import phase_cross_correlation_shift
        mon, ref = scenes[0]

        def timed(fn, reps_=3):
            fn()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            for _ in range(reps_):
                out = fn()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / reps_, out

        ms_lo, off = timed(lambda: phase_cross_correlation_shift(mon, ref), 2)
        # algorithmic traffic of the three float64 transforms is library-internal; the two
        # own kernels stream the half spectrum twice and the correlation once
        spec_bytes = size * (size // 2 + 1) * 16
        next_rows["large_offset"] = {
            "what": "LargeOffsetMatcher.match: 2 rfft2 + cross-power + irfft2 + argmax, float64 (cuFFT + 2 own kernels)",
            "ms": round(ms_lo, 3), "offset_yx": [float(off[0]), float(off[1])],
            "own_kernel_alg_bytes": 3 * spec_bytes + P * 8, "bound": "cuFFT (library) dominates"}
        ms_sh, _ = timed(lambda: N.shift_image(mon, 52, -37))
        next_rows["shift_image"] = {"ms": round(ms_sh, 4), "alg_bytes": 2 * P * 2,
                                    "gbs": round(2 * P * 2 / (ms_sh * 1e6), 1), "bound": "hbm"}
        ms_pc, pct = timed(lambda: kapi.percentiles_2_98(DeviceRaster(mon)))
        next_rows["percentiles_2_98"] = {"ms": round(ms_pc, 3), "value": [float(pct[0]), float(pct[1])],
                                         "alg_bytes": 2 * P * 2, "gbs": round(2 * P * 2 / (ms_pc * 1e6), 1),
                                         "bound": "shared-memory atomics (coarse pass) + HBM (refinement pass)"}
        ms_ss, ss = timed(lambda: kapi.scene_scan(DeviceRaster(mon)))
        next_rows["scene_scan"] = {"what": "percentiles + valid-pixel count of one raster: one fused pass "
                                           "(kr_histogram_count) + one refinement pass",
                                   "ms": round(ms_ss, 3), "value": [float(ss[0][0]), float(ss[0][1]), int(ss[1])],
                                   "alg_bytes": 2 * P * 2, "gbs": round(2 * P * 2 / (ms_ss * 1e6), 1)}
        ms_cv, nvalid = timed(lambda: kapi.count_valid_pixels(DeviceRaster(mon)))
        next_rows["count_valid"] = {"ms": round(ms_cv, 4), "value": int(nvalid), "alg_bytes": P * 2,
                                    "gbs": round(P * 2 / (ms_cv * 1e6), 1), "bound": "hbm"}
        # config 4 (--enable-large-shift-detection) on the full scene: monitored shifted by whole
        # pixels (+37 columns, -52 rows), detection, shift_image, KLT on the shifted raster
        mon_far = N.shift_image(mon, 52, -37)

        def cfg4():
            return kapi.match_pair_large_shift(mon_far, ref, None, conf, offset_threshold=10)
        cfg4()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        df4, applied = cfg4()
        torch.cuda.synchronize()
        next_rows["config4_large_shift"] = {
            "what": "match_images with enable_large_shift_detection on the full scene, monitored displaced by "
                    "(+37, -52) px: phase correlation, shift_image, KLT on the shifted raster, offsets added back",
            "ms": round(1e3 * (time.perf_counter() - t0), 2), "applied_offset_xy": [float(v) for v in applied]
            if applied else None, "rows": int(len(df4)), "mean_dx": float(df4.dx.mean()),
            "mean_dy": float(df4.dy.mean()),
            "fft_peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)}
        del mon_far, df4
    except Exception as e:  # noqa: BLE001
        next_rows["scene_passes"] = {"error": repr(e)}

    # ---- automatic kernel-size search on the full scene -----------
    try:
        if args.quick:
            raise RuntimeError("skipped (--quick)")
        from karios_b200.core.image import DeviceRaster
        from karios_b200.matcher.klt import KLT
        mon, ref = scenes[0]
        aconf = default_conf(KLTConfiguration, laplacian_kernel_size="auto")
        res = {}
        for label, batched in (("device_search", True), ("host_loop", False)):
            k = KLT(aconf)
            k._batched_auto = batched
            list(k.match(DeviceRaster(mon), DeviceRaster(ref), None))          # warm-up (context, scratch)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            frames = list(k.match(DeviceRaster(mon), DeviceRaster(ref), None))
            torch.cuda.synchronize()
            res[label] = {"ms": round(1e3 * (time.perf_counter() - t0), 2), "rows": int(sum(len(f) for f in frames)),
                          "selected": list(k.auto_selected_ksize or ())}
        next_rows["auto_ksize"] = {
            "what": "KLT.match with laplacian_kernel_size='auto': 5+5 Laplacians, 5 corner sets, 25 LK round trips "
                    "(kr_auto_ksize: one launch sequence, one sync) vs one kr_klt_track per pair", **res}
    except Exception as e:  # noqa: BLE001
        next_rows["auto_ksize"] = {"error": repr(e)}

    # ---- end to end: host rasters -> rows on the host ------------------------------
    e2e, e2e_dropin, h2d = None, None, None
    if not args.no_e2e:
        n_e2e = max(args.steps, 1)
        host_pairs = [(m.cpu().pin_memory(), r.cpu().pin_memory()) for m, r in scenes]
        seq = [host_pairs[i % len(host_pairs)] for i in range(n_e2e)]
        bytes_pair = 2 * P * 2

        def wall(fn):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            out = fn()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            (dt,) = _max_over_ranks(dist, dev, world, dt)
            return dt, out

        # (0) the copy ceiling: the same pinned rasters, the same double-buffered H2D copies,
        #     no kernel at all -- what the host side of this box can deliver to N GPUs at once
        pipe = ScenePipeline(size, size, torch.uint16, conf, 0.4, device=dev)
        pipe.copy_only(seq[:2])
        dt_copy, _ = wall(lambda: pipe.copy_only(seq))
        h2d = {"what": "copy-only run of the e2e leg: same pinned rasters and buffers, no kernels",
               "scene_pairs_per_sec": world * len(seq) / dt_copy, "gbs_per_gpu": bytes_pair * len(seq) / dt_copy / 1e9,
               "gbs_aggregate": world * bytes_pair * len(seq) / dt_copy / 1e9, "numa": numa}
        # (1) ScenePipeline: pinned rasters, upload of pair i+1 overlapped with the matching of pair i
        pipe.run(seq[: max(1, min(args.warmup, 2))])
        dt, m2 = wall(lambda: pipe.run(seq))
        (m2,) = _sum_over_ranks(dist, dev, world, float(m2))
        e2e = {"value": m2 / dt, "unit": "matches/s", "h2d_bytes_per_step": bytes_pair,
               "d2h_bytes_per_step": int(st.n_kept) * (5 * 4 + 8),
               "scene_pairs_per_sec": world * len(seq) / dt, "ms_per_step": 1e3 * dt / len(seq),
               "api": "karios_b200.api.ScenePipeline.run (pinned host rasters -> rows in pinned host memory)",
               "frac_of_copy_ceiling": round(dt_copy / dt, 4)}
        pipe.close()
        del pipe

        # (2) the reference's plugin API, as KariosAPI._compute_matches/_handle_klt_results call it
        #     (api/core.py:845-891): NumPy rasters -> KLT.match -> DataFrame -> compute_zncc of the
        #     score >= 0.4 rows.  Fresh raster objects every step, so every step uploads its pair.
        import pandas as pd
        from karios_b200.core import image as kimg
        from karios_b200.matcher.klt import KLT
        from karios_b200.matcher.zncc_service import ZNCCService

        def as_np(t):
            return t.view(torch.int16).numpy().view(np.uint16)

        zs = ZNCCService()

        def dropin_step(mon_np, ref_np):
            mon_img, ref_img = kimg.ArrayRaster(mon_np), kimg.ArrayRaster(ref_np)
            all_frame = pd.DataFrame()
            for dataframe in KLT(conf).match(mon_img, ref_img, None):
                cand = dataframe[dataframe["score"] >= 0.4]
                dataframe["zncc_score"] = np.nan
                z = zs.compute_zncc(cand, mon_img, ref_img)
                dataframe.loc[cand.index, "zncc_score"] = z
                all_frame = pd.concat([all_frame, dataframe])
            return len(all_frame)

        def dropin_run(pairs_np):
            return sum(dropin_step(m, r) for m, r in pairs_np)

        def dropin_step_imgs(mon_img, ref_img):
            all_frame = pd.DataFrame()
            for dataframe in KLT(conf).match(mon_img, ref_img, None):
                cand = dataframe[dataframe["score"] >= 0.4]
                dataframe["zncc_score"] = np.nan
                z = zs.compute_zncc(cand, mon_img, ref_img)
                dataframe.loc[cand.index, "zncc_score"] = z
                all_frame = pd.concat([all_frame, dataframe])
            return len(all_frame)

        def dropin_run_prefetch(pairs_np):
            """The same calls plus core.image.prefetch of the NEXT pair's rasters before the
            current pair is matched (one added line in the caller's loop)."""
            imgs = [(kimg.ArrayRaster(m), kimg.ArrayRaster(r)) for m, r in pairs_np]
            total = 0
            for im in imgs[0]:
                kimg.prefetch(im)
            for i, (mon_img, ref_img) in enumerate(imgs):
                if i + 1 < len(imgs):
                    for im in imgs[i + 1]:
                        kimg.prefetch(im)
                total += dropin_step_imgs(mon_img, ref_img)
                kimg.release_device(mon_img)
                kimg.release_device(ref_img)
            return total

        pinned_np = [(as_np(m), as_np(r)) for m, r in host_pairs]
        seq_np = [pinned_np[i % len(pinned_np)] for i in range(n_e2e)]
        dropin_run(seq_np[:2])
        up0 = dict(kimg.uploads)
        dt_d, m3 = wall(lambda: dropin_run(seq_np))
        (m3,) = _sum_over_ranks(dist, dev, world, float(m3))
        uploads_per_step = (kimg.uploads["count"] - up0["count"]) / len(seq_np)
        e2e_dropin = {"value": m3 / dt_d, "unit": "matches/s", "scene_pairs_per_sec": world * len(seq_np) / dt_d,
                      "ms_per_step": 1e3 * dt_d / len(seq_np), "h2d_bytes_per_step": bytes_pair,
                      "d2h_bytes_per_step": int(st.n_kept) * (5 * 4 + 8),
                      "raster_uploads_per_step": uploads_per_step,
                      "api": "karios_b200.matcher.klt.KLT.match + ZNCCService.compute_zncc on NumPy rasters "
                             "(pinned-memory backed), pandas DataFrames out: the calls of karios/api/core.py:845-891",
                      "vs_scene_pipeline": round(dt_d / dt, 3)}
        dropin_run_prefetch(seq_np[:3])
        dt_pf, m4 = wall(lambda: dropin_run_prefetch(seq_np))
        e2e_dropin["with_prefetch"] = {
            "what": "the same calls + karios_b200.core.image.prefetch() of the next pair's rasters",
            "ms_per_step": round(1e3 * dt_pf / len(seq_np), 3),
            "scene_pairs_per_sec": round(world * len(seq_np) / dt_pf, 2), "vs_scene_pipeline": round(dt_pf / dt, 3)}
        if rank == 0 and world == 1:
            # the same through ordinary pageable NumPy arrays (what GDAL hands out today)
            pag = [(np.array(m, copy=True), np.array(r, copy=True)) for m, r in pinned_np[:1]]
            dropin_run(pag)
            n_p = min(5, n_e2e)
            dt_p, _ = wall(lambda: dropin_run(pag * n_p))
            e2e_dropin["pageable_ms_per_step"] = round(1e3 * dt_p / n_p, 3)
        del host_pairs, pinned_np, seq_np

    # ---- CPU baseline beside it (rank 0, N = 1 only): the unmodified reference ------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        mon_h, ref_h = (t.cpu().view(torch.int16).numpy().view(np.uint16) for t in scenes[0])
        m, s, how = reference_pass(ref_h, mon_h)
        cpu = {"value": m / s, "unit": "matches/s", "cores": os.cpu_count(), "kind": "reference",
               "sample": f"1 full {size}x{size} scene pair (scene 0 of the CUDA arm), default config, one pass, "
                         f"{how}; cv2.setNumThreads({os.cpu_count()})",
               "seconds": round(s, 3), "matches": m}
        from oracle import oracle as O
        if "rows" in next_rows.get("mutual_info", {}):
            k = min(1500, int(st.n_kept))
            c4 = [c[:k].cpu().numpy() for c in cols]
            t0 = time.perf_counter()
            O.mutual_info(*c4, mon_h, ref_h)
            dt_mi = time.perf_counter() - t0
            next_rows["mutual_info"]["cpu_rows_per_sec"] = round(k / dt_mi, 1)
            next_rows["mutual_info"]["cpu_sample"] = f"{k} rows, oracle.mutual_info (np.histogram2d per row, 1 core)"
        if "ms" in next_rows.get("large_offset", {}):
            side_ = min(size, 2048)
            t0 = time.perf_counter()
            O.phase_cross_correlation_shift(mon_h[:side_, :side_], ref_h[:side_, :side_])
            dt_lo = time.perf_counter() - t0
            next_rows["large_offset"]["cpu_seconds_sample"] = round(dt_lo, 3)
            next_rows["large_offset"]["cpu_sample"] = (f"{side_}x{side_} crop, oracle (numpy.fft float64, 1 core); "
                                                       f"the full frame is {(size / side_) ** 2:.0f}x the pixels")
            t0 = time.perf_counter()
            O.percentiles_2_98(mon_h)
            next_rows["percentiles_2_98"]["cpu_seconds"] = round(time.perf_counter() - t0, 3)

    if rank == 0:
        line = {
            "metric": "matches_per_sec", "value": value, "unit": "matches/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / max(1, args.steps),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": bench_config(size, world),
            "run": {"scenes_per_rank": len(scenes), "pairs_in_flight": args.depth,
                    "timed_batches": [round(b["ms"], 4) for b in batches], "reported": "median batch",
                    "numa": numa},
            "scene_pairs_per_sec": pairs_per_sec,
            "matches_per_scene": matches / max(1, world * args.steps),
            "single_pair_latency_ms": round(single_pair_ms, 4),
            "roofline": roofline,
            "path_hbm": {"alg_bytes_per_pair": total_alg, "kernel_ms_per_pair": round(total_ms, 4),
                         "gbs": round(total_alg / (total_ms * 1e6), 1) if total_ms > 0 else None,
                         "frac": round(total_alg / (total_ms * 1e6) / hbm_peak, 4) if total_ms > 0 else None},
            "stages": stages,
            "cpu_baseline": cpu,
            "e2e": e2e,
            "e2e_dropin": e2e_dropin,
            "h2d_ceiling": h2d,
            "next_rows": next_rows,
            "pipeline": {"redo_units": max(b["redo_units"] for b in batches),
                         "max_unit_gap_ms": max((b["max_unit_gap_ms"] or 0) for b in batches)},
            "exchange_ms": round(exchange_ms, 3), "exchange_moments": diag,
            "gpu_launches": launches,      # kr_launch_count() over one timed batch (this rank)
            "clocks": sampler.summary(),
            "stats_last": {k: v for k, v in st.as_dict().items() if k.startswith("n_") or k == "nms_rounds"},
            "corner_cut": {"est_cut_bits": int(st.est_cut_bits), "rows_skipped": int(st.rows_skipped),
                           "row_pieces": ((size + 103) // 104) * size},
        }
        print(json.dumps(line), flush=True)
    sm.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        reference_arm(args)
    else:
        cuda_arm(args)


if __name__ == "__main__":
    main()
