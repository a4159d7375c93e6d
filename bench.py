#!/usr/bin/env python
"""bench.py -- nuclei/s of the per-nucleus feature pipeline on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (libnfx.so)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (oracle port)

A "step" is one pass of the hot path over one batch of synthetic input: config[1] of BASELINE.json
(colour/intensity feature set, 100k nuclei, 64x64 windows, one B200) unless --workload says
otherwise. Under torchrun every rank owns one GPU, its own synthetic tile and its own contiguous
range of nuclei (weak scaling, no data-path collective); torch.distributed is used only for the
barrier and the max-over-ranks of the device time.

value : nuclei/s with tile and polygons already resident in HBM (CUDA events on the context stream)
e2e   : nuclei/s through the public host API with HOST buffers: pinned H2D of the tile and the
        polygons, kernels, D2H of the feature matrix, every step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "nuclei-feature-extraction_b200"))

WORKLOADS = {
    # name: (feature sets, nuclei, tile side, patch, polygon kwargs)
    "color": (["color"], 100_000, 16384, 64, {}),
    "shape": (["geometry"], 10_000, 4096, 64, {}),
    "glcm": (["glcm"], 1_000_000, 49152, 64, {}),
    "all": (["all"], 100_000, 16384, 64, {}),
    # the two remaining sets of `texture` / `all` on their own (BASELINE metric: nuclei/s PER feature set)
    "glrlm": (["glrlm"], 100_000, 16384, 64, {}),
    "gabor": (["gabor"], 100_000, 16384, 64, {}),
    # BASELINE config 5: large irregular nuclei, 256x256 windows, 500-vertex polygons
    "stress": (["all"], 20_000, 16384, 256,
               dict(r0_range=(40.0, 110.0), v_range=(500, 500), harmonics=(3, 7, 19))),
    # BASELINE config 4: one slide resident in HBM, written tile by tile; nuclei split over the ranks (strong scaling)
    "slide": (["all"], 5_000_000, 100_000, 64, {}),
}
V_MEAN = 30.0   # mean ring length of the synthetic polygons (12..48 vertices + closing duplicate)


def algorithmic_bytes(workload: str, P: int, F: int) -> float:
    """SURVEY.md 8d, fused per-feature-set pipeline: 3*P^2 (u8 window, read once) + polygon ring
    + 4*F output bytes per nucleus."""
    b_px = 3 * P * P
    b_poly = 8 * (V_MEAN + 1) + 8
    if workload == "shape":
        return b_poly + 4 * F
    return b_px + b_poly + 4 * F


def kernel_bytes(name: str, P: int, slabs: int) -> float:
    """Per-nucleus compulsory traffic of one kernel."""
    px, bm, info = 3 * P * P, P * P / 8, 16
    if name in ("k_color", "k_color_warp"):
        return px + bm + info + 4 * 17
    if name == "k_hue_batch":
        return px + bm + info + 8 * slabs
    if name == "k_glcm":
        return px + bm + info + 4 * 224
    if name.startswith("k_geom"):
        return 8 * (V_MEAN + 1) + bm + info + 8 + (4 * 12 if "shape" in name else 0)
    if name == "k_hue_finalize":
        return 8 * slabs + 8
    if name == "k_glrlm":
        return px + bm + info + 4 * 68
    if name == "k_gabor":
        return px + bm + info + 4 * 96
    return 0.0


class ClockSampler:
    """`nvidia-smi -lms 20` running from the first warm-up step to the end of the timed region (the
    warm-up is the same kernel sequence, so every sample is taken under the measured load)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        rows = []
        if self.proc is not None:
            try:
                self.proc.terminate()
                out, _ = self.proc.communicate(timeout=5)
                rows = [[x.strip() for x in ln.split(",")] for ln in out.strip().splitlines()]
            except Exception:
                pass
        rows = [r for r in rows if len(r) >= 6]
        sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in rows:
            for nm, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_inputs(workload, nuclei, side, P, seed, pinned):
    from nfx import pinned_empty, synth
    kw = WORKLOADS[workload][4]
    block = min(side, 4096)
    base = synth.synth_tile(block, block, seed)
    tile = pinned_empty((side, side, 3), np.uint8) if pinned else np.empty((side, side, 3), np.uint8)
    for r in range(0, side, block):
        for c in range(0, side, block):
            tile[r:r + block, c:c + block] = base[:min(block, side - r), :min(block, side - c)]
    xy, off = synth.synth_polygons(nuclei, side, side, seed, patch=P, **kw)
    if pinned:
        pxy = pinned_empty(xy.shape, np.float32)
        pxy[...] = xy
        poff = pinned_empty(off.shape, np.int64)
        poff[...] = off
        xy, off = pxy, poff
    return tile, xy, off


def run_staged(args, local_rank):
    """North-star kernels (1) gather and (2) raster on their own: batched window gather tile -> u8 patch
    array (k_gather) and polygon -> 1-bit mask (k_geom<raster>), timed with CUDA events per launch."""
    import nfx
    sets, nuclei, side, P, kw = WORKLOADS["color"]
    nuclei = args.nuclei or nuclei
    side = args.tile or side
    tile, xy, off = make_inputs("color", nuclei, side, P, 2, pinned=True)
    ex = nfx.Extractor(local_rank, P, args.batch_size)
    ex.upload_tile(tile)
    ex.upload_polygons(xy, off)
    for _ in range(3):
        ex.rasterize_device()
        ex.gather_patches(want=False)
    ex.profile(True)
    ex.profile_reset()
    K = max(args.steps, 1)
    for _ in range(K):
        ex.rasterize_device()
        ex.gather_patches(want=False)
    prof = ex.profile_get()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    bytes_per = {"k_gather": 2 * 3 * P * P + 16, "k_geom<raster>": 8 * (V_MEAN + 1) + 8 + P * P / 8 + 16 + 8}
    kern = {}
    for k, (n_l, tot) in prof.items():
        avg = tot / max(n_l, 1)
        gbs = bytes_per.get(k, 0) * nuclei / (avg * 1e-3) / 1e9 if avg > 0 else 0.0
        kern[k] = {"launches": n_l, "avg_ms": avg, "gbs": gbs, "frac_of_hbm_peak": gbs / hbm_peak,
                   "bytes_per_nucleus": bytes_per.get(k, 0)}
    g = kern.get("k_gather", {"avg_ms": float("nan"), "gbs": 0.0})
    print(json.dumps({
        "metric": "nuclei/sec", "value": nuclei / (g["avg_ms"] * 1e-3), "unit": "nuclei/s", "n_gpus": 1, "steps": K, "warmup": 3,
        "ms_per_step": g["avg_ms"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic",
        "config": {"workload": f"staged: k_gather (tile -> u8 patch array) and k_geom<raster> alone, {nuclei} nuclei, {P}x{P} windows, "
                               f"tile {side}x{side}"},
        "roofline": {"bound": "hbm", "kernel": "k_gather", "achieved": g["gbs"], "peak": hbm_peak, "unit": "GB/s",
                     "frac": g["gbs"] / hbm_peak, "traffic": None},
        "kernels": kern, "gpu_launches": int(sum(v["launches"] for v in kern.values())),
    }))
    ex.close()


def cpu_reference_rate(sets, tile, xy, off, P, batch, sample, workers, budget_s=25.0):
    """The reference's CPU path (oracle = op-for-op torch-CPU restatement): patch_loader +
    compute_features_batched per chunk, chunks in parallel over `workers` host threads like the
    reference's rayon par_chunks (src/main.rs:146-158), on the first `sample` nuclei. Chunks are issued
    in waves of `workers`; no new wave starts after `budget_s` seconds (the rate is nuclei done / time)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from concurrent.futures import ThreadPoolExecutor
    import torch
    import nfx_oracle as o
    from nfx import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(max(1, cores // workers))
    rings = synth.rings_of(np.asarray(xy), np.asarray(off)[:sample + 1])
    tile = np.asarray(tile)
    step = batch if P <= 64 else max(1, batch // 25)      # 256x256 windows: 16x the pixels per nucleus
    chunks = [rings[k:k + step] for k in range(0, len(rings), step)]
    done = 0
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=workers) as pool:
        for w0 in range(0, len(chunks), workers):
            wave = chunks[w0:w0 + workers]
            list(pool.map(lambda ch: o.extract(ch, tile, sets, P, batch), wave))
            done += sum(len(c) for c in wave)
            if time.perf_counter() - t0 > budget_s:
                break
    dt = time.perf_counter() - t0
    return done / dt, dt, done


def workload_config(args, sets, nuclei, side, P):
    """The `config` object: identical for this repo's arm and for the reference arm."""
    return {"workload": f"{args.workload}: {'+'.join(sets)} feature set(s), {nuclei} nuclei per GPU, "
                        f"{P}x{P} windows, tile {side}x{side} u8 RGB per GPU, batch_size={args.batch_size}",
            "l2": f"inputs > L2: tile {3 * side * side / 1e6:.0f} MB per GPU (L2 126 MB)",
            "partition": "contiguous index ranges per GPU, no collective"}


def cpu_workers():
    return max(1, min(os.cpu_count() or 1, 16))   # bounded: each colour chunk holds ~1 GB of [N,N,P,P] f32


def bind_to_gpu_numa(local_rank: int):
    """Pin this rank to the CPUs of its GPU's NUMA node before any pinned buffer is allocated, so that first-touch
    places the staging memory next to the GPU's PCIe root (matters for the end-to-end number at N > 1, where eight
    H2D streams otherwise cross the socket interconnect). Returns what it did (for the JSON line)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dev = "/sys/bus/pci/devices/" + bus.lower()[-12:]
        cpulist = open(dev + "/local_cpulist").read().strip()
        node = open(dev + "/numa_node").read().strip()
        cpus = set()
        for part in cpulist.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return {"numa_node": int(node), "cpus": len(cpus)}
    except Exception as e:  # no NVML, no sysfs entry, single-node box: nothing to do
        return {"numa_node": None, "note": str(e)[:80]}
    return {"numa_node": None}


def run_trait(args, local_rank):
    """The drop-in route that keeps the reference's loader: `FeatureSet::compute_features_batched` per chunk of
    `batch_size` nuclei with the reference's own Batch layout (patchs [n,3,P,P] f32, masks [n,1,P,P] f32 in host memory,
    src/utils.rs:17, src/features/mod.rs:12-28) through nfx_compute_features_batched. Wall clock over whole chunks."""
    import nfx
    sets = args.sets.split(",") if args.sets else ["color"]
    P, B = 64, args.batch_size
    nchunks = max(1, (args.nuclei or 20_000) // B)
    tile, xy, off = make_inputs("color", B, 2048, P, 2, pinned=False)
    ex = nfx.Extractor(local_rank, P, B)
    ex.upload_tile(tile)
    ex.upload_polygons(xy, off)
    masks_u8 = ex.rasterize()
    patch_u8 = ex.gather_patches()
    cents = np.zeros((B, 2), np.float32)   # only used for the key column, which the host formats
    patchs = nfx.pinned_empty((B, 3, P, P), np.float32)
    masks = nfx.pinned_empty((B, 1, P, P), np.float32)
    patchs[:] = np.transpose(patch_u8.reshape(B, P, P, 3), (0, 3, 1, 2)).astype(np.float32) / np.float32(255.0)
    masks[:, 0] = masks_u8.reshape(B, P, P)
    rings = synth_rings(xy, off)
    bits = {"geometry": nfx.FS_GEOMETRY, "color": nfx.FS_COLOR, "glcm": nfx.FS_GLCM, "glrlm": nfx.FS_GLRLM, "gabor": nfx.FS_GABOR}
    todo = [nfx.FS_ALL] if sets == ["union"] else [bits[s] for s in sets]   # "union": all five sets in ONE call (one upload)
    def chunk():
        return [ex.compute_features_batched(b, cents, rings, patchs, masks) for b in todo]
    for _ in range(max(args.warmup, 3)):
        chunk()
    t0 = time.perf_counter()
    for _ in range(nchunks):
        chunk()
    dt = time.perf_counter() - t0
    h2d = patchs.nbytes + masks.nbytes
    print(json.dumps({
        "metric": "nuclei/sec", "value": nchunks * B / dt, "unit": "nuclei/s", "n_gpus": 1, "steps": nchunks, "warmup": max(args.warmup, 3),
        "ms_per_step": dt / nchunks * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 in, u8/f32 inside",
        "data": "synthetic",
        "config": {"workload": f"trait: compute_features_batched({'+'.join(sets)}) per chunk of {B} nuclei, {P}x{P} f32 patches and masks from pinned host "
                               f"memory ({h2d / 1e6:.1f} MB per call and set), {nchunks} chunks", "timing": "host wall clock"},
        "e2e": {"value": nchunks * B / dt, "unit": "nuclei/s", "h2d_bytes_per_step": h2d * len(todo), "d2h_bytes_per_step": 4 * B * sum(len(nfx.feature_names(b)) for b in todo)},
        "gpu_launches": int(ex.launch_count()),
    }))
    ex.close()


def synth_rings(xy, off):
    return [xy[off[i]:off[i + 1]] for i in range(len(off) - 1)]


def run_pipeline(args, local_rank):
    """The reference's whole main() minus file I/O (src/main.rs:110-190) on one GPU: GeoJSON text in host memory ->
    CSR polygons (nfx_geojson_parse, host threads) -> features (tile uploaded from pinned host memory every step) ->
    CSV text in pinned host memory (nfx_csv_rows: cells formatted on the GPU). Wall-clock per stage, whole steps."""
    import nfx
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import bench_geojson
    sets = args.sets.split(",") if args.sets else ["color"]
    nuclei = args.nuclei or 200_000
    side = args.tile or 16384
    P = 64
    tile, xy, off = make_inputs("color", nuclei, side, P, 2, pinned=True)
    text = np.frombuffer(bench_geojson.make_text_from(xy, off), dtype=np.uint8)
    mask = nfx.parse_feature_sets(sets)
    F = len(nfx.feature_names(mask))
    ex = nfx.Extractor(local_rank, P, args.batch_size)
    block = max(1, min(nuclei, (256 << 20) // (12 * (F + 2))))
    buf = nfx.pinned_empty((block * (F + 2) * 14,), np.uint8)
    stages = {"geojson_parse (tile H2D in flight)": 0.0, "rest of tile H2D + compute": 0.0, "csv_format+d2h": 0.0}
    K = max(args.steps, 1)
    csv_bytes = 0
    for it in range(args.warmup + K):
        t0 = time.perf_counter()
        ex.upload_tile(tile)                       # asynchronous H2D from pinned memory: overlaps the host-side parse
        pxy, poff, _bbox, _rings = nfx.geojson_pack(text, 0)
        t1 = time.perf_counter()
        ex.upload_polygons(pxy, poff)
        ex.compute(mask)
        ex.sync()
        t2 = time.perf_counter()
        nb = 0
        for lo in range(0, nuclei, block):
            nb += len(ex.csv_rows(lo, min(lo + block, nuclei), buf, view=True))
        t3 = time.perf_counter()
        if it >= args.warmup:
            stages["geojson_parse (tile H2D in flight)"] += t1 - t0
            stages["rest of tile H2D + compute"] += t2 - t1
            stages["csv_format+d2h"] += t3 - t2
            csv_bytes = nb
    total = sum(stages.values())
    print(json.dumps({
        "metric": "nuclei/sec", "value": nuclei * K / total, "unit": "nuclei/s", "n_gpus": 1, "steps": K, "warmup": args.warmup,
        "ms_per_step": total / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/f32",
        "data": "synthetic",
        "config": {"workload": f"pipeline: GeoJSON text ({text.size / 1e6:.0f} MB) -> {'+'.join(sets)} -> CSV text ({csv_bytes / 1e6:.0f} MB), "
                               f"{nuclei} nuclei, {P}x{P} windows, tile {side}x{side} uploaded every step, host threads = {os.cpu_count()}",
                   "timing": "host wall clock around whole stages (they include host work and PCIe)"},
        "stages_ms": {k: v / K * 1e3 for k, v in stages.items()},
        "gpu_launches": int(ex.launch_count()),
    }))
    ex.close()


# ======================================================================================================
# Building blocks of the default run
# ======================================================================================================
def hbm_peak():
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback 6650 GB/s (B200_PROFILING.md)"


# kernel sources behind the captured kernels of each workload in profiles/r2_traffic.json (shared headers included)
TRAFFIC_SOURCES = {"color": ("color.cu", "geom.cu", "nfx_device.cuh", "nfx_kernels.h"),
                   "staged": ("staged.cu", "geom.cu", "nfx_device.cuh", "nfx_kernels.h")}


def source_hash(files=None):
    """sha256 (16 hex digits) over kernel sources (all of csrc/ by default): ties the ncu-derived DRAM traffic in
    profiles/r2_traffic.json to the kernels it was captured from (a stale entry reads as null, never as a number)."""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "nuclei-feature-extraction_b200", "csrc")
    for f in sorted(files if files is not None else os.listdir(d)):
        if f.endswith((".cu", ".cuh", ".h")):
            h.update(f.encode())
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def measured_traffic(workload, kernel, nuclei, P):
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
        wl = tj.get(workload, {})
        if wl.get("src_sha16") != source_hash(TRAFFIC_SOURCES[workload]):
            return None
        ent = wl.get(kernel)
        if ent and ent["nuclei"] == nuclei and ent["patch"] == P:
            return ent["dram_bytes_per_launch"]
    except Exception:
        pass
    return None


def resident(ex, mask, K, W, min_s=0.0, barrier=lambda: None):
    """W (or more, for >= min_s seconds under load) untimed steps, then exactly K timed steps with the inputs already in
    HBM: CUDA events on the context stream around the K steps and around every kernel launch."""
    t_w, nw = time.perf_counter(), 0
    while nw < W or (time.perf_counter() - t_w < min_s and nw < 2000):
        ex.compute(mask)
        nw += 1
        if nw % 8 == 0:
            ex.sync()
    ex.sync()
    barrier()
    ex.profile(True)
    ex.profile_reset()
    l0 = ex.launch_count()
    ex.sync()
    barrier()
    ex.timer_start()
    for _ in range(K):
        ex.compute(mask)
    ms = ex.timer_stop() / K
    barrier()
    launches = ex.launch_count() - l0
    prof = ex.profile_get()
    ex.profile(False)
    kern = {k: {"launches": v[0], "avg_ms": v[1] / max(v[0], 1)} for k, v in prof.items()}
    tot = sum(v["avg_ms"] for v in kern.values()) or 1.0
    for v in kern.values():
        v["share"] = v["avg_ms"] / tot
    return {"ms": ms, "kern": kern, "launches": launches, "warmup_steps_run": nw}


def roofline_of(run, workload, nuclei, P, F, peak, peak_kind, world=1):
    """HBM roofline of the dominant kernel: algorithmic bytes per launch / its CUDA-event time."""
    kern = run["kern"]
    if not kern:
        return None
    R = max(1, min(P, 256 // ((P + 3) // 4)))      # hue_slab_rows(P) of color.cu
    slabs = (P + R - 1) // R
    for k, v in kern.items():
        v["gbs"] = kernel_bytes(k, P, slabs) * nuclei / (v["avg_ms"] * 1e-3) / 1e9 if v["avg_ms"] > 0 else 0.0
    dom = max(kern, key=lambda k: kern[k]["avg_ms"])
    value = nuclei / (run["ms"] * 1e-3)
    return {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["gbs"], "peak": peak, "unit": "GB/s",
            "frac": kern[dom]["gbs"] / peak, "traffic": measured_traffic(workload, dom, nuclei, P), "peak_source": peak_kind,
            "bytes_per_nucleus": kernel_bytes(dom, P, slabs), "avg_launch_ms": kern[dom]["avg_ms"],
            "pipeline_bytes_per_nucleus": algorithmic_bytes(workload, P, F),
            "pipeline_frac": value * algorithmic_bytes(workload, P, F) / 1e9 / peak}


def e2e_double_buffered(ctxs, outs, submit, n_steps):
    """n_steps whole steps (pinned H2D of every input, kernels, D2H of the results) through two contexts used alternately,
    like the reference's rayon workers overlap their batches: step k's copies overlap step k-1's kernels; the timed region
    contains the full H2D + kernels + D2H of every step. Returns ms per step (host wall clock)."""
    for c, (oc, of) in zip(ctxs, outs):     # warm both contexts
        submit(c)
        c.download(oc, of)
    t0 = time.perf_counter()
    submit(ctxs[0])
    for k in range(1, n_steps):
        submit(ctxs[k & 1])
        ctxs[(k - 1) & 1].download(*outs[(k - 1) & 1])
    ctxs[(n_steps - 1) & 1].download(*outs[(n_steps - 1) & 1])
    return 1e3 * (time.perf_counter() - t0) / n_steps


SET_NAMES = {"geometry": "shape", "color": "color", "glcm": "glcm", "glrlm": "glrlm", "gabor": "gabor", "all": "all"}

# Shared-memory atomics of the GLCM kernel (BASELINE north_star: "shared-memory atomic throughput against its peak for GLCM").
# Per nucleus of the bench's synthetic set, from the ncu capture of k_glcm64 in profiles/r2_tex_ncu.txt (20 000 nuclei):
#   smsp__inst_executed_op_shared_atom.sum = 32.23 M warp-instructions x 28.49 active lanes  -> lane-atomics
#   l1tex__data_pipe_lsu_wavefronts_mem_shared.sum = 276.8 M                                 -> shared-memory wavefronts
# Peaks: scripts/atoms_probe.cu (profiles/r1_smem_atomics_probe.txt): 3.5e12 random-address, 9.1e12 conflict-free
# shared-memory atomics per second and B200; the shared-memory data pipe moves one 128-byte wavefront per clock and SM.
GLCM_LANE_ATOMICS_PER_NUCLEUS = 32230563 * 28.49 / 20000
GLCM_SMEM_WAVEFRONTS_PER_NUCLEUS = 276781036 / 20000
SMEM_ATOMICS_PEAK_RANDOM, SMEM_ATOMICS_PEAK_CONFLICT_FREE = 3.5e12, 9.1e12
B200_SMS = 148


def glcm_atomics(nuclei, kernel_ms, sm_mhz=1965.0):
    rate = GLCM_LANE_ATOMICS_PER_NUCLEUS * nuclei / (kernel_ms * 1e-3)
    wf = GLCM_SMEM_WAVEFRONTS_PER_NUCLEUS * nuclei / (kernel_ms * 1e-3)
    wf_peak = B200_SMS * sm_mhz * 1e6
    return {"lane_atomics_per_nucleus": GLCM_LANE_ATOMICS_PER_NUCLEUS, "source": "ncu capture of k_glcm64, P = 64 (profiles/r2_tex_ncu.txt)",
            "achieved_per_s": rate, "peak_random_address_per_s": SMEM_ATOMICS_PEAK_RANDOM,
            "peak_conflict_free_per_s": SMEM_ATOMICS_PEAK_CONFLICT_FREE, "frac_of_random_address_peak": rate / SMEM_ATOMICS_PEAK_RANDOM,
            "smem_pipe": {"wavefronts_per_nucleus": GLCM_SMEM_WAVEFRONTS_PER_NUCLEUS, "achieved_per_s": wf,
                          "peak_per_s": wf_peak, "frac": wf / wf_peak,
                          "note": "all shared-memory wavefronts of the kernel (atomics, their bank-conflict replays, loads, clears) "
                                  "against one wavefront per clock and SM at the maximum SM clock"}}


def per_set_block(args, ex, ex2, tile, xy, off, nuclei, P, peak, peak_kind, cpu):
    """BASELINE metric "nuclei/sec per feature set": every set on the headline's resident inputs (100 000 nuclei, 64 x 64
    windows, 16384^2 tile): device-resident rate, dominant kernel and its HBM fraction, end-to-end rate, CPU oracle rate."""
    import nfx
    out = {}
    for name in ("geometry", "color", "glcm", "glrlm", "gabor", "all"):
        mask = nfx.parse_feature_sets([name])
        F = len(nfx.feature_names(mask))
        wl = SET_NAMES[name]
        run = resident(ex, mask, 10 if name in ("geometry", "color") else 4, 3)
        roof = roofline_of(run, wl, nuclei, P, F, peak, peak_kind)
        rec = {"columns": F, "ms_per_step": run["ms"], "value": nuclei / (run["ms"] * 1e-3), "unit": "nuclei/s",
               "dominant_kernel": roof["kernel"], "frac": roof["frac"], "pipeline_frac": roof["pipeline_frac"],
               "kernels_ms": {k: v["avg_ms"] for k, v in run["kern"].items()}, "gpu_launches": run["launches"]}
        if "k_glcm" in run["kern"] and P <= 64:
            rec["smem_atomics"] = glcm_atomics(nuclei, run["kern"]["k_glcm"]["avg_ms"])
        if not args.no_e2e:
            outs = [(nfx.pinned_empty((nuclei, 2), np.float32), nfx.pinned_empty((nuclei, F), np.float32)) for _ in range(2)]

            def submit(c):
                c.upload_tile(tile)
                c.upload_polygons(xy, off)
                c.compute(mask)
            ms = e2e_double_buffered([ex, ex2], outs, submit, 4)
            rec["e2e"] = {"value": nuclei / (ms * 1e-3), "unit": "nuclei/s", "ms_per_step": ms,
                          "h2d_bytes_per_step": tile.nbytes + xy.nbytes + off.nbytes, "d2h_bytes_per_step": 8 * nuclei + 4 * F * nuclei}
        if cpu:
            workers = cpu_workers()
            rate, dt, done = cpu_reference_rate([name], tile, xy, off, P, args.batch_size, workers * args.batch_size, workers, budget_s=5.0)
            rec["cpu_baseline"] = {"value": rate, "unit": "nuclei/s", "cores": workers, "kind": "port",
                                   "sample": f"first {done} nuclei ({dt:.1f} s), oracle on {workers} chunk-parallel host threads"}
        out[name] = rec
    return out


def fill_slide(ex, side, stage, T, y_lo=0, y_hi=None):
    """Stream rows [y_lo, y_hi) of a side x side slide from the two pinned staging tiles; returns the bytes sent."""
    y_hi = side if y_hi is None else y_hi
    k, sent = 0, 0
    for y in range(y_lo, y_hi, T):
        for x in range(0, side, T):
            h, w = min(T, y_hi - y), min(T, side - x)
            ex.write_tile(stage[k & 1][:h, :w], x, y)
            sent += 3 * h * w
            k += 1
    return sent


def staging_tiles(T, seed):
    import nfx
    from nfx import synth
    block = synth.synth_tile(min(T, 4096), min(T, 4096), seed)
    stage = [nfx.pinned_empty((T, T, 3), np.uint8) for _ in range(2)]
    for b in stage:
        for r in range(0, T, block.shape[0]):
            for c in range(0, T, block.shape[1]):
                b[r:r + block.shape[0], c:c + block.shape[1]] = block[:min(block.shape[0], T - r), :min(block.shape[1], T - c)]
    return stage


def slide_job(args, sets, nuclei, side, P, rank, local_rank, world, dist, seed, poly_kw, steps, label):
    """One slide resident in HBM on every GPU, nuclei split over the ranks in contiguous index ranges aligned to batch_size
    (index order is not spatial order, so every range touches every tile). Each rank receives only ITS 1/N of the rows
    from the host; the rest comes from the peers over NVLink (nfx_slide_export / nfx_slide_import_rows).
    Device-resident step = all kernels over the rank's range; end-to-end step = own rows H2D + peers' rows over NVLink +
    polygons H2D + kernels + D2H of the features. Times are the max over ranks."""
    import nfx
    from nfx import synth
    mask = nfx.parse_feature_sets(sets)
    F = len(nfx.feature_names(mask))
    T = min(8192, side)
    stage = staging_tiles(T, seed)
    xy, off = synth.synth_polygons_pool(nuclei, side, side, seed, patch=P, **poly_kw)
    bounds = nfx.partition(nuclei, args.batch_size, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    n = hi - lo
    pxy = nfx.pinned_empty((int(off[hi] - off[lo]), 2), np.float32)
    pxy[...] = xy[off[lo]:off[hi]]
    poff = nfx.pinned_empty((n + 1,), np.int64)
    poff[...] = off[lo:hi + 1] - off[lo]
    del xy
    cents = nfx.pinned_empty((n, 2), np.float32)
    feats = nfx.pinned_empty((n, F), np.float32)
    ex = nfx.Extractor(local_rank, P, args.batch_size)
    ex.slide_alloc(side, side)
    rows = [(side * r) // world for r in range(world + 1)]     # rank r owns rows [rows[r], rows[r+1])

    def barrier():
        if dist is not None:
            dist.barrier()

    def place_slide():
        sent = fill_slide(ex, side, stage, T, rows[rank], rows[rank + 1])
        nvl = 0
        if world > 1:
            ex.sync()
            handles = [None] * world
            dist.all_gather_object(handles, ex.slide_export())      # also the barrier: every rank's rows are in its HBM
            for q in range(world):
                if q != rank:
                    ex.slide_import_rows(handles[q], rows[q], rows[q + 1] - rows[q])
                    nvl += 3 * side * (rows[q + 1] - rows[q])
        return sent, nvl

    place_slide()
    ex.upload_polygons(pxy, poff)
    run = resident(ex, mask, steps, 1, barrier=barrier)
    t0 = time.perf_counter()
    sent, nvl = place_slide()
    ex.upload_polygons(pxy, poff)
    ex.compute(mask)
    ex.download(cents, feats)
    e2e_ms = 1e3 * (time.perf_counter() - t0)
    if world > 1:
        ex.sync()
        barrier()           # peers may still be reading this rank's rows
    ms = run["ms"]
    if dist is not None:
        import torch
        t = torch.tensor([ms, e2e_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms = float(t[0]), float(t[1])
    checksum = float(np.nansum(feats[:: max(1, n // 997)]))
    rec = {"workload": f"{label}: {'+'.join(sets)}, {nuclei} nuclei over a {side}x{side} u8 RGB slide resident in HBM "
                       f"({3 * side * side / 1e9:.1f} GB per GPU), P={P}, batch_size={args.batch_size}",
           "partition": f"contiguous index ranges aligned to batch_size ({n} nuclei on rank 0); slide rows split over the ranks "
                        f"for the host upload, peers' rows fetched over NVLink" if world > 1 else "one GPU",
           "n_gpus": world, "scaling": "strong", "nuclei": nuclei, "columns": F, "steps": steps,
           "ms_per_step": ms, "value": nuclei / (ms * 1e-3), "unit": "nuclei/s",
           "e2e": {"value": nuclei / (e2e_ms * 1e-3), "unit": "nuclei/s", "ms_per_step": e2e_ms,
                   "h2d_bytes_per_step_per_gpu": sent + pxy.nbytes + poff.nbytes, "nvlink_bytes_per_step_per_gpu": nvl,
                   "d2h_bytes_per_step_per_gpu": cents.nbytes + feats.nbytes},
           "kernels_ms": {k: v["avg_ms"] for k, v in run["kern"].items()}, "gpu_launches": run["launches"], "checksum": checksum}
    if run["kern"]:
        dom = max(run["kern"], key=lambda k: run["kern"][k]["avg_ms"])
        rec["dominant_kernel"] = dom
        R = max(1, min(P, 256 // ((P + 3) // 4)))
        peak, _ = hbm_peak()
        rec["frac"] = kernel_bytes(dom, P, (P + R - 1) // R) * n / (run["kern"][dom]["avg_ms"] * 1e-3) / 1e9 / peak
        if "k_glcm" in run["kern"] and P <= 64:
            rec["smem_atomics"] = glcm_atomics(n, run["kern"]["k_glcm"]["avg_ms"])
    ex.close()
    return rec


def config_records(args, local_rank, peak, peak_kind, cpu):
    """BASELINE.json configs 1, 3, 4 and 5 at their stated sizes (config 2 is the headline)."""
    import nfx
    out = {}
    # ---- config 1: shape set, 10k polygons, 4096^2 tile (the reference's own CPU-runnable case) ----
    sets, nuclei, side, P, kw = WORKLOADS["shape"]
    tile, xy, off = make_inputs("shape", nuclei, side, P, 1, pinned=True)
    mask = nfx.parse_feature_sets(sets)
    F = len(nfx.feature_names(mask))
    ex, ex2 = nfx.Extractor(local_rank, P, args.batch_size), nfx.Extractor(local_rank, P, args.batch_size)
    ex.upload_tile(tile)
    ex.upload_polygons(xy, off)
    run = resident(ex, mask, 20, 3)
    roof = roofline_of(run, "shape", nuclei, P, F, peak, peak_kind)
    rec = {"workload": f"shape: geometry set, {nuclei} nuclei, {P}x{P} windows, tile {side}x{side}", "ms_per_step": run["ms"],
           "value": nuclei / (run["ms"] * 1e-3), "unit": "nuclei/s", "dominant_kernel": roof["kernel"], "frac": roof["frac"],
           "note": "the geometry set reads no pixel: the HBM fraction is reported for completeness only",
           "kernels_ms": {k: v["avg_ms"] for k, v in run["kern"].items()}, "gpu_launches": run["launches"]}
    if not args.no_e2e:
        outs = [(nfx.pinned_empty((nuclei, 2), np.float32), nfx.pinned_empty((nuclei, F), np.float32)) for _ in range(2)]

        def submit(c):
            c.upload_tile(tile)
            c.upload_polygons(xy, off)
            c.compute(mask)
        ms = e2e_double_buffered([ex, ex2], outs, submit, 6)
        rec["e2e"] = {"value": nuclei / (ms * 1e-3), "unit": "nuclei/s", "ms_per_step": ms,
                      "h2d_bytes_per_step": tile.nbytes + xy.nbytes + off.nbytes, "d2h_bytes_per_step": 8 * nuclei + 4 * F * nuclei}
    if cpu:
        workers = cpu_workers()
        rate, dt, done = cpu_reference_rate(sets, tile, xy, off, P, args.batch_size, min(nuclei, workers * args.batch_size * 2), workers, budget_s=8.0)
        rec["cpu_baseline"] = {"value": rate, "unit": "nuclei/s", "cores": workers, "kind": "port",
                               "sample": f"first {done} nuclei of the same workload ({dt:.1f} s), oracle = torch-CPU restatement of the tch path"}
    out["1"] = rec
    ex.close()
    ex2.close()
    # ---- config 3: GLCM set, 1M nuclei ----
    sets, nuclei, side, P, kw = WORKLOADS["glcm"]
    nuclei = int(os.environ.get("NFX_BENCH_C3_NUCLEI", nuclei))
    out["3"] = slide_job(args, sets, nuclei, side, P, 0, local_rank, 1, None, 3, kw, 3, "glcm (config 3)")
    out["3"]["note"] = ("BASELINE words the set as '32 grey levels, distances 1/2, 4 angles' (8 matrices); the reference's GlcmFeatureSet "
                        "is levels {32,64,128,254} x 4 offsets at distance 1 = 16 matrices / 224 columns (src/features/texture.rs:19-20): "
                        "the drop-in computes and this line times the reference's set, the larger of the two")
    # ---- config 5: stress, 256x256 windows, 500-vertex polygons, all sets ----
    sets, nuclei, side, P, kw = WORKLOADS["stress"]
    nuclei = int(os.environ.get("NFX_BENCH_C5_NUCLEI", nuclei))
    out["5"] = slide_job(args, sets, nuclei, side, P, 0, local_rank, 1, None, 5, kw, 2, "stress (config 5)")
    # ---- config 4: all sets, 5M nuclei, 100k x 100k slide ----
    sets, nuclei, side, P, kw = WORKLOADS["slide"]
    nuclei = int(os.environ.get("NFX_BENCH_C4_NUCLEI", nuclei))
    side = int(os.environ.get("NFX_BENCH_C4_SIDE", side))
    out["4"] = slide_job(args, sets, nuclei, side, P, 0, local_rank, 1, None, 4, kw, 1, "slide (config 4)")
    return out


def h2d_rate(ex, tile, barrier):
    """GB/s of ONE pinned upload of this rank's tile while every other rank does the same (what bounds the weak-scaling
    end-to-end number: the ranks share the host's PCIe root complexes)."""
    ex.sync()
    barrier()
    t0 = time.perf_counter()
    ex.upload_tile(tile)
    ex.sync()
    return tile.nbytes / (time.perf_counter() - t0) / 1e9


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="nfx", choices=["nfx", "reference"])
    ap.add_argument("--workload", default="color", choices=sorted(WORKLOADS) + ["staged", "pipeline", "trait"])
    ap.add_argument("--sets", default="", help="pipeline workload: comma separated feature sets (default color)")
    ap.add_argument("--nuclei", type=int, default=0)
    ap.add_argument("--tile", type=int, default=0)
    ap.add_argument("--batch-size", type=int, default=100)
    ap.add_argument("--cpu-sample", type=int, default=0, help="nuclei of the CPU-baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--quick", action="store_true", help="headline only: no per_set / configs / strong blocks")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.workload == "staged":
        if rank == 0:
            run_staged(args, local_rank)
        return
    if args.workload == "pipeline":
        if rank == 0:
            run_pipeline(args, local_rank)
        return
    if args.workload == "trait":
        if rank == 0:
            run_trait(args, local_rank)
        return
    sets, nuclei, side, P, poly_kw = WORKLOADS[args.workload]
    nuclei = args.nuclei or nuclei
    side = args.tile or side
    W = max(args.warmup, 3) if args.impl == "nfx" else args.warmup
    K = max(args.steps, 1)
    cores = os.cpu_count() or 1
    quick = args.quick or bool(os.environ.get("NFX_BENCH_QUICK")) or args.workload != "color" or bool(args.nuclei) or bool(args.tile)

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        workers = cpu_workers()
        sample = args.cpu_sample or workers * args.batch_size * {"color": 2, "shape": 4}.get(args.workload, 1)
        if args.workload == "slide":       # the 100k^2 slide is never held on the host: same nuclei density on one 8192^2 tile
            tile, xy, off = make_inputs(args.workload, max(sample, 1000), 8192, P, 4, pinned=False)
        else:                              # exactly the GPU arm's rank-0 inputs; the oracle runs on their first `sample` nuclei
            tile, xy, off = make_inputs(args.workload, nuclei, side, P, 2, pinned=False)
        sample = min(sample, len(off) - 1)
        rates = []
        for it in range(args.warmup + K):
            r, dt, done = cpu_reference_rate(sets, tile, xy, off, P, args.batch_size, sample, workers)
            if it >= args.warmup:
                rates.append((done, dt))
        total_t = sum(d for _, d in rates)
        value = sum(n_ for n_, _ in rates) / total_t
        line = {
            "impl": "reference", "metric": "nuclei/sec", "value": value, "unit": "nuclei/s", "n_gpus": args.gpus,
            "steps": K, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / len(rates), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, sets, nuclei, side, P),
            "cpu_baseline": {"value": value, "unit": "nuclei/s", "cores": workers, "kind": "port",
                             "sample": f"each step = the first {sample} nuclei of that workload (same tile, same polygons, same seed as the "
                                       f"GPU arm's rank 0) x {K} steps, oracle (torch-CPU restatement of the tch path), {workers} "
                                       f"chunk-parallel host threads of {cores} cores"},
            "e2e": {"value": value, "unit": "nuclei/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ this repo's CUDA path
    import nfx
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()

    peak, peak_kind = hbm_peak()
    if args.workload == "slide":
        rec = slide_job(args, sets, nuclei, side, P, rank, local_rank, world, dist, 4, poly_kw, max(1, min(args.steps, 3)), "slide (config 4)")
        if rank == 0:
            line = {"metric": "nuclei/sec", "value": rec["value"], "unit": "nuclei/s", "n_gpus": world, "steps": rec["steps"], "warmup": 1,
                    "ms_per_step": rec["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                    "dtype": "u8/f32", "data": "synthetic", "config": {"workload": rec["workload"], "partition": rec["partition"]},
                    "e2e": rec["e2e"], "gpu_launches": rec["gpu_launches"], "kernels_ms": rec["kernels_ms"], "checksum": rec["checksum"]}
            print(json.dumps(line))
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return
    mask = nfx.parse_feature_sets(sets)
    F = len(nfx.feature_names(mask))
    numa = bind_to_gpu_numa(local_rank) if world > 1 else None
    tile, xy, off = make_inputs(args.workload, nuclei, side, P, 2 + rank, pinned=True)
    ex = nfx.Extractor(local_rank, P, args.batch_size)
    ex.upload_tile(tile)
    ex.upload_polygons(xy, off)
    cents = nfx.pinned_empty((nuclei, 2), np.float32)
    feats = nfx.pinned_empty((nuclei, F), np.float32)

    # ---- device-resident throughput ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    exact = bool(os.environ.get("NFX_BENCH_EXACT_WARMUP"))   # profiling runs: exactly W warm-up steps
    run = resident(ex, mask, K, W, min_s=0.0 if exact else 0.25, barrier=barrier)
    clocks = sampler.stop()
    ex.download(cents, feats)
    checksum = float(np.nansum(feats[:: max(1, nuclei // 997)]))

    # ---- end to end through the host API with host buffers ----
    e2e_ms = None
    h2d = tile.nbytes + xy.nbytes + off.nbytes
    d2h = cents.nbytes + feats.nbytes
    ex2 = None
    h2d_gbs = None
    if not args.no_e2e:
        ex2 = nfx.Extractor(local_rank, P, args.batch_size)
        outs = [(cents, feats), (nfx.pinned_empty((nuclei, 2), np.float32), nfx.pinned_empty((nuclei, F), np.float32))]

        def submit(c):
            c.upload_tile(tile)
            c.upload_polygons(xy, off)
            c.compute(mask)

        barrier()
        e2e_ms = e2e_double_buffered([ex, ex2], outs, submit, max(4, min(K, 6)))
        h2d_gbs = h2d_rate(ex, tile, barrier)

    # ---- max over ranks ----
    step_ms = run["ms"]
    if dist is not None:
        import torch
        t = torch.tensor([step_ms, e2e_ms or 0.0], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        step_ms, e2e_max = float(t[0]), float(t[1])
        e2e_ms = e2e_max if e2e_ms is not None else None
        if h2d_gbs is not None:
            g = torch.tensor([h2d_gbs], device="cuda")
            allg = [torch.zeros_like(g) for _ in range(world)]
            dist.all_gather(allg, g)
            h2d_gbs = [float(x[0]) for x in allg]
    total = nuclei * world
    value = total / (step_ms * 1e-3)
    run["ms"] = step_ms

    line = None
    if rank == 0:
        roof = roofline_of(run, args.workload, nuclei, P, F, peak, peak_kind)
        if roof:
            roof["pipeline_frac"] = value / world * algorithmic_bytes(args.workload, P, F) / 1e9 / peak
        cpu = None
        if not args.no_cpu_baseline:
            workers = cpu_workers()
            sample = args.cpu_sample or workers * args.batch_size * {"color": 3, "shape": 6}.get(args.workload, 1)
            sample = min(sample, nuclei)
            rate, dt, done = cpu_reference_rate(sets, tile, xy, off, P, args.batch_size, sample, workers)
            cpu = {"value": rate, "unit": "nuclei/s", "cores": workers, "kind": "port",
                   "sample": f"first {done} nuclei of the same workload ({dt:.1f} s), oracle = torch-CPU restatement of the "
                             f"tch path, {workers} chunk-parallel host threads of {cores} cores"}
        line = {
            "metric": "nuclei/sec", "value": value, "unit": "nuclei/s", "n_gpus": world, "steps": K, "warmup": W, "warmup_steps_run": run["warmup_steps_run"],
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/f32", "data": "synthetic",
            "config": workload_config(args, sets, nuclei, side, P),
            "e2e": None if e2e_ms is None else {"value": total / (e2e_ms * 1e-3), "unit": "nuclei/s",
                                                 "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                                                 "ms_per_step": e2e_ms,
                                                 "h2d_gbs_per_rank_all_ranks_uploading": h2d_gbs},
            "gpu_launches": run["launches"],
            "clocks": clocks,
            "roofline": roof,
            "kernels": run["kern"],
            "cpu_baseline": cpu,
            "checksum": checksum,
            "csrc_sha16": source_hash(),
        }
        if numa is not None:
            line["host_numa_binding_rank0"] = numa
    # ---- the other feature sets and BASELINE configs, measured in the same run (one GPU) ----
    if not quick and world == 1:
        if ex2 is None:
            ex2 = nfx.Extractor(local_rank, P, args.batch_size)
        line["per_set"] = per_set_block(args, ex, ex2, tile, xy, off, nuclei, P, peak, peak_kind, not args.no_cpu_baseline)
    if ex2 is not None:
        ex2.close()
    ex.close()
    del tile, cents, feats
    if not quick and world == 1:
        line["configs"] = config_records(args, local_rank, peak, peak_kind, not args.no_cpu_baseline)
        line["configs"]["2"] = "the headline of this line"
    # ---- N > 1: BASELINE config 4 split over the ranks (strong scaling) ----
    if not quick and world > 1:
        s_sets, s_nuclei, s_side, s_P, s_kw = WORKLOADS["slide"]
        s_nuclei = int(os.environ.get("NFX_BENCH_C4_NUCLEI", s_nuclei))
        s_side = int(os.environ.get("NFX_BENCH_C4_SIDE", s_side))
        rec = slide_job(args, s_sets, s_nuclei, s_side, s_P, rank, local_rank, world, dist, 4, s_kw, 1, "slide (config 4)")
        if rank == 0:
            line["strong"] = rec
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
