"""The other BASELINE.json configs as `bench.py --workload <name>` (the default line stays configs[1]):

    zstack256   north_star target: the full pipeline on one 2048 x 2048 x 256 uint16 stack per GPU
    refine4096  configs[2]: refine_boundaries (threshold -> EDT -> local maxima -> label -> watershed) on the
                4096 x 4096 touching-particle boundary map (~10k dense touching particles)
    class2048   the reference's real main path (tiff_analysis.py:627-671): the seed-1234 2048 x 2048 ilastik-style
                class image through the tiff_analysis mirrors (median -> label/classify/merge -> counts -> particle
                area recreation)
    nanosims    configs[3]: the 5-isotope 256 x 256 ROI stack: per-ROI sums / activities / distances + binning

Every line carries the same keys as the default one: value (device-resident inputs, CUDA events), e2e (numpy in,
numpy out through the public API, copies inside the timed region), roofline of the kernel with the largest share of
the step, cpu_baseline (the oracle on one thread -- how the reference runs -- and on all cores where the workload
has independent units).  N = 1 only: these are single-image workloads.
"""

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic bytes per pixel and launch for the kernels that dominate these workloads (inputs read once + outputs
# written once; the watershed sweep streams its whole state)
EXTRA_BYTES = {
    "k_ws_sweep": 8.0 + 1.0 + 16.0 + 16.0,  # image f64 + kind + (b f64, hops i32, label i32) in and out
    "k_ws_init": 8.0 + 4.0 + 0.125 + 16.0 + 1.0,
    "k_conn_planes": 4.0 + 0.875,
    "k_median_u8": 2.0,
    "k_dilate_bits": 0.25,
    "k_member_u8": 1.125,
    "k_compare": 4.125,
    "k_roi_sums": 4.0 + 8.0 * 5,
}


def _events(fn, steps, warmup):
    import torch

    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def _wall(fn, steps, warmup):
    import torch

    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps * 1e3


def _profile(lib, fn, steps, units_per_launch, bench):
    """Per-kernel CUDA-event times of `steps` runs -> roofline dict of the dominant kernel."""
    import torch

    lib.pcs_profile_enable(1)
    for _ in range(steps):
        fn()
    torch.cuda.synchronize()
    lib.pcs_profile_enable(0)
    prof = bench.collect_profile(lib)
    if not prof:
        return None
    peak, peak_src = bench.measured_peak()
    tot = sum(v[0] for v in prof.values())
    name, (kms, kcnt) = max(prof.items(), key=lambda kv: kv[1][0])
    bpv = EXTRA_BYTES.get(name, bench.KERNEL_BYTES_PER_VOXEL.get(name))
    avg_ms = kms / kcnt
    achieved = None if bpv is None else bpv * units_per_launch / (avg_ms * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": None if achieved is None else achieved / peak, "traffic": None,
            "bytes_per_pixel": bpv, "avg_launch_ms": avg_ms, "launches": kcnt, "peak_source": peak_src,
            "kernel_time_share": {k: round(v[0] / tot, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])},
            "kernel_ms_per_step": {k: round(v[0] / steps, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])},
            "timed": "per-kernel CUDA events on the launching stream over a second pass (eager launches)"}


def _line(metric, unit, value, ms, args, workload, dtype, extra):
    line = {"metric": metric, "value": value, "unit": unit, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic", "config": {"workload": workload}}
    line.update(extra)
    return line


_CPU_FN = None


def _cpu_call(_i):
    t = time.perf_counter()
    _CPU_FN()
    return time.perf_counter() - t


def _cpu_time(fn, copies=1):
    """Seconds per call of `fn` on one thread; with copies > 1, that many independent calls on that many processes
    (fork, before CUDA is initialised) -> seconds for the batch."""
    global _CPU_FN
    if copies <= 1:
        t = time.perf_counter()
        fn()
        return time.perf_counter() - t
    import multiprocessing as mp

    _CPU_FN = fn
    with mp.get_context("fork").Pool(copies) as pool:
        t = time.perf_counter()
        pool.map(_cpu_call, range(copies), chunksize=1)
        dt = time.perf_counter() - t
    _CPU_FN = None
    return dt


# ---------------------------------------------------------------------------------------------- refine4096
def run_refine4096(args, bench):
    from oracle import refine as oref
    from particle_col_image_segmentation_b200 import synth

    S = args.size if args.size != 2048 else 4096
    _, prob = synth.touching_particles(S, S, seed=1003, pitch=40.96)
    npx = float(S) * S
    cpu = None
    if not args.no_cpu:  # the restated script on one thread (one image: no independent units to spread over cores)
        crop = prob[:1024, :1024]  # bounded sample: a 1024^2 crop (the sequential flood takes ~1 min on the full map)
        t1 = _cpu_time(lambda: oref.refine_boundaries(crop, run_watershed=True))
        cpu = {"value": crop.size / t1 / 1e6, "unit": "Mpixel/s", "cores": 1, "kind": "port",
               "sample": "refine_boundaries incl. watershed on the top-left 1024x1024 crop of the seed-1003 map, one thread (a single image has no independent units)"}
    import torch

    from particle_col_image_segmentation_b200 import _lib, refine_boundaries

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    lib = _lib.load()
    t_dev = torch.from_numpy(prob).to(dev)
    sweeps = []

    def run_dev():
        out = refine_boundaries.refine_boundaries(t_dev, run_watershed=True, return_sweeps=True)
        sweeps.append(out["sweeps"])
        return out

    sampler = bench.ClockSampler(0)
    sampler.start()
    ms = _events(run_dev, args.steps, args.warmup)
    roof = None if args.no_profile else _profile(lib, run_dev, max(1, min(args.steps, 3)), npx, bench)
    clocks = sampler.stop()
    pin = torch.from_numpy(prob).pin_memory()
    ms_e2e = _wall(lambda: refine_boundaries.refine_boundaries(pin.numpy(), run_watershed=True), max(1, args.e2e_steps), 1)
    d2h = int(npx * (1 + 8 + 1 + 4 + 4))  # binary_mask, distance, local_max, markers, labels as numpy
    return _line("Mpixel/s refine_boundaries (threshold->EDT->local maxima->label->watershed)", "Mpixel/s", npx / ms / 1e3, ms, args,
                 f"refine_boundaries morphology + EDT + watershed on a {S}x{S} boundary-probability map with ~10k dense touching particles (BASELINE.json configs[2])", "f32/f64",
                 {"clocks": clocks, "watershed_sweeps": sweeps[-1] if sweeps else None, "roofline": roof, "cpu_baseline": cpu,
                  "e2e": {"value": npx / ms_e2e / 1e3, "unit": "Mpixel/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": int(npx * 4), "d2h_bytes_per_step": d2h},
                  "gpu_launches": None})


# ---------------------------------------------------------------------------------------------- class2048
def run_class2048(args, bench):
    from oracle import l2 as ol2
    from particle_col_image_segmentation_b200 import synth

    S = args.size
    img = synth.class_image(S, S, seed=1234)
    npx = float(S) * S
    cell_types = {1: "Cells", 2: "Particle", 3: "Background"}

    def cpu_once():  # tiff_analysis.py:642-651, restated call for call in oracle/l2.py
        a = ol2.denoise(ol2.normalize_ds_arr(img, side=None))
        pos, clusters, parea, merged = ol2.get_cell_positions_and_areas(a, cell_types, merged=True)
        ol2.get_cell_counts_and_densities(pos, clusters, parea)
        ol2.recreate_particle_area(a, cell_types, parea)

    cpu = None
    if not args.no_cpu:
        t1 = _cpu_time(cpu_once)
        cores = os.cpu_count() or 1
        n = max(1, min(cores, 16))
        tn = _cpu_time(cpu_once, copies=n)
        cpu = {"value": n * npx / tn / 1e6, "unit": "Mpixel/s", "cores": n, "kind": "port", "sample": f"{n} copies of the seed-1234 {S}x{S} class image, one process each",
               "single_thread": {"value": npx / t1 / 1e6, "unit": "Mpixel/s", "cores": 1, "seconds_per_image": t1, "sample": "one image, one thread: how the reference runs (BASELINE.md: ~7.8 s per image in the survey container)"}}
    import torch

    from particle_col_image_segmentation_b200 import _lib, tiff_analysis

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    lib = _lib.load()
    t_dev = torch.from_numpy(img).to(dev)
    def run(x):
        return tiff_analysis.process_single_array(x, cell_types)

    sampler = bench.ClockSampler(0)
    sampler.start()
    ms = _wall(lambda: run(t_dev), args.steps, args.warmup)  # the L2 functions return Python objects: wall clock around the call
    # kernel share of the call: sum of the per-kernel event times of one profiled pass
    roof = None if args.no_profile else _profile(lib, lambda: run(t_dev), max(1, min(args.steps, 3)), npx, bench)
    clocks = sampler.stop()
    kernel_ms = None if roof is None else sum(roof["kernel_ms_per_step"].values())
    ms_e2e = _wall(lambda: run(img), max(1, args.e2e_steps), 1)
    return _line("Mpixel/s tiff_analysis single-file path (median->label/classify/merge->counts->particle area)", "Mpixel/s", npx / ms / 1e3, ms, args,
                 f"tiff_analysis.process_single_h5_file pixel work (tiff_analysis.py:642-651) on the seed-1234 {S}x{S} uint8 class image", "u8",
                 {"clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "host_share": None if kernel_ms is None else max(0.0, 1.0 - kernel_ms / ms),
                  "timing": "wall clock around the drop-in call with the class image resident on the device (the functions return Python region lists, so they synchronise)",
                  "e2e": {"value": npx / ms_e2e / 1e3, "unit": "Mpixel/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": int(npx), "d2h_bytes_per_step": int(npx)},
                  "gpu_launches": None})


# ---------------------------------------------------------------------------------------------- nanosims
def run_nanosims(args, bench):
    from oracle import nanosims as onano
    from particle_col_image_segmentation_b200 import synth

    planes, roi, set_id, agg = synth.nanosims_stack(256, 5, seed=1004)
    red = np.isin(roi, np.nonzero(set_id == 1)[0] + 1)
    green = np.isin(roi, np.nonzero(set_id == 2)[0] + 1)
    edges = np.linspace(0, 5, 11)
    npx = float(planes.shape[1]) * planes.shape[2]

    def cpu_once():
        rows = onano.analyse(planes, red, green, agg)
        return onano.activity_vs_distance(rows[:, 2 + 5], rows[:, -1], edges)

    cpu = None
    if not args.no_cpu:
        reps = 5
        t1 = _cpu_time(lambda: [cpu_once() for _ in range(reps)]) / reps
        cpu = {"value": npx / t1 / 1e6, "unit": "Mpixel/s", "cores": 1, "kind": "port", "sample": "the restated MATLAB script on the 5 x 256 x 256 stack, one thread, mean of 5 (parity with MATLAB itself is unpinned)"}
    import torch

    from particle_col_image_segmentation_b200 import _lib, nanosims

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    lib = _lib.load()
    d = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (planes, red, green, agg)]

    def run(p, r, g, a):
        rows = nanosims.analyse(p, r, g, a)
        return nanosims.activity_vs_distance(rows[:, 2 + 5], rows[:, -1], edges)

    sampler = bench.ClockSampler(0)
    sampler.start()
    ms = _wall(lambda: run(*d), args.steps * 5, args.warmup)
    roof = None if args.no_profile else _profile(lib, lambda: run(*d), 3, npx, bench)
    clocks = sampler.stop()
    ms_e2e = _wall(lambda: run(planes, red, green, agg), args.steps * 5, 1)
    return _line("Mpixel/s NanoSIMS per-ROI activity vs boundary distance (5 isotopes)", "Mpixel/s", npx / ms / 1e3, ms, args,
                 "NanoSIMS-style 5-isotope 256x256 ROI stack: per-ROI sums, activities, nearest-neighbour and boundary distances, distance binning (BASELINE.json configs[3])", "f64",
                 {"clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "note": "a 65 536-pixel problem: launch latency bound (a dozen kernels of a few microseconds), not bandwidth bound",
                  "e2e": {"value": npx / ms_e2e / 1e3, "unit": "Mpixel/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": int(planes.nbytes + red.nbytes + green.nbytes + agg.nbytes), "d2h_bytes_per_step": None},
                  "gpu_launches": None})


RUNNERS = {"refine4096": run_refine4096, "class2048": run_class2048, "nanosims": run_nanosims}


def run(args, bench):
    print(json.dumps(RUNNERS[args.workload](args, bench)))
