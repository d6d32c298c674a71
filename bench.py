#!/usr/bin/env python3
"""Benchmark of the segmentation hot path (BASELINE.json metric: Mvoxel/s of the full
segment pipeline, % of the HBM roofline, CPU reference beside it).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # reference CPU path (oracle port)

One "step" = one pass of threshold -> median -> label -> regionprops -> refine -> EDT
over one synthetic (Z, 2048, 2048) uint16 z-stack per GPU (BASELINE.json configs[1]:
Z = 64).  Weak scaling: every rank owns its own stack; the only exchange is the
gather of the per-label tables.  Rank 0 prints ONE JSON line.
"""

import argparse
import ctypes
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "Mvoxel/s full segment pipeline (threshold->label->refine->EDT->regionprops)"
PIPELINE_BYTES_PER_VOXEL = 16.0  # u16 in 2 + mask 1 + labels 4 + refined 1 + EDT f64 8
SUM_OF_STAGES_BYTES_PER_VOXEL = 27.0
# algorithmic bytes per voxel of each kernel (DESIGN.md section "Kernels"): inputs that must be
# read once + outputs that must be written once, bit images counted as 1/8 B
KERNEL_BYTES_PER_VOXEL = {
    "k_hist_u16": 2.0,
    "k_seg_threshold_tile": 2.0 + 1.0 + 0.125,  # uint16 in, uint8 mask + bit rows out (parents / run sums / word list are sparse)
    "k_seg_merge_list": 0.125,
    "k_seg_flatten_list": 0.125,
    "k_seg_rank_list": 0.125,
    "k_seg_label_list": 0.125,
    "k_seg_relabel_table": 4.0 + 0.125,
    "k_compare": 2.0 + 0.125,
    "k_majority5_bits": 0.25 + 1.0,
    "k_unpack": 1.0 + 0.125,
    "k_ccl_tile": 0.125,
    "k_ccl_merge_edges": 0.125,
    "k_ccl_flatten": 0.125,
    "k_ccl_rank": 0.125,
    "k_ccl_relabel": 4.0 + 0.125,
    "k_ccl_mark": 0.125,
    "k_ccl_select": 0.25 + 1.0,
    "k_refine_rows": 0.375,
    "k_hole_select": 0.375 + 1.0,
    "k_bbox_raster": 0.125,
    "k_hole_candidates": 0.5,
    "k_region_table": 4.0 + 2.0,
    "k_region_table_bits": 4.0 + 2.0,
    "k_select_by_area": 0.25,
    "k_edt_transpose": 0.25,
    "k_edt_carry": 0.25,
    "k_edt_near": 8.0 + 0.25,  # every float64 of the tile (background zeros included) is written by this kernel, once
    "k_edt_far": 8.0 + 0.25,
}


# DRAM traffic per voxel of the dominant kernels from the `ncu --set full` capture of profiles/prof_step.py
# (16 slices of 2048^2 per launch; dram__bytes_read.sum + dram__bytes_write.sum over 67.1 Mvoxel), see
# profiles/r2_ncu_full_summary.txt.  Below the algorithmic figure where part of the output is still in L2
# when the kernel ends.
NCU_TRAFFIC_SOURCE = "profiles/r2_ncu_full_summary.txt"
NCU_TRAFFIC_BYTES_PER_VOXEL = {
    "k_edt_near": (16.81 + 481.60) / 67.109,
    "k_hist_u16": (136.03 + 4.01) / 67.109,
    "k_seg_threshold_tile": (142.32 + 59.99) / 67.109,
    "k_ccl_relabel": (18.31 + 210.57) / 67.109,
    "k_refine_rows": (18.48 + 25.75) / 67.109,
}
NCU_TRAFFIC_WHOLE_STEP_BYTES_PER_VOXEL = (431.3 + 782.0) / 67.109  # every kernel of the step: 18.1 B/voxel (round 1: 21.8; algorithmic: 16)


def write_only_probe(dev, lib):
    """Best-of-5 fill of a 1 GiB buffer with the library's 128-bit store kernel: the store-only ceiling of this
    GPU, next to the copy figure of MEASURED_PEAKS.json (which counts read + write bytes)."""
    import torch

    x = torch.empty(1 << 28, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    best = 1e9
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lib.pcs_fill_u32(x.data_ptr(), 1, x.numel(), st)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del x
    return (1 << 30) / (best * 1e-3) / 1e9


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--slices", type=int, default=64, help="slices per GPU (configs[1]: 64; north_star target: 256)")
    ap.add_argument("--size", type=int, default=2048)
    ap.add_argument("--chunk", type=int, default=32, help="slices per batched launch")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--e2e-chunk", type=int, default=4, help="slices per pipeline stage of the host-buffer path (H2D / kernels / D2H overlap chunk-wise)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="slices in the CPU baseline sample (0 = one per host core, at most 32)")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--streams", type=int, default=2, help="streams the chunks of a step are spread over (inside the captured graph)")
    ap.add_argument("--no-gather", action="store_true", help="diagnosis: skip the table gather of an N > 1 step")
    ap.add_argument("--workload", default="zstack64", choices=["zstack64", "zstack256", "refine4096", "class2048", "nanosims"],
                    help="zstack64 = BASELINE.json configs[1] (the default line); the others: bench_workloads.py")
    ap.add_argument("--seed", type=int, default=1002, help="seed of the synthetic stack (SURVEY 8d: 1002 for configs[1])")
    ap.add_argument("--seed-per-rank", action="store_true", help="rank r renders seed + r instead of the same stack: the step time then depends on the content each rank drew (about +-5 %), which the max over ranks turns into an apparent scaling loss")
    ap.add_argument("--no-graph", action="store_true", help="launch eagerly instead of replaying a captured CUDA graph")
    return ap.parse_args()


# ---------------------------------------------------------------------- CPU reference arm
_SAMPLE = None


def _cpu_one(i):
    from oracle import pipeline as opipe

    t = time.perf_counter()
    r = opipe.segment_slice(_SAMPLE[i])
    return time.perf_counter() - t, int(r["labels"].max())


def cpu_reference(size, n_sample, steps=1, warmup=0, procs_cap=None):
    """Times the oracle (the reference's scipy / scikit-image composition) on host cores.
    Returns (Mvoxel/s, cores used, description, ms per step)."""
    import multiprocessing as mp

    from particle_col_image_segmentation_b200 import synth

    global _SAMPLE
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, n_sample, procs_cap or cores))
    _SAMPLE = synth.zstack_u16(n_sample, size, size, seed=1002)
    ctx = mp.get_context("fork")  # workers inherit the sample; CUDA is not initialised yet
    with ctx.Pool(procs) as pool:
        for _ in range(warmup):
            pool.map(_cpu_one, range(n_sample), chunksize=1)
        t0 = time.perf_counter()
        for _ in range(steps):
            pool.map(_cpu_one, range(n_sample), chunksize=1)
        dt = (time.perf_counter() - t0) / steps
    _SAMPLE = None
    mvox = n_sample * size * size / dt / 1e6
    return mvox, procs, f"{n_sample} slices of the {size}x{size} uint16 stack (seed 1002), one process per slice on {procs} of {cores} host cores", dt * 1e3


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_sample = args.cpu_sample or max(1, min(cores, 32))
    mvox, procs, desc, ms = cpu_reference(args.size, n_sample, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": mvox,
        "unit": "Mvoxel/s",
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": ms,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "u16",
        "data": "synthetic",
        "config": {"workload": f"split_zstack + segment a synthetic {args.size}x{args.size}x{args.slices} uint16 z-stack (configs[1]); each step = a bounded sample of {n_sample} slices", "size": args.size, "slices": args.slices},
        "cpu_baseline": {"value": mvox, "unit": "Mvoxel/s", "cores": procs, "kind": "port", "sample": desc},
        "e2e": {"value": mvox, "unit": "Mvoxel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference = the repo's own scipy/scikit-image call sequence (oracle port; scikit-image calls restated, it is not installable offline); no C/C++ reference sources exist, so oracle/_ref does not apply",
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------- clocks
def bind_near_gpu(index):
    """Restrict this process to the CPUs NVML reports as local to the GPU, so that first-touch places the
    pinned staging buffers of the host-buffer path in that socket's memory.  Best effort."""
    try:
        import pynvml

        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(_nvml_handle(pynvml, index))
        return len(os.sched_getaffinity(0))
    except Exception:  # noqa: BLE001
        return None


def _nvml_handle(nv, cuda_index):
    """NVML handle of a CUDA device: by UUID, because CUDA_VISIBLE_DEVICES renumbers CUDA devices but not NVML's."""
    try:
        import torch

        uuid = str(torch.cuda.get_device_properties(cuda_index).uuid)
        return nv.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
    except Exception:  # noqa: BLE001
        return nv.nvmlDeviceGetHandleByIndex(cuda_index)


class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.ok, self.stop_flag = [], set(), False, False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = _nvml_handle(pynvml, index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:  # noqa: BLE001
            self.max_mhz = None

    def _loop(self):
        while not self.stop_flag:
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                try:
                    r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.002)  # a few hundred Hz (an NVML query takes about a millisecond itself), on rank 0 only

    def start(self):
        if self.ok:
            self.th = threading.Thread(target=self._loop, daemon=True)
            self.th.start()

    def stop(self):
        self.stop_flag = True
        if self.ok:
            self.th.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---------------------------------------------------------------------- CUDA arm
def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json, copy bandwidth)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def collect_profile(lib):
    names = ctypes.create_string_buffer(64 * 64)
    ms = (ctypes.c_double * 64)()
    cnt = (ctypes.c_int32 * 64)()
    n = lib.pcs_profile_collect(ctypes.cast(names, ctypes.c_void_p), ctypes.cast(ms, ctypes.c_void_p), ctypes.cast(cnt, ctypes.c_void_p), 64)
    out = {}
    for i in range(n):
        out[names.raw[64 * i : 64 * i + 64].split(b"\0")[0].decode()] = (ms[i], cnt[i])
    return out


def run_b200(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    Z, S = args.slices, args.size

    cpu = None
    if rank == 0 and args.gpus == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        n_sample = args.cpu_sample or max(1, min(cores, 32))
        mvox, procs, desc, _ = cpu_reference(S, n_sample)  # before CUDA is initialised (fork)
        mv1, _, d1, _ = cpu_reference(S, 2, procs_cap=1)      # how the reference itself runs: one thread
        cpu = {"value": mvox, "unit": "Mvoxel/s", "cores": procs, "kind": "port", "sample": desc,
               "single_thread": {"value": mv1, "unit": "Mvoxel/s", "cores": 1, "sample": d1}}

    import torch
    import torch.distributed as dist

    from oracle import pipeline as opipe
    from particle_col_image_segmentation_b200 import _lib, split_zstack, synth
    from particle_col_image_segmentation_b200 import dist as pdist

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_near_gpu(local)  # pinned host buffers land on the GPU's own NUMA node (matters when ranks share the host)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    # weak scaling: every rank owns a stack of identical shape AND content, so the work per GPU is exactly fixed as N
    # grows (the slices still count as distinct units: z0 = rank * Z)
    stack = synth.zstack_u16_device(Z, S, S, seed=args.seed + (rank if args.seed_per_rank else 0), device=dev)
    # the pipeline bound to this stack; eager launches for the profiled pass, a captured CUDA
    # graph (same kernels, same order) for the timed pass
    plan_eager = split_zstack.SegmentPlan(stack, chunk=args.chunk, z0=rank * Z)
    res = plan_eager.out
    plan = split_zstack.SegmentPlan(stack, chunk=args.chunk, z0=rank * Z, out=None, graph=True, streams=args.streams) if not args.no_graph else plan_eager

    # N > 1: one gather-to-root per step (header with the row counts + the rows of every chunk).  Two captured graphs that
    # share every buffer except the float64 table alternate: each finalises its rows straight into its own message buffer
    # of the gather, so the exchange of step n (side stream) overlaps the kernels of step n + 1 and the pipeline's stream
    # carries nothing for it but an event record.
    gatherer = pdist.TableGather()
    last_exchange = [None]
    duo, turn = [], [0]
    if world > 1 and not args.no_gather and not args.no_graph:
        r0 = plan()
        torch.cuda.synchronize()
        seen = torch.stack([off[-1].to(torch.int64) for off, _ in r0.table_padded()])
        dist.all_reduce(seen, op=dist.ReduceOp.MAX)
        stages = gatherer.make_staging(seen.cpu().tolist(), [int(ft.shape[0]) for _, ft in r0.table_padded()], 13, dev, n=2)
        duo = [(split_zstack.SegmentPlan(stack, chunk=args.chunk, z0=rank * Z, out=None, graph=True, streams=args.streams, staging=st), st) for st in stages]

    def step(p=None):
        if duo and p is None:
            pl, st = duo[turn[0] % 2]
            turn[0] += 1
            gatherer.wait_free(st)  # the exchange two steps back has read this message buffer
            r = pl()
            table = gatherer.exchange(st)
            last_exchange[0] = table
            return r, table
        r = (p or plan)()
        if world > 1 and not args.no_gather:  # eager / profiled pass: the copying form of the same exchange
            table = gatherer(r.table_padded())
            last_exchange[0] = table
        else:
            table = r.table_padded()  # finished float64 table in HBM (row count in offsets[-1]); no host sync
        return r, table

    # parity spot check against the oracle on one slice of this very stack (outside the timed region)
    parity = None
    if rank == 0:
        r = plan()  # local only: no collective may run on a single rank
        table = r.table_device()
        torch.cuda.synchronize()
        zi = Z // 2
        want = opipe.segment_slice(stack[zi].cpu().numpy(), z=zi)
        got_tab = table[table[:, 0] == zi].cpu().numpy()
        parity = bool(
            int(r.threshold[zi]) == want["threshold"]
            and np.array_equal(r.mask[zi].cpu().numpy().astype(bool), want["mask"])
            and np.array_equal(r.labels[zi].cpu().numpy(), want["labels"])
            and np.array_equal(r.refined[zi].cpu().numpy().astype(bool), want["refined"])
            and np.array_equal(r.edt[zi].cpu().numpy(), want["edt"])
            and np.array_equal(got_tab, want["table"])
        )
        if not parity:
            raise SystemExit("bench: CUDA pipeline differs from the oracle on the spot-check slice; refusing to time it")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(p, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            step(p)
        if last_exchange[0] is not None and last_exchange[0].ready is not None:
            torch.cuda.current_stream().wait_event(last_exchange[0].ready)  # the interval covers the last step's completed gather
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / steps

    sampler = ClockSampler(local)  # samples from the warm-up steps through the timed and the profiled pass (same load)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step()
    launches0 = lib.pcs_kernel_launches()
    ms_step = timed(None if duo else plan, args.steps)
    # kernels per step: counted on an eager pass (a graph replay re-issues the captured launches)
    for _ in range(1):
        l0 = lib.pcs_kernel_launches()
        step(plan_eager)
        launches = (lib.pcs_kernel_launches() - l0) * args.steps
    gather_check = None
    if world > 1 and not args.no_gather:
        # outside the timed region: the table the root gathered must equal, byte for byte, the concatenation of the
        # ranks' own compact tables (exchanged here the plain way: counts, then padded rows)
        r_chk, g_chk = step()
        got = g_chk.compact()
        want = pdist.gather_tables(r_chk.table_device())
        if rank == 0:
            gather_check = bool(got.shape == want.shape and torch.equal(got.view(torch.int64), want.view(torch.int64)))
            if not gather_check:
                raise SystemExit("bench: gathered table differs from the concatenation of the per-rank tables")
    voxels = float(Z) * S * S * world
    value = voxels / (ms_step * 1e-3) / 1e6

    # per-kernel CUDA-event timing: a second timed pass of the same K steps, launched eagerly with an
    # event pair around every kernel on the launching stream (the events add a few percent to the step)
    prof, ms_step_profiled = {}, None
    if not args.no_profile:
        lib.pcs_profile_enable(1)
        ms_step_profiled = timed(plan_eager, args.steps)
        lib.pcs_profile_enable(0)
        prof = collect_profile(lib)
    clocks = sampler.stop() if rank == 0 else None
    del launches0

    # end to end through the public API with host buffers (pinned), copies inside the timed region
    e2e = None
    e2e_ok = not args.no_e2e
    if e2e_ok:
        # 4.3 GB of pinned host memory per rank: if any rank cannot get it, every rank skips the e2e leg
        # (a lone failing rank would leave the others waiting in the barrier)
        try:
            host_in = torch.empty((Z, S, S), dtype=torch.uint16).pin_memory()
            host_in.copy_(stack)
            host_out = split_zstack.alloc_host_outputs(Z, S, S)
        except RuntimeError as e:  # noqa: BLE001
            sys.stderr.write(f"bench: rank {rank}: no pinned host buffers ({str(e)[:120]}); e2e skipped\n")
            e2e_ok = False
        if world > 1:
            flag = torch.tensor([1 if e2e_ok else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            e2e_ok = bool(flag.item())
    if e2e_ok:
        split_zstack.segment_zstack_pinned(host_in, host_out, chunk=args.e2e_chunk)  # warm-up
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            nrows = split_zstack.segment_zstack_pinned(host_in, host_out, chunk=args.e2e_chunk)
        barrier()
        dt = (time.perf_counter() - t0) / args.e2e_steps
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        d2h = sum(v.numel() * v.element_size() for k, v in host_out.items() if k != "table") + nrows * 13 * 8
        e2e = {"value": voxels / float(tt.item()) / 1e6, "unit": "Mvoxel/s", "h2d_bytes_per_step": host_in.numel() * 2, "d2h_bytes_per_step": int(d2h), "ms_per_step": float(tt.item()) * 1e3, "chunk": args.e2e_chunk, "overlap": "H2D, kernels and D2H of consecutive chunks on three streams", "cpus_bound_near_gpu": numa,
               "outputs": "mask, labels, refined, edt + table (everything the pipeline produces: 14 B/voxel back over PCIe, 8 of them the float64 EDT)"}
        # the same call for a caller that needs the region table and the label image only: 4 B/voxel back
        sel = ("labels",)
        host_sel = {k: host_out[k] for k in sel + ("threshold", "counts")}
        split_zstack.segment_zstack_pinned(host_in, host_sel, chunk=args.e2e_chunk, outputs=sel)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            nrows = split_zstack.segment_zstack_pinned(host_in, host_sel, chunk=args.e2e_chunk, outputs=sel)
        barrier()
        ts = torch.tensor([(time.perf_counter() - t0) / args.e2e_steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        e2e["table_and_labels"] = {"value": voxels / float(ts.item()) / 1e6, "unit": "Mvoxel/s", "ms_per_step": float(ts.item()) * 1e3, "h2d_bytes_per_step": host_in.numel() * 2,
                                   "d2h_bytes_per_step": int(host_sel["labels"].numel() * 4 + 2 * Z * 4 + nrows * 13 * 8)}
        if world > 1:
            e2e["note"] = f"{world} ranks share the host: the aggregate pinned traffic ({world} x {int(d2h + host_in.numel() * 2) >> 20} MiB per step) is bounded by the host's memory / PCIe root complexes, not by the GPUs"

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    roof = None
    shares = {}
    if prof:
        tot = sum(v[0] for v in prof.values())
        shares = {k: round(v[0] / tot, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])}
        name, (kms, kcnt) = max(prof.items(), key=lambda kv: kv[1][0])
        per_launch_vox = float(min(args.chunk, Z)) * S * S  # every launch covers one chunk of slices
        bpv = KERNEL_BYTES_PER_VOXEL.get(name, PIPELINE_BYTES_PER_VOXEL)
        avg_ms = kms / kcnt
        achieved = bpv * per_launch_vox / (avg_ms * 1e-3) / 1e9
        tpv = NCU_TRAFFIC_BYTES_PER_VOXEL.get(name)
        roof = {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": (tpv * per_launch_vox if tpv else None), "traffic_source": f"ncu --set full, {NCU_TRAFFIC_SOURCE} (per voxel, scaled to this launch)" if tpv else None,
                "write_only_gbs_live": write_only_probe(dev, lib),
                "bytes_per_voxel": bpv, "avg_launch_ms": avg_ms, "launches": kcnt, "peak_source": peak_src, "kernel_time_share": shares,
                "kernel_ms_per_step": {k: round(v[0] / args.steps, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])},
                "timed": "per-kernel CUDA events on the launching stream over a second pass of the same K steps (eager launches, one stream): the pairs inflate the step, shares are exact, absolute per-kernel ms are high by about event_inflation",
                "ms_per_step_with_events": ms_step_profiled, "event_inflation": (sum(v[0] for v in prof.values()) / args.steps) / ms_step}
    per_gpu = value / world * 1e6
    line = {
        "metric": METRIC,
        "value": value,
        "unit": "Mvoxel/s",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": ms_step,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "u16",
        "data": "synthetic",
        "config": {"workload": f"split_zstack + segment a synthetic {S}x{S}x{Z} uint16 z-stack per GPU (BASELINE.json configs[1])", "size": S, "slices_per_gpu": Z, "chunk": args.chunk, "streams": args.streams, "launch": "eager" if args.no_graph else "cuda graph replay (one graph per step)",
                   "l2": "input stack (%.0f MiB) and outputs are larger than L2; no flush needed" % (Z * S * S * 2 / 2**20), "parity_spot_check": parity,
                   "table_gather": None if world == 1 else {"collectives_per_step": 1, "kind": "gather to rank 0 (NCCL send/recv group), overlapped with the next step", "byte_equal_to_per_rank_tables": gather_check}},
        "clocks": clocks,
        "e2e": e2e,
        "gpu_launches": int(launches),
        "roofline": roof,
        "pipeline_roofline": {"bytes_per_voxel": PIPELINE_BYTES_PER_VOXEL, "achieved_gbs_per_gpu": per_gpu * PIPELINE_BYTES_PER_VOXEL / 1e9, "frac_of_peak": per_gpu * PIPELINE_BYTES_PER_VOXEL / 1e9 / peak,
                              "sum_of_stages_frac": per_gpu * SUM_OF_STAGES_BYTES_PER_VOXEL / 1e9 / peak,
                              "dram_traffic_bytes_per_voxel_ncu": NCU_TRAFFIC_WHOLE_STEP_BYTES_PER_VOXEL, "traffic_source": NCU_TRAFFIC_SOURCE},
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.workload == "zstack256":
        args.slices = 256  # the north_star's own target; same pipeline, same line
    elif args.workload != "zstack64" and args.impl != "reference":
        if int(os.environ.get("RANK", "0")) == 0:
            import bench_workloads

            bench_workloads.run(args, sys.modules[__name__])
        return
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
