"""Round-2 probe A (run on a B200): two questions that decide the pipeline layout.

1. L2 residency: is `k_compare` faster when it runs right after `k_hist_u16` on a sub-batch small enough
   to stay in L2?  hist -> otsu -> compare over sub-batches of S slices, total per 64 slices.
2. Stream overlap: does the atomics-bound histogram of one chunk overlap the store-bound kernels of another
   when they are issued on two streams?  hist(32 slices) beside a 1 GiB fill, and beside the EDT.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from particle_col_image_segmentation_b200 import _lib, ops, synth  # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.load()
Z, S0 = 64, 2048
stack = synth.zstack_u16_device(Z, S0, S0, 1002, dev)
hist = torch.empty((Z, 65536), dtype=torch.int32, device=dev)
thr = torch.empty(Z, dtype=torch.int32, device=dev)
bits = ops.new_bits(Z, S0, S0, dev)
P = ops._p


def ev():
    return torch.cuda.Event(enable_timing=True)


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = ev(), ev()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def sub(S, do_hist=True, do_otsu=True, do_cmp=True):
    st = torch.cuda.current_stream().cuda_stream
    for a in range(0, Z, S):
        b = min(Z, a + S)
        B = b - a
        if do_hist:
            lib.pcs_histogram_u16(P(stack[a:b]), P(hist[a:b]), B, S0, S0, st)
        if do_otsu:
            lib.pcs_otsu_u16(P(hist[a:b]), P(thr[a:b]), 0, B, S0 * S0, st)
        if do_cmp:
            lib.pcs_compare_u16(P(stack[a:b]), 0, P(thr[a:b]), 0, P(bits[a:b]), 0, B, S0, S0, st)


print("== 1. hist -> otsu -> compare over sub-batches of S slices (ms per 64 slices)")
for S in ((64, 32, 16, 8, 4, 2, 1) if "--part1" in sys.argv else (64, 32)):
    t_all = timeit(lambda: sub(S))
    t_h = timeit(lambda: sub(S, True, False, False))
    t_ho = timeit(lambda: sub(S, True, True, False))
    t_c = timeit(lambda: sub(S, False, False, True))
    print(f"S={S:3d}  all {t_all:.4f}  hist {t_h:.4f}  hist+otsu {t_ho:.4f}  compare alone {t_c:.4f}  -> compare after hist costs {t_all - t_ho:.4f}")

# same with a CUDA graph (launch gaps out of the picture)
for S in ((8, 4) if "--part1" in sys.argv else ()):
    g = torch.cuda.CUDAGraph()
    sub(S)
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        sub(S)
    print(f"graph S={S}: {timeit(g.replay):.4f} ms per 64 slices")

print("== 2. overlap on two streams")
big = torch.empty(1 << 28, dtype=torch.int32, device=dev)  # 1 GiB
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
half = 32


def hist_only():
    lib.pcs_histogram_u16(P(stack[:half]), P(hist[:half]), half, S0, S0, torch.cuda.current_stream().cuda_stream)


def fill_only():
    lib.pcs_fill_u32(P(big), 1, big.numel(), torch.cuda.current_stream().cuda_stream)


def both(first_hist=True):
    main = torch.cuda.current_stream()
    s1.wait_stream(main)
    s2.wait_stream(main)
    order = [(s1, hist_only), (s2, fill_only)] if first_hist else [(s2, fill_only), (s1, hist_only)]
    for s, f in order:
        with torch.cuda.stream(s):
            f()
    main.wait_stream(s1)
    main.wait_stream(s2)


th, tf = timeit(hist_only), timeit(fill_only)
print(f"hist(32) {th:.4f}  fill 1GiB {tf:.4f}  sum {th + tf:.4f}  two streams (hist first) {timeit(lambda: both(True)):.4f}  (fill first) {timeit(lambda: both(False)):.4f}")

# hist beside the EDT of another chunk
ref_bits = ops.new_bits(half, S0, S0, dev)
lib.pcs_compare_u16(P(stack[half:]), 5000, 0, 0, P(ref_bits), 0, half, S0, S0, torch.cuda.current_stream().cuda_stream)
edt_out = torch.empty((half, S0, S0), dtype=torch.float64, device=dev)
nws = lib.pcs_edt_workspace_bytes(half, S0, S0)
ews = torch.empty(nws, dtype=torch.uint8, device=dev)


def edt_only():
    lib.pcs_edt_bits(P(ref_bits), 0, half, S0, S0, P(edt_out), 0, 0, 0, P(ews), nws, torch.cuda.current_stream().cuda_stream)


def both_edt(first_hist=True):
    main = torch.cuda.current_stream()
    s1.wait_stream(main)
    s2.wait_stream(main)
    order = [(s1, hist_only), (s2, edt_only)] if first_hist else [(s2, edt_only), (s1, hist_only)]
    for s, f in order:
        with torch.cuda.stream(s):
            f()
    main.wait_stream(s1)
    main.wait_stream(s2)


te = timeit(edt_only)
print(f"hist(32) {th:.4f}  edt(32) {te:.4f}  sum {th + te:.4f}  two streams (hist first) {timeit(lambda: both_edt(True)):.4f}  (edt first) {timeit(lambda: both_edt(False)):.4f}")
