"""Round-2 probe D: a small persistent zero-fill grid on a side stream beside the histogram / the whole chunk."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from particle_col_image_segmentation_b200 import _lib, ops, split_zstack, synth
dev = torch.device("cuda:0"); lib = _lib.load(); P = ops._p
Z, S0 = 32, 2048
stack = synth.zstack_u16_device(Z, S0, S0, 1002, dev)
hist = torch.empty((Z, 65536), dtype=torch.int32, device=dev)
big = torch.empty(Z * S0 * S0 * 12, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def ev(): return torch.cuda.Event(enable_timing=True)
def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); a, b = ev(), ev(); a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n
def hist_only(): lib.pcs_histogram_u16(P(stack), P(hist), Z, S0, S0, torch.cuda.current_stream().cuda_stream)
def both(a, b):
    main = torch.cuda.current_stream(); s1.wait_stream(main); s2.wait_stream(main)
    with torch.cuda.stream(s2): b()      # the fill first: all its CTAs become resident
    with torch.cuda.stream(s1): a()
    main.wait_stream(s1); main.wait_stream(s2)
plan = split_zstack.SegmentPlan(stack, chunk=32)
th, tp = timeit(hist_only), timeit(plan)
print(f"hist(32) {th:.4f}   pipeline(32) {tp:.4f}")
for cps in (1, 2, 4, 8):
    fill = lambda: lib.pcs_zero_background(big.data_ptr(), big.numel(), cps, torch.cuda.current_stream().cuda_stream)
    tf = timeit(fill)
    print(f"fill {cps} CTA/SM alone {tf:.4f} ({big.numel()/tf/1e6:.0f} GB/s) | beside hist {timeit(lambda: both(hist_only, fill)):.4f} (sum {th+tf:.4f}) | beside pipeline {timeit(lambda: both(plan, fill)):.4f} (sum {tp+tf:.4f})")
