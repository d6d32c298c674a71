"""Round-2 probe B: is k_hist_u16 bound by shared-memory bank conflicts?  Same kernel, three inputs:
real stack; conflict-free values (lane l of a warp always hits bank l); all lanes same bank different word."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from particle_col_image_segmentation_b200 import _lib, ops, synth
dev = torch.device("cuda:0"); lib = _lib.load(); P = ops._p
Z, S0 = 32, 2048
real = synth.zstack_u16_device(Z, S0, S0, 1002, dev)
n = Z * S0 * S0
i = torch.arange(n, device=dev, dtype=torch.int64)
lane = (i >> 3) & 31
g = torch.Generator(device=dev); g.manual_seed(1)
r = torch.randint(0, 8, (n,), device=dev, generator=g)
free = (2 * lane + 64 * r).to(torch.int32).to(torch.uint16).view(Z, S0, S0)          # bank = lane: conflict-free, 256 distinct words
r2 = torch.randint(0, 256, (n,), device=dev, generator=g)
same_bank = (64 * r2).to(torch.int32).to(torch.uint16).view(Z, S0, S0)                # every lane hits bank 0, 256 distinct words
uni = torch.randint(300, 700, (n,), device=dev, generator=g).to(torch.int32).to(torch.uint16).view(Z, S0, S0)  # uniform over 400 values
hist = torch.empty((Z, 65536), dtype=torch.int32, device=dev)
def run(x):
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3): lib.pcs_histogram_u16(P(x), P(hist), Z, S0, S0, st)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): lib.pcs_histogram_u16(P(x), P(hist), Z, S0, S0, st)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / 10
for name, x in (("real", real), ("conflict-free", free), ("same-bank", same_bank), ("uniform 400 values", uni)):
    print(f"{name:20s} {run(x):.4f} ms per {Z} slices")
