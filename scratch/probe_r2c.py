"""Round-2 probe C: can the dense zero stores of the label / EDT outputs hide under the kernels that leave DRAM idle?
cudaMemsetAsync (driver memset: copy engine or kernel?) on a second stream beside (a) the histogram, (b) the whole
labelling stage of a chunk."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from particle_col_image_segmentation_b200 import _lib, ops, split_zstack, synth
dev = torch.device("cuda:0"); lib = _lib.load(); P = ops._p
rt = ctypes.CDLL("libcudart.so.12")
rt.cudaMemsetAsync.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p]
Z, S0 = 32, 2048
stack = synth.zstack_u16_device(Z, S0, S0, 1002, dev)
hist = torch.empty((Z, 65536), dtype=torch.int32, device=dev)
big = torch.empty(Z * S0 * S0 * 12, dtype=torch.uint8, device=dev)  # labels (4 B) + EDT (8 B) of the chunk: 1.6 GB
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def ev(): return torch.cuda.Event(enable_timing=True)
def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); a, b = ev(), ev(); a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n
def hist_only(): lib.pcs_histogram_u16(P(stack), P(hist), Z, S0, S0, torch.cuda.current_stream().cuda_stream)
def memset_only(): rt.cudaMemsetAsync(big.data_ptr(), 0, big.numel(), torch.cuda.current_stream().cuda_stream)
def both(a, b, first=0):
    main = torch.cuda.current_stream(); s1.wait_stream(main); s2.wait_stream(main)
    order = [(s1, a), (s2, b)] if first == 0 else [(s2, b), (s1, a)]
    for s, f in order:
        with torch.cuda.stream(s): f()
    main.wait_stream(s1); main.wait_stream(s2)
th, tm = timeit(hist_only), timeit(memset_only)
print(f"hist(32) {th:.4f}  memset 1.6 GB {tm:.4f} ({big.numel()/tm/1e6:.0f} GB/s)  sum {th+tm:.4f}  two streams: hist first {timeit(lambda: both(hist_only, memset_only)):.4f}  memset first {timeit(lambda: both(hist_only, memset_only, 1)):.4f}")
# the whole pipeline of a 32-slice chunk beside the memset
plan = split_zstack.SegmentPlan(stack, chunk=32)
tp = timeit(plan)
print(f"pipeline(32) {tp:.4f}  + memset sequential {tp+tm:.4f}  two streams: pipeline first {timeit(lambda: both(plan, memset_only)):.4f}  memset first {timeit(lambda: both(plan, memset_only, 1)):.4f}")
