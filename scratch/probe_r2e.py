"""Round-2 probe E: what bounds k_edt_near -- its stores or its instructions?  Times pcs_edt_bits (transpose + carry +
near + far) on the refined mask of a 32-slice chunk of the bench stack for the library selected by PCS_LIB_PATH
(build variants: -DEDT_STORE_ONLY = every tile leaves after its zero fill, -DEDT_NO_EARLY_EXIT, -DEDT_ER=8), next to
a cudaMemset of the same 1.07 GB."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from particle_col_image_segmentation_b200 import _lib, ops, split_zstack, synth
dev = torch.device("cuda:0"); lib = _lib.load(); P = ops._p
rt = ctypes.CDLL("libcudart.so.12")
rt.cudaMemsetAsync.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p]
Z, S0 = 32, 2048
stack = synth.zstack_u16_device(Z, S0, S0, 1002, dev)
plan = split_zstack.SegmentPlan(stack, chunk=32); plan(); torch.cuda.synchronize()
refined = plan.out.refined
bits = ops.pack(refined)
def ev(): return torch.cuda.Event(enable_timing=True)
def timeit(fn, n=20, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); a, b = ev(), ev(); a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n
dist = torch.empty((Z, S0, S0), dtype=torch.float64, device=dev)
n = lib.pcs_edt_workspace_bytes(Z, S0, S0); ws = torch.empty(n, dtype=torch.uint8, device=dev)
def edt(): _lib.call("pcs_edt_bits", P(bits), 0, Z, S0, S0, P(dist), None, None, 0, P(ws), n, torch.cuda.current_stream().cuda_stream)
def memset(): rt.cudaMemsetAsync(dist.data_ptr(), 0, dist.numel() * 8, torch.cuda.current_stream().cuda_stream)
te, tm = timeit(edt), timeit(memset)
print(f"{os.environ.get('PCS_LIB_PATH', 'default'):40s} edt(32 slices) {te:.4f} ms   memset {tm:.4f} ms ({dist.numel()*8/tm/1e6:.0f} GB/s)  fg {float(refined.float().mean()):.4f}")
