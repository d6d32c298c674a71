import torch
def t(f, n=10):
    f(); torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    best=1e9
    for _ in range(n):
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); best=min(best,e0.elapsed_time(e1))
    return best
x=torch.empty(2*1024**3, dtype=torch.uint8, device='cuda'); y=torch.empty_like(x)
ms=t(lambda: x.zero_()); print("zero_ 2GiB", ms, "ms", x.numel()/ms/1e6, "GB/s")
ms=t(lambda: torch.cuda.memset if False else x.fill_(3)); print("fill_ 2GiB", ms, "ms", x.numel()/ms/1e6, "GB/s")
ms=t(lambda: y.copy_(x)); print("copy 2GiB", ms, "ms", 2*x.numel()/ms/1e6, "GB/s (r+w)")
xs=x.view(torch.float64)
ms=t(lambda: xs.sum()); print("read(sum f64) 2GiB", ms, "ms", x.numel()/ms/1e6, "GB/s")
