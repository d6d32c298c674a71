import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from particle_col_image_segmentation_b200 import ops, synth
dev=torch.device("cuda:0")
stack=synth.zstack_u16_device(64,2048,2048,1002,dev)
ref=None
for it in range(3):
    thr,hist=ops.otsu_u16(stack, return_hist=True)
torch.cuda.synchronize()
e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): ops.otsu_u16(stack)
e1.record(); torch.cuda.synchronize()
want=torch.stack([torch.bincount(stack[i].to(torch.int32).flatten(), minlength=65536) for i in (0,63)])
ok=torch.equal(hist[[0,63]].to(torch.int64), want)
print("hist+otsu ms per 64 slices:", e0.elapsed_time(e1)/10, "exact:", ok)
