import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from particle_col_image_segmentation_b200 import split_zstack, synth, dist as pdist
dev=torch.device("cuda:0")
stack=synth.zstack_u16_device(64,2048,2048,1002,dev)
plan=split_zstack.SegmentPlan(stack, chunk=32, graph=True, streams=2)
g=pdist.TableGather()
def run(mode, n=10):
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    t0=time.perf_counter(); e0.record()
    for _ in range(n):
        r=plan()
        if mode>=1:
            pads=r.table_padded()
        if mode==2:
            x=pads[0][0][-1:].to(torch.int64)
        if mode==3:
            t=g(pads[0][1], pads[0][0])
    e1.record(); t1=time.perf_counter(); torch.cuda.synchronize()
    print(mode, "gpu ms/step", e0.elapsed_time(e1)/n, "cpu enqueue ms/step", (t1-t0)/n*1e3)
for m in (0,0,1,2,3,3,0): run(m)
