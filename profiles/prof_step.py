#!/usr/bin/env python3
"""One eager step of the segment pipeline over a 16-slice 2048x2048 chunk, for Nsight Compute:

    ncu --set full --clock-control none --import-source on --profile-from-start off \
        -o gpurun_out/prof -f python profiles/prof_step.py

The first (warm-up) step runs outside the profiled range; the second one is captured."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from particle_col_image_segmentation_b200 import split_zstack, synth  # noqa: E402

Z = int(os.environ.get("PROF_SLICES", "16"))
stack = synth.zstack_u16_device(Z, 2048, 2048, 1002, torch.device("cuda:0"))
plan = split_zstack.SegmentPlan(stack, chunk=Z)
plan()
torch.cuda.synchronize()
torch.cuda.profiler.start()
plan()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one step of", Z, "slices")
