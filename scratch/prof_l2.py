"""cProfile of the tiff_analysis single-file path on the seed-1234 2048^2 class image (host share of the call)."""
import cProfile, pstats, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from particle_col_image_segmentation_b200 import tiff_analysis, synth
img = synth.class_image(2048, 2048, seed=1234)
t = torch.from_numpy(img).cuda()
ct = {1: "Cells", 2: "Particle", 3: "Background"}
for _ in range(3):
    tiff_analysis.process_single_array(t, ct)
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    tiff_analysis.process_single_array(t, ct)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(30)
