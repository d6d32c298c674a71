import sys; sys.path.insert(0, '/root/repo')
import numpy as np, torch
from oracle import nanosims as on
from particle_col_image_segmentation_b200 import synth, nanosims, ops, _io
planes, roi, set_id, agg = synth.nanosims_stack(256, 5, 120, seed=1004)
want = on.boundary_pixels(agg)
got = nanosims.boundary_pixels(agg).cpu().numpy()
print('bd', want.shape, got.shape, np.array_equal(want, got))
red = np.isin(roi, np.nonzero(set_id == 1)[0] + 1)
lab, n = on.matlab_label(red); xy = on.roi_centroids_xy(lab, n)
w = on.min_dist_to_points(xy, want)
a_d = torch.from_numpy(xy).cuda(); bd = torch.from_numpy(want).cuda()
g = ops.min_dist(a_d, bd).cpu().numpy()
print('mindist eq', np.array_equal(w, g), np.abs(w-g).max())
g2 = ops.min_dist(a_d, nanosims.boundary_pixels(agg)).cpu().numpy()
print('mindist2 eq', np.array_equal(w, g2), np.abs(w-g2).max())
bp = nanosims.boundary_pixels(agg); print(bp.dtype, bp.is_contiguous(), bp.stride(), bp.shape)
