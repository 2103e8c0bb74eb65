"""Seeded inputs for deformable PS-RoI pooling (shared by the CPU oracle tests and the GPU parity tests)."""
import numpy as np

# name: B, output_dim, group, H, W, n_rois, pooled, part, samples, spatial_scale, trans_std, no_trans, num_classes
CASES = {
    "plain": (2, 3, 1, 12, 16, 5, 3, 3, 2, 0.25, 0.0, True, 1),
    "trans": (2, 4, 1, 14, 18, 6, 3, 3, 4, 0.5, 0.1, False, 1),
    "groups": (1, 2, 3, 20, 20, 4, 3, 3, 2, 1.0, 0.2, False, 2),
    "parts": (2, 4, 2, 16, 16, 7, 4, 2, 3, 0.5, 0.3, False, 4),
    "outside": (1, 2, 1, 8, 8, 5, 2, 2, 4, 1.0, 0.5, False, 1),
}


def make(name, dtype=np.float64):
    B, od, g, H, W, n, P, part, sp, scale, tstd, no_trans, ncls = CASES[name]
    rng = np.random.RandomState(abs(hash(name)) % (2 ** 31) if False else sum(map(ord, name)))
    C = od * g * g
    data = rng.standard_normal((B, C, H, W)).astype(dtype)
    rois = np.zeros((n, 5), dtype)
    img_w, img_h = W / scale, H / scale
    for i in range(n):
        x1, y1 = rng.uniform(-0.1 * img_w, 0.7 * img_w), rng.uniform(-0.1 * img_h, 0.7 * img_h)
        x2, y2 = x1 + rng.uniform(0.05 * img_w, 0.6 * img_w), y1 + rng.uniform(0.05 * img_h, 0.6 * img_h)
        rois[i] = (rng.randint(B), x1 + 0.3, y1 + 0.3, x2 + 0.3, y2 + 0.3)       # keep round() away from .5 ties
    if name == "outside":
        rois[0, 1:] = (-40.3, -40.3, -20.3, -20.3)                                 # entirely outside: count == 0 everywhere
        rois[1, 1:] = (3.3, 3.3, 3.3, 3.3)                                         # degenerate roi: width clamps
    trans = None if no_trans else rng.uniform(-1, 1, (n, 2 * ncls, part, part)).astype(dtype)
    args = dict(no_trans=no_trans, spatial_scale=scale, output_dim=od, group_size=g, pooled=P, part_size=part,
                sample_per_part=sp, trans_std=tstd)
    return data, rois, trans, args
