#!/usr/bin/env python3
"""Extract known-answer pixels from the reference's own committed render
(/root/reference/output.png, 1140x950, crate-default features + simd_render) into
tests/golden/output_png_kats.json.  Run in the build container (the GPU box has no /root/reference).

`smooth` pixels lie in penumbra-free, silhouette-free regions, where the render does not depend on
the (OS-entropy seeded, hence irreproducible) Poisson samplers: a centre-sample, M = 1 evaluation
of the algorithm must reproduce them to the LSB.  `stochastic` pixels are
compared with a tolerance under the full default+simd_render feature set with our seeded sampler.
`mask_grid` is the non-black mask on a 12-pixel grid; `dense_grid` the RGB values on a 25-pixel grid
(1748 pixels) for a distributional comparison.
"""
import json
from pathlib import Path

import numpy as np
from PIL import Image

img = np.asarray(Image.open("/root/reference/output.png").convert("RGB")).astype(int)
H, W, _ = img.shape
smooth = [(60, 100), (150, 150), (400, 120), (640, 150), (100, 600), (900, 300), (1000, 500), (420, 690), (175, 620), (480, 300),
          (60, 400), (570, 745), (300, 800)]
stochastic = [(560, 350), (640, 600), (700, 820), (300, 560), (800, 700), (36, 70), (35, 68)]
ys, xs = np.mgrid[6:H:12, 6:W:12]
gy, gx = np.mgrid[12:H:25, 12:W:25]  # dense grid for the distributional comparison under the full default feature set
out = {
    "source": "/root/reference/output.png", "width": W, "height": H,
    "non_black_fraction": float((img.sum(-1) > 0).mean()),
    "smooth": [{"x": x, "y": y, "rgb": img[y, x].tolist()} for x, y in smooth],
    "stochastic": [{"x": x, "y": y, "rgb": img[y, x].tolist()} for x, y in stochastic],
    "mask_grid": {"x0": 6, "y0": 6, "step": 12, "mask": (img[ys, xs].sum(-1) > 0).astype(int).tolist()},
    "dense_grid": {"x0": 12, "y0": 12, "step": 25, "rgb": img[gy, gx].reshape(-1, 3).tolist()},
}
Path(__file__).with_name("output_png_kats.json").write_text(json.dumps(out))
print("wrote", len(smooth), "+", len(stochastic), "pixels, grid", ys.shape)
