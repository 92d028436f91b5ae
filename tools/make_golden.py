#!/usr/bin/env python
"""Generate tests/golden/ fixtures from the reference's own data files.

Runs ONLY in the build container (needs /root/reference); the fixtures it writes are
committed so that the GPU box (which has no /root/reference) can use them.

blurxy: utils/images/rgb.png -> utils/images/reference_blurxy.png is the reference's
own golden pair for blurxy_tiramisu (benchmarks/halide/blurxy_tiramisu.cpp; loaded by
tests/wrapper_test_14.cpp:11-12).  The blur is local, so an input crop
[y0:y0+h+8, x0:x0+w+8] maps to the output crop [y0:y0+h, x0:x0+w]; we commit two crops
(one at the image origin, one interior, ragged sizes) in the generated function's
layout u8[C][H][W].
"""
import os
import sys

import numpy as np
from PIL import Image

REF = "/root/reference/utils/images"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def main():
    os.makedirs(OUT, exist_ok=True)
    rgb = np.array(Image.open(os.path.join(REF, "rgb.png")))            # H, W, C
    ref = np.array(Image.open(os.path.join(REF, "reference_blurxy.png")))
    assert rgb.shape == (3520, 2112, 3) and ref.shape == (3512, 2104, 3)
    crops = {"a": (0, 0, 96, 160), "b": (1777, 1031, 131, 203)}          # y0, x0, h, w
    for name, (y0, x0, h, w) in crops.items():
        i = np.ascontiguousarray(rgb[y0:y0 + h + 8, x0:x0 + w + 8].transpose(2, 0, 1))
        o = np.ascontiguousarray(ref[y0:y0 + h, x0:x0 + w].transpose(2, 0, 1))
        np.save(os.path.join(OUT, f"blurxy_{name}_in.npy"), i)
        np.save(os.path.join(OUT, f"blurxy_{name}_out.npy"), o)
        print(name, i.shape, o.shape)
    # whole-image digest so the full-size oracle run can be checked in this container
    import hashlib
    with open(os.path.join(OUT, "blurxy_full_sha256.txt"), "w") as f:
        f.write(hashlib.sha256(np.ascontiguousarray(ref.transpose(2, 0, 1)).tobytes()).hexdigest() + "\n")


if __name__ == "__main__":
    sys.exit(main())
