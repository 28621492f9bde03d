#!/usr/bin/env python
"""Regenerate bbm_b200/data/epd_g1.f32 from the reference's precomputed Holzschuch-Pacanowski G1 table.

The EPD model is *defined* through this table: `ndf::epd::G1` is `G1.interpolate(p, tanTheta*beta)`
(include/ndf/epd.h:142-152) over 100 x 1000 floats printed by the reference's own generator
(precompute/HolzschuchPacanowski/G1.cpp) into include/precomputed/holzschuchpacanowski/G1.h.
They are model constants (numbers, like the MERL measurements), not code; this script parses the numbers
out of the header where it lies under /root/reference and stores them as raw little-endian float32,
row-major [p index 0..99][tan index 0..999], p index = 5/p - 1, tan index = exp(-(ln 1/t)^0.05)*1000 - 1.
Run only in the build container (the GPU box has no /root/reference; it uses the committed asset).
"""
import os
import re
import sys

import numpy as np

REF = os.environ.get("BBM_REF", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "bbm_b200", "data", "epd_g1.f32")


def main():
    src = open(os.path.join(REF, "include/precomputed/holzschuchpacanowski/G1.h")).read()
    body = src[src.index("G1 = {") + len("G1 = {"):]
    body = body[:body.index("};")]
    body = re.sub(r"//[^\n]*", "", body)
    vals = np.array([float(t) for t in body.replace("\n", " ").split(",") if t.strip()], np.float32)
    assert vals.size == 100 * 1000, vals.size
    vals.tofile(OUT)
    print(f"wrote {OUT}: {vals.size} floats, min {vals.min()} max {vals.max()}")


if __name__ == "__main__":
    sys.exit(main())
