#!/usr/bin/env python
"""Copy the reference's demo INPUTS for BASELINE configs[0] into tests/golden/ as arrays (the GPU box has no
/root/reference): the Middlebury Urban3 frame pair runme.m:74 feeds to FlowEminHS_elin_2D_v10, and yosemite.mat
(frame pair + ground-truth flow Utrue/Vtrue) runme.m:90 feeds to FlowEminNDFASFMG_elin_2D_v10.

    python tests/golden/make_fixtures.py            # run where /root/reference exists

These are data fixtures, not source. Stored losslessly (uint8 frames; the ground truth as float32, which holds the
file's values to 1e-7 relative)."""
import os

import numpy as np
import scipy.io as sio
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/images/middlebury"

f7 = np.asarray(Image.open(os.path.join(REF, "Urban3_frame07.png")).convert("RGB"))
f8 = np.asarray(Image.open(os.path.join(REF, "Urban3_frame08.png")).convert("RGB"))
assert f7.shape == (480, 640, 3) and f7.dtype == np.uint8
np.savez_compressed(os.path.join(HERE, "urban3_pair.npz"), frame07=f7, frame08=f8,
                    source="images/middlebury/Urban3_frame07.png, Urban3_frame08.png (runme.m:74)")
y = sio.loadmat(os.path.join(REF, "yosemite.mat"))
assert y["I"].shape == (252, 316, 2) and y["I"].dtype == np.uint8
np.savez_compressed(os.path.join(HERE, "yosemite.npz"), I=y["I"], Utrue=y["Utrue"].astype(np.float32), Vtrue=y["Vtrue"].astype(np.float32),
                    source="images/middlebury/yosemite.mat (runme.m:88-90)")
for n in ("urban3_pair.npz", "yosemite.npz"):
    print(n, os.path.getsize(os.path.join(HERE, n)), "bytes")
