"""Seeded inputs of tests/golden/input_ref.npz (regenerated on both sides instead of stored): decoded uint8 HWC images of
assorted sizes and ground-truth strings.  Must stay in sync with oracle/make_golden.py::input_cases."""
import numpy as np


def input_cases():
    rng = np.random.RandomState(31)
    sizes = [(37, 91), (64, 512), (20, 15), (32, 128), (11, 300), (48, 33), (5, 7), (33, 1)]
    images = [rng.randint(0, 256, size=(h, w, 3)).astype(np.uint8) for h, w in sizes]
    texts = ["hello", "MegReader2019", "a", "x" * 40, "B200 sm_100a!", "0123456789", "Zz", "ctc"]
    return images, texts


MODES = {"resize": (32, 128), "pad": (32, 160)}
# cv2.resize (third-party, OpenCV 4.13 with IPP) differs from the published-algorithm restatement by up to ~3e-3 on the 0..255
# scale for non-integer ratios (measured: 2.7e-3 worst over these cases); after the /255 of NormalizeImage that is ~1.1e-5.
# Tolerance used against the reference goldens (absolute, on the normalised [-0.5, 0.6] scale):
TOL = 2e-5
