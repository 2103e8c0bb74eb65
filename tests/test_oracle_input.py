"""CPU: the input-step restatement (oracle/input_port.py) against goldens recorded from the UNMODIFIED reference processes
(ResizeImage modes resize / pad, NormalizeImage, MakeRecognitionLabel; oracle/make_golden.py input) and, where OpenCV is
importable, against cv2.resize itself."""
import os

import numpy as np
import pytest

from megreader_b200.charset import EnglishCharset
from oracle import input_port
from tests.input_cases import MODES, TOL, input_cases

GOLD = os.path.join(os.path.dirname(__file__), "golden", "input_ref.npz")


@pytest.mark.parametrize("mode", sorted(MODES))
def test_resize_normalize_reproduces_reference(mode):
    g = np.load(GOLD)
    images, _ = input_cases()
    got = np.stack([input_port.normalize(input_port.resize_or_pad(im.astype(np.float32), MODES[mode], mode)) for im in images])
    assert got.shape == g["image." + mode].shape and got.dtype == np.float32
    np.testing.assert_allclose(got, g["image." + mode], rtol=0, atol=TOL)


def test_label_packing_reproduces_reference():
    g = np.load(GOLD)
    _, texts = input_cases()
    cs = EnglishCharset()
    packed = [input_port.pack_label(t, cs.index, 32) for t in texts]
    assert np.array_equal(np.stack([p[0] for p in packed]), g["labels"])
    assert np.array_equal(np.array([p[1] for p in packed], np.int32), g["lengths"])


def test_resize_restatement_vs_opencv():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.RandomState(0)
    for (h, w, dh, dw) in [(37, 91, 32, 128), (64, 512, 32, 128), (20, 15, 32, 128), (100, 33, 48, 160), (5, 7, 32, 100), (33, 1, 32, 32)]:
        img = (rng.rand(h, w, 3) * 255).astype(np.float32)
        np.testing.assert_allclose(input_port.resize_bilinear_f32(img, dh, dw), cv2.resize(img, (dw, dh)), rtol=0, atol=2.5e-3)
