"""CPU: the PRODUCT's per-pixel input-step routines (megreader_b200/csrc/input_core.cuh -- the code the CUDA kernels run)
compiled for the host by tests/host_harness/input_core_host.cpp and compared with the oracle and the reference goldens, so that
only the launch glue of csrc/input_pipeline.cu is left to the GPU tests."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

from megreader_b200.charset import EnglishCharset
from megreader_b200.input_pipeline import RGB_MEAN, charset_lut, resized_width
from oracle import input_port
from tests.input_cases import MODES, TOL, input_cases

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "input_ref.npz")


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("g++ not available")
    so = str(tmp_path_factory.mktemp("harness") / "libinput_core_host.so")
    subprocess.check_call([gxx, "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off",
                           "-I", os.path.join(HERE, "..", "megreader_b200", "csrc"),
                           os.path.join(HERE, "host_harness", "input_core_host.cpp"), "-o", so])
    return ctypes.CDLL(so)


def _run(lib, images, size, mode, as_u8):
    dt = np.uint8 if as_u8 else np.float32
    flat = np.concatenate([np.ascontiguousarray(im, dtype=dt).reshape(-1) for im in images])
    offsets = np.zeros(len(images), np.int64)
    offsets[1:] = np.cumsum([im.size for im in images])[:-1]
    hs = np.array([im.shape[0] for im in images], np.int32)
    ws = np.array([im.shape[1] for im in images], np.int32)
    valid = np.array([resized_width(mode, size, im.shape[0], im.shape[1]) for im in images], np.int32)
    out = np.empty((len(images), 3, size[0], size[1]), np.float32)
    mean = np.array(RGB_MEAN, np.float64)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    lib.host_resize_normalize(p(flat), ctypes.c_int(int(as_u8)), p(offsets), p(hs), p(ws), p(valid), ctypes.c_int(len(images)),
                              ctypes.c_int(size[0]), ctypes.c_int(size[1]), p(mean), p(out))
    return out


@pytest.mark.parametrize("mode", sorted(MODES))
@pytest.mark.parametrize("as_u8", [True, False])
def test_core_routine_vs_oracle_and_reference(harness, mode, as_u8):
    images, _ = input_cases()
    got = _run(harness, images, MODES[mode], mode, as_u8)
    want = np.stack([input_port.normalize(input_port.resize_or_pad(im.astype(np.float32), MODES[mode], mode)) for im in images])
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-7)            # same arithmetic, same order
    np.testing.assert_allclose(got, np.load(GOLD)["image." + mode], rtol=0, atol=TOL)


def test_core_label_packing(harness):
    _, texts = input_cases()
    raw = [t.encode("latin-1") for t in texts]
    offsets = np.zeros(len(raw) + 1, np.int64)
    offsets[1:] = np.cumsum([len(r) for r in raw])
    blob = np.frombuffer(b"".join(raw), np.uint8).copy()
    lut = charset_lut(EnglishCharset())
    labels = np.empty((len(raw), 32), np.int32)
    lengths = np.empty((len(raw),), np.int32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    harness.host_pack_labels(p(blob), p(offsets), ctypes.c_int(len(raw)), p(lut), ctypes.c_int(32), p(labels), p(lengths))
    g = np.load(GOLD)
    assert np.array_equal(labels, g["labels"]) and np.array_equal(lengths, g["lengths"])
