"""MatNet selector, natively: the HDF5 reader against the reference's own weight files (when the reference
checkout is present -- this container) and an independent Python reader; the forward pass against a NumPy
restatement of CPU/MatNet.py.  Host only, no GPU needed.  There is no Keras in the image, so inference is
pinned on restated layers, not on Keras output."""
import os

import numpy as np
import pytest

from ia_spgemm_b200.engine import MatNet, load_library
from ia_spgemm_b200 import workloads as W
import matnet_numpy as MN
from h5mini import H5File

REF = "/root/reference"
FILES = {"Intel": ("IA-SPGEMM-CPU_release/NetWeights/Intel_weights.h5", 26, 5, 43589),
         "Amd": ("IA-SPGEMM-CPU_release/NetWeights/Amd_weights.h5", 26, 5, 43589),
         "P100": ("IA-SPGEMM-GPU_release/NetWeights/P100_weights.h5", 18, 3, 43023)}


@pytest.fixture(scope="module")
def lib():
    return load_library()


def _images(oracle, seed):
    A = W.random_sparse(300, 300, 0.02, seed=seed)
    B = W.banded(500, [-3, 0, 1, 40], seed=seed)
    return oracle.density(A[0], A[1], A[2], A[3]), oracle.density(B[0], B[1], B[2], B[3])


@pytest.mark.parametrize("n_features,n_classes", [(26, 5), (18, 3)])
def test_forward_matches_numpy_restatement(lib, oracle, n_features, n_classes):
    for seed in range(4):
        w = MN.random_weights(n_features, n_classes, seed=seed, scale=1.0 + seed)
        net = MatNet.from_arrays(w, lib)
        assert net.shape()[:2] == (n_features, n_classes)
        img1, img2 = _images(oracle, seed + 1)
        feats = np.random.default_rng(seed).uniform(0, 3, size=26)
        cls, probs = net.predict(img1, img2, feats)
        want_cls, want_probs = MN.predict(w, img1, img2, feats)
        assert np.allclose(probs, want_probs, rtol=2e-3, atol=2e-5)           # fp32 engine vs float64 restatement
        assert probs.sum() == pytest.approx(1.0, abs=1e-9)
        if np.sort(want_probs)[-1] - np.sort(want_probs)[-2] > 1e-3:
            assert cls == want_cls
        net.close()


def test_same_padding_and_pooling_shapes():
    """128 -(3x3 valid)-> 126 -pool-> 63 -(5x5 s2 same)-> 32 -pool-> 16 -(5x5 s2 same)-> 8 -pool-> 4; 4*4*16 = 256."""
    x = np.random.default_rng(0).normal(size=(128, 128, 1))
    w = MN.random_weights(26, 5)
    a = MN.maxpool2(MN.conv2d(x, w["conv2d_1/kernel"].astype(float), w["conv2d_1/bias"].astype(float), 1, False))
    assert a.shape == (63, 63, 16)
    b = MN.maxpool2(MN.conv2d(a, w["conv2d_2/kernel"].astype(float), w["conv2d_2/bias"].astype(float), 2, True))
    assert b.shape == (16, 16, 16)
    c = MN.maxpool2(MN.conv2d(b, w["conv2d_3/kernel"].astype(float), w["conv2d_3/bias"].astype(float), 2, True))
    assert c.shape == (4, 4, 16)


def test_missing_file_and_garbage(lib, tmp_path):
    import ctypes as C
    h = C.c_void_p()
    assert lib.ias_matnet_load(str(tmp_path / "nope.h5").encode(), C.byref(h)) == 6          # IAS_E_IO
    p = tmp_path / "junk.h5"
    p.write_bytes(b"\x89HDF\r\n\x1a\n" + b"\0" * 200)
    assert lib.ias_matnet_load(str(p).encode(), C.byref(h)) == 6


@pytest.mark.parametrize("which", sorted(FILES))
def test_reads_the_reference_weight_files(lib, oracle, which):
    rel, nf, nc, nparams = FILES[which]
    path = os.path.join(REF, rel)
    if not os.path.exists(path):
        pytest.skip("reference checkout not present (GPU box)")
    net = MatNet.load(path, lib)
    assert net.shape() == (nf, nc, nparams)                 # 43 589 parameters for the 26-feature / 5-class net
    ref = {k.split("/", 1)[1].replace(":0", ""): v for k, v in H5File(path).tensors().items()}
    assert sorted(ref) == sorted(MatNet.TENSORS)
    for name in MatNet.TENSORS:
        assert np.array_equal(net.tensor(name), ref[name]), name       # bit-identical to the independent reader
    assert net.tensor("conv2d_1/kernel").shape == (3, 3, 1, 16) and net.tensor("dense_4/kernel").shape == (64 + nf, nc)
    # the real weights through both forward passes
    img1, img2 = _images(oracle, 7)
    A = W.poisson2d(24)
    feats = oracle.features26(A, A)
    cls, probs = net.predict(img1, img2, feats)
    want_cls, want_probs = MN.predict(ref, img1, img2, feats)
    assert np.allclose(probs, want_probs, rtol=2e-3, atol=2e-5) and 0 <= cls < nc
    net.close()


def test_known_answer_from_the_reference_screenshots(lib, golden):
    """CPU/1.jpg: `./spgemm-cpu Inputs/dia.mtx` prints "The Chosen One = Algorithm 3" (and "Correct Prediction")
    for exactly these inputs: density(dia.mtx), density(dia.mtx^T) (the shipped imgs/ files) and the 26 features
    printed in the screenshot.  The native reader + forward pass reproduce that pick with the reference's weights."""
    from util import decode_img
    path = os.path.join(REF, FILES["Intel"][0])
    if not os.path.exists(path):
        pytest.skip("reference checkout not present (GPU box)")
    ship = golden["shipped_imgs"]
    feats = np.array(golden["screenshot_features_dia"])
    img1, img2 = decode_img(ship["gpu_img1_dia"]), decode_img(ship["gpu_img2_diaT"])
    for which in ("Intel", "Amd"):
        net = MatNet.load(os.path.join(REF, FILES[which][0]), lib)
        cls, probs = net.predict(img1, img2, feats)
        assert cls + 1 == 3 and probs[cls] > 0.9                       # Algorithm 3 = DIA
        net.close()
    net = MatNet.load(os.path.join(REF, FILES["P100"][0]), lib)
    cls, probs = net.predict(img1, img2, feats[:18])
    assert cls == 2 and probs[cls] > 0.9                               # GPU/2.jpg: Algorithm 3 (NSPARSE) is the fastest
    net.close()


_FUZZ = r'''
import ctypes as C, random, sys
sys.path.insert(0, sys.argv[1])
from ia_spgemm_b200 import engine as E
lib = E.load_library()
src = open(sys.argv[2], "rb").read()
rng = random.Random(int(sys.argv[4]))
loaded = 0
for t in range(int(sys.argv[5])):
    b = bytearray(src)
    kind = rng.randrange(4)
    if kind == 0:                                   # a few flipped bytes anywhere
        for _ in range(rng.randrange(1, 8)): b[rng.randrange(len(b))] = rng.randrange(256)
    elif kind == 1:                                 # truncated
        b = b[:rng.randrange(len(b))]
    elif kind == 2:                                 # the superblock / root group / heaps at the start
        for _ in range(rng.randrange(1, 40)): b[rng.randrange(min(4096, len(b)))] = rng.randrange(256)
    else:                                           # an 8-byte offset or length replaced by a random value
        o = rng.randrange(len(b) // 8) * 8
        b[o:o + 8] = rng.getrandbits(64).to_bytes(8, "little")
    open(sys.argv[3], "wb").write(b)
    h = C.c_void_p()
    if lib.ias_matnet_load(sys.argv[3].encode(), C.byref(h)) == 0 and h.value:
        loaded += 1
        lib.ias_matnet_free(h)
print("survived", loaded)
'''


@pytest.mark.timeout(300)
def test_weight_reader_survives_corrupted_files(tmp_path):
    """The HDF5 reader bounds-checks every offset and length it follows: 300 corrupted copies of a shipped weight file
    (flipped bytes, truncation, random offsets) are refused or loaded, never a crash (run in a child process so that a
    crash would be seen as one)."""
    import subprocess
    import sys
    path = os.path.join(REF, FILES["Intel"][0])
    if not os.path.exists(path):
        pytest.skip("reference weights not present")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for seed in (11, 12):
        r = subprocess.run([sys.executable, "-c", _FUZZ, root, path, str(tmp_path / "fuzz.h5"), str(seed), "150"],
                           capture_output=True, text=True, timeout=280)
        assert r.returncode == 0, "the reader crashed on a corrupted file (seed %d): %s" % (seed, r.stderr[-500:])
        assert r.stdout.startswith("survived")
