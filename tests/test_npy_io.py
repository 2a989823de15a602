"""CPU: the native .npy reader / writer (SURVEY.md §8 row f4 — the on-disk formats either side of the
path) against numpy itself: written files are byte-identical to numpy.save, numpy.save'd files of
every supported dtype / format version read back as float32(array)."""
import os

import numpy as np
import pytest
import torch

import cp360_b200
from cp360_b200 import _lib


@pytest.mark.parametrize("shape", [(14, 28), (6, 1000, 7, 7), (16, 32), (5,), (), (0, 3), (1, 1, 1, 1, 2)])
def test_writer_is_byte_identical_to_numpy_save(tmp_path, shape):
    rng = np.random.default_rng(len(shape))
    a = rng.standard_normal(shape).astype(np.float32)
    ours, theirs = str(tmp_path / "ours.npy"), str(tmp_path / "theirs.npy")
    cp360_b200.save_npy(ours, a)
    np.save(theirs, a)
    assert open(ours, "rb").read() == open(theirs, "rb").read()
    np.testing.assert_array_equal(np.load(ours), a)
    assert not os.path.exists(ours + ".tmp~")


def test_writer_header_padding_matches_numpy_for_every_header_length(tmp_path):
    """numpy pads in two steps (21-digit growth room for the first extent, then 1..64 spaces up to a 64 B boundary,
    never 0): sweep header lengths across the boundaries with zero-size arrays of 1..14 dims and extents of 1..10
    digits, including the exactly-aligned case and headers longer than 128 bytes."""
    rng = np.random.default_rng(11)
    seen = set()
    for i in range(300):
        nd = int(rng.integers(1, 15))
        shape = [int(rng.integers(1, 4)) for _ in range(nd)]            # product stays far below 2**63
        shape[int(rng.integers(nd))] = int(10 ** int(rng.integers(0, 10)) * int(rng.integers(1, 10)))
        shape[int(rng.integers(nd))] = 0                       # no data: extents are free
        a = np.empty(tuple(shape), np.float32)
        ours, theirs = str(tmp_path / "o.npy"), str(tmp_path / "t.npy")
        cp360_b200.save_npy(ours, a)
        np.save(theirs, a)
        bo, bt = open(ours, "rb").read(), open(theirs, "rb").read()
        assert bo == bt, shape
        seen.add(len(bt))
        assert cp360_b200.npy_header(ours)[1] == tuple(shape)
    assert {128, 192} <= seen


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.float16, np.uint8, np.int32, np.int64])
def test_reader_converts_like_float_tensor(tmp_path, dtype):
    rng = np.random.default_rng(7)
    a = (rng.standard_normal((6, 10, 7, 7)) * 50).astype(dtype)
    if dtype == np.float16:                      # subnormals, zeros and infinities of binary16 too
        a.reshape(-1)[:6] = np.array([0.0, -0.0, 6e-8, -6e-5, np.inf, -np.inf], dtype=np.float16)
    path = str(tmp_path / "a.npy")
    np.save(path, a)
    descr, shape, off, fortran = cp360_b200.npy_header(path)
    assert descr == np.dtype(dtype).str.replace("=", "<") and shape == a.shape and not fortran and off % 64 == 0
    got = cp360_b200.load_npy(path)
    assert got.dtype == torch.float32 and tuple(got.shape) == a.shape
    np.testing.assert_array_equal(got.numpy(), a.astype(np.float32))       # == torch.FloatTensor(np.load(path))
    buf = torch.empty(a.size)
    assert cp360_b200.load_npy(path, out=buf) is buf
    with pytest.raises(ValueError):                    # same element count, different layout: refused
        cp360_b200.load_npy(path, out=torch.empty((6, 7, 7, 10)))
    np.testing.assert_array_equal(buf.numpy().reshape(a.shape), a.astype(np.float32))


def test_reader_format_versions_and_cube_feat(tmp_path):
    a = np.random.default_rng(1).standard_normal((6, 5, 7, 7)).astype(np.float32)
    for version in ((1, 0), (2, 0), (3, 0)):
        path = str(tmp_path / ("v%d.npy" % version[0]))
        with open(path, "wb") as f:
            np.lib.format.write_array(f, a, version=version)
        np.testing.assert_array_equal(cp360_b200.load_cube_feat(path).numpy(), a)
    bad = str(tmp_path / "bad.npy")
    np.save(bad, np.zeros((5, 5, 7, 7), np.float32))
    with pytest.raises(ValueError):
        cp360_b200.load_cube_feat(bad)


def test_reader_errors(tmp_path):
    lib = _lib.lib()
    assert lib.cp360_npy_read_f32(str(tmp_path / "missing.npy").encode(), None, 0) == 1
    junk = tmp_path / "junk.npy"
    junk.write_bytes(b"not an npy file at all")
    with pytest.raises(_lib.CP360Error):
        cp360_b200.load_npy(str(junk))
    f = str(tmp_path / "f.npy")
    np.save(f, np.asfortranarray(np.zeros((3, 4), np.float32)))
    with pytest.raises(_lib.CP360Error):
        cp360_b200.load_npy(f)
    c = str(tmp_path / "c.npy")
    np.save(c, np.zeros((3, 4), np.complex64))
    with pytest.raises(_lib.CP360Error):
        cp360_b200.load_npy(c)
    t = str(tmp_path / "t.npy")
    np.save(t, np.zeros((30, 40), np.float32))
    data = open(t, "rb").read()
    open(t, "wb").write(data[:-100])                                        # truncated data
    with pytest.raises(_lib.CP360Error):
        cp360_b200.load_npy(t)
    buf = torch.empty(5)
    with pytest.raises(ValueError):
        cp360_b200.load_npy(c, out=buf)


@pytest.mark.gpu
def test_backproject_files_matches_oracle(tmp_path):
    from oracle import c2e as oc2e
    rng = np.random.default_rng(3)
    feats, outs, cubes = [], [], []
    for i in range(5):
        cube = rng.standard_normal((6, 40, 7, 7)).astype(np.float32)
        p = str(tmp_path / ("%06d.npy" % (i + 1)))
        np.save(p, cube)                                                    # dataset_feat_extractor.py:187-189
        feats.append(p)
        outs.append(str(tmp_path / ("%05d.npy" % i)))
        cubes.append(cube)
    assert cp360_b200.backproject_files(feats, outs, batch=2) == 5
    face, coord = oc2e.build_maps(7)
    for cube, o in zip(cubes, outs):
        got = np.load(o)
        assert got.shape == (14, 28) and got.dtype == np.float32            # test_temporal.py:86-88
        assert np.abs(got - oc2e.to_equi_max(cube, face, coord)).max() <= 1e-5


def test_reader_survives_mutated_headers(tmp_path):
    """Robustness of the native header parser (csrc/npy.cpp): 600 seeded byte-level mutations of valid files
    (flips, insertions, deletions, truncations, in the header region or anywhere; hostile shapes) never crash the process;
    whenever the native reader accepts a file numpy accepts it too and both yield the same values."""
    rng = np.random.default_rng(20261017)
    base = []
    for shape, dtype, version in [((6, 10, 7, 7), np.float32, (1, 0)), ((14, 28), np.float64, (2, 0)),
                                  ((5,), np.int32, (1, 0)), ((), np.float32, (1, 0)), ((3, 0, 2), np.uint8, (3, 0))]:
        import numpy.lib.format as fmt
        a = (rng.standard_normal(shape) * 10).astype(dtype)
        p = tmp_path / ("base_%d.npy" % len(base))
        with open(p, "wb") as f:
            fmt.write_array(f, a, version=version)
        base.append(open(p, "rb").read())
    hostile = [b"\x93NUMPY\x01\x00\x46\x00{'descr': '<f4', 'fortran_order': False, 'shape': (4611686018427387904, 4), }   \n",
               b"\x93NUMPY\x01\x00\x40\x00{'descr': '<f4', 'fortran_order': False, 'shape': (" + b"1," * 40 + b"), }\n",
               b"\x93NUMPY\x01\x00\xff\xff{'descr': '<f4'", b"\x93NUMPY\x02\x00\xff\xff\xff\x7f", b"\x93NUMPY\x01\x00\x00\x00",
               b"\x93NUMPY\x01\x00\x10\x00{'shape': (2,), }"]
    accepted = 0
    path = str(tmp_path / "mut.npy")
    for i in range(600):
        if i < len(hostile):
            blob = hostile[i]
        else:
            blob = bytearray(base[int(rng.integers(len(base)))])
            region = min(len(blob), 128) if i % 2 else len(blob)       # header region / anywhere (data bytes too)
            for _ in range(int(rng.integers(1, 4))):
                if len(blob) < 2:
                    break
                kind, pos = int(rng.integers(4)), int(rng.integers(min(region, len(blob))))
                if kind == 0:
                    blob[pos] = int(rng.integers(256))
                elif kind == 1:
                    blob.insert(pos, int(rng.integers(256)))
                elif kind == 2:
                    del blob[pos]
                else:
                    del blob[int(rng.integers(len(blob))):]
            blob = bytes(blob)
        with open(path, "wb") as f:
            f.write(blob)
        try:
            got = cp360_b200.load_npy(path)
        except (ValueError, RuntimeError, MemoryError, OverflowError):
            continue
        want = np.load(path, allow_pickle=False)                 # accepted natively => numpy must accept it too
        assert tuple(got.shape) == want.shape, (i, blob[:96])
        assert np.array_equal(got.numpy(), want.astype(np.float32), equal_nan=True), (i, blob[:96])
        accepted += 1
    assert accepted >= 20                                        # mutations in padding / data leave files valid
