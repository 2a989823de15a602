"""CPU: the C-ABI boundary and the host map builders (no kernel launches).

 * libcp360.so loads and exports every symbol include/cp360.h declares
 * host-built integer maps == oracle == golden fixtures (bit-exact)
 * error behaviour mirrors the reference's (size mismatch, 2:1 assert, square faces)
"""
import ctypes
import hashlib
import os
import re

import numpy as np
import pytest
import torch

import cp360_b200
from cp360_b200 import _lib
from oracle import c2e as oc2e
from oracle import cubepad as ocp
from oracle import e2c as oe2c

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_header_symbols_exported():
    hdr = open(os.path.join(ROOT, "include", "cp360.h")).read()
    declared = set(re.findall(r"CP360_API\s+[\w\s\*]+?\b(cp360_\w+)\s*\(", hdr))
    assert len(declared) >= 16
    handle = ctypes.CDLL(cp360_b200.LIB_PATH)
    for name in declared:
        assert hasattr(handle, name), "libcp360.so does not export %s" % name
    assert declared == set(_lib.SIGNATURES), "ctypes table and header disagree"
    assert _lib.lib().cp360_version() == 100


def test_cubepad_map_matches_oracle_and_golden(golden_meta, golden_small):
    for key, info in golden_meta["cubepad_maps"].items():
        m = cp360_b200.cubepad_index_map(info["H"], info["H"], info["pad"])
        assert m.dtype == np.int32 and list(m.shape) == info["shape"]
        assert sha(m) == info["sha256"], key                    # == reference output on arange
        if info["H"] <= 64:
            np.testing.assert_array_equal(m, ocp.index_map(info["H"], info["H"], info["pad"]))


@pytest.mark.parametrize("H,pad", [(1, 1), (2, 2), (3, [3, 0, 1, 2]), (10, [0, 0, 0, 0]), (11, 5),
                                   (12, [5, 1, 2, 4]), (13, [1, 5, 4, 2])])
def test_cubepad_map_more_shapes_vs_oracle(H, pad):
    np.testing.assert_array_equal(cp360_b200.cubepad_index_map(H, H, pad), ocp.index_map(H, H, pad))


def test_cubepad_errors():
    lib = _lib.lib()
    ho, wo = ctypes.c_int(), ctypes.c_int()
    assert lib.cp360_cubepad_out_shape(4, 5, 1, 1, 1, 1, ctypes.byref(ho), ctypes.byref(wo)) == 3
    assert lib.cp360_cubepad_out_shape(4, 4, 5, 1, 1, 1, ctypes.byref(ho), ctypes.byref(wo)) == 3
    assert lib.cp360_cubepad_out_shape(4, 4, 1, 2, 3, 0, ctypes.byref(ho), ctypes.byref(wo)) == 0
    assert (ho.value, wo.value) == (7, 7)
    # batch not a multiple of 6: status 2 before any device is touched (cube_pad.py:33-35)
    assert lib.cp360_cubepad_fwd(None, None, 5, 1, 4, 4, 1, 1, 1, 1, 4, None) == 2
    assert b"size mismatch" in lib.cp360_last_error()
    assert lib.cp360_cubepad_fwd(None, None, 6, 1, 4, 4, 1, 1, 1, 1, 3, None) == 1   # elem size
    assert lib.cp360_cubepad_fwd(None, None, 0, 8, 4, 4, 1, 1, 1, 1, 4, None) == 0   # empty batch
    with pytest.raises(ValueError, match="size mismatch"):
        _lib.check(2)
    with pytest.raises(RuntimeError, match="CUDA"):
        cp360_b200.CubePad(1)(torch.zeros(6, 1, 4, 4))                               # CPU tensor
    with pytest.raises(RuntimeError):
        cp360_b200.CubePad(1, use_gpu=False)
    assert cp360_b200.get_pad_size(2) == (2, 2, 2, 2)
    assert cp360_b200.get_pad_size([1, 2, 3, 0]) == (1, 2, 3, 0)
    assert len(list(cp360_b200.CubePad(1).parameters())) == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-device error path")
def test_no_cpu_fallback_without_device():
    buf = np.zeros(6 * 36, np.float32)
    rc = _lib.lib().cp360_cubepad_fwd(buf.ctypes.data, buf.ctypes.data, 6, 1, 4, 4, 1, 1, 1, 1, 4, None)
    assert rc == 5 and b"no CPU fallback" in _lib.lib().cp360_last_error()
    with pytest.raises(RuntimeError, match="CUDA"):
        cp360_b200.Equi2Cube(8, np.zeros((32, 64, 3), np.float32)).to_cube(np.zeros((32, 64, 3), np.float32))


def test_e2c_maps_bit_exact(golden_meta, golden_small):
    for key, info in golden_meta["e2c"].items():
        w, H, W = info["w"], info["H"], info["W"]
        obj = cp360_b200.Equi2Cube(w, np.empty((H, W, 3), np.float32), vfov=info["vfov"])
        assert sha(np.concatenate([obj.sx.reshape(-1), obj.sy.reshape(-1)])) == info["sxsy_sha256"], key
        np.testing.assert_array_equal(obj.packed.reshape(6, w, w), oe2c.pack_map(obj.sx, obj.sy))
        inx32 = np.stack(obj.inXs).astype(np.float32)
        iny32 = np.stack(obj.inYs).astype(np.float32)
        assert sha(np.concatenate([inx32.reshape(-1), iny32.reshape(-1)])) == info["in32_sha256"], key
        if key + "_inX" in golden_small.files:
            np.testing.assert_allclose(np.stack(obj.inXs), golden_small[key + "_inX"], rtol=0, atol=1e-9)
            np.testing.assert_allclose(np.stack(obj.inYs), golden_small[key + "_inY"], rtol=0, atol=1e-9)
            assert len(obj.inXs) == 6 and obj.inXs[0].shape == (w * w,) and obj.inXs[0].dtype == np.float64


def test_e2c_errors():
    with pytest.raises(AssertionError):
        cp360_b200.Equi2Cube(8, np.zeros((32, 60, 3), np.float32))          # equi_to_cube.py:15
    lib = _lib.lib()
    assert lib.cp360_e2c_build_map(8, 32, 60, 90.0, None, None, None, None, None) == 3
    # frames beyond 2047 x 1023 take the wide (two-word) map; only sizes beyond 65535 x 32767 are out of range
    assert lib.cp360_e2c_map_words(4, 960, 1920) == 6 * 16 and lib.cp360_e2c_map_words(4, 2048, 4096) == 12 * 16
    big = np.empty(12 * 4 * 4, np.uint32)
    assert lib.cp360_e2c_build_map(4, 2048, 4096, 90.0, big.ctypes.data, None, None, None, None) == 0
    assert lib.cp360_e2c_build_map(4, 40000, 80000, 90.0, None, None, None, None, None) == 4
    assert lib.cp360_e2c_fwd(None, None, None, 1, 32, 60, 3, 8, 0, None, None, None) == 3


def test_c2e_maps_bit_exact(golden_meta, golden_small):
    for key, info in golden_meta["c2e"].items():
        w = info["w"]
        obj = cp360_b200.Cube2Equi(w)
        assert obj.face_map.dtype == np.float64 and obj.face_map.shape == (2 * w, 4 * w)
        assert obj.out_coord.shape == (2 * w, 4 * w, 2)
        assert sha(obj.face_map.astype(np.int64))[:16] == info["face_sha256_16"], key
        assert sha(obj.out_coord.astype(np.float32))[:16] == info["coord32_sha256_16"], key
        assert obj.M == info["M"]
        face, coord = oc2e.build_maps(w)
        np.testing.assert_allclose(obj.out_coord, coord, rtol=0, atol=1e-12)
        for ac in (False, True):
            o = cp360_b200.Cube2Equi(w, align_corners=ac)
            x0, y0, wts, M = oc2e.sample_plan(coord, w, ac)
            np.testing.assert_array_equal(o.weights.reshape(2 * w, 4 * w, 4), wts)   # fp32 bit-exact
            np.testing.assert_array_equal((o.taps & 0x3fff).astype(np.int32).reshape(2 * w, 4 * w) - 1, x0)
            np.testing.assert_array_equal(((o.taps >> 14) & 0x3fff).astype(np.int32).reshape(2 * w, 4 * w) - 1, y0)
            np.testing.assert_array_equal((o.taps >> 28).reshape(2 * w, 4 * w), face.astype(np.uint32))
            assert o.M == M


@pytest.mark.parametrize("w", [2, 7, 8, 16, 64, 256])
def test_c2e_cubic_plan_bit_exact(w):
    """cp360_c2e_build_cubic_plan == the oracle's cv2 fixed-point conversion of out_coord (row f4)."""
    taps = np.empty(8 * w * w, dtype=np.uint32)
    _lib.check(_lib.lib().cp360_c2e_build_cubic_plan(w, taps.ctypes.data))
    face, coord = oc2e.build_maps(w)
    x0, y0, fx, fy = oc2e.cubic_plan(coord)
    t = taps.reshape(2 * w, 4 * w)
    np.testing.assert_array_equal((t & 0x1ff).astype(np.int32) - 1, x0)
    np.testing.assert_array_equal(((t >> 9) & 0x1ff).astype(np.int32) - 1, y0)
    np.testing.assert_array_equal(((t >> 18) & 31).astype(np.int32), fx)
    np.testing.assert_array_equal(((t >> 23) & 31).astype(np.int32), fy)
    np.testing.assert_array_equal(t >> 28, face.astype(np.uint32))


def test_c2e_cubic_errors():
    lib = _lib.lib()
    assert lib.cp360_c2e_build_cubic_plan(0, None) == 1
    buf = np.empty(8, np.uint32)
    assert lib.cp360_c2e_build_cubic_plan(513, buf.ctypes.data) == 4
    assert lib.cp360_c2e_cubic_fwd(None, None, None, 1, 4, 8, None) == 1       # null pointers
    assert lib.cp360_c2e_cubic_fwd(None, None, None, 0, 4, 8, None) == 0       # empty batch


@pytest.mark.parametrize("H,pad", [(1, 1), (2, 1), (2, 2), (3, 3), (4, 1), (7, 1), (8, 3), (5, 2), (6, [1, 2, 3, 0]),
                                   (6, [2, 1, 1, 3]), (7, [3, 3, 1, 1]), (7, [1, 1, 3, 3]), (5, [0, 2, 1, 0]),
                                   (5, [2, 0, 0, 1]), (5, [0, 0, 2, 2]), (4, [0, 0, 0, 0]), (9, [4, 2, 3, 5]),
                                   (9, [9, 9, 9, 9]), (32, 3), (56, 1)])
def test_cubepad_inverse_map_is_transpose_of_forward_map(H, pad):
    """cp360_cubepad_build_inverse_map (what the backward kernel sums over) is exactly the transpose
    of the forward index map: every output pixel appears once, under the source it copies."""
    pl, pr, pt, pd = cp360_b200.get_pad_size(pad)
    fwd = cp360_b200.cubepad_index_map(H, H, pad).reshape(-1)                  # [6*Ho*Wo] -> source index
    n_src, n_out = 6 * H * H, fwd.size
    offs = np.empty(n_src + 1, dtype=np.int32)
    ents = np.full(n_out, -1, dtype=np.int32)
    _lib.check(_lib.lib().cp360_cubepad_build_inverse_map(H, H, pl, pr, pt, pd, offs.ctypes.data, ents.ctypes.data))
    assert offs[0] == 0 and offs[-1] == n_out
    assert np.array_equal(np.sort(ents), np.arange(n_out))                     # a permutation of the outputs
    src_of_entry = np.repeat(np.arange(n_src), np.diff(offs))
    np.testing.assert_array_equal(fwd[ents], src_of_entry)
    np.testing.assert_array_equal(np.diff(offs), np.bincount(fwd, minlength=n_src))
    # count-only form
    offs2 = np.empty_like(offs)
    _lib.check(_lib.lib().cp360_cubepad_build_inverse_map(H, H, pl, pr, pt, pd, offs2.ctypes.data, None))
    np.testing.assert_array_equal(offs, offs2)


@pytest.mark.parametrize("w,align", [(2, 0), (3, 1), (7, 0), (7, 1), (8, 0), (16, 0), (20, 1)])
def test_c2e_backward_plan_is_transpose_of_forward_plan(w, align):
    """cp360_c2e_build_bwd_plan: CSR over cube pixels of the same (equirect pixel, weight) pairs the forward plan
    applies — checked as dense matrices, and its application against torch autograd of grid_sample-style sampling
    (the gradient Cube2Equi.to_equi_nn needs, train_temporal.py:167-170)."""
    lib = _lib.lib()
    P, NC = 8 * w * w, 6 * w * w
    taps = np.empty(P, np.uint32)
    wts = np.empty((P, 4), np.float32)
    _lib.check(lib.cp360_c2e_build_plan(w, align, taps.ctypes.data, wts.ctypes.data, None))
    fwd = np.zeros((P, NC), np.float64)
    for i in range(P):
        face, y0, x0 = int(taps[i] >> 28), int((taps[i] >> 14) & 0x3fff) - 1, int(taps[i] & 0x3fff) - 1
        for k, (dy, dx) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
            yy, xx = y0 + dy, x0 + dx
            if 0 <= yy < w and 0 <= xx < w:
                fwd[i, (face * w + yy) * w + xx] += wts[i, k]
    offs = np.empty(NC + 1, np.int32)
    _lib.check(lib.cp360_c2e_build_bwd_plan(w, align, offs.ctypes.data, None, None))
    n = int(offs[-1])
    assert offs[0] == 0 and np.all(np.diff(offs) >= 0) and int((fwd != 0).sum()) <= n <= 4 * P
    pix, bw = np.empty(n, np.int32), np.empty(n, np.float32)
    _lib.check(lib.cp360_c2e_build_bwd_plan(w, align, offs.ctypes.data, pix.ctypes.data, bw.ctypes.data))
    bwd = np.zeros((NC, P), np.float64)
    for s in range(NC):
        seg = pix[offs[s]:offs[s + 1]]
        assert np.all(np.diff(seg) > 0), "contributors must be listed once each, in increasing order"
        for e in range(offs[s], offs[s + 1]):
            bwd[s, pix[e]] += bw[e]
    np.testing.assert_array_equal(bwd, fwd.T)
    # every equirect pixel's valid weights sum to <= 1 (bilinear), the plan has no entry for out-of-face taps
    assert float(fwd.sum(axis=1).max()) <= 1.0 + 1e-6
