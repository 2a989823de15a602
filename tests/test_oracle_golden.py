"""CPU: pin the numpy oracle against fixtures produced by executing the reference itself
(tests/golden/make_golden.py). CubePad + integer maps bit-exact; float maps <= 1e-6."""
import hashlib

import numpy as np
import pytest

from oracle import c2e as oc2e
from oracle import cubepad as ocp
from oracle import e2c as oe2c


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_cubepad_index_maps_all(golden_meta, golden_small):
    for key, info in golden_meta["cubepad_maps"].items():
        m = ocp.index_map(info["H"], info["H"], info["pad"]).astype(np.int32)
        assert list(m.shape) == info["shape"], key
        assert sha(m) == info["sha256"], key
        if key in golden_small.files:
            np.testing.assert_array_equal(m, golden_small[key])


def test_cubepad_kat(golden_meta):
    for kat in golden_meta["cubepad_kat"]:
        x = np.arange(int(np.prod(kat["shape"])), dtype=np.float32).reshape(kat["shape"])
        y = ocp.cubepad(x, kat["pad"])
        assert list(y.shape) == kat["out_shape"]
        assert sha(y) == kat["sha256"]
        assert float(y.astype(np.float64).sum()) == kat["sum"]


def test_cubepad_worked_example():
    # SURVEY.md §8c worked example, 4x4 p=1
    x = np.arange(96, dtype=np.float32).reshape(6, 1, 4, 4)
    y = ocp.cubepad(x, 1)[:, 0].astype(int)
    assert y[0].tolist() == [[83, 83, 82, 81, 80, 80], [67, 0, 1, 2, 3, 48], [71, 4, 5, 6, 7, 52],
                             [75, 8, 9, 10, 11, 56], [79, 12, 13, 14, 15, 60], [31, 31, 30, 29, 28, 28]]
    assert y[5].tolist() == [[3, 3, 2, 1, 0, 0], [48, 80, 81, 82, 83, 67], [49, 84, 85, 86, 87, 66],
                             [50, 88, 89, 90, 91, 65], [51, 92, 93, 94, 95, 64], [32, 32, 33, 34, 35, 35]]


def test_cubepad_random_multigroup(golden_meta, golden_small):
    info = golden_meta["cubepad_rand"]
    x = np.random.default_rng(info["seed"]).standard_normal(info["shape"]).astype(np.float32)
    y = ocp.cubepad(x, info["pad"])
    np.testing.assert_array_equal(y, golden_small["cubepad_rand_12x5x6x6_p2-1-1-3"])


def test_cubepad_errors():
    with pytest.raises(ValueError):
        ocp.cubepad(np.zeros((5, 1, 4, 4), np.float32), 1)
    with pytest.raises(ValueError):
        ocp.index_map(4, 5, 1)


def test_e2c_maps_and_faces(golden_meta, golden_small):
    for key, info in golden_meta["e2c"].items():
        w, H, W = info["w"], info["H"], info["W"]
        sx, sy = oe2c.fixed_maps(w, H, W, info["vfov"])
        assert sha(np.concatenate([sx.reshape(-1), sy.reshape(-1)])) == info["sxsy_sha256"], key
        img = np.random.default_rng(info["seed"]).random((H, W, 3), dtype=np.float32)
        faces = oe2c.to_cube(img, sx, sy)
        # cv2.remap fp32 result reproduced bit for bit
        assert sha(faces) == info["faces_sha256"], key
        if key + "_sx" in golden_small.files:
            np.testing.assert_array_equal(sx, golden_small[key + "_sx"])
            np.testing.assert_array_equal(sy, golden_small[key + "_sy"])
            np.testing.assert_array_equal(faces, golden_small[key + "_faces"])
            xs, ys = oe2c.build_maps(w, H, W, info["vfov"])
            np.testing.assert_allclose(np.stack(xs), golden_small[key + "_inX"], rtol=0, atol=1e-9)
            np.testing.assert_allclose(np.stack(ys), golden_small[key + "_inY"], rtol=0, atol=1e-9)
        else:
            np.testing.assert_array_equal(faces[:, ::17, ::13, :], golden_small[key + "_faces_probe"])


def test_e2c_pack_roundtrip():
    sx, sy = oe2c.fixed_maps(16, 64, 128)
    p = oe2c.pack_map(sx, sy)
    assert p.dtype == np.uint32
    np.testing.assert_array_equal((p >> 20) & 2047, sx >> 5)
    np.testing.assert_array_equal((p >> 10) & 1023, sy >> 5)
    np.testing.assert_array_equal((p >> 5) & 31, sx & 31)
    np.testing.assert_array_equal(p & 31, sy & 31)


def test_c2e_maps_and_output(golden_meta, golden_small):
    for key, info in golden_meta["c2e"].items():
        w, C = info["w"], info["C"]
        face, coord = oc2e.build_maps(w)
        assert sha(face.astype(np.int64))[:16] == info["face_sha256_16"], key
        assert sha(coord.astype(np.float32))[:16] == info["coord32_sha256_16"], key
        assert sha(coord) == info["coord64_sha256"], key       # float64 bit-for-bit
        _, _, _, M = oc2e.sample_plan(coord, w)
        assert M == info["M"]
        cube = np.random.default_rng(info["seed"]).standard_normal((6, C, w, w)).astype(np.float32)
        out = oc2e.to_equi(cube, face, coord)
        assert abs(float(out.astype(np.float64).sum()) - info["out_sum"]) < 1e-2
        if key + "_out" in golden_small.files:
            np.testing.assert_array_equal(face.astype(np.int8), golden_small[key + "_face"])
            ref = golden_small[key + "_out"]
            assert out.shape == ref.shape
            assert np.abs(out - ref).max() <= 4e-6, key   # tolerance: CPU grid_sample vs restatement
        else:
            assert np.abs(out[:, :, ::7, ::11] - golden_small[key + "_out_probe"]).max() <= 4e-6


# ---------------------------------------------------------------------------------------------
# bicubic back-projection (row f4): fixture = the reference's own to_equi_cv2 output
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["w7", "w8", "w16"])
def test_c2e_cubic_golden_bit_exact(golden_cubic, tag):
    w, seed = (int(v) for v in golden_cubic[tag + "_meta"])
    cube = np.random.default_rng(seed).standard_normal((6, 1000, w, w)).astype(np.float32)
    face, coord = oc2e.build_maps(w)
    out = oc2e.to_equi_cv2(cube, face, coord)
    assert out.shape == (1000, 2 * w, 4 * w) and out.dtype == np.float32
    np.testing.assert_array_equal(out[golden_cubic["keep"]], golden_cubic[tag + "_planes"])
    assert hashlib.sha256(out.tobytes()).digest() == golden_cubic[tag + "_sha256"].tobytes()


def test_c2e_cubic_table_against_cv2():
    """Every 1/32-pixel fraction pair, interior and border windows, against cv2.remap itself."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    w = 9
    src = rng.standard_normal((w, w, 4)).astype(np.float32)
    # a synthetic out_coord sweeping x,y over [0, w-1] in 1/32 steps (plus off-grid values that round)
    xs = np.arange(0, (w - 1) * 32 + 1, dtype=np.float64) / 32
    gx, gy = np.meshgrid(xs, xs[::5] + 1e-3)
    coord = np.stack([gx, np.minimum(gy, w - 1)], axis=-1)
    want = cv2.remap(src, coord[..., 0].astype(np.float32), coord[..., 1].astype(np.float32), cv2.INTER_CUBIC)
    cube = np.ascontiguousarray(np.broadcast_to(src.transpose(2, 0, 1)[None], (6, 4, w, w)))
    got = oc2e.to_equi_cv2(cube, np.zeros(coord.shape[:2]), coord)
    np.testing.assert_array_equal(got, want.transpose(2, 0, 1))


# ---------------------------------------------------------------------------------------------
# oracle.ref_port — the library-call port timed as bench.py's CPU baseline
# ---------------------------------------------------------------------------------------------
def test_ref_port_cubepad(golden_meta, golden_small):
    import torch
    from oracle import ref_port
    for key, info in golden_meta["cubepad_maps"].items():
        H = info["H"]
        if H > 64:
            continue
        x = torch.arange(6 * H * H, dtype=torch.float64).reshape(6, 1, H, H)
        y = ref_port.CubePadPort(info["pad"])(x)[:, 0].numpy().astype(np.int32)
        assert sha(y) == info["sha256"], key
    for kat in golden_meta["cubepad_kat"]:
        x = torch.arange(int(np.prod(kat["shape"])), dtype=torch.float32).reshape(kat["shape"])
        assert sha(ref_port.CubePadPort(kat["pad"])(x).numpy()) == kat["sha256"]
    with pytest.raises(ValueError):
        ref_port.CubePadPort(1)(torch.zeros(5, 1, 4, 4))


def test_ref_port_e2c_c2e(golden_meta, golden_small):
    import torch
    from oracle import ref_port
    cv2 = pytest.importorskip("cv2")
    for key, info in golden_meta["e2c"].items():
        if info["w"] > 64:
            continue
        img = np.random.default_rng(info["seed"]).random((info["H"], info["W"], 3), dtype=np.float32)
        faces = ref_port.Equi2CubePort(info["w"], info["H"], info["W"], info["vfov"]).to_cube(img)
        assert sha(np.stack([faces[i] for i in range(6)])) == info["faces_sha256"], key
    for key, info in golden_meta["c2e"].items():
        if key + "_out" not in golden_small.files:
            continue
        cube = np.random.default_rng(info["seed"]).standard_normal((6, info["C"], info["w"], info["w"])).astype(np.float32)
        port = ref_port.Cube2EquiPort(info["w"])
        out = port.to_equi_nn(torch.from_numpy(cube)).numpy()
        assert np.abs(out - golden_small[key + "_out"]).max() <= 4e-6, key
        np.testing.assert_array_equal(port.to_equi_max(torch.from_numpy(cube)).numpy(), out[0].max(axis=0))


def test_e2c_integer_frames_within_one_lsb_of_cv2():
    """Equi2Cube.to_cube on uint8 frames: the host mirror resamples in float32 and rounds to nearest + saturates
    (equi_to_cube.py docstring). cv2.remap on 8-bit sources uses its own fixed-point path; pin how far the two
    are apart on the same maps: never more than 1 LSB, equal on the large majority of pixels."""
    import cv2
    H, W, w = 96, 192, 24
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    inX, inY = oe2c.build_maps(w, H, W)
    sx, sy = oe2c.fixed_maps(w, H, W)
    faces = oe2c.to_cube(img.astype(np.float32), sx, sy).reshape(6, w, w, 3)
    ours = np.clip(np.rint(faces), 0, 255).astype(np.uint8)
    for f in range(6):
        mx = inX[f].astype(np.float32).reshape(w, w)
        my = inY[f].astype(np.float32).reshape(w, w)
        want = np.stack([cv2.remap(np.ascontiguousarray(img[:, :, c]), mx, my, cv2.INTER_LINEAR) for c in range(3)], -1)
        d = np.abs(ours[f].astype(np.int32) - want.astype(np.int32))
        assert int(d.max()) <= 1
        assert float((d == 0).mean()) > 0.9
