"""The .npy files either side of the hot path, through libcp360's native reader / writer.

    cube_feat/%06d.npy   [6,1000,7,7] float32 cube class scores   dataset_feat_extractor.py:187-189 (write)
                                                                  test_temporal.py:64,70, data/dataset.py:65 (read)
    %05d.npy             [14,28] float32 equirect saliency map     test_temporal.py:86-88 (write)

``load_npy`` fills a (pinned) float32 host tensor, which is what the reference builds with
``torch.FloatTensor(np.load(path))`` before its ``.cuda()``; ``save_npy`` writes files byte-identical
to ``numpy.save``. ``backproject_files`` is the file-to-file form of the path's last step:
score files -> Cube2Equi + channel max on the GPU -> result files.
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib


def npy_header(path):
    """(dtype string, shape tuple, data offset, fortran_order) of a .npy file."""
    descr = ctypes.create_string_buffer(32)
    ndim, fortran = ctypes.c_int(), ctypes.c_int()
    shape = (ctypes.c_int64 * 32)()
    off = ctypes.c_int64()
    _lib.check(_lib.lib().cp360_npy_read_header(os.fsencode(path), descr, 32, ctypes.byref(ndim), shape, 32,
                                               ctypes.byref(off), ctypes.byref(fortran)))
    return descr.value.decode(), tuple(int(shape[i]) for i in range(ndim.value)), int(off.value), bool(fortran.value)


def load_npy(path, out=None, pin=False):
    """Read a .npy array as a float32 host tensor of the file's shape (``out``: reuse this buffer)."""
    _, shape, _, _ = npy_header(path)
    n = int(np.prod(shape, dtype=np.int64)) if shape else 1
    if out is None:
        out = torch.empty(shape, dtype=torch.float32, pin_memory=bool(pin and torch.cuda.is_available()))
    elif out.dtype != torch.float32 or not out.is_contiguous() or out.is_cuda:
        raise ValueError("out must be a contiguous float32 host tensor")
    elif tuple(out.shape) != tuple(shape) and not (out.dim() == 1 and out.numel() == n):
        # a flat buffer of the right size is fine; the same element count in another layout (e.g. a
        # [6,7,7,1000] file into a [6,1000,7,7] buffer) must not be reinterpreted silently
        raise ValueError("%s holds an array of shape %s, out has shape %s" % (path, tuple(shape), tuple(out.shape)))
    _lib.check(_lib.lib().cp360_npy_read_f32(os.fsencode(path), out.data_ptr(), n))
    return out


def save_npy(path, array):
    """Write a float32 array (numpy or host/CUDA tensor) as .npy, byte-identical to numpy.save."""
    if isinstance(array, torch.Tensor):
        array = array.detach().to("cpu", torch.float32).contiguous().numpy()
    a = np.asarray(array, dtype=np.float32, order="C")      # (ascontiguousarray would promote 0-d to 1-d)
    shape = (ctypes.c_int64 * max(1, a.ndim))(*a.shape)
    _lib.check(_lib.lib().cp360_npy_write_f32(os.fsencode(path), a.ctypes.data, a.ndim, shape))
    return path


def load_cube_feat(path, out=None, pin=False):
    """`cube_feat/%06d.npy` -> float32 host tensor [6,C,w,w] (test_temporal.py:64,70)."""
    t = load_npy(path, out=out, pin=pin)
    if t.dim() != 4 or t.shape[0] != 6 or t.shape[2] != t.shape[3]:
        raise ValueError("%s: expected cube scores [6,C,w,w], got %s" % (path, tuple(t.shape)))
    return t


def backproject_files(feat_paths, out_paths, device=None, batch=16, square=False, align_corners=False):
    """Score files -> saliency files: for every `cube_feat` file, Cube2Equi + channel max
    (test_temporal.py:82-88 without the ConvLSTM in between; dataset_feat_extractor.py:174-176 with
    ``square=True``). Files are read into pinned buffers, moved and processed ``batch`` at a time on the
    GPU, and written back with the native writer. Returns the number of maps written.
    align_corners: grid_sample convention of the back-projection (Cube2Equi docstring); every file's own
    header is checked against the batch's shape."""
    from .cube_to_equi import Cube2Equi
    if len(feat_paths) != len(out_paths):
        raise ValueError("feat_paths and out_paths differ in length")
    if not torch.cuda.is_available():
        raise RuntimeError("backproject_files needs a CUDA device (no CPU fallback)")
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    done, c2e, stage = 0, None, None
    for i0 in range(0, len(feat_paths), batch):
        chunk = feat_paths[i0:i0 + batch]
        _, shape, _, _ = npy_header(chunk[0])
        if len(shape) != 4 or shape[0] != 6 or shape[2] != shape[3]:
            raise ValueError("%s: expected cube scores [6,C,w,w], got %s" % (chunk[0], shape))
        if stage is None or tuple(stage.shape[1:]) != shape or stage.shape[0] < len(chunk):
            stage = torch.empty((batch,) + shape, dtype=torch.float32).pin_memory()
            c2e = Cube2Equi(shape[2], align_corners=align_corners) if c2e is None or c2e.input_w != shape[2] else c2e
        for j, p in enumerate(chunk):
            load_cube_feat(p, out=stage[j])
        x = stage[:len(chunk)].to(dev, non_blocking=True).reshape((6 * len(chunk),) + shape[1:])
        sal = c2e.to_equi_max(x)
        if square:
            sal = sal * sal
        sal = sal.cpu()
        for j in range(len(chunk)):
            save_npy(out_paths[i0 + j], sal[j])
            done += 1
    return done
