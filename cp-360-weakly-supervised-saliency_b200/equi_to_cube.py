"""Equi2Cube — host mirror of the reference's utils/equi_to_cube.py:11-129.

    e2c = Equi2Cube(output_width, in_image, vfov=90)      # builds the sampling maps (host, once)
    faces = e2c.to_cube(in_image)                         # dict {0..5: ndarray[w,w,C]}  (reference API)
    t = e2c.to_cube_tensor(frames)                        # [B,H,W,C] cuda -> [6B,C,w,w]  (addition)
    x = e2c.to_padded_cube_tensor(frames, 3, mean, std)   # ... -> CubePad(3)(im_norm(faces)), one kernel

Map construction is cp360_e2c_build_map (C++, float64, same operation order as the numpy
code); resampling is cp360_e2c_fwd on the GPU with cv2.remap's fixed-point arithmetic.
"""
import numpy as np
import torch

from . import _lib


class Equi2Cube:
    def __init__(self, output_width, in_image, vfov=90, device=None):
        shape = tuple(in_image.shape)
        assert shape[0] * 2 == shape[1]                      # equi_to_cube.py:15
        self.output_width = self.output_height = int(output_width)
        self.input_height, self.input_width = int(shape[0]), int(shape[1])
        self.vfov = vfov
        w = self.output_width
        n = 6 * w * w
        # one uint32 per output pixel, two for frames beyond 2047 x 1023 (cp360.h: cp360_e2c_build_map)
        self.packed = np.empty(int(_lib.lib().cp360_e2c_map_words(w, self.input_height, self.input_width)), dtype=np.uint32)
        self.sx = np.empty(n, dtype=np.int32)
        self.sy = np.empty(n, dtype=np.int32)
        inx = np.empty(n, dtype=np.float64)
        iny = np.empty(n, dtype=np.float64)
        _lib.check(_lib.lib().cp360_e2c_build_map(
            w, self.input_height, self.input_width, float(vfov), self.packed.ctypes.data,
            self.sx.ctypes.data, self.sy.ctypes.data, inx.ctypes.data, iny.ctypes.data))
        # the reference's attributes: lists of 6 float64 [w*w] arrays, 1-based coordinates
        self.inXs = [inx[i * w * w:(i + 1) * w * w] for i in range(6)]
        self.inYs = [iny[i * w * w:(i + 1) * w * w] for i in range(6)]
        self.sx = self.sx.reshape(6, w, w)
        self.sy = self.sy.reshape(6, w, w)
        self._device = device
        self._packed_dev = {}

    # ------------------------------------------------------------------ device plumbing
    def _map_on(self, device):
        key = (device.type, device.index)
        m = self._packed_dev.get(key)
        if m is None:
            # uint32 payload carried in an int32 tensor (same bits)
            m = torch.from_numpy(self.packed.view(np.int32)).to(device)
            self._packed_dev[key] = m
        return m

    def _pick_device(self):
        if self._device is not None:
            return torch.device(self._device)
        if not torch.cuda.is_available():
            raise RuntimeError("Equi2Cube.to_cube needs a CUDA device (sm_100a); no CPU fallback")
        return torch.device("cuda", torch.cuda.current_device())

    def to_cube_tensor(self, frames, layout="NCHW", mean=None, std=None, out=None, denom=255.0):
        """frames [B,H,W,C] (or [H,W,C]) cuda -> faces [6B,C,w,w] ('NCHW') or [6B,w,w,C], float32.

        float32 frames are resampled as they are. uint8 frames (a decoded video frame before the
        reference's "/255.0", dataset_feat_extractor.py:131,142) are converted on the fly as
        float32(u8)/denom, bit-identical to converting first, at a quarter of the traffic.
        mean/std (sequences of C floats): fuse utils/utils.py:28-33 im_norm into the store."""
        if not isinstance(frames, torch.Tensor) or not frames.is_cuda:
            raise RuntimeError("to_cube_tensor expects a CUDA tensor (no CPU fallback)")
        if frames.dim() == 3:
            frames = frames.unsqueeze(0)
        is_u8 = frames.dtype == torch.uint8
        if not is_u8 and frames.dtype != torch.float32:
            frames = frames.float()
        frames = frames.contiguous()
        b, h, wi, c = frames.shape
        if (h, wi) != (self.input_height, self.input_width):
            raise ValueError("frame is %dx%d, maps were built for %dx%d" % (wi, h, self.input_width, self.input_height))
        w = self.output_width
        nchw = layout.upper() == "NCHW"
        shape = (6 * b, c, w, w) if nchw else (6 * b, w, w, c)
        if out is None:
            out = torch.empty(shape, dtype=torch.float32, device=frames.device)
        elif tuple(out.shape) != shape or not out.is_contiguous() or out.dtype != torch.float32:
            raise ValueError("out must be a contiguous float32 tensor of shape %s" % (shape,))
        mp = sp = None
        if mean is not None or std is not None:
            m_arr = np.ascontiguousarray(mean, dtype=np.float32)
            s_arr = np.ascontiguousarray(std, dtype=np.float32)
            if m_arr.size != c or s_arr.size != c:
                raise ValueError("mean/std need %d entries" % c)
            mp, sp = m_arr.ctypes.data, s_arr.ctypes.data
        pm = self._map_on(frames.device)
        lay = _lib.LAYOUT_NCHW if nchw else _lib.LAYOUT_NHWC
        with torch.cuda.device(frames.device):
            st = torch.cuda.current_stream().cuda_stream
            if is_u8:
                _lib.check(_lib.lib().cp360_e2c_fwd_u8(
                    frames.data_ptr(), pm.data_ptr(), out.data_ptr(), b, h, wi, c, w, lay, float(denom), mp, sp, st))
            else:
                _lib.check(_lib.lib().cp360_e2c_fwd(
                    frames.data_ptr(), pm.data_ptr(), out.data_ptr(), b, h, wi, c, w, lay, mp, sp, st))
        return out

    def to_padded_cube_tensor(self, frames, lrtd_pad, mean=None, std=None, out=None, denom=255.0):
        """frames [B,H,W,C] cuda (float32 or uint8) -> CubePad(lrtd_pad)(faces) [6B,C,w+pt+pd,w+pl+pr]
        in one kernel (cp360_e2c_cubepad_fwd): to_cube + im_norm + NHWC->NCHW + the CubePad(3) in front
        of conv1 (dataset_feat_extractor.py:145-157, resnet_cubic.py:116-117) without writing the
        unpadded faces. Bit-identical to CubePad(lrtd_pad)(to_cube_tensor(frames, mean=..., std=...))."""
        from .cube_pad import get_pad_size
        if not isinstance(frames, torch.Tensor) or not frames.is_cuda:
            raise RuntimeError("to_padded_cube_tensor expects a CUDA tensor (no CPU fallback)")
        if frames.dim() == 3:
            frames = frames.unsqueeze(0)
        is_u8 = frames.dtype == torch.uint8
        if not is_u8 and frames.dtype != torch.float32:
            frames = frames.float()
        frames = frames.contiguous()
        b, h, wi, c = frames.shape
        if (h, wi) != (self.input_height, self.input_width):
            raise ValueError("frame is %dx%d, maps were built for %dx%d" % (wi, h, self.input_width, self.input_height))
        pl, pr, pt, pd = get_pad_size(lrtd_pad)
        w = self.output_width
        shape = (6 * b, c, w + pt + pd, w + pl + pr)
        if out is None:
            out = torch.empty(shape, dtype=torch.float32, device=frames.device)
        elif tuple(out.shape) != shape or not out.is_contiguous() or out.dtype != torch.float32:
            raise ValueError("out must be a contiguous float32 tensor of shape %s" % (shape,))
        mp = sp = None
        if mean is not None or std is not None:
            m_arr = np.ascontiguousarray(mean, dtype=np.float32)
            s_arr = np.ascontiguousarray(std, dtype=np.float32)
            if m_arr.size != c or s_arr.size != c:
                raise ValueError("mean/std need %d entries" % c)
            mp, sp = m_arr.ctypes.data, s_arr.ctypes.data
        pm = self._map_on(frames.device)
        with torch.cuda.device(frames.device):
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.lib().cp360_e2c_cubepad_fwd(
                frames.data_ptr(), int(is_u8), pm.data_ptr(), out.data_ptr(), b, h, wi, c, w, pl, pr, pt, pd,
                float(denom), mp, sp, st))
        return out

    # ------------------------------------------------------------------ reference API
    def to_cube(self, in_image):
        """in_image ndarray [H,W,C] -> {0..5: ndarray[w,w,C] of in_image.dtype} (equi_to_cube.py:112-129).

        The GPU path computes in float32 (the BASELINE configuration); float64 input is rounded
        to float32 first, so results agree with the reference's float64 run to ~1e-7. Integer input is
        rounded to nearest and saturated on the way back (cv2 semantics to within 1 LSB)."""
        dev = self._pick_device()
        if isinstance(in_image, torch.Tensor):
            src, np_dtype = in_image, None
        else:
            arr = np.asarray(in_image)
            np_dtype = arr.dtype
            src = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float32))
        if src.dim() == 2:
            src = src.unsqueeze(-1)
        faces = self.to_cube_tensor(src.to(dev, non_blocking=True), layout="NHWC")
        if np_dtype is None:
            return {i: faces[i] for i in range(6)}
        host = faces.cpu().numpy()
        if np.issubdtype(np_dtype, np.integer):
            # integer frames (e.g. uint8): resampled in float32, then rounded to nearest and saturated like
            # cv2's saturate_cast — within 1 LSB of cv2.remap's own 8-bit fixed-point path (pinned by
            # tests/test_oracle_golden.py::test_e2c_integer_frames_within_one_lsb_of_cv2); the reference itself
            # only ever passes float frames (dataset_feat_extractor.py:131,142)
            info = np.iinfo(np_dtype)
            host = np.clip(np.rint(host), info.min, info.max)
        host = host.astype(np_dtype, copy=False)
        return {i: host[i] for i in range(6)}
