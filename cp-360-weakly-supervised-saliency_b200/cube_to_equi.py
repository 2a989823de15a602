"""Cube2Equi — host mirror of the reference's utils/cube_to_equi.py:11-91.

    c2e = Cube2Equi(input_w)                   # builds face_map / out_coord (host, once)
    equi = c2e.to_equi_nn(cube)                # [6,C,w,w] -> cuda [1,C,2w,4w]     (reference API)
    sal  = c2e.to_equi_max(cube)               # [6B,C,w,w] -> [B,2w,4w] fused channel max (addition)
    equi = c2e.to_equi_cv2(cube_numpy)         # [6,C,w,w] -> numpy [C,2w,4w], bicubic (:68-91)

The unmodified reference calls F.grid_sample without ``align_corners``; under the installed
torch (2.11) that means align_corners=False, which is therefore the default here. Pass
``align_corners=True`` for the torch<=1.2 behaviour the reference was written against.
"""
import os

import numpy as np
import torch

from . import _lib


class Cube2Equi:
    def __init__(self, input_w, align_corners=False):
        w = int(input_w)
        self.input_w = w
        self.align_corners = bool(align_corners)
        face = np.empty((2 * w, 4 * w), dtype=np.int8)
        coord = np.empty((2 * w, 4 * w, 2), dtype=np.float64)
        _lib.check(_lib.lib().cp360_c2e_build_map(w, face.ctypes.data, coord.ctypes.data))
        self.out_coord = coord                       # reference attribute (float64 [2w,4w,2])
        self.face_map = face.astype(np.float64)      # reference attribute (float64 [2w,4w], 0..5)
        self.taps = np.empty(8 * w * w, dtype=np.uint32)
        self.weights = np.empty((8 * w * w, 4), dtype=np.float32)
        m = np.zeros(1, dtype=np.float32)
        _lib.check(_lib.lib().cp360_c2e_build_plan(
            w, int(self.align_corners), self.taps.ctypes.data, self.weights.ctypes.data, m.ctypes.data))
        self.M = float(m[0])                         # torch.max(gridf), cube_to_equi.py:58
        self._plan_dev = {}

    def _plan_on(self, device):
        key = (device.type, device.index)
        p = self._plan_dev.get(key)
        if p is None:
            p = (torch.from_numpy(self.taps.view(np.int32)).to(device),
                 torch.from_numpy(self.weights).to(device))
            self._plan_dev[key] = p
        return p

    def _bwd_plan_on(self, device):
        """Transposed plan (cp360_c2e_build_bwd_plan) on `device`: (offsets, pixels, weights)."""
        key = ("bwd", device.type, device.index)
        p = self._plan_dev.get(key)
        if p is None:
            w = self.input_w
            offs = np.empty(6 * w * w + 1, dtype=np.int32)
            lib = _lib.lib()
            _lib.check(lib.cp360_c2e_build_bwd_plan(w, int(self.align_corners), offs.ctypes.data, None, None))
            n = int(offs[-1])
            self._bwd_entries = n
            pix, wts = np.empty(max(n, 1), dtype=np.int32), np.empty(max(n, 1), dtype=np.float32)
            _lib.check(lib.cp360_c2e_build_bwd_plan(w, int(self.align_corners), offs.ctypes.data, pix.ctypes.data, wts.ctypes.data))
            p = tuple(torch.from_numpy(a).to(device) for a in (offs, pix, wts))
            self._plan_dev[key] = p
        return p

    def _prepare(self, input_data):
        if isinstance(input_data, np.ndarray):
            # dataset_feat_extractor.py:174 hands over a numpy array
            if not torch.cuda.is_available():
                raise RuntimeError("Cube2Equi needs a CUDA device (sm_100a); no CPU fallback")
            input_data = torch.from_numpy(np.ascontiguousarray(input_data, dtype=np.float32)).cuda()
        if not isinstance(input_data, torch.Tensor) or not input_data.is_cuda:
            raise RuntimeError("Cube2Equi expects a CUDA tensor (no CPU fallback)")
        if input_data.dim() != 4 or input_data.shape[0] % 6 or input_data.shape[2] != self.input_w \
                or input_data.shape[3] != self.input_w:
            raise ValueError("expected [6B, C, %d, %d], got %s" % (self.input_w, self.input_w, tuple(input_data.shape)))
        if input_data.dtype != torch.float32:
            input_data = input_data.float()
        return input_data.contiguous()

    def _launch(self, name, src, dst, b, c):
        taps, wts = self._plan_on(src.device)
        with torch.cuda.device(src.device):
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(getattr(_lib.lib(), name)(
                src.data_ptr(), taps.data_ptr(), wts.data_ptr(), dst.data_ptr(), b, c, self.input_w, st))
        return dst

    def to_equi_nn(self, input_data):
        """[6B,C,w,w] -> [B,C,2w,4w] float32 on the input's device (B=1 in the reference)."""
        x = self._prepare(input_data)
        if x.requires_grad and torch.is_grad_enabled():
            return _C2EFn.apply(x, self)
        return self._forward(x)

    def _forward(self, x):
        b, c, w = x.shape[0] // 6, x.shape[1], self.input_w
        out = torch.empty((b, c, 2 * w, 4 * w), dtype=torch.float32, device=x.device)
        return self._launch("cp360_c2e_fwd", x, out, b, c)

    def _backward(self, gout):
        b, c, w = gout.shape[0], gout.shape[1], self.input_w
        g = gout.float().contiguous()
        gx = torch.empty((6 * b, c, w, w), dtype=torch.float32, device=g.device)
        offs, pix, wts = self._bwd_plan_on(g.device)
        with torch.cuda.device(g.device):
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.lib().cp360_c2e_bwd(g.data_ptr(), offs.data_ptr(), pix.data_ptr(), wts.data_ptr(), self._bwd_entries,
                                                gx.data_ptr(), b, c, w, st))
        return gx

    def to_equi_max(self, input_data, out=None):
        """Fused back-projection + channel max: [6B,C,w,w] -> [B,2w,4w]
        (== torch.max(to_equi_nn(x), 1)[0], test_temporal.py:82-84). Differentiable: when the
        input requires grad the arg-max channel is recorded and the backward scatters the map's
        gradient to that channel's four taps (train_temporal.py:105-107)."""
        x = self._prepare(input_data)
        if x.requires_grad and torch.is_grad_enabled():
            y = _C2EMaxFn.apply(x, self)
            return y if out is None else out.copy_(y)
        b, c, w = x.shape[0] // 6, x.shape[1], self.input_w
        if out is None:
            out = torch.empty((b, 2 * w, 4 * w), dtype=torch.float32, device=x.device)
        return self._launch("cp360_c2e_max_fwd", x, out, b, c)

    def to_equi_max_with_indices(self, input_data):
        """(sal [B,2w,4w] float32, argmax [B,2w,4w] int32) == torch.max(to_equi_nn(x), 1)."""
        x = self._prepare(input_data)
        b, c, w = x.shape[0] // 6, x.shape[1], self.input_w
        sal = torch.empty((b, 2 * w, 4 * w), dtype=torch.float32, device=x.device)
        arg = torch.empty((b, 2 * w, 4 * w), dtype=torch.int32, device=x.device)
        # the cluster kernel (w <= 16, 16 B-aligned input, channel count a multiple of the bulk-copy quantum) needs no
        # scratch; the atomic-key path every other case takes does
        q = 1
        while (q * w * w) % 4:
            q <<= 1
        clustered = w <= 16 and x.data_ptr() % 16 == 0 and c % q == 0 and os.environ.get("CP360_C2E_CLUSTER", "1") != "0"
        scratch = None if clustered else torch.empty((b, 2 * w, 4 * w), dtype=torch.int64, device=x.device)
        taps, wts = self._plan_on(x.device)
        with torch.cuda.device(x.device):
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.lib().cp360_c2e_max_arg_fwd(
                x.data_ptr(), taps.data_ptr(), wts.data_ptr(), sal.data_ptr(), arg.data_ptr(),
                scratch.data_ptr() if scratch is not None else None, b, c, w, st))
        return sal, arg

    def _max_backward(self, gsal, arg, c):
        b, w = gsal.shape[0], self.input_w
        g = gsal.float().contiguous()
        gx = torch.empty((6 * b, c, w, w), dtype=torch.float32, device=g.device)
        offs, pix, wts = self._bwd_plan_on(g.device)
        with torch.cuda.device(g.device):
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.lib().cp360_c2e_max_bwd(
                g.data_ptr(), arg.data_ptr(), offs.data_ptr(), pix.data_ptr(), wts.data_ptr(), gx.data_ptr(), b, c, w, st))
        return gx

    def _cubic_plan_on(self, device):
        key = ("cubic", device.type, device.index)
        p = self._plan_dev.get(key)
        if p is None:
            taps = np.empty(8 * self.input_w * self.input_w, dtype=np.uint32)
            _lib.check(_lib.lib().cp360_c2e_build_cubic_plan(self.input_w, taps.ctypes.data))
            p = torch.from_numpy(taps.view(np.int32)).to(device)
            self._plan_dev[key] = p
        return p

    def to_equi_cv2(self, input_data):
        """Bicubic back-projection with cv2.remap(INTER_CUBIC) arithmetic (cube_to_equi.py:68-91).

        numpy [6,C,w,w] -> numpy float32 [C,2w,4w] (the reference's signature; it hard-codes
        C = 1000, here any C). A CUDA tensor [6B,C,w,w] returns a CUDA tensor [B,C,2w,4w]."""
        as_numpy = isinstance(input_data, np.ndarray)
        x = self._prepare(input_data)
        b, c, w = x.shape[0] // 6, x.shape[1], self.input_w
        taps = self._cubic_plan_on(x.device)
        out = torch.empty((b, c, 2 * w, 4 * w), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.lib().cp360_c2e_cubic_fwd(x.data_ptr(), taps.data_ptr(), out.data_ptr(), b, c, w, st))
        if as_numpy:
            if b != 1:
                raise ValueError("numpy input is one cube [6,C,w,w] (cube_to_equi.py:68-74)")
            return out[0].cpu().numpy()
        return out


class _C2EFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, op):
        ctx.op = op
        return op._forward(x)

    @staticmethod
    def backward(ctx, gout):
        return ctx.op._backward(gout), None


class _C2EMaxFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, op):
        sal, arg = op.to_equi_max_with_indices(x.detach())
        ctx.op, ctx.channels = op, x.shape[1]
        ctx.save_for_backward(arg)
        ctx.mark_non_differentiable(arg)
        return sal

    @staticmethod
    def backward(ctx, gsal):
        (arg,) = ctx.saved_tensors
        return ctx.op._max_backward(gsal, arg, ctx.channels), None
