"""CubePad — host mirror of the reference's model/cube_pad.py (same names, same arguments).

    CubePad(lrtd_pad, use_gpu=True)            cube_pad.py:23-42
    CubePadding(lrtd_pad, use_gpu=True)        cube_pad.py:45-216   (one 6-face group)
    get_pad_size(lrtd_pad)                     cube_pad.py:12-20

forward(x[6N,C,H,W]) -> [6N,C,H+p_t+p_d,W+p_l+p_r] runs ONE kernel launch of libcp360
(cp360_cubepad_fwd) on the tensor's device and current stream, for the whole batch — the
reference's per-group Python loop, ~41 ATen launches and 8 synchronous index uploads per group
are gone. Differentiable (cp360_cubepad_bwd_f32) because temporal_model/train_temporal.py
back-propagates through it. CPU tensors are rejected: there is no fallback path.
"""
import numbers

import torch
import torch.nn as nn

from . import _lib


def get_pad_size(lrtd_pad):
    """int -> same pad on all sides; else [p_l, p_r, p_t, p_d] (cube_pad.py:12-20)."""
    if isinstance(lrtd_pad, numbers.Integral):
        p = int(lrtd_pad)
        return p, p, p, p
    p_l, p_r, p_t, p_d = (int(v) for v in lrtd_pad)
    return p_l, p_r, p_t, p_d


def cubepad_index_map(H, W, lrtd_pad):
    """Host: int32 [6,Ho,Wo] source index (face*H*W + y*W + x) of every output pixel."""
    import numpy as np
    import ctypes
    p_l, p_r, p_t, p_d = get_pad_size(lrtd_pad)
    lib = _lib.lib()
    ho, wo = ctypes.c_int(), ctypes.c_int()
    _lib.check(lib.cp360_cubepad_out_shape(H, W, p_l, p_r, p_t, p_d, ctypes.byref(ho), ctypes.byref(wo)))
    m = np.empty((6, ho.value, wo.value), dtype=np.int32)
    _lib.check(lib.cp360_cubepad_build_map(H, W, p_l, p_r, p_t, p_d, m.ctypes.data))
    return m


def _require_cuda(t, who):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s expects a torch.Tensor, got %s" % (who, type(t).__name__))
    if not t.is_cuda:
        raise RuntimeError("%s: tensor is on %s; this build runs only on CUDA (sm_100a) and has "
                           "no CPU fallback" % (who, t.device))


def cubepad_forward(x, pads, algo=_lib.ALGO_AUTO):
    """Raw launch: x [6N,C,H,W] cuda -> new contiguous [6N,C,Ho,Wo]."""
    _require_cuda(x, "CubePad")
    if x.dim() != 4:
        raise ValueError("CubePad expects [6N, C, H, W], got %s" % (tuple(x.shape),))
    p_l, p_r, p_t, p_d = pads
    n, c, h, w = x.shape
    if n % 6 != 0:
        # the reference prints this and calls exit() (cube_pad.py:33-35); an exception is kinder
        raise ValueError("CubePad size mismatch! batch %d is not a multiple of 6" % n)
    x = x.contiguous()
    y = torch.empty((n, c, h + p_t + p_d, w + p_l + p_r), dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib().cp360_cubepad_fwd_algo(
            x.data_ptr(), y.data_ptr(), n, c, h, w, p_l, p_r, p_t, p_d, x.element_size(), algo, st))
    return y


def _needs_grad(*ts):
    return torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in ts)


def autotune_cubepad(x, lrtd_pad, effort=1):
    """EXPLICIT tuning of one CubePad problem (cp360_cubepad_autotune): times candidate tilings on `x`, remembers the
    winner for later calls with the same shape on this device, returns (padded tensor, description of the choice).
    Allocates a flush buffer and synchronises — call it once per shape at start-up, never inside a CUDA graph
    capture. Shapes of the cubic ResNet-50 / ConvLSTM are already covered by the built-in table."""
    import ctypes
    _require_cuda(x, "autotune_cubepad")
    if x.dim() != 4 or x.dtype != torch.float32:
        raise ValueError("autotune_cubepad expects a float32 [6N, C, H, W] tensor")
    pads = get_pad_size(lrtd_pad)
    p_l, p_r, p_t, p_d = pads
    n, c, h, w = x.shape
    x = x.contiguous()
    y = torch.empty((n, c, h + p_t + p_d, w + p_l + p_r), dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib().cp360_cubepad_autotune(x.data_ptr(), y.data_ptr(), n, c, h, w, p_l, p_r, p_t, p_d,
                                                     int(effort), st))
        buf = ctypes.create_string_buffer(256)
        _lib.check(_lib.lib().cp360_cubepad_tune_info(n, c, h, w, p_l, p_r, p_t, p_d, buf, 256))
    return y, buf.value.decode()


def cubepad_fused(x, pads, scale=None, shift=None, relu=False, out=None, out_channel_offset=0):
    """CubePad(act(x * scale[c] + shift[c])) in ONE pass (fp32): the eval-mode BatchNorm affine and
    ReLU that precede CubePad in the cubic ResNet (model/resnet_cubic.py:89-92) never cost a tensor
    round trip of their own. `out` [6N,Cout,Ho,Wo] with Cout >= C lets several sources land in one
    padded tensor (see cubepad_cat). Separate multiply and add: bit-exact against numpy fp32.

    Differentiable in x, scale and shift (backward = cp360_cubepad_bwd_f32, then the ReLU mask and
    the affine's chain rule) — except with a caller-supplied `out` window, which autograd cannot
    track: that form raises when a gradient is required (use cubepad_cat, which is differentiable)."""
    if _needs_grad(x, scale, shift):
        if out is not None:
            raise RuntimeError("cubepad_fused(out=...) writes a channel window in place and is not differentiable; "
                               "use cubepad_cat([...]) or call it under torch.no_grad()")
        as_t = lambda v: None if v is None else (v if isinstance(v, torch.Tensor) else  # noqa: E731
                                                 torch.as_tensor(v, dtype=torch.float32, device=x.device))
        return _CubePadFusedFn.apply(x, as_t(scale), as_t(shift), tuple(pads), bool(relu))
    return _cubepad_fused_raw(x, pads, scale, shift, relu, out, out_channel_offset)


def _cubepad_fused_raw(x, pads, scale=None, shift=None, relu=False, out=None, out_channel_offset=0):
    _require_cuda(x, "cubepad_fused")
    if x.dim() != 4 or x.dtype != torch.float32:
        raise ValueError("cubepad_fused expects a float32 [6N, C, H, W] tensor")
    p_l, p_r, p_t, p_d = pads
    n, c, h, w = x.shape
    if n % 6 != 0:
        raise ValueError("CubePad size mismatch! batch %d is not a multiple of 6" % n)
    x = x.contiguous()
    ho, wo = h + p_t + p_d, w + p_l + p_r
    if out is None:
        if out_channel_offset:
            raise ValueError("out_channel_offset needs out")
        out = torch.empty((n, c, ho, wo), dtype=x.dtype, device=x.device)
    if (out.dtype != torch.float32 or not out.is_contiguous() or out.device != x.device or out.dim() != 4 or
            out.shape[0] != n or out.shape[2] != ho or out.shape[3] != wo):
        raise ValueError("out must be a contiguous float32 [%d, Cout, %d, %d] tensor on %s" % (n, ho, wo, x.device))

    def vec(v, name):
        if v is None:
            return None
        v = torch.as_tensor(v, dtype=torch.float32, device=x.device).contiguous()
        if v.numel() != c:
            raise ValueError("%s needs %d entries, got %d" % (name, c, v.numel()))
        return v
    scale, shift = vec(scale, "scale"), vec(shift, "shift")
    with torch.cuda.device(x.device):
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib().cp360_cubepad_fused_fwd(
            x.data_ptr(), out.data_ptr(), n, c, h, w, p_l, p_r, p_t, p_d,
            scale.data_ptr() if scale is not None else None, shift.data_ptr() if shift is not None else None,
            int(bool(relu)), out.shape[1], int(out_channel_offset), st))
    return out


def cubepad_cat(tensors, lrtd_pad):
    """CubePad(torch.cat(tensors, 1)) without materialising the concatenation (model/clstm.py:57-58):
    every source is read once and written straight into its channel window of the padded tensor.
    Differentiable: the ConvLSTM trains through this site (train_temporal.py:100-107,167-170), so each
    source receives cp360_cubepad_bwd_f32 of its channel window of the output gradient."""
    pads = get_pad_size(lrtd_pad)
    tensors = list(tensors)
    if _needs_grad(*tensors):
        return _CubePadCatFn.apply(pads, *tensors)
    return _cubepad_cat_raw(tensors, pads)


def _cubepad_cat_raw(tensors, pads):
    p_l, p_r, p_t, p_d = pads
    n, _, h, w = tensors[0].shape
    ctot = sum(int(t.shape[1]) for t in tensors)
    out = torch.empty((n, ctot, h + p_t + p_d, w + p_l + p_r), dtype=torch.float32, device=tensors[0].device)
    off = 0
    for t in tensors:
        if t.shape[0] != n or t.shape[2] != h or t.shape[3] != w:
            raise ValueError("cubepad_cat: tensors must agree in every dimension but channels")
        _cubepad_fused_raw(t, pads, out=out, out_channel_offset=off)
        off += int(t.shape[1])
    return out


def cubepad_bn_relu(x, bn, lrtd_pad, relu=True):
    """CubePad(relu(bn(x))) for an eval-mode nn.BatchNorm2d, folded into the pad kernel's epilogue."""
    if bn.training or bn.running_mean is None:
        raise ValueError("cubepad_bn_relu folds running statistics: call bn.eval() first")
    inv = torch.rsqrt(bn.running_var.float() + bn.eps)
    scale = inv * (bn.weight.float() if bn.affine else 1.0)
    shift = (bn.bias.float() if bn.affine else 0.0) - bn.running_mean.float() * scale
    return cubepad_fused(x, get_pad_size(lrtd_pad), scale=scale, shift=shift, relu=relu)


def cubepad_backward(gy, pads, in_hw):
    _require_cuda(gy, "CubePad.backward")
    p_l, p_r, p_t, p_d = pads
    h, w = in_hw
    n, c = gy.shape[0], gy.shape[1]
    g32 = gy.contiguous() if gy.dtype == torch.float32 else gy.float().contiguous()
    gx = torch.empty((n, c, h, w), dtype=torch.float32, device=gy.device)
    with torch.cuda.device(gy.device):
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib().cp360_cubepad_bwd_f32(
            g32.data_ptr(), gx.data_ptr(), n, c, h, w, p_l, p_r, p_t, p_d, st))
    return gx if gy.dtype == torch.float32 else gx.to(gy.dtype)


class _CubePadFusedFn(torch.autograd.Function):
    """y = CubePad(act(x * scale + shift)); dL/dx, dL/dscale, dL/dshift through the pad's transpose."""

    @staticmethod
    def forward(ctx, x, scale, shift, pads, relu):
        ctx.pads, ctx.relu, ctx.in_hw = pads, relu, (x.shape[2], x.shape[3])
        ctx.save_for_backward(x, scale if isinstance(scale, torch.Tensor) else None,
                              shift if isinstance(shift, torch.Tensor) else None)
        return _cubepad_fused_raw(x.detach(), pads, None if scale is None else scale.detach(),
                                  None if shift is None else shift.detach(), relu)

    @staticmethod
    def backward(ctx, gy):
        x, scale, shift = ctx.saved_tensors
        ga = cubepad_backward(gy, ctx.pads, ctx.in_hw)              # gradient w.r.t. act(z), z = x * scale + shift
        c = x.shape[1]
        sc = None if scale is None else scale.float().reshape(1, c, 1, 1)
        if ctx.relu:
            z = x if sc is None else x * sc
            if shift is not None:
                z = z + shift.float().reshape(1, c, 1, 1)
            ga = ga * (z > 0)
        gx = gscale = gshift = None
        if ctx.needs_input_grad[0]:
            gx = ga if sc is None else ga * sc
        if scale is not None and ctx.needs_input_grad[1]:
            gscale = (ga * x).sum((0, 2, 3)).reshape(scale.shape).to(scale.dtype)
        if shift is not None and ctx.needs_input_grad[2]:
            gshift = ga.sum((0, 2, 3)).reshape(shift.shape).to(shift.dtype)
        return gx, gscale, gshift, None, None


class _CubePadCatFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pads, *tensors):
        ctx.pads = pads
        ctx.in_hw = (tensors[0].shape[2], tensors[0].shape[3])
        ctx.channels = [int(t.shape[1]) for t in tensors]
        return _cubepad_cat_raw([t.detach() for t in tensors], pads)

    @staticmethod
    def backward(ctx, gy):
        grads, off = [], 0
        for i, c in enumerate(ctx.channels):
            g = None
            if ctx.needs_input_grad[1 + i]:
                g = cubepad_backward(gy[:, off:off + c].contiguous(), ctx.pads, ctx.in_hw)
            grads.append(g)
            off += c
        return (None, *grads)


class _CubePadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, pads):
        ctx.pads = pads
        ctx.in_hw = (x.shape[2], x.shape[3])
        return cubepad_forward(x, pads)

    @staticmethod
    def backward(ctx, gy):
        return cubepad_backward(gy, ctx.pads, ctx.in_hw), None


class CubePad(nn.Module):
    """Drop-in for the reference's CubePad (cube_pad.py:23-42). No parameters or buffers."""

    def __init__(self, lrtd_pad, use_gpu=True):
        super().__init__()
        if not use_gpu:
            raise RuntimeError("CubePad(use_gpu=False): this build is CUDA-only (no CPU fallback)")
        self.pads = get_pad_size(lrtd_pad)
        self.p_l, self.p_r, self.p_t, self.p_d = self.pads

    def forward(self, x):
        """x [6N,C,H,W] -> [6N,C,H+p_t+p_d,W+p_l+p_r]; face order B,D,F,L,R,T per group of 6."""
        if x.requires_grad and torch.is_grad_enabled():
            return _CubePadFn.apply(x, self.pads)
        return cubepad_forward(x, self.pads)

    def extra_repr(self):
        return "lrtd_pad=[%d, %d, %d, %d]" % self.pads


class CubePadding(CubePad):
    """The reference's inner module (cube_pad.py:45-216) padded exactly one 6-face group; the
    kernel handles any number of groups, so this is the same operator under the old name."""

    def forward(self, x):
        if x.dim() == 4 and x.shape[0] != 6:
            raise ValueError("CubePadding expects exactly 6 faces, got %d" % x.shape[0])
        return super().forward(x)
