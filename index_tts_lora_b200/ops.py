"""Per-op wrappers over the C ABI (include/bvg.h) taking torch CUDA tensors.

``activation1d`` supersedes ``anti_alias_activation_cuda.forward`` /
``FusedAntiAliasActivation`` (alias_free_activation/cuda/activation1d.py:13-31): same argument
meaning (input [B,C,T], up/down filter, alpha, beta), runs on the current stream, raises
RuntimeError on failure — but never falls back to a torch path."""
from __future__ import annotations

import torch

from . import _lib


def _req_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("B200-native op called with a non-CUDA tensor; there is no CPU fallback")


def _f32(t, dev):
    return t.detach().to(device=dev, dtype=torch.float32).contiguous()


def activation1d(x, up_filter, down_filter, alpha, beta, logscale=True):
    _req_cuda(x)
    lib = _lib.load()
    x = x.contiguous()
    B, C, T = x.shape
    y = torch.empty_like(x)
    up, dn = _f32(up_filter, x.device).reshape(-1), _f32(down_filter, x.device).reshape(-1)
    a, b = _f32(alpha, x.device), _f32(beta, x.device)
    if up.numel() != 12 or dn.numel() != 12:
        raise ValueError("filters must have 12 taps")
    with torch.cuda.device(x.device):
        _lib.check(lib.bvg_activation1d(x.data_ptr(), y.data_ptr(), _lib.torch_dtype_code(x.dtype),
                                        B, C, T, up.data_ptr(), dn.data_ptr(), a.data_ptr(),
                                        b.data_ptr(), int(bool(logscale)), _lib.stream_ptr(x.device)),
                   "bvg_activation1d")
    return y


def amp_layer(x, conv, act=None, resid=None, precision="fp32"):
    """y = conv1d(act1d(x)) [+ resid] as ONE fused launch.  ``conv`` is a (weight-normed)
    torch Conv1d holding the parameters, ``act`` an Activation1d container or None."""
    from .models import folded_weight

    _req_cuda(x, resid)
    lib = _lib.load()
    dev = x.device
    xf = x.float().contiguous()
    w = folded_weight(conv).to(dev).contiguous()
    bias = _f32(conv.bias, dev)
    B, Cin, T = xf.shape
    Cout, _, k = w.shape
    y = torch.empty(B, Cout, T, device=dev, dtype=torch.float32)
    r = resid.float().contiguous() if resid is not None else None
    if act is not None:
        up = _f32(act.upsample.filter, dev).reshape(-1)
        dn = _f32(act.downsample.lowpass.filter, dev).reshape(-1)
        a, b = _f32(act.act.alpha, dev), _f32(act.act.beta, dev)
        ls = int(bool(act.act.alpha_logscale))
        ptrs = (up.data_ptr(), dn.data_ptr(), a.data_ptr(), b.data_ptr())
    else:
        ptrs, ls = (None, None, None, None), 1
    with torch.cuda.device(dev):
        _lib.check(lib.bvg_amp_layer(xf.data_ptr(), y.data_ptr(), r.data_ptr() if r is not None else None,
                                     B, Cin, Cout, T, w.data_ptr(), bias.data_ptr(), k,
                                     int(conv.dilation[0]), int(act is not None), *ptrs, ls,
                                     {"fp32": _lib.PREC_F32, "bf16": _lib.PREC_BF16}[precision],
                                     _lib.stream_ptr(dev)), "bvg_amp_layer")
    return y.to(x.dtype)


def conv_transpose1d(x, convt, precision="fp32"):
    from .models import folded_weight

    _req_cuda(x)
    lib = _lib.load()
    dev = x.device
    xf = x.float().contiguous()
    w = folded_weight(convt).to(dev).contiguous()          # [C_in, C_out, k]
    bias = _f32(convt.bias, dev)
    B, Cin, T = xf.shape
    _, Cout, k = w.shape
    u = int(convt.stride[0])
    y = torch.empty(B, Cout, T * u, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        _lib.check(lib.bvg_conv_transpose1d(xf.data_ptr(), y.data_ptr(), B, Cin, Cout, T, w.data_ptr(),
                                            bias.data_ptr(), k, u,
                                            {"fp32": _lib.PREC_F32, "bf16": _lib.PREC_BF16}[precision],
                                            _lib.stream_ptr(dev)), "bvg_conv_transpose1d")
    return y.to(x.dtype)
