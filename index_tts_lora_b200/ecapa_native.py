"""Host side of the native speaker encoder (csrc/ecapa.cu): turns the parameters of the PyTorch ``ECAPA_TDNN`` module
(the checkpoint's ``speaker_encoder.*`` tensors, unchanged) into the kernel-ready form ``bvg_ecapa_desc`` asks for and
calls ``bvg_ecapa_forward``.  ``ECAPA_TDNN.forward(mel_ref, lens=None)`` (ECAPA_TDNN.py:543-581) -> ``[B, 1, emb]``."""
from __future__ import annotations

import ctypes as C
from typing import List

import torch

from . import _lib


def _bn_fold(bn: torch.nn.BatchNorm1d, dev):
    w, b = bn.weight.detach().float(), bn.bias.detach().float()
    m, v = bn.running_mean.detach().float(), bn.running_var.detach().float()
    sc = w / torch.sqrt(v + bn.eps)
    return sc.to(dev).contiguous(), (b - m * sc).to(dev).contiguous()


class NativeSpeakerEncoder:
    """One native handle per (module parameters, device).  Rebuilt by the owner when the parameters change."""

    def __init__(self, enc, device: torch.device):
        self.lib = _lib.load()
        self.device = device
        self.keep: List[torch.Tensor] = []          # the descriptor holds raw pointers into these
        d = _lib.BvgEcapaDesc()
        b0 = enc.blocks[0]
        d.in_channels = b0.conv.conv.in_channels
        d.channels = b0.conv.conv.out_channels
        d.scale = enc.blocks[1].res2net_block.scale
        d.se_channels = enc.blocks[1].se_block.conv1.conv.out_channels
        d.att_channels = enc.asp.tdnn.conv.conv.out_channels
        d.mfa_channels = enc.mfa.conv.conv.out_channels
        d.emb_dim = enc.fc.conv.out_channels
        if len(enc.blocks) != 4 or d.scale != 8:
            raise NotImplementedError("native speaker encoder: 1 TDNN + 3 SE-Res2Net blocks of scale 8 (ECAPA_TDNN.py:443-489)")
        self._tdnn(d.block0, b0)
        for i in range(3):
            blk, o = enc.blocks[i + 1], d.blocks[i]
            if blk.shortcut is not None:
                raise NotImplementedError("native speaker encoder: SE-Res2Net blocks with equal in / out channels")
            self._tdnn(o.tdnn1, blk.tdnn1)
            for k in range(7):
                self._tdnn(o.res2[k], blk.res2net_block.blocks[k])
            self._tdnn(o.tdnn2, blk.tdnn2)
            se = blk.se_block
            o.se_w1 = self._t(se.conv1.conv.weight.squeeze(-1))
            o.se_b1 = self._t(se.conv1.conv.bias)
            o.se_w2 = self._t(se.conv2.conv.weight.squeeze(-1))
            o.se_b2 = self._t(se.conv2.conv.bias)
        self._tdnn(d.mfa, enc.mfa)
        # ASP: tdnn's 1x1 weight [att][3*mfa] = [x part | mean part | std part] (cat order, ECAPA_TDNN.py:318)
        mf = d.mfa_channels
        w = enc.asp.tdnn.conv.conv.weight.detach().float().squeeze(-1)
        self._tdnn(d.asp_tdnn, enc.asp.tdnn, weight=w[:, :mf].unsqueeze(-1))
        d.asp_ctx_w = self._t(w[:, mf:])
        self._conv(d.asp_conv, enc.asp.conv.conv, None, relu=0)
        sc, sh = _bn_fold(enc.asp_bn.norm, device)
        self.keep += [sc, sh]
        d.asp_bn_scale, d.asp_bn_shift = sc.data_ptr(), sh.data_ptr()
        d.fc_w = self._t(enc.fc.conv.weight.squeeze(-1))
        d.fc_b = self._t(enc.fc.conv.bias)
        self.desc = d
        self.emb_dim = int(d.emb_dim)
        h = C.c_void_p()
        _lib.check(self.lib.bvg_ecapa_create(C.byref(d), device.index or 0, C.byref(h)), "bvg_ecapa_create")
        self.handle = h

    def _t(self, t: torch.Tensor) -> int:
        x = t.detach().float().to(self.device).contiguous()
        self.keep.append(x)
        return x.data_ptr()

    def _conv(self, o, conv, bn, relu, weight=None):
        w = (conv.weight.detach().float() if weight is None else weight)          # [cout][cin][k]
        o.w = self._t(w.permute(1, 2, 0))                                           # [cin][k][cout]
        o.bias = self._t(conv.bias)
        if bn is not None:
            sc, sh = _bn_fold(bn, self.device)
            self.keep += [sc, sh]
            o.bn_scale, o.bn_shift = sc.data_ptr(), sh.data_ptr()
        o.cin, o.cout, o.k = int(w.shape[1]), int(w.shape[0]), int(w.shape[2])
        o.dil, o.relu = int(conv.dilation[0]), int(relu)

    def _tdnn(self, o, blk, weight=None):
        self._conv(o, blk.conv.conv, blk.norm.norm, relu=1, weight=weight)

    def __call__(self, mel: torch.Tensor) -> torch.Tensor:
        """mel [B, Tm, num_mels] on self.device -> [B, 1, emb] fp32."""
        if mel.device != self.device or mel.dim() != 3 or mel.shape[-1] != int(self.desc.in_channels):
            raise ValueError(f"mel_ref must be [B, Tm, {int(self.desc.in_channels)}] on {self.device}, got {tuple(mel.shape)} on {mel.device}")
        if mel.dtype not in (torch.float32, torch.bfloat16, torch.float16):
            mel = mel.float()
        mel = mel.contiguous()
        B, Tm, _ = mel.shape
        out = torch.empty(B, 1, self.emb_dim, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.bvg_ecapa_forward(self.handle, mel.data_ptr(), _lib.torch_dtype_code(mel.dtype), B, Tm,
                                                  out.data_ptr(), _lib.stream_ptr(self.device)), "bvg_ecapa_forward")
        return out

    def __del__(self):
        try:
            if getattr(self, "handle", None) is not None:
                self.lib.bvg_ecapa_destroy(self.handle)
                self.handle = None
        except Exception:
            pass
