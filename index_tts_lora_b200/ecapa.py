"""ECAPA-TDNN speaker encoder, PyTorch, state-dict compatible with the reference.

Row a13 of SURVEY.md §8: ``speaker_encoder(mel_ref, lens)`` at ``models.py:204`` maps
``[B, Tm, num_mels] -> [B, 1, 512]``.  north_star keeps it outside the CUDA kernel list (7.2 M
parameters, ~21 ms of 5.3 s on CPU, once per utterance, independent of audio length), so it
stays PyTorch here; what matters is that ``speaker_encoder.*`` checkpoint keys load unchanged
and the numbers match ``indextts/BigVGAN/ECAPA_TDNN.py:429-581``.

Written from the model's published structure (TDNN -> 3x SE-Res2Net -> MFA -> attentive
statistics pooling -> BN -> FC); every conv is "same"-padded with reflect padding
(``nnet/CNN.py:367,458-487``) and every TDNN block is conv -> ReLU -> BatchNorm
(``ECAPA_TDNN.py:122-124``).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class _SameConv(nn.Module):
    """[B,C,T] conv with reflect 'same' padding; parameters live under ``.conv`` (CNN.py:372)."""

    def __init__(self, cin, cout, k, dilation=1):
        super().__init__()
        self.pad = dilation * (k - 1) // 2
        self.conv = nn.Conv1d(cin, cout, k, dilation=dilation)

    def forward(self, x):
        if self.pad:
            x = F.pad(x, (self.pad, self.pad), mode="reflect")
        return self.conv(x)


class _BN(nn.Module):
    """BatchNorm1d over channels of [B,C,T]; parameters under ``.norm`` (normalization.py:62)."""

    def __init__(self, c):
        super().__init__()
        self.norm = nn.BatchNorm1d(c)

    def forward(self, x):
        return self.norm(x)


class _TDNN(nn.Module):
    def __init__(self, cin, cout, k, dilation):
        super().__init__()
        self.conv = _SameConv(cin, cout, k, dilation)
        self.activation = nn.ReLU()
        self.norm = _BN(cout)

    def forward(self, x):
        return self.norm(self.activation(self.conv(x)))


class _Res2Net(nn.Module):
    def __init__(self, c, scale, k, dilation):
        super().__init__()
        self.scale = scale
        w = c // scale
        self.blocks = nn.ModuleList(_TDNN(w, w, k, dilation) for _ in range(scale - 1))

    def forward(self, x):
        outs, prev = [], None
        for i, xi in enumerate(torch.chunk(x, self.scale, dim=1)):
            if i == 0:
                prev = xi
            elif i == 1:
                prev = self.blocks[0](xi)
            else:
                prev = self.blocks[i - 1](xi + prev)
            outs.append(prev)
        return torch.cat(outs, dim=1)


def _len_mask(lengths, L, device):
    # ECAPA_TDNN.py:16-61 — relative lengths in (0,1] -> boolean [B,1,L]
    n = lengths.to(device) * L
    return (torch.arange(L, device=device).unsqueeze(0) < n.unsqueeze(1)).unsqueeze(1)


class _SE(nn.Module):
    def __init__(self, c, se):
        super().__init__()
        self.conv1 = _SameConv(c, se, 1)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = _SameConv(se, c, 1)
        self.sigmoid = nn.Sigmoid()

    def forward(self, x, lengths=None):
        if lengths is not None:
            m = _len_mask(lengths, x.shape[-1], x.device).to(x.dtype)
            s = (x * m).sum(2, keepdim=True) / m.sum(2, keepdim=True)
        else:
            s = x.mean(2, keepdim=True)
        return self.sigmoid(self.conv2(self.relu(self.conv1(s)))) * x


class _SERes2Net(nn.Module):
    def __init__(self, cin, cout, scale, se, k, dilation):
        super().__init__()
        self.tdnn1 = _TDNN(cin, cout, 1, 1)
        self.res2net_block = _Res2Net(cout, scale, k, dilation)
        self.tdnn2 = _TDNN(cout, cout, 1, 1)
        self.se_block = _SE(cout, se)
        self.shortcut = _SameConv(cin, cout, 1) if cin != cout else None

    def forward(self, x, lengths=None):
        r = self.shortcut(x) if self.shortcut is not None else x
        x = self.tdnn2(self.res2net_block(self.tdnn1(x)))
        return self.se_block(x, lengths) + r


class _ASP(nn.Module):
    """Attentive statistics pooling with global context (ECAPA_TDNN.py:245-338)."""

    def __init__(self, c, att):
        super().__init__()
        self.eps = 1e-12
        self.tdnn = _TDNN(c * 3, att, 1, 1)
        self.tanh = nn.Tanh()
        self.conv = _SameConv(att, c, 1)

    def _stats(self, x, w):
        mean = (w * x).sum(2)
        std = torch.sqrt((w * (x - mean.unsqueeze(2)).pow(2)).sum(2).clamp(self.eps))
        return mean, std

    def forward(self, x, lengths=None):
        L = x.shape[-1]
        if lengths is None:
            lengths = torch.ones(x.shape[0], device=x.device)
        mask = _len_mask(lengths, L, x.device)
        mf = mask.to(x.dtype)
        mean, std = self._stats(x, mf / mf.sum(2, keepdim=True).float())
        ctx = torch.cat([x, mean.unsqueeze(2).expand(-1, -1, L), std.unsqueeze(2).expand(-1, -1, L)], 1)
        att = self.conv(self.tanh(self.tdnn(ctx)))
        att = F.softmax(att.masked_fill(~mask, float("-inf")), dim=2)
        mean, std = self._stats(x, att)
        return torch.cat([mean, std], 1).unsqueeze(2)


class ECAPA_TDNN(nn.Module):
    """``ECAPA_TDNN(num_mels, lin_neurons=speaker_embedding_dim)`` as built at models.py:193."""

    def __init__(self, input_size, lin_neurons=192, channels=(512, 512, 512, 512, 1536),
                 kernel_sizes=(5, 3, 3, 3, 1), dilations=(1, 2, 3, 4, 1),
                 attention_channels=128, res2net_scale=8, se_channels=128):
        super().__init__()
        self.blocks = nn.ModuleList([_TDNN(input_size, channels[0], kernel_sizes[0], dilations[0])])
        for i in range(1, len(channels) - 1):
            self.blocks.append(_SERes2Net(channels[i - 1], channels[i], res2net_scale,
                                          se_channels, kernel_sizes[i], dilations[i]))
        self.mfa = _TDNN(channels[-2] * (len(channels) - 2), channels[-1], kernel_sizes[-1],
                         dilations[-1])
        self.asp = _ASP(channels[-1], attention_channels)
        self.asp_bn = _BN(channels[-1] * 2)
        self.fc = _SameConv(channels[-1] * 2, lin_neurons, 1)

    def forward(self, x, lengths=None):
        x = x.transpose(1, 2)
        feats = []
        for i, blk in enumerate(self.blocks):
            x = blk(x) if i == 0 else blk(x, lengths)
            feats.append(x)
        x = self.mfa(torch.cat(feats[1:], dim=1))
        x = self.fc(self.asp_bn(self.asp(x, lengths)))
        return x.transpose(1, 2)
