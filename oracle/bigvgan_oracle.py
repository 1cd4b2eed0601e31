"""CPU ORACLE for the BigVGAN decode path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this file.  The product (``index_tts_lora_b200``) never
does; it fails loudly when ``libbvg.so`` is missing instead of falling back to this.

What it is: a functional fp32 (or fp64) restatement, on torch CPU tensors, of
``indextts/BigVGAN/models.py:203-252`` and the ops under it.  It is *not* a copy of the
reference modules: the anti-aliased activation is evaluated through the closed polyphase form
of SURVEY.md §7 (gather + dot products) rather than the reference's
pad -> grouped conv_transpose1d -> crop -> ... -> pad -> grouped strided conv1d pipeline, the
dense convolutions call ``torch.nn.functional`` directly on folded weights, and everything is
keyed by the reference's state-dict names so the same ``generator`` dict drives both.

Parity pin (the reference's own tests hold no vectors for this path): the
oracle is checked against golden outputs of the REAL reference module
(``indextts.BigVGAN.models.BigVGAN``, ``use_cuda_kernel=False``) generated in the build
container by ``tests/golden/make_golden.py`` and committed under ``tests/golden/``.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------- filter design
def kaiser_sinc_filter(cutoff: float = 0.25, half_width: float = 0.3, taps: int = 12,
                       dtype=torch.float32) -> Tensor:
    """alias_free_torch/filter.py:29-58 — Kaiser-windowed sinc low-pass, unit DC gain.

    UpSample1d / DownSample1d both call it with cutoff 0.5/ratio, half_width 0.6/ratio,
    kernel 12 (resample.py:19-21, 42-45), so the up and down taps are identical.
    """
    half = taps // 2
    atten = 2.285 * (half - 1) * math.pi * (4 * half_width) + 7.95
    if atten > 50.0:
        beta = 0.1102 * (atten - 8.7)
    elif atten >= 21.0:
        beta = 0.5842 * (atten - 21.0) ** 0.4 + 0.07886 * (atten - 21.0)
    else:
        beta = 0.0
    win = torch.kaiser_window(taps, periodic=False, beta=beta, dtype=torch.float32)
    if taps % 2 == 0:
        t = torch.arange(-half, half, dtype=torch.float32) + 0.5
    else:
        t = torch.arange(taps, dtype=torch.float32) - half
    f = 2 * cutoff * win * torch.sinc(2 * cutoff * t)
    f = f / f.sum()
    return f.to(dtype)


# --------------------------------------------------------------------------- activation
def snake_beta(y: Tensor, alpha: Tensor, beta: Tensor, logscale: bool = True) -> Tensor:
    """activations.py:109-122:  y + sin(y*a)^2 / (b + 1e-9), a = e^alpha, b = e^beta."""
    a = alpha.view(1, -1, 1)
    b = beta.view(1, -1, 1)
    if logscale:
        a, b = torch.exp(a), torch.exp(b)
    return y + (1.0 / (b + 1e-9)) * torch.sin(y * a).pow(2)


def upsample2x(x: Tensor, f: Tensor) -> Tensor:
    """resample.py:25-33 in polyphase form.

    y[2m]   = 2 * sum_{i<6} f[11-2i] * x[clamp(m-3+i)]
    y[2m+1] = 2 * sum_{i<6} f[10-2i] * x[clamp(m-2+i)]
    (replicate pad 5/5, zero-stuffed transposed conv with stride 2, x2 gain, crop 15/15).
    """
    B, C, T = x.shape
    xp = F.pad(x, (3, 3), mode="replicate")            # xp[j] = x[clamp(j-3)]
    fe = f[[11, 9, 7, 5, 3, 1]]
    fo = f[[10, 8, 6, 4, 2, 0]]
    ye = torch.zeros_like(x)
    yo = torch.zeros_like(x)
    for i in range(6):
        ye = ye + fe[i] * xp[..., i:i + T]              # x[m-3+i]
        yo = yo + fo[i] * xp[..., i + 1:i + 1 + T]      # x[m-2+i]
    y = torch.stack((ye, yo), dim=-1).reshape(B, C, 2 * T)
    return 2.0 * y


def downsample2x(s: Tensor, f: Tensor) -> Tensor:
    """filter.py:87-96 + resample.py:46-49:  z[m] = sum_k f[k] * s[clamp(2m + k - 5)]
    (replicate pad 5 left / 6 right — of the ACTIVATED signal — then stride-2 FIR)."""
    n2 = s.shape[-1]
    sp = F.pad(s, (5, 6), mode="replicate")
    T = n2 // 2
    z = torch.zeros(s.shape[:-1] + (T,), dtype=s.dtype)
    for k in range(12):
        z = z + f[k] * sp[..., k:k + 2 * T:2]
    return z


def activation1d(x: Tensor, alpha: Tensor, beta: Tensor, up_f: Tensor, dn_f: Tensor,
                 logscale: bool = True) -> Tensor:
    """alias_free_torch/act.py:24-29: up-FIR x2 -> SnakeBeta -> down-FIR x2."""
    return downsample2x(snake_beta(upsample2x(x, up_f.reshape(-1)), alpha, beta, logscale),
                        dn_f.reshape(-1))


# --------------------------------------------------------------------------- weights
def fold_weight_norm(g: Tensor, v: Tensor) -> Tensor:
    """torch weight_norm with dim=0: w = g * v / ||v||, norm over all dims but 0.
    (For ConvTranspose1d dim 0 is C_in, so g is [C_in,1,1] — SURVEY §7.)"""
    n = v.reshape(v.shape[0], -1).norm(dim=1).view(-1, *([1] * (v.dim() - 1)))
    return v * (g / n)


def folded(sd: Dict[str, Tensor], prefix: str) -> Tensor:
    """Return `<prefix>.weight` from either key set (models.py:254-262 / infer.py:409)."""
    if prefix + ".weight" in sd:
        return sd[prefix + ".weight"]
    return fold_weight_norm(sd[prefix + ".weight_g"], sd[prefix + ".weight_v"])


def fold_state_dict(sd: Dict[str, Tensor]) -> Dict[str, Tensor]:
    out = {}
    for k, v in sd.items():
        if k.endswith(".weight_g"):
            p = k[: -len(".weight_g")]
            out[p + ".weight"] = fold_weight_norm(v, sd[p + ".weight_v"])
        elif k.endswith(".weight_v"):
            continue
        else:
            out[k] = v
    return out


# --------------------------------------------------------------------------- layers
def amp_layer(x: Tensor, sd, act_prefix: str, conv_prefix: str, k: int, d: int,
              logscale: bool = True) -> Tensor:
    """One `xt = conv(act(x))` step of AMPBlock1.forward (models.py:68-71)."""
    z = activation1d(x, sd[act_prefix + ".act.alpha"], sd[act_prefix + ".act.beta"],
                     sd[act_prefix + ".upsample.filter"],
                     sd[act_prefix + ".downsample.lowpass.filter"], logscale)
    return F.conv1d(z, folded(sd, conv_prefix), sd[conv_prefix + ".bias"], dilation=d,
                    padding=(k * d - d) // 2)          # get_padding, utils.py:59-60


def amp_block1(x: Tensor, sd, n: int, k: int, dils, logscale: bool = True) -> Tensor:
    """AMPBlock1.forward (models.py:65-74) for resblocks[n]."""
    p = f"resblocks.{n}"
    for m, d in enumerate(dils):
        xt = amp_layer(x, sd, f"{p}.activations.{2 * m}", f"{p}.convs1.{m}", k, d, logscale)
        xt = amp_layer(xt, sd, f"{p}.activations.{2 * m + 1}", f"{p}.convs2.{m}", k, 1, logscale)
        x = xt + x
    return x


def generator_forward(sd: Dict[str, Tensor], h, latent: Tensor, spk_emb: Tensor,
                      collect: Optional[dict] = None) -> Tensor:
    """models.py:212-252 with the speaker embedding [B,1,D] given.  latent [B,T,gpt_dim]."""
    logscale = bool(h["snake_logscale"])
    nk = len(h["resblock_kernel_sizes"])
    e = spk_emb.transpose(1, 2)                                       # :212  [B,D,1]
    x = latent.transpose(1, 2)                                        # :222
    x = F.conv1d(x, folded(sd, "conv_pre"), sd["conv_pre.bias"], padding=3)   # :226
    x = x + F.conv1d(e, sd["cond_layer.weight"], sd["cond_layer.bias"])       # :228
    for i, (u, ku) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
        x = F.conv_transpose1d(x, folded(sd, f"ups.{i}.0"), sd[f"ups.{i}.0.bias"], stride=u,
                               padding=(ku - u) // 2)                          # :232-233
        if h["cond_d_vector_in_each_upsampling_layer"]:
            x = x + F.conv1d(e, sd[f"conds.{i}.weight"], sd[f"conds.{i}.bias"])  # :235-236
        if collect is not None:
            collect[f"ups{i}"] = x
        xs = None
        for j, (k, dils) in enumerate(zip(h["resblock_kernel_sizes"], h["resblock_dilation_sizes"])):
            r = amp_block1(x, sd, i * nk + j, k, dils, logscale)
            xs = r if xs is None else xs + r                                   # :239-244
        x = xs / nk                                                            # :245
        if collect is not None:
            collect[f"stage{i}"] = x
    x = activation1d(x, sd["activation_post.act.alpha"], sd["activation_post.act.beta"],
                     sd["activation_post.upsample.filter"],
                     sd["activation_post.downsample.lowpass.filter"], logscale)  # :248
    x = F.conv1d(x, folded(sd, "conv_post"), sd["conv_post.bias"], padding=3)  # :249
    return torch.tanh(x)                                                       # :250


def generator_forward_ragged(sd, h, latent: Tensor, lengths, spk_emb: Tensor) -> Tensor:
    """Per-utterance decode of a padded batch: utterance b uses latent[b, :lengths[b]] — the
    semantics of B independent reference calls; tail is zero-filled."""
    B, Tmax, _ = latent.shape
    up = 1
    for u in h["upsample_rates"]:
        up *= u
    out = torch.zeros(B, 1, Tmax * up, dtype=latent.dtype)
    for b in range(B):
        L = int(lengths[b])
        out[b:b + 1, :, : L * up] = generator_forward(sd, h, latent[b:b + 1, :L], spk_emb[b:b + 1])
    return out


def to_int16(wav: Tensor) -> Tensor:
    """infer.py:892 + :911: clamp(32767*wav, -32767, 32767) -> int16 (truncation toward 0)."""
    return torch.clamp(32767 * wav, -32767.0, 32767.0).to(torch.int16)


def snr_db(ref: Tensor, test: Tensor) -> float:
    ref = ref.double().flatten()
    err = test.double().flatten() - ref
    return float(10 * torch.log10(ref.pow(2).sum() / err.pow(2).sum().clamp_min(1e-300)))
