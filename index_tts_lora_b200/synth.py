"""Deterministic synthetic weights and inputs (there are no checkpoints and no network).

The same generator is used (a) in the build container to drive the REAL reference module when
``tests/golden/make_golden.py`` records golden outputs, and (b) on the GPU box to rebuild the
identical state dict for the CUDA path — the state dict is regenerated from a seed, not shipped
(126.7 M parameters).  Every tensor gets its own ``torch.Generator`` seeded from a stable hash
of (seed, key), so the values do not depend on key order or on which keys are present.

Profiles
  "init"    the distribution of a freshly constructed reference module (models.py:132-199):
            generator conv weight_v ~ N(0, 0.01^2) (utils.py:47-50), weight_g = ||v||, torch
            default U(+-1/sqrt(fan_in)) for biases / cond convs / conv_pre / ECAPA,
            alpha = beta = 0.  This is BASELINE.json's "random-init weights".
  "stress"  O(1) signal everywhere: fan-in-scaled weights, weight_g != ||v||, alpha, beta ~
            N(0, 0.3^2), non-trivial BatchNorm statistics (SURVEY §7 hard part 8).
"""
from __future__ import annotations

import hashlib
import math
from typing import Dict

import torch


def _gen(seed: int, key: str) -> torch.Generator:
    hsh = hashlib.sha256(f"{seed}:{key}".encode()).digest()
    g = torch.Generator(device="cpu")
    g.manual_seed(int.from_bytes(hsh[:8], "little") & 0x7FFFFFFFFFFFFFFF)
    return g


def _randn(shape, g):
    return torch.randn(shape, generator=g, dtype=torch.float32)


def _uniform(shape, bound, g):
    return (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * bound


def synth_state_dict(template: Dict[str, torch.Tensor], seed: int = 1234,
                     profile: str = "init") -> Dict[str, torch.Tensor]:
    """Fill a state dict shaped like ``template`` (``module.state_dict()`` of either the
    reference or the drop-in module, weight-norm attached)."""
    assert profile in ("init", "stress")
    stress = profile == "stress"
    out: Dict[str, torch.Tensor] = {}
    shapes = {k: tuple(v.shape) for k, v in template.items()}

    def fan_in_of(prefix: str) -> int:
        for suffix in (".weight_v", ".weight"):
            if prefix + suffix in shapes:
                s = shapes[prefix + suffix]
                if prefix.startswith("ups."):           # ConvTranspose1d [C_in, C_out, k], stride u
                    return max(1, s[0] * s[2] // max(1, int(round(s[2] / 2))))  # ~ C_in * k / u
                return s[1] * s[2] if len(s) == 3 else s[1]
        return 1

    for k, s in shapes.items():
        g = _gen(seed, k)
        is_gen_conv = k.startswith(("ups.", "resblocks.", "conv_post."))
        if k.endswith("num_batches_tracked"):
            out[k] = torch.zeros((), dtype=torch.int64)
        elif k.endswith(".filter"):
            out[k] = template[k].detach().clone().float()           # kaiser-sinc buffers
        elif k.endswith(".act.alpha") or k.endswith(".act.beta"):
            out[k] = 0.3 * _randn(s, g) if stress else torch.zeros(s)
        elif k.endswith(".weight_v"):
            p = k[: -len(".weight_v")]
            if stress:
                out[k] = _randn(s, g)
            elif is_gen_conv:
                out[k] = 0.01 * _randn(s, g)
            else:                                                    # conv_pre: torch default init
                out[k] = _uniform(s, 1.0 / math.sqrt(fan_in_of(p)), g)
        elif k.endswith(".weight_g"):
            continue                                                 # second pass
        elif k.endswith("running_mean"):
            out[k] = 0.1 * _randn(s, g) if stress else torch.zeros(s)
        elif k.endswith("running_var"):
            out[k] = 0.5 + torch.rand(s, generator=g) if stress else torch.ones(s)
        elif ".norm.norm." in k or k.startswith("speaker_encoder.asp_bn.norm."):
            if k.endswith(".weight"):
                out[k] = 0.5 + torch.rand(s, generator=g) if stress else torch.ones(s)
            else:
                out[k] = 0.1 * _randn(s, g) if stress else torch.zeros(s)
        elif k.endswith(".weight"):                                  # cond convs, ECAPA convs
            fi = s[1] * (s[2] if len(s) == 3 else 1)
            out[k] = _uniform(s, 1.0 / math.sqrt(fi), g)
        elif k.endswith(".bias"):
            p = k[: -len(".bias")]
            out[k] = 0.05 * _randn(s, g) if stress else _uniform(s, 1.0 / math.sqrt(fan_in_of(p)), g)
        else:
            raise KeyError(f"synth_state_dict: unhandled key {k} {s}")

    for k, s in shapes.items():
        if not k.endswith(".weight_g"):
            continue
        p = k[: -len(".weight_g")]
        v = out[p + ".weight_v"]
        norm = v.reshape(v.shape[0], -1).norm(dim=1).view(s)
        if stress:
            # folded weight rows get std gain/sqrt(fan_in) with a per-row jitter so g != ||v||
            gain = 0.5 if p.startswith("resblocks.") else (0.3 if p == "conv_post" else 1.0)
            n_per_row = v[0].numel()
            jitter = 1.0 + 0.2 * _randn(s, _gen(seed, k))
            if p.startswith("ups."):
                # dim 0 is C_in: every row contributes to all outputs; same target std
                target = gain / math.sqrt(fan_in_of(p))
            else:
                target = gain / math.sqrt(fan_in_of(p))
            out[k] = target * math.sqrt(n_per_row) * jitter.abs()
        else:
            out[k] = norm.clone()                                    # weight_norm init: g = ||v||
    return out


def synth_latent(B: int, F: int, gpt_dim: int = 1280, seed: int = 0) -> torch.Tensor:
    """GPT final_norm output stand-in (gpt/model.py:459-474): unit-variance Gaussian."""
    return _randn((B, F, gpt_dim), _gen(seed, f"latent:{B}x{F}x{gpt_dim}"))


def synth_mel(B: int, Tm: int = 300, num_mels: int = 100, seed: int = 1) -> torch.Tensor:
    return _randn((B, Tm, num_mels), _gen(seed, f"mel:{B}x{Tm}x{num_mels}"))


def synth_lengths(n: int, lo: int, hi: int, seed: int = 2):
    """Utterance lengths in latent frames ~ U{lo..hi} (BASELINE config 4: 47..469 = 2-20 s)."""
    g = _gen(seed, f"len:{n}:{lo}:{hi}")
    return torch.randint(lo, hi + 1, (n,), generator=g).tolist()


def checksum(t: torch.Tensor) -> float:
    """Order-sensitive fp64 fingerprint used to detect RNG drift between containers."""
    f = t.detach().double().flatten()
    w = torch.arange(1, f.numel() + 1, dtype=torch.float64)
    return float((f * torch.cos(w * 0.001)).sum())
