"""Generate the golden fixtures in this directory by running the REAL reference.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

It imports ``indextts.BigVGAN.models.BigVGAN`` (use_cuda_kernel=False — the torch path is the
only valid oracle at sequence edges) from /root/reference, with an empty
``matplotlib`` stub because ``BigVGAN/utils.py:7-13`` imports it at module top, drives it with
the deterministic weights/inputs of ``index_tts_lora_b200.synth`` and stores inputs (when small),
seeds, input checksums and OUTPUTS.  Nothing from the reference is copied into the repo; only
its numeric outputs are.
"""
from __future__ import annotations

import json
import os
import sys
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("BVG_REFERENCE", "/root/reference")
sys.path.insert(0, REF)

_m = types.ModuleType("matplotlib")
_m.use = lambda *a, **k: None
sys.modules.setdefault("matplotlib", _m)
sys.modules.setdefault("matplotlib.pylab", types.ModuleType("matplotlib.pylab"))
warnings.filterwarnings("ignore")

from indextts.BigVGAN import activations as ref_act  # noqa: E402
from indextts.BigVGAN.alias_free_torch import Activation1d as RefAct1d  # noqa: E402
from indextts.BigVGAN.models import AMPBlock1 as RefAMP, BigVGAN as RefBigVGAN  # noqa: E402

from index_tts_lora_b200 import synth  # noqa: E402
from index_tts_lora_b200.config import AttrDict, load_yaml_config, tiny_config  # noqa: E402

torch.set_grad_enabled(False)


def ref_generator(h, seed, profile):
    g = RefBigVGAN(AttrDict(dict(h)), use_cuda_kernel=False)
    sd = synth.synth_state_dict(g.state_dict(), seed=seed, profile=profile)
    g.load_state_dict(sd)
    keys_wn = {k: list(v.shape) for k, v in g.state_dict().items()}
    g.remove_weight_norm()
    g.eval()
    return g, sd, keys_wn


def golden_act1d():
    out = {}
    for name, (B, C, T) in {"a": (2, 5, 37), "t1": (1, 3, 1), "t2": (1, 2, 2), "t7": (2, 4, 7),
                            "t12": (1, 3, 12), "long": (1, 2, 301)}.items():
        g = synth._gen(7, f"act:{name}")
        x = 1.5 * torch.randn(B, C, T, generator=g)
        alpha = 0.4 * torch.randn(C, generator=g)
        beta = 0.4 * torch.randn(C, generator=g)
        act = RefAct1d(activation=ref_act.SnakeBeta(C, alpha_logscale=True))
        act.act.alpha.data.copy_(alpha)
        act.act.beta.data.copy_(beta)
        y = act(x)
        out.update({f"{name}_x": x.numpy(), f"{name}_alpha": alpha.numpy(),
                    f"{name}_beta": beta.numpy(), f"{name}_y": y.numpy()})
        out["filter"] = act.upsample.filter.reshape(-1).numpy()
        assert torch.equal(act.upsample.filter, act.downsample.lowpass.filter)
    np.savez(os.path.join(HERE, "act1d.npz"), **out)
    print("act1d.npz", len(out))


def golden_ampblock():
    out = {}
    h = AttrDict(snake_logscale=True)
    for k in (3, 7, 11):
        C, T, B = 16, 61, 2
        blk = RefAMP(h, C, k, (1, 3, 5), activation="snakebeta")
        sd = synth.synth_state_dict(blk.state_dict(), seed=11 + k, profile="stress")
        blk.load_state_dict(sd)
        blk.remove_weight_norm()
        x = torch.randn(B, C, T, generator=synth._gen(5, f"amp:{k}"))
        y = blk(x)
        # also the first half-layer alone: xt = c1(a1(x))
        xt = blk.convs1[0](blk.activations[0](x))
        out.update({f"k{k}_x": x.numpy(), f"k{k}_y": y.numpy(), f"k{k}_xt0": xt.numpy()})
    np.savez(os.path.join(HERE, "ampblock.npz"), **out)
    print("ampblock.npz")


def golden_model(tag, h, seed, profile, B, F, Tm, mel_B=None, store_inputs=False):
    g, sd, keys_wn = ref_generator(h, seed, profile)
    lat = synth.synth_latent(B, F, h["gpt_dim"], seed=0)
    mel = synth.synth_mel(mel_B or B, Tm, h["num_mels"], seed=1)
    emb = g.speaker_encoder(mel, None)
    wav, loss = g(lat, mel)
    assert loss is None
    rec = {"wav": wav.numpy(), "spk_emb": emb.numpy(),
           "latent_checksum": np.float64(synth.checksum(lat)),
           "mel_checksum": np.float64(synth.checksum(mel)),
           "sd_checksum": np.float64(sum(synth.checksum(v) for k, v in sorted(sd.items())
                                         if v.dtype.is_floating_point and v.numel() < 2_000_000)),
           "meta": np.array(json.dumps({"seed": seed, "profile": profile, "B": B, "F": F, "Tm": Tm,
                                        "mel_B": mel_B or B}))}
    if store_inputs:
        rec["latent"] = lat.numpy()
        rec["mel"] = mel.numpy()
    np.savez(os.path.join(HERE, f"{tag}.npz"), **rec)
    print(tag, tuple(wav.shape), float(wav.abs().max()), float(wav.std()))
    return g, keys_wn


def main():
    golden_act1d()
    golden_ampblock()
    ht = tiny_config()
    golden_model("tiny_init", ht, 1234, "init", B=2, F=9, Tm=40, store_inputs=True)
    golden_model("tiny_stress", ht, 1234, "stress", B=2, F=9, Tm=40, store_inputs=True)
    hf = load_yaml_config(os.path.join(REF, "finetune_models", "config.yaml"))
    g, keys_wn = golden_model("full_f157_init", hf, 1234, "init", B=1, F=157, Tm=300)
    keys_folded = {k: list(v.shape) for k, v in g.state_dict().items()}
    with open(os.path.join(HERE, "state_dict_keys.json"), "w") as f:
        json.dump({"weight_norm": keys_wn, "folded": keys_folded}, f)
    print("state dict entries", len(keys_wn), len(keys_folded))
    del g
    golden_model("full_f157_stress", hf, 1234, "stress", B=1, F=157, Tm=300)
    golden_model("full_f20_b2_stress", hf, 1234, "stress", B=2, F=20, Tm=120)


if __name__ == "__main__":
    main()
