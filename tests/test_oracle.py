"""CPU tests (-m "not gpu"): the oracle against the golden vectors of the REAL reference, the
plain-C oracle against the torch oracle, known-answer facts of SURVEY.md §8c."""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN, build_case, load_golden
from oracle import bigvgan_oracle as O
from oracle import c_oracle


def test_filter_taps_known_answer():
    """kaiser_sinc_filter1d(0.25, 0.3, 12) — symmetric, unit DC gain, known values."""
    f = O.kaiser_sinc_filter()
    g, _ = load_golden("act1d")
    assert torch.equal(f, torch.tensor(g["filter"]))
    assert torch.allclose(f, f.flip(0))
    assert abs(float(f.sum()) - 1.0) < 1e-6
    known = [0.0020289647, 0.0093894657, -0.0255434588, -0.0576573834, 0.1285725832, 0.4432097971]
    assert np.allclose(f[:6].numpy(), known, atol=1e-9)


@pytest.mark.parametrize("name", ["a", "t1", "t2", "t7", "t12", "long"])
def test_activation1d_oracle_vs_reference_golden(name):
    g, _ = load_golden("act1d")
    f = torch.tensor(g["filter"])
    x, al, be = (torch.tensor(g[f"{name}_{k}"]) for k in ("x", "alpha", "beta"))
    ref = torch.tensor(g[name + "_y"])
    y = O.activation1d(x, al, be, f, f)
    assert (y - ref).abs().max().item() < 2e-6
    yc = torch.tensor(c_oracle.activation1d(x.numpy(), f.numpy(), f.numpy(), al.numpy(), be.numpy()))
    assert (yc - ref).abs().max().item() < 2e-6


def test_activation1d_receptive_field_is_5():
    """Activation1d output m depends on x[m-5..m+5] only."""
    torch.manual_seed(0)
    f = O.kaiser_sinc_filter()
    x = torch.randn(1, 1, 64)
    a = b = torch.zeros(1)
    y0 = O.activation1d(x, a, b, f, f)
    x2 = x.clone()
    x2[0, 0, 30] += 1.0
    d = (O.activation1d(x2, a, b, f, f) - y0).abs()[0, 0]
    idx = torch.nonzero(d > 1e-7).flatten()
    assert idx.min().item() == 25 and idx.max().item() == 35


@pytest.mark.parametrize("k", [3, 7, 11])
def test_ampblock_oracle_vs_reference_golden(k):
    from index_tts_lora_b200 import synth
    from index_tts_lora_b200.config import AttrDict
    from index_tts_lora_b200.models import AMPBlock1
    g, _ = load_golden("ampblock")
    blk = AMPBlock1(AttrDict(snake_logscale=True), 16, k, (1, 3, 5), activation="snakebeta")
    sd = synth.synth_state_dict(blk.state_dict(), seed=11 + k, profile="stress")
    x = torch.tensor(g[f"k{k}_x"])
    # fold once, rename to the resblocks.N prefix the oracle expects
    sdf = {"resblocks.0." + kk: v for kk, v in O.fold_state_dict(sd).items()}
    y = O.amp_block1(x, sdf, 0, k, (1, 3, 5))
    assert (y - torch.tensor(g[f"k{k}_y"])).abs().max().item() < 2e-5   # O(1..10) values, fp32 order effects
    xt = O.amp_layer(x, sdf, "resblocks.0.activations.0", "resblocks.0.convs1.0", k, 1)
    assert (xt - torch.tensor(g[f"k{k}_xt0"])).abs().max().item() < 5e-6
    # plain-C restatement of the same half layer
    f = sd["activations.0.upsample.filter"].reshape(-1).numpy()
    z = c_oracle.activation1d(x.numpy(), f, f, sd["activations.0.act.alpha"].numpy(),
                              sd["activations.0.act.beta"].numpy())
    xc = c_oracle.conv1d(z, sdf["resblocks.0.convs1.0.weight"].numpy(), sdf["resblocks.0.convs1.0.bias"].numpy(), 1)
    assert np.abs(xc - g[f"k{k}_xt0"]).max() < 5e-6


@pytest.mark.parametrize("u,k", [(4, 8), (4, 4), (2, 4)])
def test_c_oracle_conv_transpose_vs_torch(u, k):
    torch.manual_seed(u + k)
    m = torch.nn.ConvTranspose1d(8, 6, k, u, padding=(k - u) // 2)
    x = torch.randn(2, 8, 19)
    ref = m(x).detach().numpy()
    y = c_oracle.conv_transpose1d(x.numpy(), m.weight.detach().numpy(), m.bias.detach().numpy(), u)
    assert np.abs(y - ref).max() < 2e-6


@pytest.mark.parametrize("tag", ["tiny_init", "tiny_stress"])
def test_tiny_generator_oracle_vs_reference_golden(tag):
    from index_tts_lora_b200.config import tiny_config
    h = tiny_config()
    m, sd, lat, mel, g = build_case(tag, h)
    assert np.array_equal(lat.numpy(), g["latent"]) and np.array_equal(mel.numpy(), g["mel"])
    m.load_state_dict(sd)
    m.eval()
    emb = m.speaker_encoder(mel)
    assert (emb - torch.tensor(g["spk_emb"])).abs().max().item() < 1e-5      # ECAPA port == reference
    wav = O.generator_forward(sd, h, lat, emb)
    assert (wav - torch.tensor(g["wav"])).abs().max().item() < 5e-6


def test_full_generator_oracle_vs_reference_golden():
    """BASELINE config 1 (6.7 s utterance, fp32, CPU): oracle == reference to ~1e-6."""
    from index_tts_lora_b200.config import default_config
    h = default_config()
    m, sd, lat, mel, g = build_case("full_f20_b2_stress", h)
    m.load_state_dict(sd)
    m.eval()
    emb = m.speaker_encoder(mel)
    assert (emb - torch.tensor(g["spk_emb"])).abs().max().item() < 1e-5
    wav = O.generator_forward(sd, h, lat, emb)
    ref = torch.tensor(g["wav"])
    assert wav.shape == ref.shape == (2, 1, 20 * 1024)
    assert (wav - ref).abs().max().item() < 1e-5
    # batch independence: B=2 stacked == two B=1 runs
    w0 = O.generator_forward(sd, h, lat[:1], emb[:1])
    assert (w0 - wav[:1]).abs().max().item() < 1e-5


def test_ragged_oracle_is_per_utterance():
    from index_tts_lora_b200.config import tiny_config
    from index_tts_lora_b200 import synth
    from index_tts_lora_b200.models import BigVGAN
    h = tiny_config()
    m = BigVGAN(h).eval()
    sd = synth.synth_state_dict(m.state_dict(), seed=7, profile="stress")
    lat = synth.synth_latent(2, 6, h.gpt_dim, seed=3)
    emb = torch.randn(2, 1, h.speaker_embedding_dim)
    out = O.generator_forward_ragged(sd, h, lat, [6, 3], emb)
    solo = O.generator_forward(sd, h, lat[1:2, :3], emb[1:2])
    assert torch.equal(out[1, :, :3 * 1024], solo[0])
    assert out[1, :, 3 * 1024:].abs().max().item() == 0.0


def test_int16_rule_matches_infer_py():
    w = torch.tensor([0.0, 0.5, -0.5, 1.0, -1.0, 0.99999, 1.5e-5])
    assert O.to_int16(w).tolist() == [0, 16383, -16383, 32767, -32767, 32766, 0]
