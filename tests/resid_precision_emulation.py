"""CPU emulation of the bf16 path's rounding points (weights, activation inputs / outputs, layer outputs) with the residual
stream `x = xt + x` either rounded to bf16 after every add (what the kernels do) or kept in fp32 (the bf16
hi + lo residual).  Prints the waveform SNR against the fp32 oracle.  Result (round 2): an fp32 residual stream buys
2.2-3.3 dB (tiny stress 38.9 -> 41.1 dB, full stress 42.3 -> 44.5 dB): the bf16 z / weight operands dominate the error, so
the hi + lo stream was not built.  Usage: python tests/resid_precision_emulation.py   (CPU, ~4 min)"""
import os, sys, torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import warnings; warnings.filterwarnings("ignore")
from oracle import bigvgan_oracle as O
from index_tts_lora_b200 import synth
from index_tts_lora_b200.config import tiny_config, default_config
from index_tts_lora_b200.models import BigVGAN
torch.set_grad_enabled(False)
def bf(x): return x.to(torch.bfloat16).float()
def run(h, sd, lat, emb, mode):
    # mode: 'fp32' | 'bf16' (residual stream rounded) | 'hilo' (residual stream fp32, activation input rounded)
    r_w = (lambda w: w) if mode == 'fp32' else bf
    r_a = (lambda t: t) if mode == 'fp32' else bf
    logscale = bool(h["snake_logscale"]); nk = len(h["resblock_kernel_sizes"])
    def layer(x, ap, cp, k, d):
        z = O.activation1d(r_a(x), sd[ap + ".act.alpha"], sd[ap + ".act.beta"], sd[ap + ".upsample.filter"], sd[ap + ".downsample.lowpass.filter"], logscale)
        return F.conv1d(r_a(z), r_w(O.folded(sd, cp)), sd[cp + ".bias"], dilation=d, padding=(k * d - d) // 2)
    e = emb.transpose(1, 2); x = lat.transpose(1, 2)
    x = F.conv1d(r_a(x), r_w(O.folded(sd, "conv_pre")), sd["conv_pre.bias"], padding=3) + F.conv1d(e, sd["cond_layer.weight"], sd["cond_layer.bias"])
    x = r_a(x)
    for i, (u, ku) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
        x = F.conv_transpose1d(x, r_w(O.folded(sd, f"ups.{i}.0")), sd[f"ups.{i}.0.bias"], stride=u, padding=(ku - u) // 2)
        x = x + F.conv1d(e, sd[f"conds.{i}.weight"], sd[f"conds.{i}.bias"])
        x = r_a(x)
        xs = None
        for j, (k, dils) in enumerate(zip(h["resblock_kernel_sizes"], h["resblock_dilation_sizes"])):
            p = f"resblocks.{i * nk + j}"; xb = x
            for m, d in enumerate(dils):
                xt = r_a(layer(xb, f"{p}.activations.{2*m}", f"{p}.convs1.{m}", k, d))
                xt = layer(xt, f"{p}.activations.{2*m+1}", f"{p}.convs2.{m}", k, 1)
                xb = xt + xb
                if mode == 'bf16': xb = bf(xb)
            xs = xb if xs is None else xs + xb
            if mode == 'bf16': xs = bf(xs)
        x = xs / nk
        x = r_a(x)
    x = O.activation1d(x, sd["activation_post.act.alpha"], sd["activation_post.act.beta"], sd["activation_post.upsample.filter"], sd["activation_post.downsample.lowpass.filter"], logscale)
    x = F.conv1d(r_a(x), r_w(O.folded(sd, "conv_post")), sd["conv_post.bias"], padding=3)
    return torch.tanh(x)
for cfgname, h, F_ in (("tiny", tiny_config(), 24), ("full", default_config(), 12)):
    for prof in ("init", "stress"):
        m = BigVGAN(h); sd = synth.synth_state_dict(m.state_dict(), seed=1234, profile=prof)
        sd = O.fold_state_dict(sd)
        lat = synth.synth_latent(1, F_, h.gpt_dim, seed=0); emb = torch.randn(1, 1, h.speaker_embedding_dim, generator=torch.Generator().manual_seed(3)) * 0.1
        ref = run(h, sd, lat, emb, 'fp32')
        a = run(h, sd, lat, emb, 'bf16'); b = run(h, sd, lat, emb, 'hilo')
        print(cfgname, prof, "SNR bf16 residual stream %.1f dB | fp32 residual stream %.1f dB" % (O.snr_db(ref, a), O.snr_db(ref, b)), flush=True)
