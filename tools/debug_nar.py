"""Error map of one bf16 AMP layer (k_amp_nar) against the oracle: per 8-channel group and per row range."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import warnings; warnings.filterwarnings("ignore")
import torch
from index_tts_lora_b200 import synth
from index_tts_lora_b200.config import AttrDict
from index_tts_lora_b200.models import AMPBlock1
from index_tts_lora_b200.ops import amp_layer
from oracle import bigvgan_oracle as O
torch.set_grad_enabled(False)
dev = torch.device("cuda:0")
cases = [(64, 129, 3, 1, True), (64, 129, 3, 1, False), (64, 700, 3, 1, False), (32, 129, 3, 1, False), (96, 600, 7, 3, True), (48, 1500, 11, 5, True)]
for C, T, k, d, use_r in cases:
    blk = AMPBlock1(AttrDict(snake_logscale=True), C, k, (d, d, d), activation="snakebeta")
    sd = synth.synth_state_dict(blk.state_dict(), seed=k * 100 + d, profile="stress")
    blk.load_state_dict(sd)
    x = torch.randn(2, C, T, generator=synth._gen(1, f"x{C}{T}"))
    r = torch.randn(2, C, T, generator=synth._gen(2, f"r{C}{T}"))
    ref = O.amp_layer(x, sd, "activations.0", "convs1.0", k, d) + (r if use_r else 0)
    y = amp_layer(x.to(dev), blk.convs1[0], blk.activations[0], resid=r.to(dev) if use_r else None, precision="bf16").cpu()
    err = (y - ref)
    print(f"C={C} T={T} k={k} d={d} resid={use_r}: SNR {O.snr_db(ref, y):.1f} dB")
    eg = err.pow(2).sum(dim=(0, 2)).reshape(-1, 8).sum(1) / ref.pow(2).sum(dim=(0, 2)).reshape(-1, 8).sum(1)
    print("   rel err energy per channel group:", [f"{v:.1e}" for v in eg.tolist()])
    step = max(1, T // 16)
    er = err.pow(2).sum(dim=(0, 1)); rr = ref.pow(2).sum(dim=(0, 1))
    print("   rel err energy per row range   :", [f"{(er[i:i+step].sum() / rr[i:i+step].sum()).item():.1e}" for i in range(0, T, step)])
    bad = (er / rr.clamp_min(1e-9) > 1e-2).nonzero().flatten().tolist()
    print("   rows with > 1% error energy:", bad[:40], "..." if len(bad) > 40 else "")
