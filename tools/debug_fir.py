"""Per-layer check of the tensor-core-FIR AMP kernel (k_amp_fir) against the fp32 oracle, with error localisation."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from oracle import bigvgan_oracle as O
from index_tts_lora_b200.models import AMPBlock1
from index_tts_lora_b200.ops import amp_layer
from index_tts_lora_b200 import synth
from index_tts_lora_b200.config import AttrDict

dev = torch.device("cuda:0")
cases = [(24, 700, 3, 1), (24, 700, 11, 5), (48, 1500, 7, 3), (96, 5, 3, 1), (96, 5, 11, 5), (64, 129, 7, 3), (96, 2100, 11, 1),
         (24, 3, 3, 1), (32, 256, 3, 1), (32, 257, 11, 5)]
if len(sys.argv) > 1:
    cases = [tuple(int(v) for v in s.split(",")) for s in sys.argv[1:]]
for C, T, k, d in cases:
    blk = AMPBlock1(AttrDict(snake_logscale=True), C, k, (d, d, d), activation="snakebeta")
    sd = synth.synth_state_dict(blk.state_dict(), seed=k * 100 + d, profile="stress")
    blk.load_state_dict(sd)
    x = torch.randn(2, C, T, generator=synth._gen(1, f"x{C}{T}"))
    r = torch.randn(2, C, T, generator=synth._gen(2, f"r{C}{T}"))
    ref = O.amp_layer(x, sd, "activations.0", "convs1.0", k, d) + r
    y = amp_layer(x.to(dev), blk.convs1[0], blk.activations[0], resid=r.to(dev), precision="bf16").cpu()
    torch.cuda.synchronize()
    err = (y - ref).abs()
    snr = O.snr_db(ref, y)
    pt = err.amax(dim=(0, 1))          # per time
    pc = err.amax(dim=(0, 2))          # per channel
    bad_t = (pt > 0.05 * ref.abs().max()).nonzero().flatten().tolist()
    bad_c = (pc > 0.05 * ref.abs().max()).nonzero().flatten().tolist()
    print(f"C={C} T={T} k={k} d={d}: SNR {snr:.1f} dB  max|err| {err.max():.4f} (|ref| max {ref.abs().max():.2f}) "
          f"bad rows {len(bad_t)} {bad_t[:12]}{'...' if len(bad_t) > 12 else ''} {bad_t[-4:] if len(bad_t) > 12 else ''} "
          f"bad ch {len(bad_c)} {bad_c[:8]}", flush=True)
