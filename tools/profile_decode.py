"""Run N warm decodes of a fixed batch (for ncu / compute-sanitizer).  Usage:
    python tools/profile_decode.py [B] [F] [precision] [n_decodes]
k_amp_tc launch order inside one bf16 decode (115 launches): 0 conv_pre; per stage s (0..5):
base = 1 + 19*s: +0 ConvTranspose, then 18 AMP layers ordered k in (3,7,11) x d in (1,3,5) x (A,B)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import warnings

warnings.filterwarnings("ignore")
import torch

from index_tts_lora_b200 import synth
from index_tts_lora_b200.config import default_config
from index_tts_lora_b200.models import BigVGAN

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
F = int(sys.argv[2]) if len(sys.argv) > 2 else 234
prec = sys.argv[3] if len(sys.argv) > 3 else "bf16"
n = int(sys.argv[4]) if len(sys.argv) > 4 else 2
torch.set_grad_enabled(False)
dev = torch.device("cuda:0")
h = default_config()
m = BigVGAN(h)
m.load_state_dict(synth.synth_state_dict(m.state_dict(), seed=1234, profile="init"))
m = m.to(dev)
m.remove_weight_norm()
m.eval()
m.precision = prec
lat = synth.synth_latent(B, F, h.gpt_dim, seed=0).to(dev).to(torch.bfloat16 if prec == "bf16" else torch.float32)
emb = m.speaker_embedding(synth.synth_mel(1, 300, h.num_mels, seed=1).to(dev))
for i in range(n):
    wav = m.decode(lat, emb)
torch.cuda.synchronize()
print("ok", tuple(wav.shape), float(wav.float().abs().max()))
