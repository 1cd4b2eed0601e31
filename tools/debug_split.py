import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import warnings; warnings.filterwarnings("ignore")
import torch
from index_tts_lora_b200 import synth
from index_tts_lora_b200.config import tiny_config
from index_tts_lora_b200.longform import emulate_time_split
from index_tts_lora_b200.models import BigVGAN
torch.set_grad_enabled(False)
dev = torch.device("cuda:0"); h = tiny_config(); models = []; sd = None
for r in range(3):
    m = BigVGAN(h)
    if sd is None: sd = synth.synth_state_dict(m.state_dict(), seed=21, profile="stress")
    m.load_state_dict(sd); m = m.to(dev).eval(); m.precision = "bf16"; models.append(m)
Ftot = int(sys.argv[1]) if len(sys.argv) > 1 else 130
lat = synth.synth_latent(1, Ftot, h.gpt_dim, seed=5).to(dev).to(torch.bfloat16)
emb = models[0].speaker_embedding(synth.synth_mel(1, 50, h.num_mels, seed=6).to(dev))
whole = models[0].decode(lat, emb, out_dtype=torch.float32)[0, 0]
split = emulate_time_split(models, lat, emb)
err = (split - whole).abs()
bad = torch.nonzero(err > 1e-4).flatten()
print("n bad", bad.numel(), "of", err.numel(), "max", err.max().item())
if bad.numel():
    fr = (bad // 1024).unique()
    print("bad frames:", fr.tolist()[:60])
    print("first bad sample", bad[0].item(), "last", bad[-1].item())
    # cluster
    b = bad.cpu().tolist(); cl = [[b[0], b[0]]]
    for x in b[1:]:
        if x - cl[-1][1] > 2000: cl.append([x, x])
        else: cl[-1][1] = x
    print("clusters (sample ranges):", cl[:20], "frames", [(a / 1024, c / 1024) for a, c in cl[:20]])
from index_tts_lora_b200.sharding import time_shards
print(time_shards(Ftot, 3, 0))
