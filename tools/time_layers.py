"""Time individual AMP layers of the full model via the profiling API (serialised launches)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import warnings; warnings.filterwarnings("ignore")
import ctypes as C, torch
from index_tts_lora_b200 import synth, _lib
from index_tts_lora_b200.config import default_config
from index_tts_lora_b200.models import BigVGAN
torch.set_grad_enabled(False)
dev = torch.device("cuda:0"); h = default_config()
m = BigVGAN(h); m.load_state_dict(synth.synth_state_dict(m.state_dict(), seed=1234, profile="init")); m = m.to(dev); m.remove_weight_norm(); m.eval(); m.precision = "bf16"
lat = synth.synth_latent(16, 234, h.gpt_dim, seed=0).to(dev).to(torch.bfloat16)
emb = m.speaker_embedding(synth.synth_mel(1, 300, h.num_mels, seed=1).to(dev))
lib = _lib.load(); plan = m._ensure_plan(dev)
for _ in range(2): m.decode(lat, emb)
lib.bvg_plan_set_profiling(plan, 1)
for _ in range(3): m.decode(lat, emb)
p = _lib.BvgProfile(); lib.bvg_plan_read_profile(plan, C.byref(p)); lib.bvg_plan_set_profiling(plan, 0)
print("BVG_DBG", os.environ.get("BVG_DBG"), "ms/step by class:", [round(p.ms[i] / 3, 2) for i in range(4)])
