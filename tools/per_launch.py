"""Per-launch device times of one serialised bf16 decode (BVG_PROF_DUMP lines of bvg_plan_read_profile).
Usage: BVG_PROF_DUMP=1 python tools/per_launch.py [split_min_c] 2> lines.txt"""
import os, sys
os.environ.setdefault("BVG_PROF_DUMP", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import warnings; warnings.filterwarnings("ignore")
import ctypes as C, torch
from index_tts_lora_b200 import synth, _lib
from index_tts_lora_b200.config import default_config
from index_tts_lora_b200.models import BigVGAN
torch.set_grad_enabled(False)
dev = torch.device("cuda:0"); h = default_config()
m = BigVGAN(h); m.load_state_dict(synth.synth_state_dict(m.state_dict(), seed=1234, profile="init")); m = m.to(dev); m.remove_weight_norm(); m.eval(); m.precision = "bf16"
B = int(os.environ.get("PL_B", "16")); F = int(os.environ.get("PL_F", "234"))
lat = synth.synth_latent(B, F, h.gpt_dim, seed=0).to(dev).to(torch.bfloat16)
emb = m.speaker_embedding(synth.synth_mel(1, 300, h.num_mels, seed=1).to(dev))
lib = _lib.load(); plan = m._ensure_plan(dev)
if len(sys.argv) > 1: lib.bvg_set_tc_split_min_channels(int(sys.argv[1]))
for _ in range(2): m.decode(lat, emb)
lib.bvg_plan_set_profiling(plan, 1)
m.decode(lat, emb)
p = _lib.BvgProfile(); lib.bvg_plan_read_profile(plan, C.byref(p)); lib.bvg_plan_set_profiling(plan, 0)
