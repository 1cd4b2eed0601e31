"""Pipeline timeline of k_amp_nar (experiments build): BVG_LIB=.../libbvg_exp.so BVG_TRACE_LAYER=24,3,1 python tools/nar_trace.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import warnings; warnings.filterwarnings("ignore")
import torch
from index_tts_lora_b200 import synth, _lib
from index_tts_lora_b200.config import default_config
from index_tts_lora_b200.models import BigVGAN
torch.set_grad_enabled(False)
dev = torch.device("cuda:0"); h = default_config()
m = BigVGAN(h); m.load_state_dict(synth.synth_state_dict(m.state_dict(), seed=1234, profile="init")); m = m.to(dev); m.remove_weight_norm(); m.eval(); m.precision = "bf16"
lat = synth.synth_latent(16, 234, h.gpt_dim, seed=0).to(dev).to(torch.bfloat16)
emb = m.speaker_embedding(synth.synth_mel(1, 300, h.num_mels, seed=1).to(dev))
lib = _lib.load(); plan = m._ensure_plan(dev)
lib.bvg_plan_set_profiling(plan, 1)       # serialise the blocks: the last matching launch owns the trace buffer
for _ in range(2): m.decode(lat, emb)
torch.cuda.synchronize()
import ctypes
lib.bvg_exp_dump_trace()
