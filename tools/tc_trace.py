"""Pipeline timeline of k_amp_tc, CTA 0 (experiments build):
BVG_LIB=index_tts_lora_b200/libbvg_exp.so BVG_TRACE_LAYER=768,3,1 python tools/tc_trace.py [B F]
Prints, in kilocycles since kernel start: per chunk when warp 0 started / finished activating it and when the MMA issuer
got it; per tile MMA start / end and epilogue start / end."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
if os.environ.get("TC_TRACE_CHILD") != "1":
    env = dict(os.environ, TC_TRACE_CHILD="1")
    r = subprocess.run([sys.executable, __file__] + sys.argv[1:], env=env, capture_output=True, text=True)
    ev = {}
    for ln in r.stderr.splitlines():
        if ln.startswith("nar_trace"):
            _, s, c, t = ln.split(); ev[int(s)] = (int(c), int(t))
    if 1 not in ev:
        print(r.stdout[-2000:], r.stderr[-2000:]); sys.exit(1)
    t0 = ev[1][1]
    k = lambda s: (ev[s][1] - t0) / 1000.0 if s in ev else float("nan")
    print(f"kernel body start 0, end {k(2):.1f} kcycles")
    for it in range(8):
        if 3000 + it not in ev: break
        print(f"tile {it}: mma start {k(3000+it):8.1f}  mma issued {k(3100+it):8.1f}  epi start {k(3200+it):8.1f}  epi end {k(3300+it):8.1f}")
    print("x loads issued (chunk sequence of this CTA; no stamps inside the activation warps: they cost the narrow layers 16 %):")
    print("  ".join(f"{i}:{k(2500+i):.1f}" for i in range(480) if 2500 + i in ev and i < 60))
    print("MMA issuer got chunk (tile*NCH + c):")
    print("  ".join(f"{i}:{k(1100+i):.1f}" for i in range(1900) if 1100 + i in ev and i < 80))
    sys.exit(0)
import warnings; warnings.filterwarnings("ignore")
import torch
from index_tts_lora_b200 import synth, _lib
from index_tts_lora_b200.config import default_config
from index_tts_lora_b200.models import BigVGAN
torch.set_grad_enabled(False)
dev = torch.device("cuda:0"); h = default_config()
m = BigVGAN(h); m.load_state_dict(synth.synth_state_dict(m.state_dict(), seed=1234, profile="init")); m = m.to(dev); m.remove_weight_norm(); m.eval(); m.precision = "bf16"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16; F = int(sys.argv[2]) if len(sys.argv) > 2 else 234
lat = synth.synth_latent(B, F, h.gpt_dim, seed=0).to(dev).to(torch.bfloat16)
emb = m.speaker_embedding(synth.synth_mel(1, 300, h.num_mels, seed=1).to(dev))
lib = _lib.load(); plan = m._ensure_plan(dev)
lib.bvg_plan_set_profiling(plan, 1)       # serialise the blocks: the last matching launch owns the trace buffer
for _ in range(2): m.decode(lat, emb)
torch.cuda.synchronize()
lib.bvg_exp_dump_trace()
