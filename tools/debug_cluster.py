"""Cluster mode (bvg_set_tc_cluster) vs plain launches: bit-identity on single layers and on a whole decode, then timing."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import warnings; warnings.filterwarnings("ignore")
import ctypes as C, torch
from index_tts_lora_b200 import synth, _lib
from index_tts_lora_b200.config import AttrDict, default_config
from index_tts_lora_b200.models import AMPBlock1, BigVGAN
from index_tts_lora_b200.ops import amp_layer
torch.set_grad_enabled(False)
dev = torch.device("cuda:0"); lib = _lib.load()
for C_, T, k, d in [(384, 257, 3, 1), (768, 70, 3, 1), (768, 700, 7, 3), (384, 1500, 11, 5)]:
    blk = AMPBlock1(AttrDict(snake_logscale=True), C_, k, (d, d, d), activation="snakebeta")
    blk.load_state_dict(synth.synth_state_dict(blk.state_dict(), seed=k * 100 + d, profile="stress"))
    x = torch.randn(2, C_, T, generator=synth._gen(1, f"x{C_}{T}")).to(dev)
    r = torch.randn(2, C_, T, generator=synth._gen(2, f"r{C_}{T}")).to(dev)
    lib.bvg_set_tc_cluster(0)
    y0 = amp_layer(x, blk.convs1[0], blk.activations[0], resid=r, precision="bf16")
    lib.bvg_set_tc_cluster(1)
    y1 = amp_layer(x, blk.convs1[0], blk.activations[0], resid=r, precision="bf16")
    torch.cuda.synchronize()
    print(f"layer C={C_} T={T} k={k} d={d}: cluster == plain: {bool(torch.equal(y0, y1))}  max diff {(y0 - y1).abs().max().item():.3e}", flush=True)
h = default_config()
m = BigVGAN(h); m.load_state_dict(synth.synth_state_dict(m.state_dict(), seed=1234, profile="init")); m = m.to(dev); m.remove_weight_norm(); m.eval(); m.precision = "bf16"
lat = synth.synth_latent(16, 234, h.gpt_dim, seed=0).to(dev).to(torch.bfloat16)
emb = m.speaker_embedding(synth.synth_mel(1, 300, h.num_mels, seed=1).to(dev))
plan = m._ensure_plan(dev)
outs = {}
for cl in (0, 1):
    lib.bvg_set_tc_cluster(cl)
    for _ in range(3): w = m.decode(lat, emb)
    torch.cuda.synchronize(); outs[cl] = w.clone()
    ts = []
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); m.decode(lat, emb); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    lib.bvg_plan_set_profiling(plan, 1)
    for _ in range(2): m.decode(lat, emb)
    p = _lib.BvgProfile(); lib.bvg_plan_read_profile(plan, C.byref(p)); lib.bvg_plan_set_profiling(plan, 0)
    print(f"cluster={cl}: decode ms min {min(ts):.2f} med {sorted(ts)[2]:.2f}; serialised by class {[round(p.ms[i] / 2, 2) for i in range(4)]}", flush=True)
print("whole decode cluster == plain:", bool(torch.equal(outs[0], outs[1])))
