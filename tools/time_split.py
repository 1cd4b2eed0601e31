"""A/B of the split form (bvg_set_tc_split_min_channels) on the 16 x 10 s bf16 workload: whole-decode device time with
the three AMP blocks on three streams, and the per-class serialised profile, for several thresholds in one process."""
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
B, F = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (16, 234)
lat = synth.synth_latent(B, F, h.gpt_dim, seed=0).to(dev).to(torch.bfloat16)
emb = m.speaker_embedding(synth.synth_mel(1, 300, h.num_mels, seed=1).to(dev))
lib = _lib.load(); plan = m._ensure_plan(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ref = None
for min_c in [int(v) for v in os.environ.get("SPLITS", "0,768,384,192,96,48,24").split(",")]:
    lib.bvg_set_tc_split_min_channels(min_c)
    for _ in range(2): w = m.decode(lat, emb)
    torch.cuda.synchronize()
    if ref is None: ref = w.clone()
    same = bool(torch.equal(ref, w))
    ts = []
    for _ in range(5):
        flush.zero_(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); m.decode(lat, emb); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    lib.bvg_plan_set_profiling(plan, 1)
    for _ in range(2): m.decode(lat, emb)
    p = _lib.BvgProfile(); lib.bvg_plan_read_profile(plan, C.byref(p)); lib.bvg_plan_set_profiling(plan, 0)
    print(f"split_min_c={min_c:4d}  decode ms min {min(ts):.2f} med {sorted(ts)[2]:.2f}  identical_to_fused={same}  "
          f"serialised ms by class {[round(p.ms[i] / 2, 2) for i in range(4)]} launches {[p.launches[i] // 2 for i in range(4)]}", flush=True)
