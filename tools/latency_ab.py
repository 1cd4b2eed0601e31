"""A/B of the launch-latency switches (PDL, CUDA graphs) on one shape: device-resident decode and the host-buffer call.
Usage: python tools/latency_ab.py [B] [F]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import warnings; warnings.filterwarnings("ignore")
import ctypes as C, torch
from index_tts_lora_b200 import synth, _lib
from index_tts_lora_b200.config import default_config
from index_tts_lora_b200.models import BigVGAN
torch.set_grad_enabled(False)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
F = int(sys.argv[2]) if len(sys.argv) > 2 else 157
dev = torch.device("cuda:0"); h = default_config()
m = BigVGAN(h); m.load_state_dict(synth.synth_state_dict(m.state_dict(), seed=1234, profile="init")); m = m.to(dev); m.remove_weight_norm(); m.eval(); m.precision = "bf16"
lat = synth.synth_latent(B, F, h.gpt_dim, seed=0).to(torch.bfloat16)
lat_dev = lat.to(dev); lat_host = lat.pin_memory()
wav_host = torch.empty(B, 1, F * 1024, dtype=torch.int16).pin_memory()
emb = m.speaker_embedding(synth.synth_mel(1, 300, h.num_mels, seed=1).to(dev)).reshape(1, -1).float().expand(B, -1).contiguous()
lib = _lib.load(); plan = m._ensure_plan(dev)
st = torch.cuda.current_stream(dev)
def dev_step(): m.decode(lat_dev, emb)
lens = (C.c_int32 * B)(*([F] * B)) if os.environ.get("AB_LENS") else None
def host_step():
    _lib.check(lib.bvg_decode_host(plan, lat_host.data_ptr(), _lib.BVG_BF16, lens, B, F, emb.data_ptr(), wav_host.data_ptr(), _lib.BVG_I16, _lib.PREC_BF16, st.cuda_stream), "host")
def timeit(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t) * 1e3 / n
for pdl in (0, 1):
    for gr in (0, 1):
        lib.bvg_set_pdl(pdl); lib.bvg_set_graphs(gr)
        d = timeit(dev_step); hst = timeit(host_step)
        print(f"B={B} F={F} pdl={pdl} graphs={gr}: device-resident {d[0]:.3f} ms (wall {d[1]:.3f})   host-buffer call {hst[0]:.3f} ms (wall {hst[1]:.3f})", flush=True)
