import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import warnings; warnings.filterwarnings("ignore")
import torch
from index_tts_lora_b200 import synth
from index_tts_lora_b200.config import default_config
from index_tts_lora_b200.models import BigVGAN
torch.set_grad_enabled(False)
dev = torch.device("cuda:0"); h = default_config()
m = BigVGAN(h); m.load_state_dict(synth.synth_state_dict(m.state_dict(), seed=1234, profile="init")); m = m.to(dev); m.remove_weight_norm(); m.eval(); m.precision = "bf16"
mel = synth.synth_mel(1, 300, h.num_mels, seed=1).to(dev)
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("ECAPA [1,300,100] ms:", timeit(lambda: m.speaker_embedding(mel)))
for B, F in [(16, 234), (1, 157)]:
    lat = synth.synth_latent(B, F, h.gpt_dim, seed=0).to(dev).to(torch.bfloat16)
    emb = m.speaker_embedding(mel)
    print(f"decode only B={B} F={F} ms:", timeit(lambda: m.decode(lat, emb), 10))
    import time
    t = time.perf_counter(); 
    for _ in range(10): m.decode(lat, emb)
    t1 = time.perf_counter() - t; torch.cuda.synchronize()
    print(f"   host enqueue time per decode: {t1/10*1e3:.2f} ms")
