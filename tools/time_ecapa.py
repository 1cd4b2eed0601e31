"""Device time of the speaker encoder: native (csrc/ecapa.cu) vs PyTorch / cuDNN replayed from a CUDA graph vs eager."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import warnings; warnings.filterwarnings("ignore")
import torch
from index_tts_lora_b200 import synth
from index_tts_lora_b200.config import default_config
from index_tts_lora_b200.models import BigVGAN
torch.set_grad_enabled(False)
dev = torch.device("cuda:0"); h = default_config()
m = BigVGAN(h); m.load_state_dict(synth.synth_state_dict(m.state_dict(), seed=1234, profile="init")); m = m.to(dev); m.remove_weight_norm(); m.eval()
m.cache_speaker_embedding = False
mel = synth.synth_mel(1, 300, h.num_mels, seed=1).to(dev)
def timeit(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
m.native_speaker_encoder = True
print("native (csrc/ecapa.cu, CUDA graph)   ms", round(timeit(lambda: m.speaker_embedding(mel)), 4))
m.native_speaker_encoder = False; m.graph_speaker_encoder = True
print("PyTorch / cuDNN, torch CUDA graph    ms", round(timeit(lambda: m.speaker_embedding(mel)), 4))
m.graph_speaker_encoder = False
print("PyTorch / cuDNN eager                ms", round(timeit(lambda: m.speaker_embedding(mel)), 4))
