import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import warnings; warnings.filterwarnings("ignore")
import torch
from index_tts_lora_b200 import synth
from index_tts_lora_b200.config import default_config
from index_tts_lora_b200.models import BigVGAN
torch.set_grad_enabled(False)
dev = torch.device("cuda:0"); h = default_config()
m = BigVGAN(h); m.load_state_dict(synth.synth_state_dict(m.state_dict(), seed=1234, profile="init")); m = m.to(dev); m.eval()
m.cache_speaker_embedding = False
mel = synth.synth_mel(1, 300, h.num_mels, seed=1).to(dev)
for _ in range(3): e = m.speaker_embedding(mel)
torch.cuda.synchronize(); print("ok", float(e.abs().max()))
