import json
import os
import sys
import warnings

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
warnings.filterwarnings("ignore", category=FutureWarning)
torch.set_grad_enabled(False)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def load_golden(tag):
    g = np.load(os.path.join(GOLDEN, tag + ".npz"))
    meta = json.loads(str(g["meta"])) if "meta" in g.files else {}
    return g, meta


@pytest.fixture(scope="session")
def full_template():
    """state-dict template (keys/shapes) of the full-config generator, built once."""
    from index_tts_lora_b200.config import default_config
    from index_tts_lora_b200.models import BigVGAN

    m = BigVGAN(default_config())
    return m


def build_case(tag, h, module=None):
    """Rebuild (module, state dict, latent, mel, golden) for a golden model case."""
    from index_tts_lora_b200 import synth
    from index_tts_lora_b200.models import BigVGAN

    g, meta = load_golden(tag)
    m = module if module is not None else BigVGAN(h)
    sd = synth.synth_state_dict(m.state_dict(), seed=meta["seed"], profile=meta["profile"])
    lat = synth.synth_latent(meta["B"], meta["F"], h["gpt_dim"], seed=0)
    mel = synth.synth_mel(meta["mel_B"], meta["Tm"], h["num_mels"], seed=1)
    assert abs(synth.checksum(lat) - float(g["latent_checksum"])) < 1e-9, "latent RNG drift"
    assert abs(synth.checksum(mel) - float(g["mel_checksum"])) < 1e-9, "mel RNG drift"
    return m, sd, lat, mel, g
