"""B200-native BigVGAN waveform decoder — a drop-in for the
``self.bigvgan(latent, mel_ref)`` path of CreateIntelligens/index-tts-lora
(``indextts/infer.py:748,888`` -> ``indextts/BigVGAN/models.py:203-252``).

``from index_tts_lora_b200.models import BigVGAN as Generator`` replaces
``from indextts.BigVGAN.models import BigVGAN as Generator`` (infer.py:24).
"""
from .config import AttrDict, default_config, load_yaml_config, tiny_config  # noqa: F401

__all__ = ["AttrDict", "default_config", "load_yaml_config", "tiny_config", "BigVGAN"]


def __getattr__(name):  # lazy: importing the package must not require torch to be warm
    if name == "BigVGAN":
        from .models import BigVGAN

        return BigVGAN
    raise AttributeError(name)
