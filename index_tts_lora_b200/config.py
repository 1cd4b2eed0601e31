"""Config handling for the BigVGAN decode path.

The reference passes the ``bigvgan:`` block of ``finetune_models/config.yaml:88-146`` as an
OmegaConf ``DictConfig`` (``infer.py:210,390``); ``BigVGAN.__init__`` reads it by attribute
and by ``.get`` and WRITES ``h["use_cuda_kernel"]`` (``models.py:142``).  ``AttrDict`` gives a
plain dict the same three access styles so the drop-in works with either.
"""
from __future__ import annotations

import copy


class AttrDict(dict):
    """dict with attribute access (enough of the DictConfig protocol for models.py:142-199)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover - mirrors attribute protocol
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def __deepcopy__(self, memo):
        return AttrDict(copy.deepcopy(dict(self), memo))


# The generator-relevant keys of finetune_models/config.yaml:94-107,132 (values are
# configuration, not code).  GAN-training keys of the block are irrelevant to the path.
DEFAULT_BIGVGAN_CONFIG = {
    "resblock": "1",
    "upsample_rates": [4, 4, 4, 4, 2, 2],
    "upsample_kernel_sizes": [8, 8, 4, 4, 4, 4],
    "upsample_initial_channel": 1536,
    "resblock_kernel_sizes": [3, 7, 11],
    "resblock_dilation_sizes": [[1, 3, 5], [1, 3, 5], [1, 3, 5]],
    "feat_upsample": False,
    "speaker_embedding_dim": 512,
    "cond_d_vector_in_each_upsampling_layer": True,
    "gpt_dim": 1280,
    "activation": "snakebeta",
    "snake_logscale": True,
    "num_mels": 100,
    "sampling_rate": 24000,
}


def default_config() -> AttrDict:
    return AttrDict(copy.deepcopy(DEFAULT_BIGVGAN_CONFIG))


def tiny_config(c0: int = 512, gpt_dim: int = 32, num_mels: int = 20, spk: int = 16) -> AttrDict:
    """A structurally identical but small generator for fast CPU-side tests."""
    h = default_config()
    h.update(upsample_initial_channel=c0, gpt_dim=gpt_dim, num_mels=num_mels,
             speaker_embedding_dim=spk)
    return h


def load_yaml_config(path: str) -> AttrDict:
    """Read ``bigvgan:`` from a reference-style config.yaml (infer.py:210)."""
    import yaml

    with open(path) as f:
        return AttrDict(yaml.safe_load(f)["bigvgan"])


def total_upsample(h) -> int:
    r = 1
    for u in h["upsample_rates"]:
        r *= int(u)
    return r
