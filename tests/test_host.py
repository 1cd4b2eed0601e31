"""CPU tests (-m "not gpu") of the host logic: state-dict compatibility with the reference
checkpoint layout, weight-norm folding, sharding / time-split planning (incl. a world_size-2
gloo run), the C-ABI library's exported symbols, and loud failure without a GPU."""
import ctypes
import json
import os
import re
import subprocess
import sys

import pytest
import torch

from conftest import GOLDEN, ROOT


def test_state_dict_keys_match_reference_checkpoint_layout():
    """bigvgan_generator.pth['generator'] layout: 1029 entries with weight-norm,
    913 after remove_weight_norm — names AND shapes identical to the reference module's."""
    from index_tts_lora_b200.config import default_config
    from index_tts_lora_b200.models import BigVGAN
    keys = json.load(open(os.path.join(GOLDEN, "state_dict_keys.json")))
    m = BigVGAN(default_config())
    mine = {k: list(v.shape) for k, v in m.state_dict().items()}
    assert len(mine) == 1029
    assert mine == keys["weight_norm"]
    assert list(mine) == list(keys["weight_norm"])          # same order too
    m.remove_weight_norm()
    mine2 = {k: list(v.shape) for k, v in m.state_dict().items()}
    assert len(mine2) == 913 and mine2 == keys["folded"]
    assert m.h["use_cuda_kernel"] is False                  # models.py:142 writes the flag


def test_weight_norm_folding_and_g_shapes():
    """weight-norm g is per OUTPUT channel for Conv1d, per INPUT channel for
    ConvTranspose1d; folding in fp32 equals torch's remove_weight_norm."""
    from index_tts_lora_b200.config import tiny_config
    from index_tts_lora_b200.models import BigVGAN, folded_weight
    from oracle import bigvgan_oracle as O
    m = BigVGAN(tiny_config())
    assert tuple(m.conv_pre.weight_g.shape) == (512, 1, 1)
    assert tuple(m.ups[0][0].weight_g.shape) == (512, 1, 1) and tuple(m.ups[0][0].weight_v.shape) == (512, 256, 8)
    with torch.no_grad():
        m.ups[0][0].weight_g.mul_(1.7)
    w_mine = folded_weight(m.ups[0][0])
    w_orc = O.fold_weight_norm(m.ups[0][0].weight_g.detach(), m.ups[0][0].weight_v.detach())
    m.remove_weight_norm()
    assert torch.allclose(w_mine, m.ups[0][0].weight.detach(), atol=1e-7)
    assert torch.allclose(w_orc, m.ups[0][0].weight.detach(), atol=1e-7)
    names = m.generator_tensors()
    assert "conv_pre.weight" in names and "resblocks.17.activations.5.downsample.lowpass.filter" in names
    assert not any(k.startswith("speaker_encoder") for k in names)


def test_module_protocol_used_by_infer_py():
    """infer.py:392-410: load_state_dict, .to(), dtype casts, BatchNorm iteration, eval()."""
    from index_tts_lora_b200.config import tiny_config
    from index_tts_lora_b200.models import BigVGAN
    from index_tts_lora_b200 import synth
    m = BigVGAN(tiny_config())
    sd = synth.synth_state_dict(m.state_dict(), seed=1, profile="init")
    m.load_state_dict(sd)
    assert m._weights_dirty
    m = m.to(torch.bfloat16)
    n_bn = 0
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm1d):
            mod.float()
            n_bn += 1
    assert n_bn > 10 and m.conv_post.bias.dtype == torch.bfloat16
    m.remove_weight_norm()
    m.eval()
    assert m._precision_code() == 1      # bf16 params -> tcgen05 path


def test_no_cpu_fallback_and_reference_error_behaviour():
    from index_tts_lora_b200.config import tiny_config
    from index_tts_lora_b200.models import BigVGAN
    m = BigVGAN(tiny_config()).eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(1, 4, 32), torch.randn(1, 30, 20))
    # models.py:207-211: a 2x speaker batch reaches the missing logit_scale -> AttributeError
    with pytest.raises(AttributeError, match="logit_scale"):
        m(torch.randn(1, 4, 32), torch.randn(2, 30, 20))


def test_speaker_embedding_cache_is_identity_safe():
    from index_tts_lora_b200.config import tiny_config
    from index_tts_lora_b200.models import BigVGAN
    m = BigVGAN(tiny_config()).eval()
    m.cache_speaker_embedding = True
    mel = torch.randn(1, 30, 20)
    e1 = m.speaker_embedding(mel)
    assert m.speaker_embedding(mel) is e1                     # same tensor, same version -> hit
    assert m.speaker_embedding(mel.transpose(1, 2).transpose(1, 2)) is e1   # a view of the same storage
    mel.add_(1.0)                                             # in-place edit bumps the version -> miss
    e2 = m.speaker_embedding(mel)
    assert e2 is not e1 and not torch.allclose(e1, e2)


def test_c_abi_exports_every_declared_symbol():
    """libbvg.so loads without a GPU and exports exactly what include/bvg.h declares."""
    from index_tts_lora_b200 import _lib
    from index_tts_lora_b200 import build as b
    b.build()
    hdr = open(os.path.join(ROOT, "include", "bvg.h")).read()
    declared = set(re.findall(r"\b(bvg_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    lib2 = _lib.load()
    assert lib2.bvg_version() == 100
    # argument validation works without touching a device
    assert lib2.bvg_plan_create(None, 0, None) == -1
    assert b"null" in lib2.bvg_last_error()
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T bvg_" in l}
    assert exported == declared


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: no file under the package may reference it."""
    pkg = os.path.join(ROOT, "index_tts_lora_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("test infrastructure", ""), f


# ----------------------------------------------------------------------------- sharding
def test_lpt_assignment_balances_frames():
    from index_tts_lora_b200 import synth
    from index_tts_lora_b200.sharding import lpt_assign, shard_loads
    lens = synth.synth_lengths(256, 47, 469, seed=2)            # BASELINE config 4
    for n in (1, 2, 4, 8):
        shards = lpt_assign(lens, n)
        assert sorted(i for s in shards for i in s) == list(range(256))
        loads = shard_loads(lens, shards)
        assert max(loads) - min(loads) <= max(lens)
        assert max(loads) / (sum(lens) / n) < 1.01


def test_time_shards_cover_and_clip_halos():
    from index_tts_lora_b200.sharding import time_shards
    sh = time_shards(1406, 8, 36)                               # BASELINE config 5: 60 s on 8 GPUs
    assert sh[0][0] == 0 and sh[-1][1] == 1406
    assert all(a[1] == b[0] for a, b in zip(sh, sh[1:]))
    assert sh[0][2] == 0 and sh[-1][3] == 0 and all(s[2] == 36 for s in sh[1:]) and all(s[3] == 36 for s in sh[:-1])
    with pytest.raises(ValueError):
        time_shards(3, 8, 36)


def test_length_buckets_bound_padding():
    from index_tts_lora_b200.sharding import length_buckets
    lens = [469, 47, 300, 301, 50, 48, 200]
    b = length_buckets(lens, 1000)
    assert sorted(i for g in b for i in g) == list(range(7))
    for g in b:
        assert len(g) == 1 or len(g) * max(lens[i] for i in g) <= 1000


_GLOO_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["BVG_ROOT"])
from index_tts_lora_b200 import synth
from index_tts_lora_b200.sharding import lpt_assign, time_shards
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
lens = synth.synth_lengths(64, 47, 469, seed=2)
mine = lpt_assign(lens, w)[r]
# every rank "decodes" its utterances (stand-in: frames -> samples) and the host gathers
done = torch.zeros(64, dtype=torch.int64)
for i in mine:
    done[i] = lens[i] * 1024
dist.all_reduce(done)
assert done.tolist() == [l * 1024 for l in lens]
load = torch.tensor([sum(lens[i] for i in mine)], dtype=torch.float64)
mx = load.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
tot = load.clone(); dist.all_reduce(tot)
assert mx.item() / (tot.item() / w) < 1.05
fb, fe, hl, hr = time_shards(1406, w, 36)[r]
cover = torch.zeros(1406, dtype=torch.int64); cover[fb:fe] = 1
dist.all_reduce(cover)
assert int(cover.min()) == 1 and int(cover.max()) == 1
if r == 0:
    print("GLOO_OK")
dist.destroy_process_group()
"""


def test_utterance_sharding_world_size_2_gloo(tmp_path):
    """N>1 host path on CPU: 2 ranks over gloo partition the utterances with no data-path
    collective (the all_reduce here only checks coverage / balance)."""
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, BVG_ROOT=ROOT, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "GLOO_OK" in r.stdout


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): exactly one JSON line on stdout with the
    contract's keys, the same metric / unit / workload naming as the GPU arm, and no GPU work claimed."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--workload", "utt6p7s"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "bigvgan_decode_audio_seconds_per_second"
    assert j["unit"] == "audio-s/s" and j["higher_is_better"] is True and j["gpu_launches"] == 0
    assert j["config"]["workload"] == "utt6p7s" and j["value"] > 0
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1
    assert j["e2e"] == {"value": j["value"], "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
