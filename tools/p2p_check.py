"""Multi-process check of the time-split P2P decode (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/p2p_check.py [F] [tiny|full]
Every rank decodes its frame range with per-stage NVLink halo exchange (CUDA-IPC peer stores + flags);
rank 0 also decodes the whole utterance alone and compares."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import warnings; warnings.filterwarnings("ignore")
import torch, torch.distributed as dist
from index_tts_lora_b200 import synth
from index_tts_lora_b200.config import default_config, tiny_config
from index_tts_lora_b200.longform import TimeSplitP2P
from index_tts_lora_b200.models import BigVGAN

torch.set_grad_enabled(False)
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", lr); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
F = int(sys.argv[1]) if len(sys.argv) > 1 else 1406
h = tiny_config() if (len(sys.argv) > 2 and sys.argv[2] == "tiny") else default_config()
m = BigVGAN(h)
m.load_state_dict(synth.synth_state_dict(m.state_dict(), seed=1234, profile="stress"))
m = m.to(dev).eval(); m.precision = "bf16"
lat = synth.synth_latent(1, F, h.gpt_dim, seed=5).to(dev).to(torch.bfloat16)     # same on every rank
emb = m.speaker_embedding(synth.synth_mel(1, 300, h.num_mels, seed=1).to(dev))
ts = TimeSplitP2P(m, F, rank, world).setup(dev).connect_distributed()
win = ts.window(lat)
times = []
for it in range(4):
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = ts.decode(win, emb); e1.record(); torch.cuda.synchronize()
    ts.check()
    times.append(e0.elapsed_time(e1))
n_max = max(s[1] - s[0] for s in ts.shards) * ts.up
pad = torch.zeros(n_max, device=dev); pad[: out.numel()] = out
allo = [torch.zeros(n_max, device=dev) for _ in range(world)]
dist.all_gather(allo, pad)
t = torch.tensor([min(times[1:])], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    split = torch.cat([allo[r][: (ts.shards[r][1] - ts.shards[r][0]) * ts.up] for r in range(world)])
    whole = m.decode(lat, emb, out_dtype=torch.float32)[0, 0]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); m.decode(lat, emb, out_dtype=torch.float32); e1.record(); torch.cuda.synchronize()
    err = (split - whole).abs().max().item()
    print(f"P2P time split x{world}: F={F} max-abs vs single-GPU decode {err:.3e}; "
          f"{t.item():.2f} ms per decode (max over ranks) vs {e0.elapsed_time(e1):.2f} ms on one GPU; "
          f"audio {F*1024/24000:.1f} s", flush=True)
    assert err < 1e-6, err
dist.barrier(); dist.destroy_process_group()
