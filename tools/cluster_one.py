"""One activated AMP layer, cluster launch vs plain launch: where do they differ (debugging aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import warnings; warnings.filterwarnings("ignore")
import torch
from index_tts_lora_b200 import synth, _lib
from index_tts_lora_b200.config import AttrDict
from index_tts_lora_b200.models import AMPBlock1
from index_tts_lora_b200.ops import amp_layer
torch.set_grad_enabled(False)
dev = torch.device("cuda:0"); lib = _lib.load()
C_, T, k, d = int(os.environ.get("CLC", "384")), int(os.environ.get("CLT", "257")), 3, 1
blk = AMPBlock1(AttrDict(snake_logscale=True), C_, k, (d, d, d), activation="snakebeta")
blk.load_state_dict(synth.synth_state_dict(blk.state_dict(), seed=301, profile="stress"))
x = torch.randn(2, C_, T).to(dev)
lib.bvg_set_tc_cluster(0)
y0 = amp_layer(x, blk.convs1[0], blk.activations[0], precision="bf16").float()
lib.bvg_set_tc_cluster(1)
y1 = amp_layer(x, blk.convs1[0], blk.activations[0], precision="bf16").float()
torch.cuda.synchronize()
print("plain max", float(y0.abs().max()), "cluster max", float(y1.abs().max()), "equal", bool(torch.equal(y0, y1)))
bad = ~(y0 == y1)
print("bad elements", int(bad.sum()), "of", bad.numel(), " nan", int(torch.isnan(y1).sum()))
for b in range(2):
    for ct in range(0, C_, 256):
        for tt in range(0, T, 256):
            blkb = bad[b, ct:ct + 256, tt:tt + 256]
            print(f"  b={b} cols {ct}.. rows {tt}..: bad {int(blkb.sum())}/{blkb.numel()}  nan {int(torch.isnan(y1[b, ct:ct+256, tt:tt+256]).sum())}")
yb = y1.clone(); yb[torch.isnan(yb)] = 1e9
dd = (y0 - yb).abs()
print("max diff", float(dd.max()), "median diff over bad", float(dd[bad].median()))
colbad = bad[0].any(dim=1).nonzero().flatten().tolist()
print("b=0 bad columns:", len(colbad), colbad[:20], "...", colbad[-8:])
cb = bad[0].sum(dim=1)
print("b=0 bad count per column (first 40):", cb[:40].tolist())
rb = bad[0, :256].sum(dim=0)
print("b=0 tile0 bad count per row (first 40):", rb[:40].tolist(), " rows 120..136:", rb[120:136].tolist())
print("diff per column mean (first 16):", [round(float(v), 4) for v in dd[0, :16].mean(dim=1)])
