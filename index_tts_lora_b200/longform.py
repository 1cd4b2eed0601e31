"""Long-form decode: ONE utterance split along time across GPUs (BASELINE config 5).

Two strategies, both exact on every rank's own samples:

* ``decode_overlap`` — whole-generator overlap-recompute through ``bvg_decode_shard``: each rank
  decodes its frame range plus ``bvg_receptive_field_frames`` (36) halo frames per side.  No
  communication; ~39 % redundant work at 8 x 7.5 s.
* ``TimeSplitP2P`` — per-stage halo exchange over NVLink peer memory (``bvg_shard_*``): 6
  exchanges per decode, ~6 % redundant work.  ``torch.distributed`` is used only to swap the
  CUDA-IPC handles once and as the host barrier between decodes; the data path is peer stores
  issued by a CUDA kernel plus system-scope flags — no NCCL.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from .sharding import time_shards


def decode_overlap(model, latent_full: torch.Tensor, emb: torch.Tensor, rank: int, world: int,
                   out_dtype=torch.float32) -> Tuple[torch.Tensor, Tuple[int, int]]:
    """Rank's own samples of a [1, F, D] latent by overlap-recompute. Returns (wav[n], (f_begin, f_end))."""
    lib = _lib.load()
    dev = latent_full.device
    plan = model._ensure_plan(dev)
    F = latent_full.shape[1]
    rf = lib.bvg_receptive_field_frames(plan)
    fb, fe, hl, hr = time_shards(F, world, rf)[rank]
    win = latent_full[0, fb - hl: fe + hr].contiguous()
    up = 1
    for u in model.h.upsample_rates:
        up *= int(u)
    out = torch.empty((fe - fb) * up, device=dev, dtype=out_dtype)
    e = emb.reshape(1, -1).float().contiguous()
    with torch.cuda.device(dev):
        _lib.check(lib.bvg_decode_shard(plan, win.data_ptr(), _lib.torch_dtype_code(win.dtype), fb, fe, F, hl, hr,
                                        e.data_ptr(), out.data_ptr(), _lib.torch_dtype_code(out_dtype),
                                        model._precision_code(), _lib.stream_ptr(dev)), "bvg_decode_shard")
    return out, (fb, fe)


class TimeSplitP2P:
    """One rank of a time-split decode with per-stage NVLink P2P halo exchange (bf16 path)."""

    def __init__(self, model, f_total: int, rank: int, world: int):
        self.model, self.rank, self.world, self.F = model, rank, world, f_total
        self.lib = _lib.load()
        self.shards = time_shards(f_total, world, 0)
        self.fb, self.fe = self.shards[rank][0], self.shards[rank][1]
        self.epoch = 0
        self.plan = None
        self.n_phases = len(model.h.upsample_rates) + 1
        up = 1
        for u in model.h.upsample_rates:
            up *= int(u)
        self.up = up

    # -- setup ---------------------------------------------------------------------------
    def setup(self, device: torch.device):
        own = [s[1] - s[0] for s in self.shards]
        g = _lib.BvgShardGeom(self.fb, self.fe, self.F,
                              own[self.rank - 1] if self.rank > 0 else 0,
                              own[self.rank + 1] if self.rank + 1 < self.world else 0, max(own))
        self.plan = self.model._ensure_plan(device)
        self.device = device
        with torch.cuda.device(device):
            _lib.check(self.lib.bvg_shard_setup(self.plan, C.byref(g), _lib.stream_ptr(device)), "bvg_shard_setup")
        self.halo = self.lib.bvg_shard_halo_frames(self.plan)
        return self

    def export_handles(self) -> bytes:
        buf = (C.c_uint8 * 192)()
        _lib.check(self.lib.bvg_shard_export(self.plan, buf), "bvg_shard_export")
        return bytes(buf)

    def connect_handles(self, left: Optional[bytes], right: Optional[bytes]):
        for side, h in ((0, left), (1, right)):
            if h is not None:
                arr = (C.c_uint8 * 192).from_buffer_copy(h)
                _lib.check(self.lib.bvg_shard_connect(self.plan, side, arr), "bvg_shard_connect")

    def local_ptrs(self):
        a, b, f = C.c_void_p(), C.c_void_p(), C.c_void_p()
        _lib.check(self.lib.bvg_shard_local_ptrs(self.plan, C.byref(a), C.byref(b), C.byref(f)), "bvg_shard_local_ptrs")
        return a, b, f

    def connect_ptrs(self, side: int, ptrs):
        _lib.check(self.lib.bvg_shard_connect_ptr(self.plan, side, *ptrs), "bvg_shard_connect_ptr")

    def connect_distributed(self, group=None):
        """Swap IPC handles with the neighbours through torch.distributed (setup time only)."""
        import torch.distributed as dist

        mine = self.export_handles()
        allh: List[Optional[bytes]] = [None] * self.world
        dist.all_gather_object(allh, mine, group=group)
        self.connect_handles(allh[self.rank - 1] if self.rank > 0 else None,
                             allh[self.rank + 1] if self.rank + 1 < self.world else None)
        return self

    # -- decode --------------------------------------------------------------------------
    def window(self, latent_full: torch.Tensor) -> torch.Tensor:
        """This rank's latent frames plus the conv_pre halo on every non-end side."""
        hl = self.halo if self.fb > 0 else 0
        hr = self.halo if self.fe < self.F else 0
        return latent_full[0, self.fb - hl: self.fe + hr].contiguous()

    def run_phase(self, phase: int, win: Optional[torch.Tensor], emb: Optional[torch.Tensor],
                  out: Optional[torch.Tensor], wait: bool):
        dev = self.device
        with torch.cuda.device(dev):
            _lib.check(self.lib.bvg_shard_run(
                self.plan, phase,
                win.data_ptr() if win is not None else None,
                _lib.torch_dtype_code(win.dtype) if win is not None else 0,
                emb.data_ptr() if emb is not None else None,
                out.data_ptr() if out is not None else None,
                _lib.torch_dtype_code(out.dtype) if out is not None else 0,
                self.epoch, int(wait), _lib.stream_ptr(dev)), f"bvg_shard_run(phase {phase})")

    def decode(self, win: torch.Tensor, emb: torch.Tensor, out_dtype=torch.float32) -> torch.Tensor:
        """All phases on this rank (multi-process use: neighbours run the same call concurrently).

        A phase that times out waiting for a neighbour's halo rows (~4 s) leaves a code in a mapped host word; the
        waveform of that decode is then garbage.  The word is checked here for the PREVIOUS decode (free: no device
        access) and every later phase of the plan refuses to run; call ``check()`` after synchronising the stream to
        validate the decode just enqueued."""
        self.check()
        self.epoch += 1
        out = torch.empty((self.fe - self.fb) * self.up, device=self.device, dtype=out_dtype)
        e = emb.reshape(1, -1).float().contiguous()
        for ph in range(self.n_phases):
            self.run_phase(ph, win if ph == 0 else None, e if ph == 0 else None,
                           out if ph == self.n_phases - 1 else None, wait=True)
        return out

    def check(self):
        """Raise if a phase enqueued so far timed out on a neighbour (final once the stream is synchronised)."""
        e = self.lib.bvg_shard_error(self.plan)
        if e > 0:
            raise RuntimeError("time-split decode: timed out waiting for the "
                               f"{'left' if e == 1 else 'right'} neighbour's halo rows; the waveform is invalid")


def emulate_time_split(models: Sequence, latent_full: torch.Tensor, emb: torch.Tensor,
                       out_dtype=torch.float32) -> torch.Tensor:
    """All ranks of a time split inside ONE process on one device (tests): every rank is a model
    replica with its own plan; ranks are connected by raw pointers and phase p of every rank is
    enqueued before phase p+1 of any, so stream order replaces the device-side waits."""
    world = len(models)
    dev = latent_full.device
    F = latent_full.shape[1]
    ranks = [TimeSplitP2P(m, F, r, world).setup(dev) for r, m in enumerate(models)]
    ptrs = [r.local_ptrs() for r in ranks]
    for i, r in enumerate(ranks):
        if i > 0:
            r.connect_ptrs(0, ptrs[i - 1])
        if i + 1 < world:
            r.connect_ptrs(1, ptrs[i + 1])
    outs = [torch.empty((r.fe - r.fb) * r.up, device=dev, dtype=out_dtype) for r in ranks]
    wins = [r.window(latent_full) for r in ranks]
    e = emb.reshape(1, -1).float().contiguous()
    for r in ranks:
        r.epoch += 1
    for ph in range(ranks[0].n_phases):
        for i, r in enumerate(ranks):
            r.run_phase(ph, wins[i] if ph == 0 else None, e if ph == 0 else None,
                        outs[i] if ph == r.n_phases - 1 else None, wait=False)
    torch.cuda.synchronize(dev)
    return torch.cat(outs)
