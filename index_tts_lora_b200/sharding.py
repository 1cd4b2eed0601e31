"""Host-side work partitioning for multi-GPU decode.  Pure Python, no CUDA.

* Utterance sharding (BASELINE config 4): utterances are independent (the generator has no
  cross-utterance state — SURVEY §0.4), decode cost is proportional to latent frames, so a
  longest-processing-time greedy assignment balances sum(frames) per GPU.  No collective: the
  host scatters latents and gathers waveforms.
* Time split of one long utterance (BASELINE config 5): contiguous frame shards, each decoded
  with `halo` extra frames per side (overlap-recompute, exact when halo >= the generator's
  receptive field, `bvg_receptive_field_frames`) through `bvg_decode_shard`.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple


def lpt_assign(lengths: Sequence[int], n_ranks: int) -> List[List[int]]:
    """Greedy LPT: returns, per rank, the indices of the utterances it decodes."""
    if n_ranks < 1:
        raise ValueError("n_ranks must be >= 1")
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    loads = [0] * n_ranks
    shards: List[List[int]] = [[] for _ in range(n_ranks)]
    for i in order:
        r = min(range(n_ranks), key=lambda q: (loads[q], q))
        shards[r].append(i)
        loads[r] += int(lengths[i])
    return shards


def shard_loads(lengths: Sequence[int], shards: Sequence[Sequence[int]]) -> List[int]:
    return [sum(int(lengths[i]) for i in s) for s in shards]


def time_shards(total_frames: int, n_ranks: int, halo: int) -> List[Tuple[int, int, int, int]]:
    """Split [0, total_frames) into n_ranks contiguous shards.
    Returns (f_begin, f_end, halo_left, halo_right) per rank; halos are clipped at the
    utterance ends, where the true sequence-edge rules apply instead."""
    if total_frames < n_ranks:
        raise ValueError("fewer frames than ranks")
    base, rem = divmod(total_frames, n_ranks)
    out, f = [], 0
    for r in range(n_ranks):
        n = base + (1 if r < rem else 0)
        fb, fe = f, f + n
        out.append((fb, fe, min(halo, fb), min(halo, total_frames - fe)))
        f = fe
    return out


def length_buckets(lengths: Sequence[int], max_frames_per_batch: int) -> List[List[int]]:
    """Group utterance indices (longest first) into batches whose padded size
    len(batch) * max(len) stays within `max_frames_per_batch`."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    batches: List[List[int]] = []
    cur: List[int] = []
    for i in order:
        if cur and (len(cur) + 1) * int(lengths[cur[0]]) > max_frames_per_batch:
            batches.append(cur)
            cur = []
        cur.append(i)
    if cur:
        batches.append(cur)
    return batches
