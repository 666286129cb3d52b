"""Song/window sharding over the GPUs of one box.

The Extract path has no exchange step: a window needs only its own 576 feature rows and a song only its own
windows, so ranks work on disjoint songs with the weights replicated and **no collective on the hot path**.  The
only communication is one final gather of the per-song note records (a few hundred KB per song) to the rank that
writes the results -- `gather_notes` below, a single `all_gather_object`/`gather_object` over the process group
(NCCL over NVLink on the GPU box, gloo in the CPU tests).

The reference runs one song at a time on one device (prepare.py:278-306); this module is the additive multi-GPU
layer above the drop-in class.
"""
from typing import List, Sequence

import numpy as np


def windows_of(n_samples: int, hop: int = 256, num_frame: int = 512) -> int:
    """Number of 512-frame windows `_transcript` runs for a song (extractor.py:211,227): ceil((1 + N // hop) / 512)."""
    t = 1 + int(n_samples) // hop
    return (t + num_frame - 1) // num_frame


def shard_songs(n_samples: Sequence[int], world: int) -> List[List[int]]:
    """Assigns songs to ranks, balancing the number of windows (the unit of work): longest-processing-time greedy,
    deterministic (ties broken by song index).  Returns one sorted list of song indices per rank."""
    if world < 1:
        raise ValueError("world must be >= 1")
    load = [0] * world
    out = [[] for _ in range(world)]
    order = sorted(range(len(n_samples)), key=lambda i: (-windows_of(n_samples[i]), i))
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += windows_of(n_samples[i])
    return [sorted(s) for s in out]


def window_table(n_samples: Sequence[int], feat_row_off: Sequence[int], roll_row_off: Sequence[int], num_frame: int = 512):
    """(win_rows, out_rows) for every window of every song of a shard: window i of song s reads padded feature rows
    [feat_row_off[s] + 512 i, + 576) and writes roll rows [roll_row_off[s] + 512 i, + 512) (extractor.py:227-248)."""
    win_rows, out_rows = [], []
    for s, n in enumerate(n_samples):
        for i in range(windows_of(n)):
            win_rows.append(int(feat_row_off[s]) + i * num_frame)
            out_rows.append(int(roll_row_off[s]) + i * num_frame)
    return win_rows, out_rows


def gather_notes(local_notes: List[np.ndarray], local_song_ids: Sequence[int], n_songs: int, group=None, dst: int = 0):
    """The one final gather: every rank contributes the note records of its songs; rank `dst` gets the full list in
    song order (other ranks get None).  With no initialised process group this is the identity."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        out = [None] * n_songs
        for i, r in zip(local_song_ids, local_notes):
            out[i] = r
        return out
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    payload = (list(local_song_ids), list(local_notes))
    gathered = [None] * world if rank == dst else None
    dist.gather_object(payload, gathered, dst=dst, group=group)
    if rank != dst:
        return None
    out = [None] * n_songs
    for ids, recs in gathered:
        for i, r in zip(ids, recs):
            out[i] = r
    missing = [i for i, r in enumerate(out) if r is None]
    if missing:
        raise RuntimeError(f"gather_notes: songs {missing[:8]} were not produced by any rank")
    return out


def extract_sharded(extractor, waves, group=None, dst: int = 0, as_dicts: bool = False):
    """Transcribes `waves` (the same list on every rank) with each rank running `extract_many` on its shard; returns
    the note lists in song order on rank `dst`."""
    import torch.distributed as dist

    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank(group) if world > 1 else 0
    mine = shard_songs([len(w) for w in waves], world)[rank]
    recs = extractor.extract_many([waves[i] for i in mine], as_dicts=as_dicts) if mine else []
    return gather_notes(recs, mine, len(waves), group=group, dst=dst)
