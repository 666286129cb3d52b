"""Song / window sharding over the GPUs of one box.

The Extract path has no exchange step: a window needs only its own 576 feature rows and a song only its own
windows, so ranks work on disjoint songs (or on disjoint window ranges of one long song) with the weights replicated
and **no collective on the hot path**.  The only communication is ONE FINAL GATHER to the rank that writes the results:

* song sharding (`extract_sharded`): the per-song note records, as raw bytes in one tensor per rank, received by the
  destination rank with point-to-point ops (NCCL over NVLink on the GPU box -- the payload goes device to device and
  the destination copies it to the host once; gloo on CPU tensors in the CPU tests);
* window sharding (`extract_window_sharded`): the roll rows of each rank's window range (17.6 MB per 4-minute song in
  total), received straight into the destination rank's roll tensors, which then decodes the notes.

The reference runs one song at a time on one device (prepare.py:278-306); this module is the additive multi-GPU
layer above the drop-in class.
"""
from typing import List, Sequence

import numpy as np

NOTE_DTYPE = np.dtype([("pitch", np.int32), ("velocity", np.int32), ("onset", np.float64), ("offset", np.float64)])


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


def shard_windows(n_windows: int, world: int) -> List[range]:
    """Contiguous, balanced window ranges of ONE song, one per rank (the first `n_windows % world` ranks get one more)."""
    if world < 1:
        raise ValueError("world must be >= 1")
    base, extra = divmod(int(n_windows), world)
    out, a = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append(range(a, a + n))
        a += n
    return out


def window_table(n_samples: Sequence[int], feat_row_off: Sequence[int], roll_row_off: Sequence[int], num_frame: int = 512):
    """(win_rows, out_rows) for every window of every song of a shard: window i of song s reads padded feature rows
    [feat_row_off[s] + 512 i, + 576) and writes roll rows [roll_row_off[s] + 512 i, + 512) (extractor.py:227-248)."""
    win_rows, out_rows = [], []
    for s, n in enumerate(n_samples):
        for i in range(windows_of(n)):
            win_rows.append(int(feat_row_off[s]) + i * num_frame)
            out_rows.append(int(roll_row_off[s]) + i * num_frame)
    return win_rows, out_rows


def _group_info(group):
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return None, 0, 1
    return dist, dist.get_rank(group), dist.get_world_size(group)


def _comm_device(dist, group):
    """Tensors of an NCCL group live on this rank's GPU; gloo moves CPU tensors."""
    import torch
    if dist.get_backend(group) == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def _gather_bytes(dist, group, rank, world, dst, payload, sizes, dev):
    """Point-to-point gather of one uint8 tensor per rank (sizes[r] bytes, known everywhere) to `dst`.
    Returns {rank: tensor} on dst (its own payload included), None elsewhere."""
    import torch
    if rank != dst:
        if sizes[rank]:
            dist.send(payload, dst=dist.get_global_rank(group, dst) if group is not None else dst, group=group)
        return None
    got = {dst: payload}
    reqs = []
    for r in range(world):
        if r == dst or not sizes[r]:
            continue
        buf = torch.empty(sizes[r], dtype=torch.uint8, device=dev)
        got[r] = buf
        reqs.append(dist.irecv(buf, src=dist.get_global_rank(group, r) if group is not None else r, group=group))
    for q in reqs:
        q.wait()
    return got


def gather_notes(local_notes: List[np.ndarray], local_song_ids: Sequence[int], n_songs: int, group=None, dst: int = 0):
    """The one final gather: every rank contributes the note records of its songs; rank `dst` gets the full list in
    song order (other ranks get None).  With no initialised process group this is the identity.

    Metadata (song ids, per-song counts: a few bytes) goes through `all_gather_object`; the records travel as one raw
    byte tensor per rank."""
    import torch
    dist, rank, world = _group_info(group)
    if dist is None or world == 1:
        out = [None] * n_songs
        for i, r in zip(local_song_ids, local_notes):
            out[i] = r
        return out
    dev = _comm_device(dist, group)
    counts = [int(len(r)) for r in local_notes]
    metas = [None] * world
    dist.all_gather_object(metas, (list(map(int, local_song_ids)), counts), group=group)
    sizes = [sum(m[1]) * NOTE_DTYPE.itemsize for m in metas]
    cuda = dev.type == "cuda"
    if sizes[rank] and rank != dst:
        if cuda:
            # no host copy: the records of extract_many already sit in pinned memory (engine._pinned_records), so every
            # song's array is uploaded from where it is (numpy view -> torch view -> asynchronous H2D into one payload)
            payload = torch.empty(sizes[rank], dtype=torch.uint8, device=dev)
            pos = 0
            for r in local_notes:
                b = np.ascontiguousarray(r, dtype=NOTE_DTYPE).view(np.uint8).reshape(-1)
                if b.size:
                    payload[pos : pos + b.size].copy_(torch.from_numpy(b), non_blocking=True)
                pos += b.size
        else:
            payload = torch.empty(sizes[rank], dtype=torch.uint8)
            view, pos = payload.numpy(), 0
            for r in local_notes:
                b = np.ascontiguousarray(r, dtype=NOTE_DTYPE).view(np.uint8).reshape(-1)
                view[pos : pos + b.size] = b
                pos += b.size
    else:
        payload = torch.empty(0, dtype=torch.uint8, device=dev)     # dst keeps its own records where they are
    got = _gather_bytes(dist, group, rank, world, dst, payload, [0 if r == dst else n for r, n in enumerate(sizes)], dev)
    if rank != dst:
        return None
    out = [None] * n_songs
    # one pinned buffer for everything that was received, all D2H copies in flight together, one synchronisation
    total_rx = sum(n for r, n in enumerate(sizes) if r != dst)
    host_all = torch.empty(max(1, total_rx), dtype=torch.uint8, pin_memory=cuda)
    offs, pos = {}, 0
    for r in range(world):
        if r != dst and sizes[r]:
            host_all[pos : pos + sizes[r]].copy_(got[r], non_blocking=cuda)
            offs[r] = pos
            pos += sizes[r]
    if cuda:
        torch.cuda.current_stream().synchronize()
    raw_all = host_all.numpy()
    for r, (ids, cnts) in enumerate(metas):
        if r == dst:
            recs = local_notes
        else:
            raw = raw_all[offs[r] : offs[r] + sizes[r]].view(NOTE_DTYPE) if sizes[r] else np.zeros(0, NOTE_DTYPE)
            recs, p2 = [], 0
            for c in cnts:
                recs.append(raw[p2 : p2 + c])
                p2 += c
        for i, rec in zip(ids, recs):
            out[i] = rec
    missing = [i for i, r in enumerate(out) if r is None]
    if missing:
        raise RuntimeError(f"gather_notes: songs {missing[:8]} were not produced by any rank")
    return out


def extract_sharded(extractor, waves, group=None, dst: int = 0, as_dicts: bool = False, **kw):
    """Transcribes `waves` (the same list on every rank) with each rank running `extract_many` on its shard; returns
    the note lists in song order on rank `dst` (None elsewhere)."""
    dist, rank, world = _group_info(group)
    mine = shard_songs([len(w) for w in waves], world)[rank]
    recs = extractor.extract_many([waves[i] for i in mine], as_dicts=False, **kw) if mine else []
    out = gather_notes(recs, mine, len(waves), group=group, dst=dst)
    if out is not None and as_dicts:
        from .engine import notes_to_dicts
        out = [notes_to_dicts(r) for r in out]
    return out


def extract_window_sharded(extractor, wave, group=None, dst: int = 0, as_dicts: bool = False, return_rolls: bool = False):
    """ONE long song over all ranks: every rank computes the (cheap) log-mel, runs a contiguous range of the song's
    windows through the model, and sends its roll rows to rank `dst`, which decodes the notes -- the "one final gather
    of rolls" (4 arrays, 17.6 MB per 4-minute song).  Per-window results do not depend on how windows are batched, so
    the result equals the single-GPU one bit for bit.  Returns the note records on `dst`, None elsewhere."""
    import torch
    from . import engine as _engine
    dist, rank, world = _group_info(group)
    eng = extractor.engine
    dev = extractor.device
    w = wave if isinstance(wave, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(wave, dtype=np.float32))
    w = w.to(dev, torch.float32).reshape(-1).contiguous()
    n = int(w.numel())
    F = eng.n_frame
    feat, _ = eng.logmel(w, [0], [n])
    t_pad = _engine.feature_rows(n) - 2 * _engine.MARGIN
    n_win = t_pad // F
    ranges = shard_windows(n_win, world)
    mine = ranges[rank]
    rolls = eng.alloc_rolls(t_pad, dev)
    if len(mine):
        rows = [i * F for i in mine]
        eng.forward_windows(feat, rows, rows, rolls)
    if dist is not None and world > 1:
        cdev = _comm_device(dist, group)
        torch.cuda.synchronize(dev)
        glob = (lambda r: dist.get_global_rank(group, r)) if group is not None else (lambda r: r)
        if rank == dst:
            reqs, staged = [], []
            for r, rg in enumerate(ranges):
                if r == dst or not len(rg):
                    continue
                for k, t in enumerate(rolls):
                    view = t[rg.start * F : rg.stop * F]            # contiguous rows of the destination tensor
                    if cdev.type == "cuda":
                        reqs.append(dist.irecv(view, src=glob(r), group=group))
                    else:
                        buf = torch.empty(view.shape, dtype=view.dtype)
                        staged.append((view, buf))
                        reqs.append(dist.irecv(buf, src=glob(r), group=group))
            for q in reqs:
                q.wait()
            for view, buf in staged:
                view.copy_(buf)
        elif len(mine):
            for t in rolls:
                part = t[mine.start * F : mine.stop * F]
                dist.send(part if cdev.type == "cuda" else part.cpu(), dst=glob(dst), group=group)
    if rank != dst:
        return None
    cfg = extractor.config.infer
    hop_sec = float(extractor.config.feature.hop_sample / extractor.config.feature.sr)
    rec = eng.notes(rolls[0], rolls[1], rolls[2], rolls[3], [0], [t_pad], cfg.onset_threshold, cfg.offset_threshold,
                    cfg.frame_threshold, note_min=extractor.config.midi.note_min, hop_sec=hop_sec)[0]
    out = _engine.notes_to_dicts(rec) if as_dicts else rec
    return (out, rolls) if return_rolls else out
