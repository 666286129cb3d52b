"""Multi-GPU host logic on CPU: song sharding + the one final gather, world_size 2 over gloo.
The extractor is replaced by a deterministic stand-in so no GPU is needed; the GPU tests cover the kernels."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from etude_b200 import sharding

NOTE_DT = np.dtype([("pitch", np.int32), ("velocity", np.int32), ("onset", np.float64), ("offset", np.float64)])


class FakeExtractor:
    """extract_many stand-in: one record per window of the song, derived from the wave content only."""

    def extract_many(self, waves, as_dicts=False):
        out = []
        for w in waves:
            n = sharding.windows_of(len(w))
            rec = np.zeros(n, NOTE_DT)
            rec["pitch"] = 21 + (np.arange(n) % 88)
            rec["velocity"] = int(abs(float(w[:16].sum())) * 1000) % 128
            rec["onset"] = np.arange(n) * 8.192
            rec["offset"] = rec["onset"] + float(len(w)) / 16000.0
            out.append(rec)
        return out


def _waves():
    rng = np.random.default_rng(3)
    lens = [256 * 512 * 3 - 1, 5000, 256 * 511, 256 * 512, 256 * 2000 + 17, 1025, 256 * 1024 * 2]
    return [rng.random(n).astype(np.float32) for n in lens]


def test_windows_of_matches_reference_loop():
    for n in (1025, 256 * 511, 256 * 512 - 1, 256 * 512, 480000, 3840000):
        t = 1 + n // 256
        assert sharding.windows_of(n) == len(range(0, t, 512))      # extractor.py:227
    assert sharding.windows_of(3840000) == 30 and sharding.windows_of(480000) == 4   # sizes at the configs


def test_shard_songs_partition_and_balance():
    lens = [len(w) for w in _waves()]
    for world in (1, 2, 3, 8):
        shards = sharding.shard_songs(lens, world)
        assert sorted(i for s in shards for i in s) == list(range(len(lens)))
        loads = [sum(sharding.windows_of(lens[i]) for i in s) for s in shards]
        assert max(loads) - min(loads) <= max(sharding.windows_of(n) for n in lens)
    # 256 equal songs over 8 ranks: 32 each (BASELINE config 4)
    shards = sharding.shard_songs([3840000] * 256, 8)
    assert all(len(s) == 32 for s in shards)


def test_window_table_matches_reference_padding():
    lens = [256 * 600 + 19, 256 * 512, 1025]
    rows = [((1 + n // 256 + 511) // 512) * 512 for n in lens]
    feat_off = np.concatenate([[0], np.cumsum([r + 64 for r in rows])])
    roll_off = np.concatenate([[0], np.cumsum(rows)])
    win, out = sharding.window_table(lens, feat_off, roll_off)
    assert win == [0, 512, 1088, 1600, 2176] and out == [0, 512, 1024, 1536, 2048]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        waves = _waves()
        res = sharding.extract_sharded(FakeExtractor(), waves, dst=0)
        if rank == 0:
            q.put([r.tobytes() for r in res])
        else:
            assert res is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_sharded_equals_single_process_gloo_world2():
    """Sharded over 2 ranks (gloo) == unsharded, record for record, in song order."""
    waves = _waves()
    want = [r.tobytes() for r in FakeExtractor().extract_many(waves)]
    assert [r.tobytes() for r in sharding.extract_sharded(FakeExtractor(), waves)] == want   # no process group: identity
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert got == want


def test_pipeline_groups_partition_every_song_once():
    """extract_many's group pipeline covers songs 0..n-1 exactly once, in order, with one-song groups at both ends."""
    from etude_b200.extractor import _pipeline_groups
    for n in list(range(1, 40)) + [100, 257]:
        for g in (1, 2, 4, 8):
            groups = _pipeline_groups(n, g)
            assert groups[0][0] == 0 and groups[-1][1] == n
            assert all(a < b for a, b in groups)
            assert all(groups[i][1] == groups[i + 1][0] for i in range(len(groups) - 1))
            assert max(b - a for a, b in groups) <= max(g, 1)
            if n > 4 * g and g > 1:
                assert groups[0][1] - groups[0][0] == 1 and groups[-1][1] - groups[-1][0] == 1
