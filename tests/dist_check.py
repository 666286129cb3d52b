"""Multi-GPU check of the product's sharding layer on real GPUs (run under torchrun, one rank per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py

(a) song sharding: `sharding.extract_sharded` over all ranks == `extract_many` on one GPU, record for record, after the final
    NCCL gather of the note records to rank 0;
(b) window sharding: `sharding.extract_window_sharded` splits ONE song's windows over the ranks, gathers the roll rows into
    rank 0's tensors and decodes there == the single-GPU rolls bit for bit and the same notes.
Rank 0 computes the single-GPU results itself.  Exit code 0 = all equal."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    world = dist.get_world_size()
    import gpu_diag as D
    from etude_b200 import sharding, synth
    from etude_b200 import AMTAPC_Extractor, ExtractorConfig
    from oracle import model as omodel
    sd = omodel.init_state_dict(0)
    ckpt = f"/tmp/etude_dist_sd_{rank}.pth"
    torch.save(sd, ckpt)
    ex = AMTAPC_Extractor(ExtractorConfig(), ckpt, device=f"cuda:{local}", max_windows=8)
    ok = True
    # ---- (a) songs over ranks
    waves = [synth.tones(256 * 700 + 3, 31), synth.noise(256 * 300, 32), synth.noise(256 * 1100 + 77, 33), synth.tones(256 * 20, 34),
             synth.noise(256 * 513, 35), synth.noise(256 * 2100, 36), synth.tones(256 * 512, 37)]
    got = sharding.extract_sharded(ex, waves, dst=0)
    if rank == 0:
        want = ex.extract_many(waves, as_dicts=False)
        same = len(got) == len(want) and all(a.tobytes() == b.tobytes() for a, b in zip(got, want))
        print(f"[dist_check] song sharding over {world} ranks: {len(want)} songs, {sum(len(w) for w in want)} notes, identical to one GPU: {same}", flush=True)
        ok &= same
    else:
        assert got is None
    # ---- (b) one song's windows over ranks
    wave = synth.noise(256 * 512 * 5 + 1000, 41)          # 6 windows
    res = sharding.extract_window_sharded(ex, wave, dst=0, return_rolls=True)
    if rank == 0:
        rec, rolls = res
        want_rec, want_rolls, _, rows = ex.extract_many([wave], as_dicts=False, return_rolls=True)
        same_rolls = all(torch.equal(a, b) for a, b in zip(rolls, want_rolls))
        same_notes = rec.tobytes() == want_rec[0].tobytes()
        print(f"[dist_check] window sharding over {world} ranks: {rows[0] // 512} windows, rolls bit-identical: {same_rolls}, "
              f"{len(rec)} notes identical: {same_notes}", flush=True)
        ok &= same_rolls and same_notes
    flag = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
    dist.broadcast(flag, src=0)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
