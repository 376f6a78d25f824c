#!/usr/bin/env python
"""torchrun-launched NCCL parity of the sharded forward: on N GPUs, ONE batch sharded over the
ranks must reproduce the unsharded forward of the same batch BIT FOR BIT on every row of every output -- with the
outputs left resident, gathered by the grouped NCCL operation (to every rank / to rank 0) and written into rank 0's
memory by the producing kernels (PeerGather).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/sharded_parity.py [--batch 64]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import smart_nar_fast_tts_b200 as pkg  # noqa: E402
from smart_nar_fast_tts_b200 import synthetic  # noqa: E402

IDX = (0, 1, 2, 3, 4, 5, 6, 7, 9)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    args = ap.parse_args()
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    sd = synthetic.make_state_dict(0)
    model = synthetic.build_module(sd, synthetic.STATS_NAN_BINS, device=dev)
    sp, tx, sl, L = synthetic.make_inputs(args.batch, 40, 120, seed=1)
    sp, tx, sl = sp.to(dev), tx.to(dev), sl.to(dev)
    full = model(sp, tx, sl, L)                       # the unsharded forward, on every rank's own GPU
    torch.cuda.synchronize()
    synth = pkg.ShardedSynthesizer(model)
    bounds = synth.bounds(sl)
    lo, hi = bounds[rank]
    report = {"world": world, "batch": args.batch, "bounds": bounds, "T": int(full[1].shape[1])}

    def same(a, b, what):
        bad = [i for i in IDX if not (a[i].shape == b[i].shape and torch.equal(a[i], b[i]))]
        assert not bad, f"rank {rank}: {what}: outputs {bad} differ from the unsharded forward"

    local_out = synth(sp, tx, sl, L, bounds=bounds)
    same(local_out, tuple(t[lo:hi] if isinstance(t, torch.Tensor) else t for t in full), "resident shard")
    every = synth(sp, tx, sl, L, gather=True, bounds=bounds)
    same(every, full, "all-gather (grouped NCCL)")
    root = synth(sp, tx, sl, L, gather="root", bounds=bounds)
    if rank == 0:
        same(root, full, "gather to rank 0 (grouped NCCL)")
    try:
        T = int(full[1].shape[1])
        synth.enable_peer_gather(args.batch, T + 8, 80, dst=0)
        for _ in range(2):                            # twice: the symmetric buffers are reused
            peer = synth(sp, tx, sl, L, gather="peer", bounds=bounds)
            torch.cuda.synchronize()
            if rank == 0:
                same(peer, full, "peer-memory writes (mel / mel_post stored into rank 0 by the kernels)")
        model.set_mel_post_layout(True)
        full_cm = model(sp, tx, sl, L)
        peer = synth(sp, tx, sl, L, gather="peer", bounds=bounds)
        torch.cuda.synchronize()
        if rank == 0:
            same(peer, full_cm, "peer-memory writes, channel-major mel_post")
        model.set_mel_post_layout(False)
        report["peer"] = "ok"
    except AssertionError:
        raise
    except Exception as e:                            # no symmetric memory on this box: say so, the NCCL paths still count
        report["peer"] = f"unavailable: {type(e).__name__}: {e}"[:300]
    dist.barrier()
    if rank == 0:
        report["result"] = "sharded == unsharded, bit for bit (resident, all-gather, gather-to-root" + \
                           (", peer writes)" if report["peer"] == "ok" else ")")
        print(json.dumps(report), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
