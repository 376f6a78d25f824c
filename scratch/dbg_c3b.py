import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from smart_nar_fast_tts_b200 import pipeline as P, synthetic, StreamedSynthesizer
dev = torch.device("cuda", 0)
m = synthetic.build_module(synthetic.make_state_dict(0), synthetic.STATS_NAN_BINS, device=dev)
def mk(wl, nb):
    g = np.random.Generator(np.random.PCG64(3))
    b_, lo, hi, _ = bench.WORKLOADS[wl]
    items = [(f"u{i}", 0, g.integers(1, 361, int(n)), "") for i, n in enumerate(g.integers(lo, hi + 1, b_ * nb))]
    return P.make_batches(items, b_, sort_by_length=False)[0]
def stage(batch):
    _, _, speakers, texts, src_lens, max_src_len = batch
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).long().pin_memory()
    return pin(speakers), pin(texts), pin(src_lens), int(max_src_len)
def post_for(b):
    def post(out, info):
        return (tuple(out[1].shape), info, int(out[8].sum()), int(b[4].sum()), int(out[9].sum()), int(torch.isnan(out[4]).sum()), float(out[4].abs().max()))
    return post
for wl, nb in (("c2", 24), ("c3", 24)):
    batches = mk(wl, nb)
    sp, tx, sl, L = bench.make_batch(wl, 1)
    out, info = m.forward_with_info(sp.to(dev), tx.to(dev), sl.to(dev), L)
    torch.cuda.synchronize()
    for rep in range(2):
        for b in batches:
            d = P.to_device(b, dev)
            out = m(*d[2:])
            r = [out[1][i, :3].cpu() for i in range(0, len(out[1]), 16)]
    s = StreamedSynthesizer(m, 3)
    for rep in range(2):
        jobs = [s.submit(stage(b), post=post_for(b)) for b in batches]
        res = [s.wait(j) for j in jobs]
        bad = [(i, r) for i, r in enumerate(res) if r[0][1] == 0 or r[2] != r[3]]
        print(wl, rep, "T", [r[0][1] for r in res], "bad", bad, flush=True)
    s.close()
