import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from smart_nar_fast_tts_b200 import pipeline as P, synthetic, StreamedSynthesizer
dev = torch.device("cuda", 0)
m = synthetic.build_module(synthetic.make_state_dict(0), synthetic.STATS_NAN_BINS, device=dev)
g = np.random.Generator(np.random.PCG64(3))
b_, lo, hi, _ = bench.WORKLOADS["c3"]
items = [(f"u{i}", 0, g.integers(1, 361, int(n)), "") for i, n in enumerate(g.integers(lo, hi + 1, b_ * 6))]
batches, _ = P.make_batches(items, b_, sort_by_length=False)
print("L", [int(b[5]) for b in batches])
for b in batches[:3]:
    d = P.to_device(b, dev)
    out, info = m.forward_with_info(*d[2:])
    torch.cuda.synchronize()
    print("sequential", out[1].shape, info, int(out[9].sum()), flush=True)
def stage(batch):
    _, _, speakers, texts, src_lens, max_src_len = batch
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).long().pin_memory()
    return pin(speakers), pin(texts), pin(src_lens), int(max_src_len)
s = StreamedSynthesizer(m, 3)
res = s.run([stage(b) for b in batches])
print("streamed", [tuple(r[1].shape) for r in res], flush=True)
res = s.run([stage(b) for b in batches])
print("streamed again", [tuple(r[1].shape) for r in res], flush=True)
jobs = [s.submit(stage(b), post=lambda out, info: (tuple(out[1].shape), info)) for b in batches]
print("post", [s.wait(j) for j in jobs], flush=True)
s.close()
s = StreamedSynthesizer(m, 3)
jobs = [s.submit(stage(b), post=lambda out, info: (tuple(out[1].shape), info)) for b in batches]
print("fresh synthesizer post", [s.wait(j) for j in jobs], flush=True)
s.close()
