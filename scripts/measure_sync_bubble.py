#!/usr/bin/env python
"""How much does the forward's one host synchronisation (the 8-byte read-back of T between the two stages) cost?
Times the forward as shipped against the same forward with T supplied in advance (async stage 1, no read-back).
Measured on B200 (r1r): C1 29 us (3 %), C2 29 us (1.7 %), C3 below the noise -- not worth speculating on T."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from smart_nar_fast_tts_b200 import synthetic, load_library
dev = torch.device("cuda", 0)
m = synthetic.build_module(synthetic.make_state_dict(0), synthetic.STATS_NAN_BINS, device=dev)
for wl in ("c1", "c2", "c3"):
    sp, tx, sl, L = bench.make_batch(wl, 1)
    sp, tx, sl = sp.to(dev), tx.to(dev), sl.to(dev)
    out = m(sp, tx, sl, L); torch.cuda.synchronize()
    T = out[1].shape[1]; frames = int(out[9].sum())
    def timeit(fn, n=30):
        for _ in range(5): fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n): fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / n
    t_sync = timeit(lambda: m(sp, tx, sl, L))
    # same forward, T known in advance: the device hook "knows" T, and we patch tolist() away by monkeypatching the hook path
    class Known:
        pass
    def hook(tm): pass
    m.t_max_device_hook = hook
    import smart_nar_fast_tts_b200.model as M
    orig = torch.Tensor.tolist
    def fake_tolist(self):
        if self.numel() == 2 and self.dtype == torch.int32: return [T, frames]
        return orig(self)
    torch.Tensor.tolist = fake_tolist
    t_nosync = timeit(lambda: m(sp, tx, sl, L))
    torch.Tensor.tolist = orig
    m.t_max_device_hook = None
    print(wl, "with sync %.3f ms   T known in advance (no host sync) %.3f ms   bubble %.1f us (%.1f %%)" % (t_sync, t_nosync, (t_sync - t_nosync) * 1e3, 100 * (t_sync - t_nosync) / t_sync))
