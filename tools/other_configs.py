"""Throughput of the other BASELINE.json configs on one GPU (cfg[2] shape 256 lines x 32 tokens,
cfg[3] ragged 32..512 lines x 64 tokens, cfg[4] matcher-only 1024 x 1024).  Device-resident inputs,
CUDA events, pairs/s.  Parity for these shapes is covered by tests/test_gpu_parity.py."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from linetr_b200 import LineBatch, LineTransformer, PairEngine, _native as N, _ops, synthetic as syn

dev = torch.device("cuda", 0)
sd = syn.make_state_dict(0, 1)
m = LineTransformer({"mode": "train", "max_tokens": 64}); m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}); m = m.eval().to(dev)
eng = PairEngine(m, dev)

def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

out = {}
# cfg2: 64 pairs/GPU x 256 lines x 32 tokens
P = 64
pairs = [syn.make_pair_inputs(2000 + i, 256, 32)[:2] for i in range(P)]
b = LineBatch.from_images([a for a, _ in pairs] + [c for _, c in pairs]).to(dev)
ms = timeit(lambda: eng.match_packed(b, P, 0.8))
out["cfg2_64pairs_256x32"] = {"ms_per_step": ms, "pairs_per_s": P / ms * 1e3}
del b, pairs
# cfg3: ragged 32..512 lines, 64 tokens (ragged mask), 32 pairs/GPU
P = 32
rng = np.random.Generator(np.random.PCG64(3))
pairs = [syn.make_pair_inputs(3000 + i, int(rng.integers(32, 513)), 64, n_real_tokens=(5, 64))[:2] for i in range(P)]
b = LineBatch.from_images([a for a, _ in pairs] + [c for _, c in pairs]).to(dev)
ms = timeit(lambda: eng.match_packed(b, P, 0.8), 5)
out["cfg3_32pairs_ragged32-512x64"] = {"ms_per_step": ms, "pairs_per_s": P / ms * 1e3, "total_lines": int(b.n_lines)}
del b, pairs
# cfg4: matcher only 1024 x 1024 x d256, 64 pairs per call
P = 64
d0 = torch.nn.functional.normalize(torch.randn(P * 1024, 256, device=dev), dim=1)
perm = torch.randperm(1024, device=dev)
d1 = torch.nn.functional.normalize(d0.view(P, 1024, 256)[:, perm] + 0.05 * torch.randn(P, 1024, 256, device=dev), dim=2).reshape(P * 1024, 256)
ms = timeit(lambda: _ops.match_descriptors(d0, d1, N.LAYOUT_ROWS, P, 0.8, True, n0=1024, n1=1024))
out["cfg4_matcher_64pairs_1024x1024"] = {"ms_per_step": ms, "pairs_per_s": P / ms * 1e3}
print(json.dumps(out))
