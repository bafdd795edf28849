"""Experiment: does running the batch as G independent groups of pairs on G CUDA streams fill the
wave-quantisation gaps of the persistent kernels?  (64 pairs = 128 m-tiles per GEMM on 148 SMs.)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from linetr_b200 import LineBatch, LineTransformer, PairEngine, synthetic as syn

dev = torch.device("cuda", 0)
sd = syn.make_state_dict(0, 1)
m = LineTransformer({"mode": "train"}); m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}); m = m.eval().to(dev)
eng = PairEngine(m, dev)
P, L, T = 64, 128, 21
pairs = [syn.make_pair_inputs(i, L, T) for i in range(P)]

def packed(idx):
    return LineBatch.from_images([pairs[i][0] for i in idx] + [pairs[i][1] for i in idx]).to(dev)

def run(G, iters=20):
    groups = [packed(list(range(g * P // G, (g + 1) * P // G))) for g in range(G)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(G)]
    def step():
        cur = torch.cuda.current_stream(dev)
        ev = torch.cuda.Event(); ev.record(cur)
        outs = []
        for g in range(G):
            streams[g].wait_event(ev)
            with torch.cuda.stream(streams[g]):
                outs.append(eng.match_packed(groups[g], P // G, 0.8).counts)
                e = torch.cuda.Event(); e.record(streams[g])
            cur.wait_event(e)
        return outs
    for _ in range(3): o = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): o = step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return ms, int(sum(int(x.sum()) for x in o))

base = None
for G in (1, 2, 4, 1, 2):
    ms, tot = run(G)
    print(f"groups={G}: {ms:.3f} ms/step, {P / ms * 1e3:.0f} pairs/s, matches {tot}", flush=True)
