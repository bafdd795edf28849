import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from linetr_b200 import LineBatch, LineTransformer, PairEngine, _native as N, synthetic as syn
lib = N.load()
dev = torch.device("cuda", 0)
sd = syn.make_state_dict(0, 1)
m = LineTransformer({"mode": "train"}); m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}); m = m.eval().to(dev)
eng = PairEngine(m, dev)
ims = [syn.make_image_inputs(i, 128, 21) for i in range(64)]
b = LineBatch.from_images(ims).to(dev)
eng.encode(b); torch.cuda.synchronize()
lib.ltr_debug_trace_arm(1)
eng.encode(b); torch.cuda.synchronize()
buf = (C.c_uint64 * 128)()
lib.ltr_debug_trace_read(buf)
lib.ltr_debug_trace_arm(0)
v = [buf[i] for i in range(12)]
a = [buf[i] for i in range(30, 39)]
names = {1: "ep2 done", 2: "L3 acc", 3: "ep3 done", 4: "L4 acc (nb0)", 5: "ep4 done", 6: "L5 acc (nb0)", 7: "ep5 done", 11: "tile end"}
prev = v[0]
for i in (1, 2, 3, 4, 5, 6, 7, 11):
    print(f"{names[i]:14s} +{v[i]-prev:7d} cycles")
    prev = v[i]
print("tile total", v[11] - v[0], " (softmax + pooling of a tile run inside the next tile's MMA phases)")
print("L5 second n-block complete", buf[12] - v[6], "cycles after the first")

an = ["start", "tmem alloc+sync", "kt start", "stage q,k", "stage v^T", "S mma done", "softmax+P stored", "PV mma done", "epilogue done"]
print("attention CTA (0,0,0), last layer:")
for i in range(1, 9):
    print(f"{an[i]:18s} +{a[i]-a[i-1]:7d} cycles")
print("attention CTA total", a[8] - a[0])
