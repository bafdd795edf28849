import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linetr_b200 import _native as N
lib = N.load()
names = ["mma:acc_empty", "mma:full0", "mma:issued", "epi:acc_full", "epi:ld0", "epi:done", "tma:slot_free"]
for k, bn, om in ((64, 256, 3), (64, 256, 0), (256, 256, 1), (256, 128, 3)):
    lib.ltr_gemm_bench(148 * 128 * 4, bn, k, bn, om, 1, 0)
    t = lib.ltr_gemm_trace()
    vals = [t[i] for i in range(64)]
    base = min(v for v in vals if v)
    print(f"K={k} BN={bn} out={om} (cycles since first stamp)")
    for tl in range(4):
        print(f"  tile {tl}: " + "  ".join(f"{names[j]}={vals[tl*16+j]-base if vals[tl*16+j] else -1:7d}" for j in range(7)))
