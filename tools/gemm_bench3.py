import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linetr_b200 import _native as N
lib = N.load()
for k in (64, 256):
    for bn in (128, 256):
        for om in (3, 0, 1):
            row = [lib.ltr_gemm_bench(148 * 128 * t, bn, k, bn, om, 20, 0) * 1e3 for t in (1, 2, 3, 4)]
            print(f"K={k:4d} BN={bn} out={om}: " + " ".join(f"{x:7.1f}" for x in row) + f"   per-tile {(row[3]-row[0])/3:5.2f} us")
