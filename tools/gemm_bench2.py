import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linetr_b200 import _native as N
lib = N.load()
print("tiles/CTA sweep (N=BN so n_blks=1): us per launch")
for k in (64, 128, 256, 512, 1024):
    for bn in (128, 256):
        for om in (0, 1):
            row = []
            for t in (1, 2, 3, 4, 8):
                ms = lib.ltr_gemm_bench(148 * 128 * t, bn, k, bn, om, 20, 0)
                row.append(ms * 1e3)
            slope = (row[4] - row[0]) / 7
            print(f"K={k:5d} BN={bn} out={om}: " + " ".join(f"{x:7.1f}" for x in row) + f"   per-tile {slope:5.2f} us  intercept {row[0]-slope:5.2f} us  mma-only {k/64*12*bn/256*128/1.9e3:5.2f} us")
