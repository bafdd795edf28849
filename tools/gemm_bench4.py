import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linetr_b200 import _native as N
lib = N.load()
for k in (64, 256, 1024):
    for bn in (128, 256):
        for om in (3, 0, 1):
            for t in (1, 4):
                lib.ltr_gemm_bench(148 * 128 * t, bn, k, bn, om, 1, 0)
