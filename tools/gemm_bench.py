"""Micro-benchmark of gemm_img_kernel on the layer shapes of the path (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linetr_b200 import _native as N
lib = N.load()
shapes = [("qkv", 16384, 768, 256), ("merge", 16384, 256, 256), ("mlp1", 16384, 512, 512), ("mlp2", 16384, 256, 512),
          ("ffn_w1", 16384, 1024, 256), ("ffn_w2", 16384, 256, 1024), ("big", 131072, 256, 256)]
print(f"{'layer':8s} {'M':>7s} {'N':>5s} {'K':>5s} {'bn':>4s} {'out':>4s} {'us':>8s} {'TFLOP/s(useful)':>16s} {'L2 GB/s(operands)':>18s}")
for name, m, n, k in shapes:
    for bn in (128, 256):
        for om in (0, 1):
            ms = lib.ltr_gemm_bench(m, n, k, bn, om, 20, 0)
            tiles = (m // 128) * (n // bn)
            opb = tiles * (k // 64) * (32768 + bn * 256)
            print(f"{name:8s} {m:7d} {n:5d} {k:5d} {bn:4d} {om:4d} {ms*1e3:8.1f} {2*m*n*k/ms/1e9:16.1f} {opb/ms/1e6:18.0f}")
