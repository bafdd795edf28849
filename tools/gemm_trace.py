"""Per-tile clock64 stamps of gemm_img_kernel CTA 0 on the signature-layer shapes (image output)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linetr_b200 import _native as N
lib = N.load()
names = ["mma:acc_empty", "mma:full0", "mma:issued", "epi:acc_full", "epi:ld0", "epi:done", "tma:slot_free"]
for name, m, n, k, bn, om in (("qkv", 16384, 768, 256, 256, 1), ("mlp1", 16384, 512, 512, 256, 1), ("mlp2", 16384, 256, 512, 128, 1),
                              ("mlp2", 16384, 256, 512, 256, 1), ("ffn_w1", 16384, 1024, 256, 256, 1)):
    ms = lib.ltr_gemm_bench(m, n, k, bn, om, 20, 0)
    t = lib.ltr_gemm_trace()
    vals = [t[i] for i in range(64)]
    base = min(v for v in vals if v)
    print(f"{name} M={m} N={n} K={k} BN={bn} out={om}: {ms*1e3:.1f} us/launch")
    for tl in range(3):
        print(f"  tile {tl}: " + "  ".join(f"{names[j]}={vals[tl*16+j]-base if vals[tl*16+j] else -1:7d}" for j in range(7)))
