"""CPU checks of the numeric formulas the CUDA kernels rely on (no GPU needed).

* the split-bf16 operand representation (ptx_sm100.cuh: split_bf16 / split2_bf16): hi + lo carries
  ~2^-17 relative precision, and the 3-term product a_lo*w_hi + a_hi*w_lo + a_hi*w_hi drops only lo*lo;
* the erfc-based GELU of the GEMM epilogue (gemm_img.cuh: gelu_erf) with the constants parsed from
  the source, against the exact erf GELU the reference computes (F.gelu default,
  models/line_attention.py:92).
"""
import math
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def bf16_round(x):
    """Round-to-nearest-even fp32 -> bf16 -> fp32, bit-exact emulation of cvt.rn.bf16.f32."""
    u = np.asarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32)


def split(x):
    x = np.asarray(x, dtype=np.float32)
    hi = bf16_round(x)
    lo = bf16_round((x - hi).astype(np.float32))
    return hi, lo


def test_split_bf16_precision():
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(200000) * np.exp(rng.uniform(-8, 8, 200000))).astype(np.float32)
    hi, lo = split(x)
    rel = np.abs((hi.astype(np.float64) + lo.astype(np.float64)) - x.astype(np.float64)) / np.abs(x.astype(np.float64))
    assert rel.max() < 2.0 ** -16          # 8 + 8 mantissa bits (+ hidden bits): ~2^-17 typical
    assert np.median(rel) < 2.0 ** -18


def test_three_term_product_error():
    rng = np.random.default_rng(1)
    K = 512
    a = rng.standard_normal((64, K)).astype(np.float32)
    w = (rng.standard_normal((48, K)) / math.sqrt(K)).astype(np.float32)
    ah, al = split(a)
    wh, wl = split(w)
    f = lambda m: m.astype(np.float64)
    got = f(al) @ f(wh).T + f(ah) @ f(wl).T + f(ah) @ f(wh).T
    want = f(a) @ f(w).T
    single = f(ah) @ f(wh).T
    assert np.abs(got - want).max() < 2e-5                 # what the tensor-core path computes (before fp32 accumulation)
    assert np.abs(single - want).max() > 50 * np.abs(got - want).max()   # one bf16 pass is far outside the budget


def _gelu_constants():
    src = open(os.path.join(ROOT, "linetr_b200", "csrc", "gemm_img.cuh")).read()
    body = src[src.index("float gelu_erf(float x)"):]
    body = body[:body.index("\n}\n")]
    nums = [float(v) for v in re.findall(r"(-?\d+\.\d+)f", body)]
    return body, nums


def test_gelu_erfc_formula_matches_exact_erf_gelu():
    body, nums = _gelu_constants()
    # Abramowitz-Stegun 7.1.26 constants must be the ones in the kernel
    for c in (0.3275911, 1.061405429, -1.453152027, 1.421413741, -0.284496736, 0.254829592):
        assert any(abs(c - n) < 1e-9 for n in nums), c
    x = np.linspace(-12, 12, 400001).astype(np.float32)
    f32 = np.float32
    u = np.abs(x) * f32(0.70710678118654752440)
    t = (f32(1) / (f32(0.3275911) * u + f32(1))).astype(np.float32)
    pl = f32(1.061405429) * t + f32(-1.453152027)
    pl = pl * t + f32(1.421413741)
    pl = pl * t + f32(-0.284496736)
    pl = pl * t + f32(0.254829592)
    pl = (pl * t).astype(np.float32)
    g = (f32(0.5) * x * pl * np.exp2((u * u * f32(-1.4426950408889634)).astype(np.float32))).astype(np.float32)
    got = np.where(x < 0, g, x - g).astype(np.float64)
    erf = np.vectorize(math.erf)
    want = 0.5 * x.astype(np.float64) * (1.0 + erf(x.astype(np.float64) / math.sqrt(2.0)))
    assert np.abs(got - want).max() < 5e-7
