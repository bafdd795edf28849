"""Shared test helpers: golden fixtures, regenerated inputs, drift checks."""
import json
import os

import numpy as np

from linetr_b200 import synthetic as syn

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_cache = {}


def golden():
    if "npz" not in _cache:
        _cache["npz"] = dict(np.load(os.path.join(GOLDEN_DIR, "reference_outputs.npz")))
        with open(os.path.join(GOLDEN_DIR, "reference_outputs.json")) as f:
            _cache["meta"] = json.load(f)
    return _cache["npz"], _cache["meta"]


def checksum(d):
    return {k: float(np.asarray(v, dtype=np.float64).sum()) for k, v in sorted(d.items())}


def assert_checksum(d, want):
    got = checksum(d)
    assert got.keys() == want.keys()
    for k in got:
        assert abs(got[k] - want[k]) <= 1e-6 * max(1.0, abs(want[k])), f"input generator drift in {k}"


def weights_for(tag):
    kind, *rest = tag.split(":")
    if kind == "synthetic":
        return syn.make_state_dict(int(rest[0]), int(rest[1]))
    raise KeyError(tag)


def shipped_weights_path():
    """The reference checkpoint, if a copy travelled with the repo (git-ignored) or the
    reference checkout is mounted (build container only)."""
    root = os.path.dirname(GOLDEN_DIR[:-len("/golden")])
    for p in (os.environ.get("LINETR_WEIGHTS", ""),
              os.path.join(root, "linetr_b200", "weights", "LineTR_weight.pth"),
              "/root/reference/models/weights/LineTR_weight.pth"):
        if p and os.path.exists(p):
            return p
    return None


def load_shipped_weights():
    p = shipped_weights_path()
    if p is None:
        return None
    import torch
    return {k: v.numpy() for k, v in torch.load(p, map_location="cpu").items()}


def case_inputs(case):
    ntok = case.get("ntok")
    ntok = tuple(ntok) if isinstance(ntok, list) else ntok
    return syn.make_image_inputs(case["seed"], case["L"], case["T"], ntok)


def stack(images):
    return {k: np.concatenate([im[k] for im in images], axis=0) for k in images[0]}


ENC_CASES = ["enc_L16_T21", "enc_L1_T21", "enc_L37_T5_ragged", "enc_L24_T32_ragged", "enc_L130_T21",
             "enc_L9_T21_nd2"]
