"""Seeded synthetic weights and inputs for the LineTR hot path (numpy only).

Everything here is regenerable from an integer seed on any machine, so the GPU box
(which has no copy of the reference checkout or its checkpoint) can rebuild exactly
the weights and inputs the committed golden fixtures were produced with.

Shapes/keys follow the reference checkpoint contract (SURVEY.md §8a "State-dict
contract"; reference models/line_transformer.py:93-105,139-162,203-217 and
models/line_attention.py:23-40,77-84).  Input distributions follow SURVEY.md §8(d).
"""
from __future__ import annotations

import numpy as np

D_MODEL = 256
N_HEADS = 4
D_INNER = 1024
MLP_HIDDEN = (32, 64, 128, 256)
N_SIG_LAYERS = 7


def state_dict_spec(n_desc_layers: int = 1, d_model: int = D_MODEL, d_inner: int = D_INNER):
    """[(key, shape, kind)] in the reference checkpoint's key order.

    kind in {'w','b','bn_w','bn_b','bn_mean','bn_var','bn_count','ln_w','ln_b','cls'}.
    """
    spec = [("klenc.cls_token", (1, 1, 1, d_model), "cls")]

    def mlp(prefix, chans):
        idx = 0
        for i in range(1, len(chans)):
            spec.append((f"{prefix}.{idx}.weight", (chans[i], chans[i - 1], 1), "w"))
            spec.append((f"{prefix}.{idx}.bias", (chans[i],), "b"))
            idx += 1
            if i < len(chans) - 1:
                c = chans[i]
                spec.append((f"{prefix}.{idx}.weight", (c,), "bn_w"))
                spec.append((f"{prefix}.{idx}.bias", (c,), "bn_b"))
                spec.append((f"{prefix}.{idx}.running_mean", (c,), "bn_mean"))
                spec.append((f"{prefix}.{idx}.running_var", (c,), "bn_var"))
                spec.append((f"{prefix}.{idx}.num_batches_tracked", (), "bn_count"))
                idx += 2  # BatchNorm1d + ReLU slots

    mlp("klenc.line_position_enc.encoder", [5, *MLP_HIDDEN, d_model])
    mlp("klenc.word_position_enc.encoder", [3, *MLP_HIDDEN, d_model])
    for i in range(n_desc_layers):
        p = f"klenc.desc_layers.{i}"
        for nm in ("w_qs", "w_ks", "w_vs", "fc"):
            spec.append((f"{p}.slf_attn.{nm}.weight", (d_model, d_model), "w"))
            spec.append((f"{p}.slf_attn.{nm}.bias", (d_model,), "b"))
        spec.append((f"{p}.slf_attn.layer_norm.weight", (d_model,), "ln_w"))
        spec.append((f"{p}.slf_attn.layer_norm.bias", (d_model,), "ln_b"))
        spec.append((f"{p}.pos_ffn.w_1.weight", (d_inner, d_model), "w"))
        spec.append((f"{p}.pos_ffn.w_1.bias", (d_inner,), "b"))
        spec.append((f"{p}.pos_ffn.w_2.weight", (d_model, d_inner), "w"))
        spec.append((f"{p}.pos_ffn.w_2.bias", (d_model,), "b"))
        spec.append((f"{p}.pos_ffn.layer_norm.weight", (d_model,), "ln_w"))
        spec.append((f"{p}.pos_ffn.layer_norm.bias", (d_model,), "ln_b"))
    for i in range(N_SIG_LAYERS):
        p = f"selfattn.layers.{i}"
        for nm in ("merge", "proj.0", "proj.1", "proj.2"):
            spec.append((f"{p}.attn.{nm}.weight", (d_model, d_model, 1), "w"))
            spec.append((f"{p}.attn.{nm}.bias", (d_model,), "b"))
        mlp(f"{p}.mlp", [2 * d_model, 2 * d_model, d_model])
    spec.append(("final_proj.weight", (d_model, d_model, 1), "w"))
    spec.append(("final_proj.bias", (d_model,), "b"))
    return spec


def make_state_dict(seed: int = 0, n_desc_layers: int = 1, d_inner: int = D_INNER) -> dict:
    """Random-init checkpoint of the LineTR architecture, as {key: np.ndarray}.

    Magnitudes are chosen to resemble the shipped checkpoint (weights ~ 1/sqrt(fan_in),
    non-trivial BatchNorm running stats so that the BN fold is actually exercised).
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    sd = {}
    for key, shape, kind in state_dict_spec(n_desc_layers, D_MODEL, d_inner):
        if kind == "w":
            fan_in = shape[1]
            v = rng.standard_normal(shape) * (1.0 / np.sqrt(fan_in))
        elif kind == "b":
            v = rng.standard_normal(shape) * 0.05
        elif kind in ("bn_w", "ln_w"):
            v = 1.0 + 0.1 * rng.standard_normal(shape)
        elif kind in ("bn_b", "ln_b"):
            v = 0.1 * rng.standard_normal(shape)
        elif kind == "bn_mean":
            v = 0.3 * rng.standard_normal(shape)
        elif kind == "bn_var":
            v = rng.uniform(0.5, 1.5, shape)
        elif kind == "bn_count":
            sd[key] = np.array(0, dtype=np.int64)
            continue
        elif kind == "cls":
            v = rng.standard_normal(shape)
        else:  # pragma: no cover
            raise ValueError(kind)
        sd[key] = np.ascontiguousarray(v, dtype=np.float32)
    return sd


def _unit(v, axis=-1):
    n = np.sqrt((v.astype(np.float64) ** 2).sum(axis=axis, keepdims=True))
    return (v / np.maximum(n, 1e-12)).astype(np.float32)


def make_image_inputs(seed: int, n_lines: int, n_tokens: int, n_real_tokens=None,
                      width: int = 640, height: int = 480) -> dict:
    """One image's tokenised lines in the layout the tokenizer emits
    (reference models/line_process.py:182-193), batch dim 1, float32 numpy.

    Padded token slots carry the same kind of data as real ones: the reference
    attends to them, so parity must not depend on the mask.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    L, T = n_lines, n_tokens
    wh = np.array([width, height], dtype=np.float32)
    sub = (rng.random((L, 2, 2), dtype=np.float32) * wh).astype(np.float32)
    pnt = (rng.random((L, T, 2), dtype=np.float32) * wh).astype(np.float32)
    desc = _unit(rng.standard_normal((L, T, D_MODEL)).astype(np.float32))
    score = (rng.random((L, T, 1), dtype=np.float32) * 0.1).astype(np.float32)
    resp = rng.random((L, 1), dtype=np.float32)
    theta = rng.random(L, dtype=np.float32) * np.float32(np.pi)
    angle = np.stack([np.cos(2 * theta), np.sin(2 * theta)], axis=1).astype(np.float32)
    if n_real_tokens is None:
        ntok = np.full(L, T, dtype=np.int64)
    elif np.isscalar(n_real_tokens):
        ntok = np.full(L, int(n_real_tokens), dtype=np.int64)
    else:
        lo, hi = n_real_tokens
        ntok = rng.integers(lo, hi + 1, size=L)
    mask = np.zeros((L, T + 1, 1), dtype=np.float32)
    mask[:, 0] = 1
    for i in range(L):
        mask[i, 1:1 + int(min(ntok[i], T))] = 1
    return {
        "klines": sub[None].copy(),
        "sublines": sub[None],
        "pnt_sublines": pnt[None],
        "desc_sublines": desc[None],
        "score_sublines": score[None],
        "resp_sublines": resp[None],
        "angle_sublines": angle[None],
        "mask_sublines": mask[None],
        "mat_klines2sublines": np.eye(L, dtype=np.float32)[None],
    }


def make_pair_inputs(seed: int, n_lines: int, n_tokens: int, n_lines1=None,
                     n_real_tokens=None, jitter_px: float = 2.0, desc_noise: float = 0.05):
    """An image pair whose second side is a row-permuted, jittered copy of the first
    (SURVEY.md §8(d)), so that mutual-NN matches exist with top-2 gaps >> 1e-3.

    Returns (side0, side1, perm) with side1 line j == side0 line perm[j] (+noise).
    If n_lines1 < n_lines only the first n_lines1 permuted lines are kept.
    """
    a = make_image_inputs(2 * seed, n_lines, n_tokens, n_real_tokens)
    rng = np.random.Generator(np.random.PCG64(2 * seed + 1))
    L1 = n_lines if n_lines1 is None else int(n_lines1)
    perm = rng.permutation(n_lines)[:L1]
    b = {}
    for k in ("sublines", "pnt_sublines"):
        v = a[k][0][perm]
        b[k] = (v + rng.standard_normal(v.shape).astype(np.float32) * np.float32(jitter_px))[None]
    d = a["desc_sublines"][0][perm]
    d = _unit(d + np.float32(desc_noise) * rng.standard_normal(d.shape).astype(np.float32))
    b["desc_sublines"] = d[None]
    for k in ("score_sublines", "resp_sublines", "angle_sublines", "mask_sublines"):
        b[k] = a[k][0][perm][None].copy()
    b["klines"] = b["sublines"].copy()
    b["mat_klines2sublines"] = np.eye(L1, dtype=np.float32)[None]
    return a, b, perm


def make_descriptor_pair(seed: int, n0: int, n1: int = None, d: int = D_MODEL, noise: float = 0.05):
    """Unit-norm descriptor sets [d,n0], [d,n1] for the matcher-only workload (cfg[4])."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n1 = n0 if n1 is None else n1
    d0 = _unit(rng.standard_normal((n0, d)).astype(np.float32))
    perm = rng.permutation(max(n0, n1))
    perm = perm[perm < n0][:n1] if n1 <= n0 else np.concatenate([rng.permutation(n0), rng.integers(0, n0, n1 - n0)])
    d1 = _unit(d0[perm] + np.float32(noise) * rng.standard_normal((n1, d)).astype(np.float32))
    return np.ascontiguousarray(d0.T), np.ascontiguousarray(d1.T), perm
