"""CPU oracle for the LineTR line-descriptor forward + NN matcher hot path.

TEST INFRASTRUCTURE ONLY.  This file is a plain-numpy (float32) restatement of the
reference algorithm and exists so that the CUDA path can be checked against it.  Only
`tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import it; nothing under `linetr_b200/` does (the product path fails
loudly when its CUDA library is missing - there is no CPU fallback).

Parity pin: the reference has no tests or golden vectors of its own,
so this oracle is pinned against the reference *as executed in the build container*:
`tests/golden/make_golden.py` imports /root/reference/models/* read-only, runs it on
seeded inputs and commits the outputs as fixtures; `tests/test_oracle.py` checks every
function below against those fixtures (max-abs 2e-5 on descriptors, exact on matches).

Each function cites the reference lines it restates (paths relative to the reference
checkout).  The restatement is *literal*: all T+1 query rows are computed, the dead
row-mask is applied exactly as the reference applies it, descriptive layers are not
chained, heads of the signature network are interleaved (c = d*4 + h).
"""
from __future__ import annotations

import numpy as np
from scipy.special import erf as _erf

F32 = np.float32
BN_EPS = F32(1e-5)      # torch.nn.BatchNorm1d default (line_transformer.py:17)
LN_EPS = F32(1e-6)      # line_attention.py:40,83
N_HEADS_SIG = 4         # line_transformer.py:172


def _f32(x):
    return np.ascontiguousarray(x, dtype=F32)


# ----------------------------------------------------------------------------- pieces
def mlp_1x1(sd, prefix, x):
    """MLP() of Conv1d(k=1) [+BatchNorm1d(eval)+ReLU] blocks, line_transformer.py:9-20.

    x: [R, C_in] rows (the reference runs channels-first [B,C,N]; a k=1 conv is the same
    matmul per position).  Layer indices step by 3 (conv, bn, relu) except the last.
    """
    idx = 0
    while f"{prefix}.{idx}.weight" in sd:
        w = sd[f"{prefix}.{idx}.weight"][:, :, 0]
        x = x @ w.T + sd[f"{prefix}.{idx}.bias"]
        if f"{prefix}.{idx + 1}.running_mean" in sd:
            p = f"{prefix}.{idx + 1}"
            inv = F32(1.0) / np.sqrt(sd[p + ".running_var"] + BN_EPS)
            x = (x - sd[p + ".running_mean"]) * inv * sd[p + ".weight"] + sd[p + ".bias"]
            x = np.maximum(x, F32(0))
            idx += 3
        else:
            idx += 1
    return _f32(x)


def normalize_keylines(klines, kplines, image_shape):
    """line_transformer.py:22-38: (xy - size/2) / (0.7 * max(W, H))."""
    if len(image_shape) == 2:
        height, width = image_shape
    else:
        _, _, height, width = image_shape
    size = np.array([width, height], dtype=F32)
    center = size / F32(2)
    scaling = F32(max(width, height)) * F32(0.7)
    return _f32((klines - center) / scaling), _f32((kplines - center) / scaling)


def layer_norm(x, w, b):
    mu = x.mean(axis=-1, keepdims=True, dtype=F32)
    var = ((x - mu) ** 2).mean(axis=-1, keepdims=True, dtype=F32)
    return _f32((x - mu) / np.sqrt(var + LN_EPS) * w + b)


def softmax_last(x):
    m = x.max(axis=-1, keepdims=True)
    e = np.exp(x - m)
    return _f32(e / e.sum(axis=-1, keepdims=True, dtype=F32))


def gelu_erf(x):
    """F.gelu default (exact erf), line_attention.py:89."""
    return _f32(x * F32(0.5) * (F32(1) + _erf(x * F32(0.7071067811865476))))


def multi_head_attention(sd, p, x, mask):
    """MultiHeadAttention + ScaledDotProduct, line_attention.py:42-75 and :13-21.

    x: [B,L,N,256]; mask: [B,L,N,1] -> unsqueeze(2) -> broadcast against attn
    [B,L,H,Nq,Nk]: it masks QUERY rows.  Head layout c = h*64+d.
    """
    B, L, N, D = x.shape
    H = 4
    dk = D // H
    q = (x @ sd[p + ".w_qs.weight"].T + sd[p + ".w_qs.bias"]).reshape(B, L, N, H, dk).transpose(0, 1, 3, 2, 4)
    k = (x @ sd[p + ".w_ks.weight"].T + sd[p + ".w_ks.bias"]).reshape(B, L, N, H, dk).transpose(0, 1, 3, 2, 4)
    v = (x @ sd[p + ".w_vs.weight"].T + sd[p + ".w_vs.bias"]).reshape(B, L, N, H, dk).transpose(0, 1, 3, 2, 4)
    attn = (q / F32(dk ** 0.5)) @ k.transpose(0, 1, 2, 4, 3)            # [B,L,H,N,N]
    if mask is not None:
        m = mask[:, :, None, :, :]                                       # [B,L,1,N,1]
        attn = np.where(np.broadcast_to(m, attn.shape) == 0, F32(-1e9), attn)
    attn = softmax_last(_f32(attn))
    out = (attn @ v).transpose(0, 1, 3, 2, 4).reshape(B, L, N, D)
    out = out @ sd[p + ".fc.weight"].T + sd[p + ".fc.bias"]
    out = out + x
    return layer_norm(_f32(out), sd[p + ".layer_norm.weight"], sd[p + ".layer_norm.bias"])


def feed_forward(sd, p, x):
    """FeedForward, line_attention.py:86-94."""
    h = gelu_erf(_f32(x @ sd[p + ".w_1.weight"].T + sd[p + ".w_1.bias"]))
    y = h @ sd[p + ".w_2.weight"].T + sd[p + ".w_2.bias"] + x
    return layer_norm(_f32(y), sd[p + ".layer_norm.weight"], sd[p + ".layer_norm.bias"])


def keyline_encoder(sd, klines, resp, angle, pnt, desc, score, mask):
    """KeylineEncoder.forward, line_transformer.py:107-130.  Returns sentence [B,L,256]
    (the reference keeps it channels-first [B,256,L])."""
    B, L, T, D = desc.shape
    # LinePositionalEncoder, :46-50
    mid = (klines[:, :, 0] + klines[:, :, 1]) / F32(2)
    lp_in = np.concatenate([mid, resp, angle], axis=-1).reshape(B * L, 5)
    klines_pos = mlp_1x1(sd, "klenc.line_position_enc.encoder", _f32(lp_in)).reshape(B, L, D)
    # WordPositionalEncoder, :61-73
    wp_in = np.concatenate([pnt, score], axis=-1).reshape(B * L * T, 3)
    wpe = mlp_1x1(sd, "klenc.word_position_enc.encoder", _f32(wp_in)).reshape(B, L, T, D)
    x = _f32(desc + wpe)                                                  # :117
    cls = np.broadcast_to(sd["klenc.cls_token"].reshape(1, 1, 1, D), (B, L, 1, D))
    x = _f32(np.concatenate([cls, x], axis=2))                            # :120-121
    n_layers = 0
    while f"klenc.desc_layers.{n_layers}.slf_attn.w_qs.weight" in sd:
        n_layers += 1
    enc = None
    for i in range(n_layers):                                             # :123-125 (NOT chained)
        p = f"klenc.desc_layers.{i}"
        enc = feed_forward(sd, p + ".pos_ffn", multi_head_attention(sd, p + ".slf_attn", x, mask))
    return _f32(klines_pos + enc[:, :, 0, :])                             # :128


def signature_layer(sd, p, x):
    """AttentionalPropagation + MultiHeadedAttention + attention(),
    line_transformer.py:132-136,149-154,164-166.  x: [B,L,256] rows; heads interleaved."""
    B, L, D = x.shape
    H = N_HEADS_SIG
    dim = D // H

    def proj(i):
        y = x @ sd[f"{p}.attn.proj.{i}.weight"][:, :, 0].T + sd[f"{p}.attn.proj.{i}.bias"]
        return _f32(y).reshape(B, L, dim, H)                              # channel c = d*H + h

    q, k, v = proj(0), proj(1), proj(2)
    scores = np.einsum("bndh,bmdh->bhnm", q, k) / F32(dim ** 0.5)
    prob = softmax_last(_f32(scores))
    o = np.einsum("bhnm,bmdh->bndh", prob, v).reshape(B, L, D)
    msg = _f32(o) @ sd[f"{p}.attn.merge.weight"][:, :, 0].T + sd[f"{p}.attn.merge.bias"]
    cat = np.concatenate([x, _f32(msg)], axis=-1).reshape(B * L, 2 * D)
    return mlp_1x1(sd, f"{p}.mlp", _f32(cat)).reshape(B, L, D)


def line_transformer_forward(sd, data, image_shape=(480, 640)):
    """LineTransformer.forward, line_transformer.py:225-249 -> line_desc [B,256,L]."""
    klines, pnt = normalize_keylines(_f32(data["sublines"]), _f32(data["pnt_sublines"]), image_shape)
    x = keyline_encoder(sd, klines, _f32(data["resp_sublines"]), _f32(data["angle_sublines"]), pnt,
                        _f32(data["desc_sublines"]), _f32(data["score_sublines"]),
                        _f32(data["mask_sublines"]))
    i = 0
    while f"selfattn.layers.{i}.attn.merge.weight" in sd:                 # :176-183
        x = _f32(x + signature_layer(sd, f"selfattn.layers.{i}", x))
        i += 1
    y = x @ sd["final_proj.weight"][:, :, 0].T + sd["final_proj.bias"]    # :245
    n = np.sqrt((y * y).sum(axis=-1, keepdims=True, dtype=F32))
    y = y / np.maximum(n, F32(1e-12))                                     # :246 F.normalize
    return _f32(y.transpose(0, 2, 1))


# ---------------------------------------------------------------------------- matcher
def get_dist_matrix(desc0, desc1):
    """line_process.py:198-201."""
    s = np.einsum("bdn,bdm->bnm", desc0, desc1)
    return (2.0 - 2.0 * s).clip(min=0)


def subline2keyline(dist_sublines, a0, a1):
    """line_transformer.py:277-282 (A0 @ D @ A1^T, leading batch dim added)."""
    return (np.asarray(a0) @ dist_sublines @ np.asarray(a1).T)[None]


def nn_matcher_distmat(dist_mat, nn_thresh, is_mutual_nn=True):
    """nn_matcher.py:3-31.  Only batch element 0 is matched (b = 1, :7)."""
    n0, n1 = dist_mat.shape[1], dist_mat.shape[2]
    out = np.zeros((1, n0, n1))
    if n0 == 0 or n1 == 0:
        return out
    d = dist_mat[0].clip(min=0)
    idx = np.argmin(d, axis=1)
    scores = d[np.arange(n0), idx]
    keep = scores < nn_thresh
    if is_mutual_nn:
        idx2 = np.argmin(d, axis=0)
        keep = np.logical_and(keep, np.arange(n0) == idx2[idx])
    out[0, np.arange(n0)[keep], idx[keep]] = 1
    return out


def nn_matcher(desc0, desc1, nn_thresh=0.8, is_mutual_nn=True):
    """nn_matcher.py:33-43."""
    dmat = desc0.T @ desc1
    dist = (2.0 - 2.0 * dmat).clip(min=0)[None]
    return nn_matcher_distmat(dist, nn_thresh, is_mutual_nn), dist


def match_indices(mat):
    """Dense 0/1 [1,n0,n1] -> int32 [n0] (index in side 1 or -1)."""
    m = mat[0]
    has = m.sum(axis=1) > 0
    return np.where(has, m.argmax(axis=1), -1).astype(np.int32)


def match_pair(sd, side0, side1, nn_thresh=0.8, image_shape=(480, 640)):
    """The line branch of Matching.forward, matching.py:77-81, from tokenised dicts."""
    d0 = line_transformer_forward(sd, side0, image_shape)
    d1 = line_transformer_forward(sd, side1, image_shape)
    dist = get_dist_matrix(d0, d1)[0]
    dk = subline2keyline(dist, side0["mat_klines2sublines"][0], side1["mat_klines2sublines"][0])
    return nn_matcher_distmat(dk, nn_thresh, True), dk, d0, d1


# ---------------------------------------------------------------------------- training-side matcher
def eval_dist(desc0, desc1):
    """evaluations/matcher.py:22-28,66-72: ||a||^2 + ||b||^2 - 2 ab, clipped at 0; desc [d,n]."""
    a, b = desc0.T, desc1.T
    sq0 = np.sum(np.square(a), axis=1)[:, None]
    sq1 = np.sum(np.square(b), axis=1)[:, None].T
    return (sq0 + sq1 - 2.0 * (a @ b.T)).clip(min=0)


def eval_nn_matcher_batches(desc0, desc1, nn_thresh, is_mutual_nn=False):
    """evaluations/matcher.py:51-102: batched matcher, [b,d,n0] x [b,d,n1] -> [b,n0+1,n1+1] with a
    dustbin row / column marking the unmatched lines of either image (and the corner set)."""
    b, _, n0 = desc0.shape
    n1 = desc1.shape[2]
    out = np.zeros((b, n0 + 1, n1 + 1))
    for i in range(b):
        d = eval_dist(desc0[i], desc1[i])
        idx = np.argmin(d, axis=1)
        keep = d[np.arange(n0), idx] < nn_thresh
        if is_mutual_nn:
            idx2 = np.argmin(d, axis=0)
            keep = np.logical_and(keep, np.arange(n0) == idx2[idx])
        m1, m2 = np.arange(n0)[keep], idx[keep]
        out[i, m1, m2] = 1
        out[i, np.delete(np.arange(n0 + 1), m1), -1] = 1
        out[i, -1, np.delete(np.arange(n1 + 1), m2)] = 1
        out[i, -1, -1] = 1
    return out
