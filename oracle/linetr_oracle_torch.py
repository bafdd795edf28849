"""CPU timing port of the reference path on torch-CPU functional ops.

TEST/BENCH INFRASTRUCTURE ONLY (see oracle/linetr_oracle.py for the rules).  The reference
itself is Python on top of torch-CPU aten kernels and cannot travel to the GPU box
(/root/reference does not exist there), so `bench.py --impl reference` and the
`cpu_baseline` leg time THIS restatement: it issues the same aten ops on the same shapes as
the reference (F.conv1d k=1 + batch_norm + relu MLPs, nn.Linear QKV over all T+1 rows,
masked_fill + softmax, erf-GELU FFN, einsum signature attention, numpy einsum / matmul /
argmin matcher), with the host's default torch/BLAS threading, one image per call as
`Matching.forward` does (models/matching.py:41,59,77-81).  It is pinned to the numpy
oracle and the committed reference outputs in tests/test_oracle.py.

Reference lines restated: models/line_transformer.py:9-20,22-38,40-73,107-136,149-183,
225-249,277-282; models/line_attention.py:13-21,42-94; models/line_process.py:198-201;
models/nn_matcher.py:3-31.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import linetr_oracle as _np_oracle


def prepare(sd_np: dict) -> dict:
    return {k: torch.from_numpy(np.asarray(v)) for k, v in sd_np.items()}


def _mlp(sd, prefix, x):
    """x: [B, C, N] channels-first like the reference."""
    idx = 0
    while f"{prefix}.{idx}.weight" in sd:
        x = F.conv1d(x, sd[f"{prefix}.{idx}.weight"], sd[f"{prefix}.{idx}.bias"])
        p = f"{prefix}.{idx + 1}"
        if p + ".running_mean" in sd:
            x = F.relu(F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"],
                                    sd[p + ".bias"], False, 0.1, 1e-5))
            idx += 3
        else:
            idx += 1
    return x


def _desc_layer(sd, p, x, mask):
    B, L, N, D = x.shape
    H, dk = 4, D // 4
    a = p + ".slf_attn"
    q = F.linear(x, sd[a + ".w_qs.weight"], sd[a + ".w_qs.bias"]).view(B, L, N, H, dk).transpose(2, 3)
    k = F.linear(x, sd[a + ".w_ks.weight"], sd[a + ".w_ks.bias"]).view(B, L, N, H, dk).transpose(2, 3)
    v = F.linear(x, sd[a + ".w_vs.weight"], sd[a + ".w_vs.bias"]).view(B, L, N, H, dk).transpose(2, 3)
    attn = torch.matmul(q / dk ** 0.5, k.transpose(3, 4))
    attn = attn.masked_fill(mask.unsqueeze(2) == 0, -1e9)
    o = torch.matmul(F.softmax(attn, dim=-1), v).transpose(2, 3).contiguous().view(B, L, N, D)
    o = F.linear(o, sd[a + ".fc.weight"], sd[a + ".fc.bias"])
    o = o + x
    o = F.layer_norm(o, (D,), sd[a + ".layer_norm.weight"], sd[a + ".layer_norm.bias"], 1e-6)
    f = p + ".pos_ffn"
    y = F.linear(F.gelu(F.linear(o, sd[f + ".w_1.weight"], sd[f + ".w_1.bias"])), sd[f + ".w_2.weight"], sd[f + ".w_2.bias"])
    y = y + o
    return F.layer_norm(y, (D,), sd[f + ".layer_norm.weight"], sd[f + ".layer_norm.bias"], 1e-6)


def line_transformer_forward(sd, data, image_shape=(480, 640)):
    """-> line_desc torch [B,256,L]; data: numpy or torch tensors in the tokenizer layout."""
    t = lambda k: torch.as_tensor(data[k]).float()
    klines, resp, angle, pnt = t("sublines"), t("resp_sublines"), t("angle_sublines"), t("pnt_sublines")
    desc, score, mask = t("desc_sublines"), t("score_sublines"), t("mask_sublines")
    h, w = image_shape
    size = torch.tensor([[float(w), float(h)]])
    center, scaling = size / 2, size.max(1, keepdim=True).values * 0.7
    nk = torch.zeros_like(klines)
    nk[:, :, 0] = (klines[:, :, 0] - center[:, None, :]) / scaling[:, None, :]
    nk[:, :, 1] = (klines[:, :, 1] - center[:, None, :]) / scaling[:, None, :]
    npnt = (pnt - center[:, None, None, :]) / scaling[:, None, None, :]
    B, L, T, D = desc.shape
    mid = (nk[:, :, 0] + nk[:, :, 1]) / 2.
    kpos = _mlp(sd, "klenc.line_position_enc.encoder",
                torch.cat([mid.transpose(1, 2), resp.transpose(1, 2), angle.transpose(1, 2)], dim=1))
    wp_in = torch.cat([npnt, score], dim=-1).transpose(-2, -1).reshape(B * L, 3, T)
    wpe = _mlp(sd, "klenc.word_position_enc.encoder", wp_in).transpose(-1, -2).reshape(B, L, T, D)
    x = torch.cat((sd["klenc.cls_token"].expand(B, L, 1, D), desc + wpe), dim=2)
    i, enc = 0, None
    while f"klenc.desc_layers.{i}.slf_attn.w_qs.weight" in sd:
        enc = _desc_layer(sd, f"klenc.desc_layers.{i}", x, mask)
        i += 1
    s = kpos + enc[:, :, 0, :].transpose(1, 2)
    i = 0
    while f"selfattn.layers.{i}.attn.merge.weight" in sd:
        p = f"selfattn.layers.{i}"
        q, k, v = [F.conv1d(s, sd[f"{p}.attn.proj.{j}.weight"], sd[f"{p}.attn.proj.{j}.bias"]).view(B, D // 4, 4, -1)
                   for j in range(3)]
        prob = F.softmax(torch.einsum("bdhn,bdhm->bhnm", q, k) / (D // 4) ** .5, dim=-1)
        msg = torch.einsum("bhnm,bdhm->bdhn", prob, v).contiguous().view(B, D, -1)
        msg = F.conv1d(msg, sd[f"{p}.attn.merge.weight"], sd[f"{p}.attn.merge.bias"])
        s = s + _mlp(sd, f"{p}.mlp", torch.cat([s, msg], dim=1))
        i += 1
    y = F.conv1d(s, sd["final_proj.weight"], sd["final_proj.bias"])
    return F.normalize(y, p=2, dim=1)


def match_pair(sd, side0, side1, nn_thresh=0.8, image_shape=(480, 640)):
    with torch.no_grad():
        d0 = line_transformer_forward(sd, side0, image_shape).numpy()
        d1 = line_transformer_forward(sd, side1, image_shape).numpy()
    dist = _np_oracle.get_dist_matrix(d0, d1)[0]
    dk = _np_oracle.subline2keyline(dist, np.asarray(side0["mat_klines2sublines"][0]),
                                    np.asarray(side1["mat_klines2sublines"][0]))
    return _np_oracle.nn_matcher_distmat(dk, nn_thresh, True), dk, d0, d1
