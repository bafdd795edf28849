"""`get_dist_matrix` (CUDA) and the line tokeniser that feeds the hot path.

`get_dist_matrix` is on the hot path (reference models/line_process.py:198-201) and runs on
the GPU through `ltr_match`.  The tokeniser (reference models/line_process.py:6-196; SURVEY.md
§8f row 1, the first "next" row) exists twice: `line_tokenizer_gpu` (CUDA, `ltr_tokenize`; used by
`LineTransformer.preprocess` whenever the SuperPoint outputs live on a CUDA device) and
`line_tokenizer`, numpy/PyTorch glue that reproduces the reference bit for bit - including its
quirks (in-place end-point clipping, `index[:max_keylines]` dropping the shortest line when
max_keylines == -1, the torch-version dependent `align_corners`) - and is the pin the GPU
tokeniser is tested against.  The cheap per-image filters (`remove_borders`,
`filter_by_length`, cv2 KeyLine conversion) stay vectorised numpy.
"""
from __future__ import annotations

import math

import numpy as np
import torch


# --------------------------------------------------------------------------- hot path
def get_dist_matrix(desc0, desc1):
    """(2 - 2 * desc0^T desc1).clip(0) per batch element; numpy [B,d,n],[B,d,m] -> [B,n,m] float32."""
    from . import _ops, _native as N
    desc0 = np.ascontiguousarray(desc0, dtype=np.float32)
    desc1 = np.ascontiguousarray(desc1, dtype=np.float32)
    B, d, n = desc0.shape
    m = desc1.shape[2]
    if B == 0 or n == 0 or m == 0:
        return np.zeros((B, n, m), dtype=np.float32)
    dev = _ops.current_cuda_device()
    a = torch.from_numpy(desc0).to(dev, non_blocking=True)
    b = torch.from_numpy(desc1).to(dev, non_blocking=True)
    out = _ops.match_descriptors(a, b, N.LAYOUT_CHANNEL_FIRST, B, 0.0, False, n0=n, n1=m, d=d, want_matches=False)
    return out["dist_key"].view(B, n, m).cpu().numpy()


# --------------------------------------------------------------------------- tokeniser glue
def get_angles(lines):
    """Orientation code (cos 2a, sin 2a) with a = atan2(dx, dy) folded into [0, pi)."""
    if len(lines) == 0:
        return []
    delta = lines[:, 1] - lines[:, 0]
    ang = np.arctan2(delta[:, 0], delta[:, 1])
    ang = np.where(ang < 0, ang + np.pi, ang)
    return np.stack([np.cos(2 * ang), np.sin(2 * ang)], axis=1)


def change_cv2_T_np(klines_cv):
    """cv2 KeyLine list -> {'klines' [n,2,2] (left end point first), 'length_klines', 'angles'}."""
    n = len(klines_cv)
    pts = np.zeros((n, 2, 2), dtype=np.float64)
    length = np.zeros((n,), dtype=np.float64)
    for i, ln in enumerate(klines_cv):
        a = (ln.startPointX, ln.startPointY)
        b = (ln.endPointX, ln.endPointY)
        pts[i] = (a, b) if a[0] < b[0] else (b, a)
        length[i] = ln.lineLength * (2 ** ln.octave)
    return {"klines": pts, "length_klines": length, "angles": get_angles(pts)}


def remove_borders(lines, border: int, height: int, width: int, valid_mask_given=None):
    k = lines["klines"]
    inside = np.ones(len(k), dtype=bool)
    for e in (0, 1):
        inside &= (k[:, e, 0] >= border) & (k[:, e, 0] < (width - border))
        inside &= (k[:, e, 1] >= border) & (k[:, e, 1] < (height - border))
    eps = 0.001
    k[:, :, 0] = k[:, :, 0].clip(max=width - eps - border)
    k[:, :, 1] = k[:, :, 1].clip(max=height - eps - border)
    if isinstance(valid_mask_given, np.ndarray):
        a = np.floor(k[:, 0]).astype(int)
        b = np.floor(k[:, 1]).astype(int)
        inside &= (valid_mask_given[a[:, 1], a[:, 0]] + valid_mask_given[b[:, 1], b[:, 0]]).astype(bool)
    return {key: np.asarray(v)[inside] for key, v in lines.items()}


def filter_by_length(lines, min_length, max_sublines):
    keep = lines["length_klines"] > min_length
    k, ln = lines["klines"][keep], lines["length_klines"][keep]
    order = np.argsort(ln)[::-1][:max_sublines]   # NB: max_sublines == -1 drops the shortest line
    k, ln = k[order], ln[order]
    return {"klines": k, "length_klines": ln, "angles": get_angles(k)}


def _tokens_on_line(kline, n_tokens, token_distance, width, height):
    """n_tokens-1 points every token_distance px from the start, then the (clipped, in place) end."""
    sp, ep = kline[0], kline[1]
    seg_len = math.sqrt(float(((ep - sp) ** 2).sum()))
    dists = np.arange(n_tokens - 1, dtype=np.float64) * token_distance
    assert n_tokens <= 1 or seg_len >= dists[-1], "distance should be smaller than line length!"
    vec = ep - sp
    if vec[0] != 0:
        m = vec[1] / vec[0]
        dx = np.sqrt(dists ** 2 / (1 + m ** 2))
        dy = m * dx
    else:
        dx = np.zeros_like(dists)
        dy = dists if vec[1] > 0 else -dists
    pts = np.stack([dx, dy], axis=1) + sp
    ep[0] = min(ep[0], width - 0.6)      # in place: the stored key line end point is clipped too
    ep[1] = min(ep[1], height - 0.6)
    return np.concatenate([pts, ep[None]], axis=0)


def sample_descriptors(keypoints, descriptors, s: int = 8):
    """Bilinear sampling of the dense descriptor map at pixel positions, L2-normalised."""
    b, c, h, w = descriptors.shape
    kp = keypoints - s / 2 + 0.5
    kp = kp / torch.tensor([(w * s - s / 2 - 0.5), (h * s - s / 2 - 0.5)]).to(kp)[None]
    kp = kp * 2 - 1
    # the reference keys align_corners on the third character of torch.__version__
    kwargs = {"align_corners": True} if int(torch.__version__[2]) > 2 else {}
    out = torch.nn.functional.grid_sample(descriptors, kp.view(b, 1, -1, 2), mode="bilinear", **kwargs)
    return torch.nn.functional.normalize(out.reshape(b, c, -1), p=2, dim=1)


def line_tokenizer(klines, token_distance, max_tokens, pred_superpoint, image_shape):
    """Key lines -> sublines/tokens/masks/descriptors dict consumed by LineTransformer.forward."""
    height, width = image_shape
    K = len(klines["klines"])
    sub_l, tok_l, mask_l, resp_l, ang_l, nsub = [], [], [], [], [], []
    max_len = token_distance * max_tokens
    for i in range(K):
        kline = klines["klines"][i]
        n_tok = int(math.ceil(klines["length_klines"][i] / token_distance))
        toks = _tokens_on_line(kline, n_tok, token_distance, width, height)
        n_sub = int(math.ceil(n_tok / max_tokens))
        subs = np.zeros((n_sub, 2, 2))
        subs[0, 0] = kline[0]
        subs[-1, 1] = kline[1]
        for j in range(n_sub - 1):
            cut = toks[(j + 1) * max_tokens - 1]
            subs[j, 1] = cut
            subs[j + 1, 0] = cut
        t = np.zeros((n_sub, max_tokens, 2))
        m = np.zeros((n_sub, max_tokens + 1, 1))
        m[:, 0] = 1
        for j in range(n_sub):
            part = toks[j * max_tokens:(j + 1) * max_tokens]
            t[j, :len(part)] = part
            m[j, 1:len(part) + 1] = 1
        sub_l.append(subs)
        tok_l.append(t)
        mask_l.append(m)
        resp_l.append(np.sqrt(((subs[:, 1] - subs[:, 0]) ** 2).sum(axis=1, keepdims=True)) / max_len)
        ang_l.append(np.repeat(np.asarray(klines["angles"][i])[None], n_sub, axis=0))
        nsub.append(n_sub)
    device = pred_superpoint["dense_descriptor"].device
    f = lambda arrs, shape: torch.from_numpy(np.concatenate(arrs, axis=0).reshape(shape)).float().to(device)
    slines = f(sub_l, (-1, 2, 2))
    tokens = f(tok_l, (-1, max_tokens, 2))
    masks = f(mask_l, (-1, max_tokens + 1, 1))
    responses = f(resp_l, (-1, 1))
    angles = f(ang_l, (-1, 2))
    S = slines.shape[0]
    adj = torch.zeros((K, S)).to(device)
    st = 0
    for i, n in enumerate(nsub):
        adj[i, st:st + n] = 1 / n
        st += n
    dense = pred_superpoint["dense_descriptor"]
    desc = sample_descriptors(tokens[None], dense, 8)[0].reshape(256, S, max_tokens).permute(1, 2, 0)
    score_map = pred_superpoint["dense_score"].transpose(1, 2)
    pos = torch.round(tokens).long().reshape(-1, 2)
    pos[:, 0] = pos[:, 0].clip(max=score_map.shape[1] - 1)
    pos[:, 1] = pos[:, 1].clip(max=score_map.shape[2] - 1)
    scores = score_map[0][pos[:, 0], pos[:, 1]].reshape(S, max_tokens, 1)

    klines["klines"] = torch.from_numpy(klines["klines"]).float().to(device)[None]
    klines["length_klines"] = torch.from_numpy(klines["length_klines"]).float().to(device)[None]
    klines["angles"] = torch.from_numpy(np.asarray(klines["angles"])).float().to(device)[None]
    klines["sublines"] = slines[None]
    klines["pnt_sublines"] = tokens[None]
    klines["mask_sublines"] = masks[None]
    klines["resp_sublines"] = responses[None]
    klines["angle_sublines"] = angles[None]
    klines["desc_sublines"] = desc[None]
    klines["score_sublines"] = scores[None]
    klines["mat_klines2sublines"] = adj[None]
    return klines


def line_tokenizer_gpu(klines, token_distance, max_tokens, pred_superpoint, image_shape):
    """`line_tokenizer` on the GPU (ltr_tokenize): same inputs, same output dict, no Python loops.
    Token positions, sublines, masks, responses and the adjacency are bit-identical to the CPU
    version (float64 arithmetic, rounded once to fp32); sampled descriptors agree to fp32 rounding."""
    import ctypes as C
    from . import _native as N
    dense = pred_superpoint["dense_descriptor"]
    if not dense.is_cuda:
        raise N.LtrError("line_tokenizer_gpu needs the SuperPoint outputs on a CUDA device")
    dev = dense.device
    height, width = image_shape
    T = int(max_tokens)
    kl = klines["klines"]
    K = len(kl)
    length = np.ascontiguousarray(klines["length_klines"], dtype=np.float64)
    n_tok = np.ceil(length / token_distance).astype(np.int64)
    n_sub = -(-n_tok // T)
    geo = np.sqrt(((kl[:, 1] - kl[:, 0]) ** 2).sum(axis=1))
    assert bool((geo >= np.maximum(n_tok - 2, 0) * token_distance).all()), "distance should be smaller than line length!"
    sp = np.ascontiguousarray(kl[:, 0], dtype=np.float64)
    ep = np.ascontiguousarray(kl[:, 1], dtype=np.float64).copy()
    kl[:, 1, 0] = np.minimum(kl[:, 1, 0], width - 0.6)      # in place, like the reference
    kl[:, 1, 1] = np.minimum(kl[:, 1, 1], height - 0.6)
    epc = np.ascontiguousarray(kl[:, 1], dtype=np.float64)
    sub0 = np.zeros(K + 1, dtype=np.int64)
    sub0[1:] = np.cumsum(n_sub)
    S = int(sub0[-1])
    sub2line = np.repeat(np.arange(K), n_sub)
    angles = np.ascontiguousarray(klines["angles"], dtype=np.float32)
    up = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev)
    d_sp, d_ep, d_epc, d_len = up(sp, np.float64), up(ep, np.float64), up(epc, np.float64), up(length, np.float64)
    d_ang, d_ntok, d_sub0, d_s2l = up(angles, np.float32), up(n_tok, np.int32), up(sub0, np.int32), up(sub2line, np.int32)
    f32 = dict(dtype=torch.float32, device=dev)
    slines = torch.empty((S, 2, 2), **f32)
    tokens = torch.empty((S, T, 2), **f32)
    masks = torch.empty((S, T + 1, 1), **f32)
    responses = torch.empty((S, 1), **f32)
    ang = torch.empty((S, 2), **f32)
    desc = torch.empty((S, T, 256), **f32)
    scores = torch.empty((S, T, 1), **f32)
    dense_c = dense.float().contiguous()
    score_map = pred_superpoint["dense_score"].float().contiguous()
    b, c, hc, wc = dense_c.shape
    inp = N.LtrTokenizeInput(d_sp.data_ptr(), d_ep.data_ptr(), d_epc.data_ptr(), d_len.data_ptr(), d_ang.data_ptr(),
                             d_ntok.data_ptr(), d_sub0.data_ptr(), d_s2l.data_ptr(), K, S, T, float(token_distance),
                             dense_c.data_ptr(), c, hc, wc, score_map.data_ptr(), score_map.shape[-2], score_map.shape[-1],
                             1 if int(torch.__version__[2]) > 2 else 0)
    p = lambda t: C.c_void_p(t.data_ptr())
    with torch.cuda.device(dev):
        rc = N.load().ltr_tokenize(C.byref(inp), p(slines), p(tokens), p(masks), p(responses), p(ang), p(desc), p(scores),
                                   dev.index, C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    N.check(rc, "ltr_tokenize")
    adj = np.zeros((K, S), dtype=np.float32)
    w = (1.0 / n_sub).astype(np.float64)
    adj[sub2line, np.arange(S)] = w[sub2line]
    klines["klines"] = torch.from_numpy(klines["klines"]).float().to(dev)[None]
    klines["length_klines"] = torch.from_numpy(klines["length_klines"]).float().to(dev)[None]
    klines["angles"] = torch.from_numpy(np.asarray(klines["angles"])).float().to(dev)[None]
    klines["sublines"] = slines[None]
    klines["pnt_sublines"] = tokens[None]
    klines["mask_sublines"] = masks[None]
    klines["resp_sublines"] = responses[None]
    klines["angle_sublines"] = ang[None]
    klines["desc_sublines"] = desc[None]
    klines["score_sublines"] = scores[None]
    klines["mat_klines2sublines"] = torch.from_numpy(adj).to(dev)[None]
    return klines
