"""Device-level wrappers around the C ABI: CUDA torch tensors in, CUDA torch tensors out.

PyTorch is used here only for device memory, streams and dtype plumbing; all arithmetic
happens inside liblinetr_b200.so.  Every function raises if the tensors are not on a CUDA
device (no CPU fallback).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _native as N


def _stream_ptr(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _req_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise N.LtrError(f"linetr_b200: '{name}' is on {t.device}; the B200 path needs CUDA tensors "
                         "(there is no CPU fallback)")


def current_cuda_device() -> torch.device:
    """Device the host-numpy entry points (nn_matcher*, get_dist_matrix) compute on."""
    if not torch.cuda.is_available():
        raise N.LtrError("linetr_b200 needs a CUDA device (there is no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _i32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.int32:
        t = t.to(torch.int32)
    return t.contiguous()


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class ModelHandle:
    """Owns one LtrModel* (packed weights on one GPU)."""

    def __init__(self, state_dict: dict, device_index: int, d_model=256, n_heads=4, d_inner=1024,
                 n_desc_layers=1, n_sig_layers=7):
        lib = N.load()
        keep = []
        arr = (N.LtrTensor * len(state_dict))()
        i = 0
        for k, v in state_dict.items():
            a = np.ascontiguousarray(np.asarray(v), dtype=np.float32)
            keep.append(a)
            arr[i].name = k.encode()
            arr[i].data = a.ctypes.data
            arr[i].numel = a.size
            i += 1
        cfg = N.LtrConfig(d_model, n_heads, d_inner, n_desc_layers, n_sig_layers)
        h = C.c_void_p()
        N.check(lib.ltr_create(arr, len(state_dict), C.byref(cfg), int(device_index), C.byref(h)), "ltr_create")
        self._lib = lib
        self.ptr = h
        self.device_index = int(device_index)
        self._ws = {}   # one workspace per CUDA stream: calls on one stream are ordered, streams may overlap

    def close(self):
        if getattr(self, "ptr", None):
            self._lib.ltr_destroy(self.ptr)
            self.ptr = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def workspace(self, n_images: int, n_lines: int, n_tokens: int) -> torch.Tensor:
        need = int(self._lib.ltr_encode_workspace_bytes(self.ptr, n_images, n_lines, n_tokens))
        if need < 0:
            N.check(need, "ltr_encode_workspace_bytes")
        dev = torch.device("cuda", self.device_index)
        key = torch.cuda.current_stream(dev).cuda_stream
        ws = self._ws.get(key)
        if ws is None or ws.numel() < need:
            if len(self._ws) >= 16:   # callers that keep creating streams: do not hoard one workspace per dead stream
                self._ws.clear()
            ws = self._ws[key] = torch.empty(need, dtype=torch.uint8, device=dev)
        return ws


def encode(handle: ModelHandle, sublines, resp, angle, pnt, desc, score, image_wh, *, lines_per_image=None,
           cu_lines_host=None, cu_lines_dev=None, want_cf=True, want_rows=False, want_tiles=False):
    """ltr_encode on flattened lines.

    sublines [R,2,2], resp [R,1], angle [R,2], pnt [R,T,2], desc [R,T,256], score [R,T,1]; either
    `lines_per_image` (uniform batch) or `cu_lines_host` (np.int32 [B+1]) + `cu_lines_dev`.
    Returns (desc_cf flat [256*R] or None, desc_rows [R,256] or None), plus the descriptor tile image
    (uint8 tensor, the matcher's operand format) as a third element when `want_tiles`.
    """
    for nm, t in (("sublines", sublines), ("resp_sublines", resp), ("angle_sublines", angle),
                  ("pnt_sublines", pnt), ("desc_sublines", desc), ("score_sublines", score)):
        _req_cuda(t, nm)
    dev = desc.device
    if dev.index != handle.device_index:
        raise N.LtrError(f"linetr_b200: inputs on cuda:{dev.index} but weights on cuda:{handle.device_index}")
    sublines, resp, angle, pnt, desc, score = map(_f32c, (sublines, resp, angle, pnt, desc, score))
    R, T = int(desc.shape[0]), int(desc.shape[1])
    if desc.shape[2] != 256:
        raise N.LtrError("linetr_b200: descriptor_dim must be 256")
    if cu_lines_host is not None:
        cu_lines_host = np.ascontiguousarray(cu_lines_host, dtype=np.int32)
        n_images = len(cu_lines_host) - 1
        cu_lines_dev = _i32c(cu_lines_dev)
        lpi = 0
    else:
        lpi = int(lines_per_image)
        n_images = R // lpi if lpi > 0 else 0
    out_cf = torch.empty(R * 256, dtype=torch.float32, device=dev) if want_cf else None
    out_rows = torch.empty((R, 256), dtype=torch.float32, device=dev) if want_rows else None
    out_tiles = None
    if want_tiles:
        if want_cf or not want_rows:
            raise N.LtrError("encode: want_tiles needs want_rows=True and want_cf=False")
        out_tiles = torch.empty(max(int(handle._lib.ltr_desc_tiles_bytes(R)), 1), dtype=torch.uint8, device=dev)
    if R == 0 or n_images == 0:
        return (out_cf, out_rows, out_tiles) if want_tiles else (out_cf, out_rows)
    ws = handle.workspace(n_images, R, T)
    inp = N.LtrEncodeInput(
        sublines.data_ptr(), resp.data_ptr(), angle.data_ptr(), pnt.data_ptr(), desc.data_ptr(), score.data_ptr(),
        cu_lines_host.ctypes.data if cu_lines_host is not None else None,
        cu_lines_dev.data_ptr() if cu_lines_host is not None else None,
        n_images, R, T, lpi, float(image_wh[0]), float(image_wh[1]))
    outp = N.LtrEncodeOutput(out_cf.data_ptr() if out_cf is not None else None,
                             out_rows.data_ptr() if out_rows is not None else None,
                             out_tiles.data_ptr() if out_tiles is not None else None)
    with torch.cuda.device(dev):
        rc = handle._lib.ltr_encode(handle.ptr, C.byref(inp), C.byref(outp), _ptr(ws), ws.numel(), _stream_ptr(dev))
    N.check(rc, "ltr_encode")
    return (out_cf, out_rows, out_tiles) if want_tiles else (out_cf, out_rows)


_match_ws = {}   # (device index, stream) -> uint8 scratch tensor, grown on demand


def _match_workspace(dev, need: int) -> torch.Tensor:
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    ws = _match_ws.get(key)
    if ws is None or ws.numel() < need:
        if len(_match_ws) >= 16:
            _match_ws.clear()
        ws = _match_ws[key] = torch.empty(max(need, 1), dtype=torch.uint8, device=dev)
    return ws


def match_descriptors(desc0, desc1, layout, n_pairs, nn_thresh, mutual=True, *, n0=0, n1=0, cu0=None, cu1=None,
                      max_n0=0, max_n1=0, sub_off0=None, sub_off1=None, cuk0=None, cuk1=None, max_k0=0, max_k1=0,
                      total_k0=None, total_k1=None, d=256, want_dist=True, want_matches=True, dist_mode=0,
                      tiles0=None, tiles1=None, tiles_lines=(0, 0), tiles_row0=(0, 0), gather=None):
    """ltr_match.  Returns dict(matches0, scores0, nn1, counts, dist_key, stride).

    want_dist=False (no key-line merging, d == 256) never materialises the distance matrix: the row
    argmin lives in the epilogue of the tensor-core contraction.  want_matches=False computes the
    distance matrices only (get_dist_matrix).  tiles0/tiles1: descriptor tile images from
    `encode(want_tiles=True)` (uniform batches with n % 128 == 0).  gather: an `N.LtrPeerGather` - the tail
    kernel then also publishes the per-pair counts to every rank (engine.PeerCounts)."""
    _req_cuda(desc0, "desc0")
    _req_cuda(desc1, "desc1")
    dev = desc0.device
    desc0, desc1 = _f32c(desc0), _f32c(desc1)
    seg = sub_off0 is not None
    varlen = cu0 is not None
    mx0 = max_n0 if varlen else n0
    mx1 = max_n1 if varlen else n1
    mk0, mk1 = (max_k0, max_k1) if seg else (mx0, mx1)
    total_n0 = int(desc0.numel() // d) if varlen else n_pairs * n0
    total_n1 = int(desc1.numel() // d) if varlen else n_pairs * n1
    if total_k0 is None:
        total_k0, total_k1 = total_n0, total_n1
    if seg or d != 256:
        want_dist = True
    stride = mk0 * mk1
    i32 = dict(dtype=torch.int32, device=dev)
    f32 = dict(dtype=torch.float32, device=dev)
    out = {"stride": stride}
    if want_matches:
        out["matches0"] = torch.empty(max(total_k0, 1), **i32)[:total_k0]
        out["scores0"] = torch.empty(max(total_k0, 1), **f32)[:total_k0]
        out["nn1"] = torch.empty(max(total_k1, 1), **i32)[:total_k1]
        out["counts"] = (torch.zeros if n_pairs == 0 else torch.empty)(max(n_pairs, 1), **i32)[:n_pairs]   # zeroed by the kernels
    out["dist_key"] = torch.empty(max(n_pairs * stride, 1), **f32)[:n_pairs * stride] if want_dist else None
    if n_pairs == 0:
        return out
    dist_sub = torch.empty(max(n_pairs * mx0 * mx1, 1), **f32) if seg else None
    tensors = [cu0, cu1, sub_off0, sub_off1, cuk0, cuk1]
    tensors = [(_i32c(t) if t is not None else None) for t in tensors]
    inp = N.LtrMatchInput(desc0.data_ptr(), desc1.data_ptr(), int(layout), int(d), int(n_pairs), int(n0), int(n1),
                          *[t.data_ptr() if t is not None else None for t in tensors],
                          int(max_n0), int(max_n1), int(max_k0), int(max_k1), 0, float(nn_thresh), int(bool(mutual)),
                          int(total_n0), int(total_n1), int(dist_mode),
                          tiles0.data_ptr() if tiles0 is not None else None,
                          tiles1.data_ptr() if tiles1 is not None else None,
                          int(tiles_lines[0]), int(tiles_lines[1]), int(tiles_row0[0]), int(tiles_row0[1]))
    lib = N.load()
    need = int(lib.ltr_match_workspace_bytes(C.byref(inp)))
    if need < 0:
        N.check(need, "ltr_match_workspace_bytes")
    ws = _match_workspace(dev, need)
    g = lambda k: out[k].data_ptr() if out.get(k) is not None else None
    o = N.LtrMatchOutput(g("matches0"), g("scores0"), g("nn1"), g("counts"), g("dist_key"),
                         dist_sub.data_ptr() if dist_sub is not None else None, ws.data_ptr(), ws.numel(),
                         C.pointer(gather) if gather is not None else None)
    with torch.cuda.device(dev):
        rc = lib.ltr_match(C.byref(inp), C.byref(o), dev.index, _stream_ptr(dev))
    N.check(rc, "ltr_match")
    return out


def match_distmat(dist: torch.Tensor, nn_thresh: float, mutual=True):
    """ltr_match_distmat on dist [P, n0, n1] (CUDA).  Returns dict(matches0 [P,n0], scores0, nn1, counts)."""
    _req_cuda(dist, "dist_mat")
    dev = dist.device
    dist = _f32c(dist)
    P, n0, n1 = (int(x) for x in dist.shape)
    i32 = dict(dtype=torch.int32, device=dev)
    out = {
        "matches0": torch.full((P, n0), -1, **i32),
        "scores0": torch.empty((P, n0), dtype=torch.float32, device=dev),
        "nn1": torch.full((P, n1), -1, **i32),
        "counts": torch.zeros(P, **i32),
    }
    if P == 0 or n0 == 0:
        return out
    with torch.cuda.device(dev):
        rc = N.load().ltr_match_distmat(_ptr(dist), P, n0, n1, 0, float(nn_thresh), int(bool(mutual)),
                                        _ptr(out["matches0"]), _ptr(out["scores0"]), _ptr(out["nn1"]),
                                        _ptr(out["counts"]), dev.index, _stream_ptr(dev))
    N.check(rc, "ltr_match_distmat")
    return out


def linear(x: torch.Tensor, w: torch.Tensor, bias=None, res=None, act=0) -> torch.Tensor:
    """ltr_linear: act(x @ w.T + bias) (+ res) on the fp32 CUDA-core engine (unit-test hook and
    numerics yardstick of the tensor-core engine)."""
    _req_cuda(x, "x")
    x, w = _f32c(x), _f32c(w)
    M, K = x.shape
    Nn = w.shape[0]
    y = torch.empty((M, Nn), dtype=torch.float32, device=x.device)
    bias = _f32c(bias) if bias is not None else None
    res = _f32c(res) if res is not None else None
    with torch.cuda.device(x.device):
        rc = N.load().ltr_linear(_ptr(x), K, _ptr(w), _ptr(bias), _ptr(res), Nn, _ptr(y), Nn, M, Nn, K, int(act),
                                 x.device.index, _stream_ptr(x.device))
    N.check(rc, "ltr_linear")
    return y


def linear_img(x: torch.Tensor, w: torch.Tensor, bias=None, res=None, act=0, bn_hint=0):
    """ltr_linear_img (unit-test hook of the image-operand GEMM engine): returns (y_fp32, y_from_image)."""
    _req_cuda(x, "x")
    x = _f32c(x)
    w = _f32c(w).cpu()
    M, K = x.shape
    Nn = w.shape[0]
    y = torch.empty((M, Nn), dtype=torch.float32, device=x.device)
    y2 = torch.empty((M, Nn), dtype=torch.float32, device=x.device)
    bias = _f32c(bias) if bias is not None else None
    res = _f32c(res) if res is not None else None
    with torch.cuda.device(x.device):
        rc = N.load().ltr_linear_img(_ptr(x), K, _ptr(w), _ptr(bias), _ptr(res), Nn, _ptr(y), Nn, _ptr(y2), M, Nn, K,
                                     int(act), int(bn_hint), x.device.index, _stream_ptr(x.device))
    N.check(rc, "ltr_linear_img")
    return y, y2


def linear_img_norm(x: torch.Tensor, w: torch.Tensor, bias=None, res=None, norm=1, eps=1e-6, gamma=None, beta=None, add=None):
    """ltr_linear_img_norm (unit-test hook of the row-normalising GEMM epilogue, N = 256):
    returns (y_fp32, y_from_image)."""
    _req_cuda(x, "x")
    x = _f32c(x)
    w = _f32c(w).cpu()
    M, K = x.shape
    if w.shape[0] != 256:
        raise N.LtrError("linear_img_norm: the row-norm epilogue is built for N = 256")
    y = torch.empty((M, 256), dtype=torch.float32, device=x.device)
    y2 = torch.empty((M, 256), dtype=torch.float32, device=x.device)
    f = lambda t: _f32c(t) if t is not None else None
    bias, res, gamma, beta, add = map(f, (bias, res, gamma, beta, add))
    with torch.cuda.device(x.device):
        rc = N.load().ltr_linear_img_norm(_ptr(x), K, _ptr(w), _ptr(bias), _ptr(res), 256, int(norm), float(eps), _ptr(gamma),
                                          _ptr(beta), _ptr(add), 256, _ptr(y), 256, _ptr(y2), M, K, x.device.index,
                                          _stream_ptr(x.device))
    N.check(rc, "ltr_linear_img_norm")
    return y, y2
