// GPU line tokenizer: the step that FEEDS the hot path.
// Replaces the per-line / per-token Python loops of line_tokenizer (reference
// models/line_process.py:100-196) and sample_descriptors (:86-98): token positions along each key
// line, split of long lines into sublines, masks, responses, angles, bilinear sampling +
// L2 normalisation of the dense 256-d descriptor map, score lookup.
//
// The host prepares per key line (vectorised numpy, no loops): end points in float64, the
// clipped end point, n_tokens = ceil(length / token_distance), first subline index.  Token
// positions are computed in float64 exactly as the reference's numpy code does (IEEE sqrt/div),
// then rounded to fp32, so positions, sublines, masks and responses are bit-identical; sampled
// descriptors agree to fp32 rounding.
#pragma once
#include "common.cuh"

namespace ltr {

struct TokLines {
  const double* sp;       // [K,2] start point (x, y)
  const double* ep;       // [K,2] end point BEFORE the in-place clip (direction of the line)
  const double* epc;      // [K,2] end point after clip to (width-0.6, height-0.6)
  const double* length;   // [K] length_klines (detector length, NOT the geometric one)
  const float* angle;     // [K,2]
  const int* n_tok;       // [K]
  const int* sub0;        // [K+1] first subline of each key line (prefix sum of n_sub)
  const int* sub2line;    // [S]
  int K, S, T;
  double token_distance;
};

// point at distance d from sp along the line (reference point_on_line, line_process.py:43-59)
__device__ __forceinline__ void point_on_line_f64(double spx, double spy, double epx, double epy, double d, double& x, double& y) {
  const double vx = epx - spx, vy = epy - spy;
  double dx, dy;
  if (vx != 0.0) {
    const double m = vy / vx;
    dx = sqrt(d * d / (1.0 + m * m));
    dy = m * dx;
  } else {
    dx = 0.0;
    dy = (vy > 0.0) ? d : -d;
  }
  x = dx + spx;
  y = dy + spy;
}

// token i of key line k (i < n_tok): i < n_tok-1 -> point at i*token_distance, last -> clipped end point
__device__ __forceinline__ void token_of_line(const TokLines& L, int k, int i, double& x, double& y) {
  if (i < L.n_tok[k] - 1) {
    point_on_line_f64(L.sp[2 * k], L.sp[2 * k + 1], L.ep[2 * k], L.ep[2 * k + 1], (double)i * L.token_distance, x, y);
  } else {
    x = L.epc[2 * k];
    y = L.epc[2 * k + 1];
  }
}

// one thread per (subline, token slot): pnt [S,T,2], mask [S,T+1,1]; slot 0 threads also write the
// subline end points [S,2,2], resp [S,1], angle [S,2]
__global__ void __launch_bounds__(256)
tok_points_kernel(TokLines L, float* __restrict__ pnt, float* __restrict__ mask, float* __restrict__ sublines,
                  float* __restrict__ resp, float* __restrict__ angle) {
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (long long)L.S * L.T) return;
  const int s = (int)(idx / L.T), j = (int)(idx % L.T);
  const int k = L.sub2line[s];
  const int sl = s - L.sub0[k];          // subline index within the key line
  const int n_sub = L.sub0[k + 1] - L.sub0[k];
  const int i = sl * L.T + j;
  const bool live = i < L.n_tok[k];
  double x = 0.0, y = 0.0;
  if (live) token_of_line(L, k, i, x, y);
  pnt[2 * idx] = (float)x;
  pnt[2 * idx + 1] = (float)y;
  mask[(long long)s * (L.T + 1) + 1 + j] = live ? 1.f : 0.f;
  if (j == 0) {
    mask[(long long)s * (L.T + 1)] = 1.f;
    double ax, ay, bx, by;
    if (sl == 0) { ax = L.sp[2 * k]; ay = L.sp[2 * k + 1]; } else token_of_line(L, k, sl * L.T - 1, ax, ay);
    if (sl == n_sub - 1) { bx = L.epc[2 * k]; by = L.epc[2 * k + 1]; } else token_of_line(L, k, (sl + 1) * L.T - 1, bx, by);
    sublines[4 * s + 0] = (float)ax; sublines[4 * s + 1] = (float)ay;
    sublines[4 * s + 2] = (float)bx; sublines[4 * s + 3] = (float)by;
    const double dxx = bx - ax, dyy = by - ay;
    resp[s] = (float)(sqrt(dxx * dxx + dyy * dyy) / (L.token_distance * (double)L.T));
    angle[2 * s] = L.angle[2 * k];
    angle[2 * s + 1] = L.angle[2 * k + 1];
  }
}

// one warp per token: bilinear sample of dense [C=256, Hc, Wc] at the token (torch grid_sample,
// mode bilinear, zeros padding), L2 normalise over channels, nearest score lookup.
__global__ void __launch_bounds__(256)
tok_sample_kernel(const float* __restrict__ pnt, int n_tokens, const float* __restrict__ dense, int Hc, int Wc,
                  const float* __restrict__ score_map, int H, int W, int align_corners, float* __restrict__ desc,
                  float* __restrict__ score) {
  const int t = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (t >= n_tokens) return;
  const float px = pnt[2 * t], py = pnt[2 * t + 1];
  // reference sample_descriptors: (kp - s/2 + 0.5) / (w*s - s/2 - 0.5) * 2 - 1 with s = 8, fp32 arithmetic
  const float gx = ((px - 4.0f) + 0.5f) / ((float)Wc * 8.0f - 4.0f - 0.5f) * 2.0f - 1.0f;
  const float gy = ((py - 4.0f) + 0.5f) / ((float)Hc * 8.0f - 4.0f - 0.5f) * 2.0f - 1.0f;
  float ix, iy;
  if (align_corners) {
    ix = (gx + 1.f) / 2.f * (float)(Wc - 1);
    iy = (gy + 1.f) / 2.f * (float)(Hc - 1);
  } else {
    ix = ((gx + 1.f) * (float)Wc - 1.f) / 2.f;
    iy = ((gy + 1.f) * (float)Hc - 1.f) / 2.f;
  }
  const float fx = floorf(ix), fy = floorf(iy);
  const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
  const float wnw = ((float)x1 - ix) * ((float)y1 - iy), wne = (ix - (float)x0) * ((float)y1 - iy);
  const float wsw = ((float)x1 - ix) * (iy - (float)y0), wse = (ix - (float)x0) * (iy - (float)y0);
  const bool inx0 = x0 >= 0 && x0 < Wc, inx1 = x1 >= 0 && x1 < Wc, iny0 = y0 >= 0 && y0 < Hc, iny1 = y1 >= 0 && y1 < Hc;
  const long long plane = (long long)Hc * Wc;
  float v[8];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float* ch = dense + (long long)(lane + 32 * i) * plane;
    float a = 0.f;
    if (inx0 && iny0) a += ch[y0 * Wc + x0] * wnw;
    if (inx1 && iny0) a += ch[y0 * Wc + x1] * wne;
    if (inx0 && iny1) a += ch[y1 * Wc + x0] * wsw;
    if (inx1 && iny1) a += ch[y1 * Wc + x1] * wse;
    v[i] = a;
    ss = fmaf(a, a, ss);
  }
  ss = warp_sum(ss);
  const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
  for (int i = 0; i < 8; ++i) desc[(long long)t * 256 + lane + 32 * i] = v[i] * inv;
  if (lane == 0) {
    int rx = (int)rintf(px), ry = (int)rintf(py);   // torch.round: half to even
    rx = min(rx, W - 1);
    ry = min(ry, H - 1);
    score[t] = score_map[(long long)ry * W + rx];
  }
}

}  // namespace ltr
