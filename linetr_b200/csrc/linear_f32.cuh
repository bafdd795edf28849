// fp32 CUDA-core GEMM engine: Y = act(X W^T + b) (+ R), row-major, W is [N, K].
// Used for the narrow layers that stay on CUDA cores and as the numerics yardstick of the
// tensor-core engine (tc_weight.cuh).  Replaces the aten addmm / MKLDNN conv1d(k=1)
// calls of the reference (models/line_transformer.py:9-20, models/line_attention.py:55-57,
// 69,89).
#pragma once
#include "common.cuh"

namespace ltr {

enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2 };

struct LinearArgs {
  const float* A; int lda;
  const float* W;            // [N, K] row-major
  const float* bias;         // [N] or nullptr
  const float* R; int ldr;   // residual added after the activation, or nullptr
  float* C; int ldc;
  int M, N, K;
  int act;
  // batching over blockIdx.z (element strides)
  long long sA, sW, sB, sR, sC;
  int nz;
};

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == ACT_RELU) return fmaxf(v, 0.f);
  if (act == ACT_GELU) return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
  return v;
}

constexpr int LF_BM = 128, LF_BN = 64, LF_BK = 16, LF_THREADS = 256, LF_PAD = 4;

// 128x64 CTA tile, 8x4 register micro-tile, K step 16, register-prefetch double buffering.
// Requires K % 16 == 0, N % 4 == 0, lda/ldc/ldr % 4 == 0 and 16-byte aligned bases.
__global__ void __launch_bounds__(LF_THREADS) linear_f32_kernel(LinearArgs p) {
  __shared__ __align__(16) float As[2][LF_BK][LF_BM + LF_PAD];
  __shared__ __align__(16) float Ws[2][LF_BK][LF_BN + LF_PAD];
  const int z = blockIdx.z;
  const float* __restrict__ A = p.A + z * p.sA;
  const float* __restrict__ W = p.W + z * p.sW;
  const float* __restrict__ bias = p.bias ? p.bias + z * p.sB : nullptr;
  const float* __restrict__ R = p.R ? p.R + z * p.sR : nullptr;
  float* __restrict__ C = p.C + z * p.sC;
  const int m0 = blockIdx.x * LF_BM, n0 = blockIdx.y * LF_BN;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;

  // global->register staging: A tile 128x16 = 512 float4 (2/thread), W tile 64x16 = 256 float4
  const int a_row0 = tid >> 2, a_kq = tid & 3;  // rows a_row0 and a_row0+64
  float4 ra[2], rw;
  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int r = m0 + a_row0 + i * 64;
      ra[i] = (r < p.M) ? *reinterpret_cast<const float4*>(A + (long long)r * p.lda + k0 + a_kq * 4)
                        : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    int n = n0 + a_row0;
    rw = (n < p.N) ? *reinterpret_cast<const float4*>(W + (long long)n * p.K + k0 + a_kq * 4)
                   : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int r = a_row0 + i * 64;
      As[buf][a_kq * 4 + 0][r] = ra[i].x;
      As[buf][a_kq * 4 + 1][r] = ra[i].y;
      As[buf][a_kq * 4 + 2][r] = ra[i].z;
      As[buf][a_kq * 4 + 3][r] = ra[i].w;
    }
    Ws[buf][a_kq * 4 + 0][a_row0] = rw.x;
    Ws[buf][a_kq * 4 + 1][a_row0] = rw.y;
    Ws[buf][a_kq * 4 + 2][a_row0] = rw.z;
    Ws[buf][a_kq * 4 + 3][a_row0] = rw.w;
  };

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int nk = p.K / LF_BK;
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int kb = 0; kb < nk; ++kb) {
    const int buf = kb & 1;
    if (kb + 1 < nk) load_tiles((kb + 1) * LF_BK);
#pragma unroll
    for (int k = 0; k < LF_BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
      float4 b = *reinterpret_cast<const float4*>(&Ws[buf][k][tx * 4]);
      float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kb + 1 < nk) {
      store_tiles(buf ^ 1);
      __syncthreads();
    }
  }

  const int n = n0 + tx * 4;
  if (n >= p.N) return;
  float4 bv = bias ? *reinterpret_cast<const float4*>(bias + n) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int r = m0 + ty * 8 + i;
    if (r >= p.M) continue;
    float4 v;
    v.x = apply_act(acc[i][0] + bv.x, p.act);
    v.y = apply_act(acc[i][1] + bv.y, p.act);
    v.z = apply_act(acc[i][2] + bv.z, p.act);
    v.w = apply_act(acc[i][3] + bv.w, p.act);
    if (R) {
      float4 rr = *reinterpret_cast<const float4*>(R + (long long)r * p.ldr + n);
      v.x += rr.x; v.y += rr.y; v.z += rr.z; v.w += rr.w;
    }
    *reinterpret_cast<float4*>(C + (long long)r * p.ldc + n) = v;
  }
}

inline int launch_linear_f32(const LinearArgs& a, cudaStream_t s) {
  if (a.M <= 0) return 0;
  if (a.K % LF_BK || a.N % 4 || a.lda % 4 || a.ldc % 4 || (a.R && a.ldr % 4))
    return set_error(-1, "linear_f32: K%16, N%4, ld%4 required");
  dim3 grid(cdiv(a.M, LF_BM), cdiv(a.N, LF_BN), a.nz > 0 ? a.nz : 1);
  LaunchScope ls(KC_LINEAR, s);
  linear_f32_kernel<<<grid, LF_THREADS, 0, s>>>(a);
  LTR_CUDA_TRY(cudaGetLastError());
  return 0;
}

}  // namespace ltr
