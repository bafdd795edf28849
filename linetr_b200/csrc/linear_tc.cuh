// Tensor-core GEMM engine for sm_100a: Y = act(X W^T + b) (+ R) with fp32 in/out and
// split-bf16 operands on tcgen05 (UMMA), accumulators in TMEM.
//
// Precision: every fp32 operand is split x = hi + lo (both bf16) and three MMAs are issued per
// k-step, hi*hi + lo*hi + hi*lo, accumulated in fp32 - ~2^-17 relative operand error instead of
// bf16's 2^-9 (single-pass TF32 already misses the 1e-3 descriptor bar).
//
// CTA = one 128 x BN output tile.  Warp roles:
//   warps 0-3  A producers: fp32 global -> registers -> hi/lo bf16 -> 128B-swizzled K-major
//              smem tiles (generic proxy + fence.proxy.async), then the epilogue
//              (tcgen05.ld -> bias/activation/residual -> global)
//   warp 4     TMEM allocation + MMA issue (one elected lane, tcgen05.mma / tcgen05.commit)
//   warp 5     weight loader: cp.async.bulk (TMA) of pre-swizzled bf16 hi/lo weight tiles
// mbarrier pipeline over K blocks of 64: full_a / full_w (producers -> MMA), empty (MMA ->
// producers, via tcgen05.commit), acc_full (MMA -> epilogue).
#pragma once
#include "common.cuh"
#include "linear_f32.cuh"
#include "ptx_sm100.cuh"

namespace ltr {

// W [N, K] packed for the engine: bf16 hi and lo images, tiled as [K/64][N/8] atoms of
// 8 rows x 64 k (1024 bytes, 128B swizzle) so that any (n0, kb) tile of BN rows is one
// contiguous BN*128-byte range (one bulk copy).
struct TcWeight {
  const __nv_bfloat16* hi = nullptr;
  const __nv_bfloat16* lo = nullptr;
  int N = 0, K = 0;
};

struct TcArgs {
  const float* A; int lda;
  TcWeight W;
  const float* bias;
  const float* R; int ldr;
  float* C; int ldc;
  int M, act;
  // z-batching (blockIdx.z): element strides of A / bias / R / C, packed-weight stride in elements
  long long sA, sW, sB, sR, sC;
};

template <int BN>
struct TcCfg {
  static constexpr int BM = 128, BK = 64;
  static constexpr int A_TILE = BM * BK * 2;  // bytes of one bf16 A tile (hi or lo)
  static constexpr int W_TILE = BN * BK * 2;
  static constexpr int STAGE = 2 * A_TILE + 2 * W_TILE;
  static constexpr int STAGES = BN >= 256 ? 2 : (BN >= 128 ? 3 : 4);
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM = STAGES * STAGE + BAR_BYTES + 1024;  // + alignment slack
  static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
  static constexpr int THREADS = 192;
};

template <int BN>
__global__ void __launch_bounds__(192, 1) linear_tc_kernel(TcArgs p) {
  using Cfg = TcCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE);
  uint64_t* full_a = bars;
  uint64_t* full_w = bars + Cfg::STAGES;
  uint64_t* empty = bars + 2 * Cfg::STAGES;
  uint64_t* acc_full = bars + 3 * Cfg::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * Cfg::STAGES + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int z = blockIdx.z;
  const int m0 = blockIdx.x * Cfg::BM, n0 = blockIdx.y * BN;
  const int K = p.W.K, N = p.W.N;
  const int nk = K / Cfg::BK;

  if (warp == 5 && lane == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      ptx::mbar_init(&full_a[s], 128);
      ptx::mbar_init(&full_w[s], 1);
      ptx::mbar_init(&empty[s], 1);
    }
    ptx::mbar_init(acc_full, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 4) {
    ptx::tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // ------------------------------------------------------------ A producers
    const float* __restrict__ A = p.A + z * p.sA;
    const int c = tid & 7;    // 8-float chunk of the 64-wide K block (one 16-byte bf16 chunk)
    const int r0 = tid >> 3;  // rows r0 + 16 i
    for (int kb = 0; kb < nk; ++kb) {
      const int s = kb % Cfg::STAGES;
      const uint32_t ph = (kb / Cfg::STAGES) & 1;
      float4 v[8][2];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = m0 + r0 + 16 * i;
        if (r < p.M) {
          const float4* src = reinterpret_cast<const float4*>(A + (long long)r * p.lda + kb * Cfg::BK + c * 8);
          v[i][0] = __ldg(src);
          v[i][1] = __ldg(src + 1);
        } else {
          v[i][0] = make_float4(0.f, 0.f, 0.f, 0.f);
          v[i][1] = v[i][0];
        }
      }
      ptx::mbar_wait(&empty[s], ph ^ 1);
      uint8_t* a_hi = smem + s * Cfg::STAGE;
      uint8_t* a_lo = a_hi + Cfg::A_TILE;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = r0 + 16 * i;
        const float x[8] = {v[i][0].x, v[i][0].y, v[i][0].z, v[i][0].w, v[i][1].x, v[i][1].y, v[i][1].z, v[i][1].w};
        __nv_bfloat16 h[8], l[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) ptx::split_bf16(x[j], h[j], l[j]);
        const uint32_t off = ptx::sw128_offset(r, c * 8);
        *reinterpret_cast<uint4*>(a_hi + off) =
            make_uint4(ptx::pack_bf16(h[0], h[1]), ptx::pack_bf16(h[2], h[3]), ptx::pack_bf16(h[4], h[5]), ptx::pack_bf16(h[6], h[7]));
        *reinterpret_cast<uint4*>(a_lo + off) =
            make_uint4(ptx::pack_bf16(l[0], l[1]), ptx::pack_bf16(l[2], l[3]), ptx::pack_bf16(l[4], l[5]), ptx::pack_bf16(l[6], l[7]));
      }
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(&full_a[s]);
    }
    // ------------------------------------------------------------ epilogue
    ptx::mbar_wait(acc_full, 0);
    ptx::tc_fence_after();
    const int row = m0 + warp * 32 + lane;
    const float* __restrict__ bias = p.bias ? p.bias + z * p.sB : nullptr;
    const float* __restrict__ R = p.R ? p.R + z * p.sR : nullptr;
    float* __restrict__ C = p.C + z * p.sC;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      float acc[32];
      ptx::tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, acc);
      if (row < p.M) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const int n = n0 + c0 + j;
          float4 b = bias ? *reinterpret_cast<const float4*>(bias + n) : make_float4(0.f, 0.f, 0.f, 0.f);
          float4 o;
          o.x = apply_act(acc[j] + b.x, p.act);
          o.y = apply_act(acc[j + 1] + b.y, p.act);
          o.z = apply_act(acc[j + 2] + b.z, p.act);
          o.w = apply_act(acc[j + 3] + b.w, p.act);
          if (R) {
            float4 rr = *reinterpret_cast<const float4*>(R + (long long)row * p.ldr + n);
            o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
          }
          *reinterpret_cast<float4*>(C + (long long)row * p.ldc + n) = o;
        }
      }
    }
  } else if (warp == 4) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(128, BN);
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % Cfg::STAGES;
        const uint32_t ph = (kb / Cfg::STAGES) & 1;
        ptx::mbar_wait(&full_a[s], ph);
        ptx::mbar_wait(&full_w[s], ph);
        ptx::tc_fence_after();
        const uint32_t a_hi = ptx::smem_u32(smem + s * Cfg::STAGE);
        const uint32_t a_lo = a_hi + Cfg::A_TILE;
        const uint32_t w_hi = a_hi + 2 * Cfg::A_TILE;
        const uint32_t w_lo = w_hi + Cfg::W_TILE;
#pragma unroll
        for (int k16 = 0; k16 < Cfg::BK / 16; ++k16) {
          const uint32_t ko = k16 * 32;  // 16 bf16 = 32 bytes along K inside the swizzle atom
          const uint64_t dah = ptx::make_sw128_kmajor_desc(a_hi + ko, 1024);
          const uint64_t dal = ptx::make_sw128_kmajor_desc(a_lo + ko, 1024);
          const uint64_t dwh = ptx::make_sw128_kmajor_desc(w_hi + ko, 1024);
          const uint64_t dwl = ptx::make_sw128_kmajor_desc(w_lo + ko, 1024);
          ptx::umma_bf16(tmem_base, dal, dwh, idesc, (kb | k16) != 0);
          ptx::umma_bf16(tmem_base, dah, dwl, idesc, 1);
          ptx::umma_bf16(tmem_base, dah, dwh, idesc, 1);
        }
        ptx::umma_commit(&empty[s]);  // frees the stage once the MMAs above have read it
      }
      ptx::umma_commit(acc_full);
    }
  } else {
    // ------------------------------------------------------------ weight loader (TMA bulk)
    if (lane == 0) {
      const uint8_t* whi = reinterpret_cast<const uint8_t*>(p.W.hi + z * p.sW);
      const uint8_t* wlo = reinterpret_cast<const uint8_t*>(p.W.lo + z * p.sW);
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % Cfg::STAGES;
        const uint32_t ph = (kb / Cfg::STAGES) & 1;
        ptx::mbar_wait(&empty[s], ph ^ 1);
        uint8_t* w_hi = smem + s * Cfg::STAGE + 2 * Cfg::A_TILE;
        const size_t off = ((size_t)kb * (N / 8) + n0 / 8) * 1024;
        ptx::mbar_arrive_expect_tx(&full_w[s], 2 * Cfg::W_TILE);
        ptx::bulk_g2s(w_hi, whi + off, Cfg::W_TILE, &full_w[s]);
        ptx::bulk_g2s(w_hi + Cfg::W_TILE, wlo + off, Cfg::W_TILE, &full_w[s]);
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 4) ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

template <int BN>
static int launch_linear_tc_bn(const TcArgs& a, int nz, cudaStream_t s) {
  using Cfg = TcCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    LTR_CUDA_TRY(cudaFuncSetAttribute(linear_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr_set = true;
  }
  dim3 grid(cdiv(a.M, Cfg::BM), a.W.N / BN, nz);
  LaunchScope ls(KC_LINEAR, s);
  linear_tc_kernel<BN><<<grid, Cfg::THREADS, Cfg::SMEM, s>>>(a);
  LTR_CUDA_TRY(cudaGetLastError());
  return 0;
}

inline int launch_linear_tc(const TcArgs& a, int nz, cudaStream_t s) {
  if (a.M <= 0) return 0;
  if (a.W.K % 64 || a.W.N % 64 || a.lda % 4 || a.ldc % 4 || (a.R && a.ldr % 4))
    return set_error(-1, "linear_tc: K%64, N%64, ld%4 required");
  if (nz < 1) nz = 1;
  if (a.W.N % 256 == 0) return launch_linear_tc_bn<256>(a, nz, s);
  if (a.W.N % 128 == 0) return launch_linear_tc_bn<128>(a, nz, s);
  return launch_linear_tc_bn<64>(a, nz, s);
}

// ---------------------------------------------------------------- host-side weight packing
inline uint16_t f32_to_bf16_rn(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
inline float bf16_to_f32(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
// W: [N, K] row-major (double).  hi/lo: N*K uint16 each, engine tile layout.
inline void pack_tc_weight(const double* W, int N, int K, uint16_t* hi, uint16_t* lo) {
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      const float w = (float)W[(size_t)n * K + k];
      const uint16_t h = f32_to_bf16_rn(w);
      const uint16_t l = f32_to_bf16_rn(w - bf16_to_f32(h));
      const size_t atom = (size_t)(k / 64) * (N / 8) + n / 8;
      const size_t idx = atom * 512 + ptx::sw128_offset(n & 7, k & 63) / 2;
      hi[idx] = h;
      lo[idx] = l;
    }
}

}  // namespace ltr
