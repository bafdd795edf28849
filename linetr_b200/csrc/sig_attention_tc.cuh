// Line-signature attention on tensor cores (sm_100a): softmax(q k^T) v over the L_i lines of ONE
// image, per head, no mask (attention(), models/line_transformer.py:132-136; 1/8 is folded into
// W_q and the heads are head-major, see ltr_create).
//
// CTA = (128-query tile, head, image), 128 threads, thread t <-> query row t <-> TMEM lane t.
// Per key tile of 128 lines:
//   * q, k, v fp32 rows are gathered from the qkv buffer, split into bf16 hi/lo and written as
//     K-major SWIZZLE_128B operand tiles (v transposed: [64 dims x 128 keys]); images are not
//     tile aligned in the row space, hence the CUDA-core gather instead of TMA
//   * S = Q K^T      tcgen05 (M128 N128 K64, 3 split products), accumulator in TMEM
//   * p = exp(s - m), row sums in registers; P written as the A operand of the second MMA
//     (re-using the Q/K shared memory, which is dead once S is complete)
//   * O += P V       tcgen05 (M128 N64 K128), accumulated in TMEM across key tiles
// Images with more than 128 lines run a first pass that only computes the row maxima (S is
// recomputed in the second pass - QK^T is 1/3 of the work and far from the bottleneck), so no
// rescaling of the TMEM accumulator is ever needed.  Output: split-bf16 activation image [R, 256].
#pragma once
#include "act_img.cuh"
#include "common.cuh"
#include "ptx_sm100.cuh"

namespace ltr {

struct SigAttnSmem {
  static constexpr int QK = 64 * 1024;   // Q hi/lo + K hi/lo (4 x 16 KB); later P hi/lo (2 x 32 KB)
  static constexpr int VT = 32 * 1024;   // V^T hi/lo: 2 planes x [2 k-blocks][64 d x 64 keys]
  static constexpr int OFF_BAR = QK + VT;
  static constexpr int TOTAL = OFF_BAR + 64 + 1024;
};

__global__ void __launch_bounds__(128) sig_attention_tc_kernel(const float* __restrict__ qkv, ActImg out,
                                                                const int* __restrict__ cu, int lpi) {
  using S = SigAttnSmem;
  int lb, le;
  image_range(cu, lpi, blockIdx.z, lb, le);
  const int L = le - lb;
  const int q0 = blockIdx.x * 128;
  if (q0 >= L) return;
  const int h = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint8_t* q_hi = smem;                 // [128 x 64]
  uint8_t* q_lo = smem + 16384;
  uint8_t* k_hi = smem + 32768;         // [128 keys x 64]
  uint8_t* k_lo = smem + 49152;
  uint8_t* p_hi = smem;                 // [2 k-blocks][128 x 64 keys]  (aliases q/k)
  uint8_t* p_lo = smem + 32768;
  uint8_t* vt_hi = smem + S::QK;        // [2 k-blocks][64 d x 64 keys]
  uint8_t* vt_lo = vt_hi + 16384;
  uint64_t* mma_done = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_done + 1);

  if (tid == 0) {
    ptx::mbar_init(mma_done, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(tmem_slot, 256);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t t_s = tmem_base;          // S: 128 columns
  const uint32_t t_o = tmem_base + 128;    // O: 64 columns
  const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
  uint32_t phase = 0;

  // gather 64 fp32 of row `grow` (or zeros) -> bf16 hi/lo row `r` of a [128 x 64] operand tile
  auto stage_row = [&](uint8_t* hi, uint8_t* lo, int r, const float* src, bool live) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float v[8];
      if (live) {
        const float4 a = *reinterpret_cast<const float4*>(src + c * 8);
        const float4 b = *reinterpret_cast<const float4*>(src + c * 8 + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = 0.f;
      }
      __nv_bfloat16 hh[8], ll[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) ptx::split_bf16(v[e], hh[e], ll[e]);
      const uint32_t off = ptx::sw128_offset(r, c * 8);
      *reinterpret_cast<uint4*>(hi + off) =
          make_uint4(ptx::pack_bf16(hh[0], hh[1]), ptx::pack_bf16(hh[2], hh[3]), ptx::pack_bf16(hh[4], hh[5]), ptx::pack_bf16(hh[6], hh[7]));
      *reinterpret_cast<uint4*>(lo + off) =
          make_uint4(ptx::pack_bf16(ll[0], ll[1]), ptx::pack_bf16(ll[2], ll[3]), ptx::pack_bf16(ll[4], ll[5]), ptx::pack_bf16(ll[6], ll[7]));
    }
  };
  auto stage_qk = [&](int k0) {
    const bool ql = q0 + tid < L, kl = k0 + tid < L;
    stage_row(q_hi, q_lo, tid, qkv + (long long)(lb + q0 + tid) * 768 + h * 64, ql);
    stage_row(k_hi, k_lo, tid, qkv + (long long)(lb + k0 + tid) * 768 + 256 + h * 64, kl);
  };
  auto stage_vt = [&](int k0) {   // thread = key j: V^T[d][j]
    const bool kl = k0 + tid < L;
    const float* src = qkv + (long long)(lb + k0 + tid) * 768 + 512 + h * 64;
    const uint32_t kb_off = (tid >> 6) * 8192;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      if (kl) a = *reinterpret_cast<const float4*>(src + c * 4);
      const float v[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        __nv_bfloat16 hh, ll;
        ptx::split_bf16(v[e], hh, ll);
        const uint32_t off = kb_off + ptx::sw128_offset(c * 4 + e, tid & 63);
        *reinterpret_cast<__nv_bfloat16*>(vt_hi + off) = hh;
        *reinterpret_cast<__nv_bfloat16*>(vt_lo + off) = ll;
      }
    }
  };
  auto issue_s = [&]() {   // S = Q K^T
    constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(128, 128);
    const uint32_t qh = ptx::smem_u32(q_hi), ql = ptx::smem_u32(q_lo), kh = ptx::smem_u32(k_hi), kl = ptx::smem_u32(k_lo);
#pragma unroll
    for (int k16 = 0; k16 < 4; ++k16) {
      const uint32_t ko = k16 * 32;
      ptx::umma_bf16(t_s, ptx::make_sw128_kmajor_desc(ql + ko, 1024), ptx::make_sw128_kmajor_desc(kh + ko, 1024), idesc, k16 != 0);
      ptx::umma_bf16(t_s, ptx::make_sw128_kmajor_desc(qh + ko, 1024), ptx::make_sw128_kmajor_desc(kl + ko, 1024), idesc, 1);
      ptx::umma_bf16(t_s, ptx::make_sw128_kmajor_desc(qh + ko, 1024), ptx::make_sw128_kmajor_desc(kh + ko, 1024), idesc, 1);
    }
    ptx::umma_commit(mma_done);
  };
  auto wait_mma = [&]() {
    ptx::mbar_wait(mma_done, phase & 1);
    ++phase;
    ptx::tc_fence_after();
  };

  const int n_kt = (L + 127) / 128;
  float m = -INFINITY;
  if (n_kt > 1) {
    // ---- pass 1: row maxima over all key tiles
    for (int kt = 0; kt < n_kt; ++kt) {
      const int k0 = kt * 128, kn = min(128, L - k0);
      stage_qk(k0);
      ptx::fence_proxy_async_smem();
      __syncthreads();
      if (tid == 0) { ptx::tc_fence_after(); issue_s(); }
      wait_mma();
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 32) {
        float s[32];
        ptx::tmem_ld32(t_s + lane_addr + c0, s);
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (c0 + j < kn) m = fmaxf(m, s[j]);
      }
      ptx::tc_fence_before();
      __syncthreads();
    }
  }
  float l = 0.f;
  for (int kt = 0; kt < n_kt; ++kt) {
    const int k0 = kt * 128, kn = min(128, L - k0);
    stage_qk(k0);
    stage_vt(k0);
    ptx::fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) { ptx::tc_fence_after(); issue_s(); }
    wait_mma();
    if (n_kt == 1) {
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 32) {
        float s[32];
        ptx::tmem_ld32(t_s + lane_addr + c0, s);
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (c0 + j < kn) m = fmaxf(m, s[j]);
      }
    }
    // p = exp(s - m) -> P operand (q/k shared memory is dead: the S MMAs have completed)
#pragma unroll 1
    for (int c0 = 0; c0 < 128; c0 += 32) {
      float s[32];
      ptx::tmem_ld32(t_s + lane_addr + c0, s);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        s[j] = (c0 + j < kn) ? expf(s[j] - m) : 0.f;
        l += s[j];
      }
      const uint32_t kb_off = (c0 >> 6) * 16384;
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        __nv_bfloat16 hh[8], ll[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) ptx::split_bf16(s[j + e], hh[e], ll[e]);
        const uint32_t off = kb_off + ptx::sw128_offset(tid, (c0 & 63) + j);
        *reinterpret_cast<uint4*>(p_hi + off) =
            make_uint4(ptx::pack_bf16(hh[0], hh[1]), ptx::pack_bf16(hh[2], hh[3]), ptx::pack_bf16(hh[4], hh[5]), ptx::pack_bf16(hh[6], hh[7]));
        *reinterpret_cast<uint4*>(p_lo + off) =
            make_uint4(ptx::pack_bf16(ll[0], ll[1]), ptx::pack_bf16(ll[2], ll[3]), ptx::pack_bf16(ll[4], ll[5]), ptx::pack_bf16(ll[6], ll[7]));
      }
    }
    ptx::tc_fence_before();
    ptx::fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {   // O += P V
      ptx::tc_fence_after();
      constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(128, 64);
      const uint32_t ph = ptx::smem_u32(p_hi), pl = ptx::smem_u32(p_lo), vh = ptx::smem_u32(vt_hi), vl = ptx::smem_u32(vt_lo);
#pragma unroll
      for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
        for (int k16 = 0; k16 < 4; ++k16) {
          const uint32_t pa = kb * 16384 + k16 * 32, va = kb * 8192 + k16 * 32;
          const uint32_t first = (kt | kb | k16) == 0 ? 0u : 1u;
          ptx::umma_bf16(t_o, ptx::make_sw128_kmajor_desc(pl + pa, 1024), ptx::make_sw128_kmajor_desc(vh + va, 1024), idesc, first);
          ptx::umma_bf16(t_o, ptx::make_sw128_kmajor_desc(ph + pa, 1024), ptx::make_sw128_kmajor_desc(vl + va, 1024), idesc, 1);
          ptx::umma_bf16(t_o, ptx::make_sw128_kmajor_desc(ph + pa, 1024), ptx::make_sw128_kmajor_desc(vh + va, 1024), idesc, 1);
        }
      }
      ptx::umma_commit(mma_done);
    }
    wait_mma();   // P, V^T shared memory may be overwritten by the next key tile
  }
  // ---- epilogue: o / l -> image
  {
    const float inv = 1.f / l;
    const bool live = q0 + tid < L;
#pragma unroll 1
    for (int c0 = 0; c0 < 64; c0 += 32) {
      float o[32];
      ptx::tmem_ld32(t_o + lane_addr + c0, o);
      if (live) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          const float v[8] = {o[j] * inv, o[j + 1] * inv, o[j + 2] * inv, o[j + 3] * inv,
                              o[j + 4] * inv, o[j + 5] * inv, o[j + 6] * inv, o[j + 7] * inv};
          img_store8(out, lb + q0 + tid, h * 64 + c0 + j, v);
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem_base, 256);
}

inline int launch_sig_attention_tc(const float* qkv, ActImg out, const int* cu, int lpi, int max_l, int n_images,
                                   cudaStream_t s) {
  if (max_l <= 0 || n_images <= 0) return 0;
  static bool attr_set = false;
  if (!attr_set) {
    LTR_CUDA_TRY(cudaFuncSetAttribute(sig_attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SigAttnSmem::TOTAL));
    attr_set = true;
  }
  dim3 grid(cdiv(max_l, 128), 4, n_images);
  LaunchScope ls(KC_SIG_ATTN, s);
  sig_attention_tc_kernel<<<grid, 128, SigAttnSmem::TOTAL, s>>>(qkv, out, cu, lpi);
  LTR_CUDA_TRY(cudaGetLastError());
  return 0;
}

}  // namespace ltr
