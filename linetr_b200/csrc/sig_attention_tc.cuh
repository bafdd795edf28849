// Line-signature attention on tensor cores (sm_100a): softmax(q k^T) v over the L_i lines of ONE
// image, per head, no mask (attention(), models/line_transformer.py:132-136; 1/8 is folded into
// W_q and the heads are head-major, see ltr_create).
//
// CTA = (128-query tile, head, image), 256 threads: thread pair <-> query row (TMEM lane), the two
// threads of a pair own the two halves of the columns.  Per key tile of 128 lines:
//   * q, k, v arrive as the split-bf16 tile image the qkv GEMM epilogue wrote (k-block h = q of head h,
//     4 + h = k, 8 + h = v).  Images of the batch are not 128-row aligned in the row space, so the
//     operand tiles are assembled from 16-byte chunks with cp.async (re-swizzled for the new row
//     phase, rows past the image zero-filled) - no conversion, no registers, all copies in flight;
//     v lands behind the S MMA.  V is staged exactly like K ([keys x 64 dims]) and consumed as an
//     MN-major B operand - no transpose anywhere.
//   * S = Q K^T      tcgen05 (M128 N128 K64, 3 split products), accumulator in TMEM
//   * p = exp(s - m), row sums in registers; P written as the A operand of the second MMA
//     (re-using the Q/K shared memory, which is dead once S is complete)
//   * O += P V       tcgen05 (M128 N64 K128), accumulated in TMEM across key tiles
// Images with more than 128 lines run a first pass that only computes the row maxima (S is
// recomputed in the second pass - QK^T is 1/3 of the work and far from the bottleneck), so no
// rescaling of the TMEM accumulator is ever needed.  Output: columns [out_k0, out_k0+256) of a split-bf16
// activation image (the second half of the signature MLP input; the `merge` projection is folded into W1).
#pragma once
#include "act_img.cuh"
#include "common.cuh"
#include "ptx_sm100.cuh"

namespace ltr {

struct SigAttnSmem {
  static constexpr int QK = 64 * 1024;   // Q hi/lo + K hi/lo (4 x 16 KB); later P hi/lo (2 x 32 KB)
  static constexpr int V = 32 * 1024;    // V hi/lo: [128 keys x 64 d] each
  static constexpr int OFF_RED = QK + V;             // float [2][128] pair reduction scratch
  static constexpr int OFF_BAR = OFF_RED + 2 * 128 * 4;
  static constexpr int TOTAL = OFF_BAR + 64 + 1024;
};

__global__ void __launch_bounds__(256) sig_attention_tc_kernel(ActImg qkv, ActImg out, int out_k0,
                                                                const int* __restrict__ cu, int lpi) {
  using S = SigAttnSmem;
  // Uniform batches know their line range without touching memory; for var-len batches `cu` is read
  // only AFTER griddepcontrol.wait (nothing a preceding kernel on the stream wrote is guaranteed
  // visible before it - several kernels of the chain can be resident ahead of their wait at once).
  int lb = blockIdx.z * lpi, le = lb + lpi;
  int L = lpi;
  const int q0 = blockIdx.x * 128;
  if (!cu && q0 >= L) return;
  const int h = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qd = warp & 3, half = warp >> 2;
  const int row = qd * 32 + lane;   // query row of this thread (shared with its pair thread)
  if (tid == 0) LTR_DBG_STAMP(30);
  pdl_launch_dependents();

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint8_t* q_hi = smem;                 // [128 x 64]
  uint8_t* q_lo = smem + 16384;
  uint8_t* k_hi = smem + 32768;         // [128 keys x 64]
  uint8_t* k_lo = smem + 49152;
  uint8_t* p_hi = smem;                 // [2 k-blocks][128 x 64 keys]  (aliases q/k)
  uint8_t* p_lo = smem + 32768;
  uint8_t* v_hi = smem + S::QK;         // [128 keys x 64 d]
  uint8_t* v_lo = v_hi + 16384;
  float* red = reinterpret_cast<float*>(smem + S::OFF_RED);
  uint64_t* mma_done = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);
  uint64_t* ld_qk = mma_done + 1;      // TMA path: q and k tiles landed
  uint64_t* ld_v = mma_done + 2;       // TMA path: v tile landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_done + 3);

  if (tid == 0) {
    ptx::mbar_init(mma_done, 1);
    ptx::mbar_init(ld_qk, 1);
    ptx::mbar_init(ld_v, 1);
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
  const uint32_t lane_addr = (uint32_t)(qd * 32) << 16;
  uint32_t phase = 0;
  pdl_wait();   // qkv of this layer is complete; the previous reader of the output image is done
  if (tid == 0) LTR_DBG_STAMP(31);
  if (cu) {
    lb = cu[blockIdx.z]; le = cu[blockIdx.z + 1];
    L = le - lb;
    if (q0 >= L) {   // uniform per CTA: give the TMEM columns back and leave
      ptx::tc_fence_before();
      __syncthreads();
      if (warp == 0) ptx::tmem_dealloc(tmem_base, 256);
      return;
    }
  }

  // [128 rows x 64] operand tile (hi and lo plane) <- rows row0.. of this image, k-block kb of the qkv image:
  // chunk f = tid + 256 i -> row f/8, 16-byte chunk f%8 (8 lanes read one 128-byte image line)
  auto stage_tile = [&](int row0, int kb, uint8_t* hi, uint8_t* lo) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int f = tid + 256 * i, r = f >> 3, c = f & 7;
      const bool ok = row0 + r < L;
      const int gr = lb + row0 + r;
      const size_t src = ok ? ((size_t)(gr >> 7) * qkv.kblocks + kb) * (IMG_TILE_ELEMS * 2) + ptx::sw128_offset(gr & 127, c * 8) : 0;
      const uint32_t dst = ptx::sw128_offset(r, c * 8);
      ptx::cp_async16(hi + dst, reinterpret_cast<const uint8_t*>(qkv.hi) + src, ok ? 16u : 0u);
      ptx::cp_async16(lo + dst, reinterpret_cast<const uint8_t*>(qkv.lo) + src, ok ? 16u : 0u);
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
  // row maximum over this thread's 64 columns of S, combined with the pair thread through smem
  auto row_max = [&](int kn, float m_in) {
    float m = m_in;
#pragma unroll 1
    for (int c0 = half * 64; c0 < half * 64 + 64; c0 += 32) {
      float s[32];
      ptx::tmem_ld32(t_s + lane_addr + c0, s);
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (c0 + j < kn) m = fmaxf(m, s[j]);
    }
    red[half * 128 + row] = m;
    __syncthreads();
    m = fmaxf(red[row], red[128 + row]);
    __syncthreads();
    return m;
  };

  const int n_kt = (L + 127) / 128;
  // An image of exactly 128 lines that starts on a 128-row tile of the qkv image (every uniform L = 128 batch): its
  // q, k, v operand tiles ARE tiles of that image - six 16 KB bulk copies (TMA) by one thread instead of 24 cp.async
  // per thread; v still lands behind the S MMA.  Single key tile: the shared memory is written exactly once.
  const bool tma = L == 128 && ((lb & 127) == 0);
  float m = -INFINITY;
  if (n_kt > 1) {
    // ---- pass 1: row maxima over all key tiles
    for (int kt = 0; kt < n_kt; ++kt) {
      const int k0 = kt * 128, kn = min(128, L - k0);
      stage_tile(q0, h, q_hi, q_lo);
      stage_tile(k0, 4 + h, k_hi, k_lo);
      ptx::cp_async_wait_all();
      ptx::fence_proxy_async_smem();
      __syncthreads();
      if (tid == 0) { ptx::tc_fence_after(); issue_s(); }
      wait_mma();
      m = row_max(kn, m);
      ptx::tc_fence_before();
      __syncthreads();
    }
  }
  float l = 0.f;
  constexpr float LOG2E = 1.4426950408889634f;
  for (int kt = 0; kt < n_kt; ++kt) {
    const int k0 = kt * 128, kn = min(128, L - k0);
    if (tid == 0) LTR_DBG_STAMP(32);
    if (tma) {
      if (tid == 0) {
        const size_t t0 = (size_t)(lb >> 7) * qkv.kblocks * IMG_TILE_ELEMS;
        const size_t tq = t0 + (size_t)h * IMG_TILE_ELEMS, tk = t0 + (size_t)(4 + h) * IMG_TILE_ELEMS, tv = t0 + (size_t)(8 + h) * IMG_TILE_ELEMS;
        ptx::mbar_arrive_expect_tx(ld_qk, 4 * 16384);
        ptx::bulk_g2s(q_hi, qkv.hi + tq, 16384, ld_qk);
        ptx::bulk_g2s(q_lo, qkv.lo + tq, 16384, ld_qk);
        ptx::bulk_g2s(k_hi, qkv.hi + tk, 16384, ld_qk);
        ptx::bulk_g2s(k_lo, qkv.lo + tk, 16384, ld_qk);
        ptx::mbar_arrive_expect_tx(ld_v, 2 * 16384);
        ptx::bulk_g2s(v_hi, qkv.hi + tv, 16384, ld_v);
        ptx::bulk_g2s(v_lo, qkv.lo + tv, 16384, ld_v);
      }
      if (tid == 0) LTR_DBG_STAMP(33);
      ptx::mbar_wait(ld_qk, 0);
      if (tid == 0) LTR_DBG_STAMP(34);
    } else {
      stage_tile(q0, h, q_hi, q_lo);
      stage_tile(k0, 4 + h, k_hi, k_lo);
      ptx::cp_async_commit();
      stage_tile(k0, 8 + h, v_hi, v_lo);
      ptx::cp_async_commit();
      if (tid == 0) LTR_DBG_STAMP(33);
      ptx::cp_async_wait_group<1>();   // q and k have landed; v is still in flight behind the S MMA
      if (tid == 0) LTR_DBG_STAMP(34);
      ptx::fence_proxy_async_smem();
    }
    __syncthreads();
    if (tid == 0) { ptx::tc_fence_after(); issue_s(); }
    wait_mma();
    if (tid == 0) LTR_DBG_STAMP(35);
    if (n_kt == 1) m = row_max(kn, m);
    // p = exp(s - m) -> P operand (q/k shared memory is dead: the S MMAs have completed);
    // this thread's 64 key columns are exactly k-block `half` of the P tile
    const float mb = m * LOG2E;
#pragma unroll 1
    for (int c0 = half * 64; c0 < half * 64 + 64; c0 += 32) {
      float s[32];
      ptx::tmem_ld32(t_s + lane_addr + c0, s);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        s[j] = (c0 + j < kn) ? ptx::ex2_approx(fmaf(s[j], LOG2E, -mb)) : 0.f;
        l += s[j];
      }
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint4 hh, ll;
        ptx::split8_bf16(&s[j], hh, ll);
        const uint32_t off = half * 16384 + ptx::sw128_offset(row, (c0 & 63) + j);
        *reinterpret_cast<uint4*>(p_hi + off) = hh;
        *reinterpret_cast<uint4*>(p_lo + off) = ll;
      }
    }
    if (tma) ptx::mbar_wait(ld_v, 0);
    else ptx::cp_async_wait_group<0>();   // v
    ptx::tc_fence_before();
    ptx::fence_proxy_async_smem();
    if (tid == 0) LTR_DBG_STAMP(36);
    __syncthreads();
    if (tid == 0) {   // O += P V    (B = V as an MN-major operand: [K = keys][N = 64 dims])
      ptx::tc_fence_after();
      constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(128, 64) | (1u << 16);
      const uint32_t ph = ptx::smem_u32(p_hi), pl = ptx::smem_u32(p_lo), vh = ptx::smem_u32(v_hi), vl = ptx::smem_u32(v_lo);
#pragma unroll
      for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
        for (int k16 = 0; k16 < 4; ++k16) {
          const uint32_t pa = kb * 16384 + k16 * 32;          // A: 16 keys = 32 bytes along K
          const uint32_t va = (kb * 4 + k16) * 2048;          // B: 16 key rows of 128 bytes
          const uint32_t first = (kt | kb | k16) == 0 ? 0u : 1u;
          ptx::umma_bf16(t_o, ptx::make_sw128_kmajor_desc(pl + pa, 1024), ptx::make_sw128_mnmajor_desc(vh + va, 1024, 1024), idesc, first);
          ptx::umma_bf16(t_o, ptx::make_sw128_kmajor_desc(ph + pa, 1024), ptx::make_sw128_mnmajor_desc(vl + va, 1024, 1024), idesc, 1);
          ptx::umma_bf16(t_o, ptx::make_sw128_kmajor_desc(ph + pa, 1024), ptx::make_sw128_mnmajor_desc(vh + va, 1024, 1024), idesc, 1);
        }
      }
      ptx::umma_commit(mma_done);
    }
    wait_mma();   // P, V shared memory may be overwritten by the next key tile
    if (tid == 0) LTR_DBG_STAMP(37);
  }
  // ---- epilogue: o / l -> image (this thread: 32 of the 64 head dims)
  {
    red[half * 128 + row] = l;
    __syncthreads();
    const float inv = 1.f / (red[row] + red[128 + row]);
    const bool live = q0 + row < L;
    float o[32];
    ptx::tmem_ld32(t_o + lane_addr + half * 32, o);
    if (tma) {
      // aligned 128-line image: this CTA's output (128 rows x 64 head dims) is ONE k-block tile of the output image -
      // stage hi / lo planes in tile layout (the P operand's memory is dead: the PV MMAs have completed) and store each
      // with one 16 KB bulk copy instead of eight scattered 16-byte stores per thread
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        const float v[8] = {o[j] * inv, o[j + 1] * inv, o[j + 2] * inv, o[j + 3] * inv,
                            o[j + 4] * inv, o[j + 5] * inv, o[j + 6] * inv, o[j + 7] * inv};
        uint4 hh, ll;
        ptx::split8_bf16(v, hh, ll);
        const uint32_t off = ptx::sw128_offset(row, half * 32 + j);
        *reinterpret_cast<uint4*>(smem + off) = hh;
        *reinterpret_cast<uint4*>(smem + 16384 + off) = ll;
      }
      ptx::fence_proxy_async_smem();
      __syncthreads();
      if (tid == 0) {
        const size_t toff = ((size_t)(lb >> 7) * out.kblocks + (out_k0 >> 6) + h) * IMG_TILE_ELEMS;
        ptx::bulk_s2g(out.hi + toff, smem, 16384);
        ptx::bulk_s2g(out.lo + toff, smem + 16384, 16384);
        ptx::bulk_commit();
        // the copies only have to be done READING this CTA's shared memory before it exits; their global writes are
        // complete (and visible to the next kernel behind griddepcontrol.wait) when the grid is
        ptx::bulk_wait_read_all();
      }
    } else if (live) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        const float v[8] = {o[j] * inv, o[j + 1] * inv, o[j + 2] * inv, o[j + 3] * inv,
                            o[j + 4] * inv, o[j + 5] * inv, o[j + 6] * inv, o[j + 7] * inv};
        img_store8(out, lb + q0 + row, out_k0 + h * 64 + half * 32 + j, v);
      }
    }
  }
  if (tid == 0) LTR_DBG_STAMP(38);
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem_base, 256);
}

inline int launch_sig_attention_tc(ActImg qkv, ActImg out, int out_k0, const int* cu, int lpi, int max_l,
                                   int n_images, cudaStream_t s) {
  if (max_l <= 0 || n_images <= 0) return 0;
  LTR_CUDA_TRY(ensure_dynamic_smem(sig_attention_tc_kernel, SigAttnSmem::TOTAL));
  dim3 grid(cdiv(max_l, 128), 4, n_images);
  LaunchScope ls(KC_SIG_ATTN, s);
  LTR_CUDA_TRY(launch_pdl(sig_attention_tc_kernel, grid, dim3(256), SigAttnSmem::TOTAL, s, qkv, out, out_k0, cu, lpi));
  return 0;
}

}  // namespace ltr
