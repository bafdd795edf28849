// Line-signature attention, one CTA per IMAGE with the four heads software-pipelined (sm_100a).
//
// Fast path of attention() (models/line_transformer.py:132-136) for batches of exactly 128 lines per image
// whose images coincide with the 128-row tiles of the qkv image (every uniform L = 128 batch - BASELINE
// cfg[1]).  The general kernel (sig_attention_tc.cuh: any L, var-len batches) spends a CTA per (image, head):
// 16 k cycles of strictly serial staging -> S -> softmax -> PV -> epilogue on 256 threads, 1.73 waves at 2
// CTAs/SM.  Here:
//   * q, k, v of head h ARE tiles (mt = image, k-block h / 4+h / 8+h) of the qkv image the GEMM epilogue wrote:
//     three pairs of 16 KB bulk copies (TMA) by one thread, double-buffered over heads - no gather, no registers;
//   * warp 1 issues S(h) = Q K^T and O(h) = P V on tcgen05 (S and O double-buffered in TMEM, 384 columns) in
//     the order S0 S1 PV0 S2 PV1 S3 PV2 PV3, so the tensor pipe works on head h+1 while head h is in softmax;
//   * 16 softmax / epilogue warps (4 per TMEM lane quarter x 32 key columns): row max and row sum exchanged
//     through shared memory between the 4 column parts, p = exp2(s - m) written as the split-bf16 A operand of
//     PV into the (dead) Q/K stage; epilogue: O / l -> split-bf16 image (second half of the MLP input).
#pragma once
#include "act_img.cuh"
#include "common.cuh"
#include "ptx_sm100.cuh"

namespace ltr {

struct SigImgCfg {
  static constexpr int TILE = 16384;                 // one plane of a 128 x 64 bf16 tile
  static constexpr int STAGE = 6 * TILE;             // q hi/lo, k hi/lo, v hi/lo of one head = 96 KB
  static constexpr int OFF_RED = 2 * STAGE;          // float [2 stages][2 (max, sum)][4 parts][128]
  static constexpr int OFF_BAR = OFF_RED + 2 * 2 * 4 * 128 * 4;
  static constexpr int SMEM = OFF_BAR + 256 + 1024;
  static constexpr int THREADS = 64 + 512;
};
static_assert(SigImgCfg::SMEM <= 232448, "sig_attention_img: shared memory budget");

__device__ __forceinline__ void sig_bar16() { asm volatile("bar.sync 2, 512;" ::: "memory"); }

// qkv: image [R, 768] (k-block h = q of head h (pre-scaled by 1/8), 4 + h = k, 8 + h = v); out: image whose
// k-blocks out_kb0 .. out_kb0+3 receive the four heads' outputs.  grid = n_images (image i = rows 128 i ..).
__global__ void __launch_bounds__(SigImgCfg::THREADS, 1) sig_attention_img_kernel(ActImg qkv, ActImg out, int out_kb0) {
  using Cfg = SigImgCfg;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  float* red = reinterpret_cast<float*>(smem + Cfg::OFF_RED);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* qk_full = bars;          // [2]
  uint64_t* v_full = bars + 2;       // [2]
  uint64_t* s_done = bars + 4;       // [2] S MMAs of the stage complete
  uint64_t* p_ready = bars + 6;      // [2] 16 warps wrote P
  uint64_t* o_done = bars + 8;       // [2] PV MMAs complete (P, V of the stage are dead)
  uint64_t* o_free = bars + 10;      // [2] 16 warps drained O
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int img = blockIdx.x;
  pdl_launch_dependents();
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&qk_full[s], 1);
      ptx::mbar_init(&v_full[s], 1);
      ptx::mbar_init(&s_done[s], 1);
      ptx::mbar_init(&p_ready[s], 16);
      ptx::mbar_init(&o_done[s], 1);
      ptx::mbar_init(&o_free[s], 16);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  auto t_s = [&](int s) { return tmem_base + (uint32_t)s * 128; };          // S[s]: 128 columns
  auto t_o = [&](int s) { return tmem_base + 256 + (uint32_t)s * 64; };     // O[s]: 64 columns
  pdl_wait();   // the qkv image of this layer is complete; the previous reader of `out` is done
  if (tid == 0) LTR_DBG_STAMP(112);

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer: q, k, v tiles of head h
    if (lane == 0) {
      const size_t row_tile = (size_t)img * qkv.kblocks;
      for (int h = 0; h < 4; ++h) {
        const int s = h & 1;
        if (h >= 2) ptx::mbar_wait(&o_done[s], 0);   // PV(h-2) finished reading P (= the Q/K area) and V of this stage
        uint8_t* st = smem + s * Cfg::STAGE;
        const size_t tq = (row_tile + h) * IMG_TILE_ELEMS, tk = (row_tile + 4 + h) * IMG_TILE_ELEMS, tv = (row_tile + 8 + h) * IMG_TILE_ELEMS;
        ptx::mbar_arrive_expect_tx(&qk_full[s], 4 * Cfg::TILE);
        ptx::bulk_g2s(st, qkv.hi + tq, Cfg::TILE, &qk_full[s]);
        ptx::bulk_g2s(st + Cfg::TILE, qkv.lo + tq, Cfg::TILE, &qk_full[s]);
        ptx::bulk_g2s(st + 2 * Cfg::TILE, qkv.hi + tk, Cfg::TILE, &qk_full[s]);
        ptx::bulk_g2s(st + 3 * Cfg::TILE, qkv.lo + tk, Cfg::TILE, &qk_full[s]);
        ptx::mbar_arrive_expect_tx(&v_full[s], 2 * Cfg::TILE);
        ptx::bulk_g2s(st + 4 * Cfg::TILE, qkv.hi + tv, Cfg::TILE, &v_full[s]);
        ptx::bulk_g2s(st + 5 * Cfg::TILE, qkv.lo + tv, Cfg::TILE, &v_full[s]);
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      auto issue_s = [&](int h) {   // S[s] = Q K^T (M128 N128 K64, 3 split products)
        const int s = h & 1;
        ptx::mbar_wait(&qk_full[s], (h >> 1) & 1);
        ptx::tc_fence_after();
        constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(128, 128);
        const uint32_t qh = ptx::smem_u32(smem + s * Cfg::STAGE), ql = qh + Cfg::TILE, kh = qh + 2 * Cfg::TILE, kl = qh + 3 * Cfg::TILE;
#pragma unroll
        for (int k16 = 0; k16 < 4; ++k16) {
          const uint32_t ko = k16 * 32;
          ptx::umma_bf16(t_s(s), ptx::make_sw128_kmajor_desc(ql + ko, 1024), ptx::make_sw128_kmajor_desc(kh + ko, 1024), idesc, k16 != 0);
          ptx::umma_bf16(t_s(s), ptx::make_sw128_kmajor_desc(qh + ko, 1024), ptx::make_sw128_kmajor_desc(kl + ko, 1024), idesc, 1);
          ptx::umma_bf16(t_s(s), ptx::make_sw128_kmajor_desc(qh + ko, 1024), ptx::make_sw128_kmajor_desc(kh + ko, 1024), idesc, 1);
        }
        ptx::umma_commit(&s_done[s]);
      };
      auto issue_pv = [&](int h) {  // O[s] = P V (M128 N64 K128; V is an MN-major B operand: [keys][64 dims])
        const int s = h & 1;
        ptx::mbar_wait(&p_ready[s], (h >> 1) & 1);
        ptx::mbar_wait(&v_full[s], (h >> 1) & 1);
        if (h >= 2) ptx::mbar_wait(&o_free[s], 0);   // the epilogue of head h-2 has drained O[s]
        ptx::tc_fence_after();
        constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(128, 64) | (1u << 16);
        const uint32_t ph = ptx::smem_u32(smem + s * Cfg::STAGE), pl = ph + 2 * Cfg::TILE;   // P hi: q area, P lo: k area
        const uint32_t vh = ph + 4 * Cfg::TILE, vl = ph + 5 * Cfg::TILE;
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
          for (int k16 = 0; k16 < 4; ++k16) {
            const uint32_t pa = kb * Cfg::TILE + k16 * 32;        // A: 16 keys = 32 bytes along K
            const uint32_t va = (kb * 4 + k16) * 2048;            // B: 16 key rows of 128 bytes
            ptx::umma_bf16(t_o(s), ptx::make_sw128_kmajor_desc(pl + pa, 1024), ptx::make_sw128_mnmajor_desc(vh + va, 1024, 1024), idesc, (kb | k16) != 0);
            ptx::umma_bf16(t_o(s), ptx::make_sw128_kmajor_desc(ph + pa, 1024), ptx::make_sw128_mnmajor_desc(vl + va, 1024, 1024), idesc, 1);
            ptx::umma_bf16(t_o(s), ptx::make_sw128_kmajor_desc(ph + pa, 1024), ptx::make_sw128_mnmajor_desc(vh + va, 1024, 1024), idesc, 1);
          }
        }
        ptx::umma_commit(&o_done[s]);
      };
      issue_s(0); LTR_DBG_STAMP(113);
      issue_s(1); LTR_DBG_STAMP(114);
      issue_pv(0); LTR_DBG_STAMP(115);
      issue_s(2); LTR_DBG_STAMP(116);
      issue_pv(1); LTR_DBG_STAMP(117);
      issue_s(3); LTR_DBG_STAMP(118);
      issue_pv(2); LTR_DBG_STAMP(119);
      issue_pv(3); LTR_DBG_STAMP(120);
    }
  } else {
    // ---------------------------------------------------------------- softmax + epilogue (16 warps)
    const int q = warp & 3, part = (warp - 2) >> 2;       // TMEM lane quarter, 32-key column part
    const int row = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    constexpr float LOG2E = 1.4426950408889634f;
    auto softmax = [&](int h) {
      const int s = h & 1;
      float* rmax = red + (s * 2 + 0) * 512;
      float* rsum = red + (s * 2 + 1) * 512;
      ptx::mbar_wait(&s_done[s], (h >> 1) & 1);
      ptx::tc_fence_after();
      float v[32];
      ptx::tmem_ld32(t_s(s) + lane_addr + part * 32, v);
      float m = v[0];
#pragma unroll
      for (int j = 1; j < 32; ++j) m = fmaxf(m, v[j]);
      rmax[part * 128 + row] = m;
      sig_bar16();
      m = fmaxf(fmaxf(rmax[row], rmax[128 + row]), fmaxf(rmax[256 + row], rmax[384 + row]));
      const float mb = m * LOG2E;
      float l = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        v[j] = ptx::ex2_approx(fmaf(v[j], LOG2E, -mb));
        l += v[j];
      }
      rsum[part * 128 + row] = l;
      // P operand: keys part*32 .. +32 = half of k-block part/2 of the [128 x 128] P tile; hi plane in the q area,
      // lo plane in the k area (both dead: the S MMAs of this stage have completed)
      uint8_t* p_hi = smem + s * Cfg::STAGE + (part >> 1) * Cfg::TILE;
      uint8_t* p_lo = p_hi + 2 * Cfg::TILE;
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint4 hh, ll;
        ptx::split8_bf16(&v[j], hh, ll);
        const uint32_t off = ptx::sw128_offset(row, (part & 1) * 32 + j);
        *reinterpret_cast<uint4*>(p_hi + off) = hh;
        *reinterpret_cast<uint4*>(p_lo + off) = ll;
      }
      ptx::tc_fence_before();
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&p_ready[s]);
    };
    auto epilogue = [&](int h) {
      const int s = h & 1;
      const float* rsum = red + (s * 2 + 1) * 512;
      ptx::mbar_wait(&o_done[s], (h >> 1) & 1);
      ptx::tc_fence_after();
      float o[16];
      ptx::tmem_ld16(t_o(s) + lane_addr + part * 16, o);
      ptx::tc_fence_before();
      const float inv = 1.f / (rsum[row] + rsum[128 + row] + rsum[256 + row] + rsum[384 + row]);
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&o_free[s]);
#pragma unroll
      for (int j = 0; j < 16; j += 8) {
        const float w[8] = {o[j] * inv, o[j + 1] * inv, o[j + 2] * inv, o[j + 3] * inv, o[j + 4] * inv, o[j + 5] * inv, o[j + 6] * inv, o[j + 7] * inv};
        img_store8(out, img * 128 + row, (out_kb0 + h) * 64 + part * 16 + j, w);
      }
    };
    const bool st = warp == 2 && lane == 0;
    softmax(0); if (st) LTR_DBG_STAMP(121);
    softmax(1); if (st) LTR_DBG_STAMP(122);
    epilogue(0); if (st) LTR_DBG_STAMP(123);
    softmax(2); if (st) LTR_DBG_STAMP(124);
    epilogue(1);
    softmax(3); if (st) LTR_DBG_STAMP(125);
    epilogue(2);
    epilogue(3); if (st) LTR_DBG_STAMP(126);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, 512);
}

inline int launch_sig_attention_img(ActImg qkv, ActImg out, int out_k0, int n_images, cudaStream_t s) {
  if (n_images <= 0) return 0;
  LTR_CUDA_TRY(ensure_dynamic_smem(sig_attention_img_kernel, SigImgCfg::SMEM));
  LaunchScope ls(KC_SIG_ATTN, s);
  LTR_CUDA_TRY(launch_pdl(sig_attention_img_kernel, dim3(n_images), dim3(SigImgCfg::THREADS), SigImgCfg::SMEM, s, qkv, out, out_k0 / 64));
  return 0;
}

}  // namespace ltr
