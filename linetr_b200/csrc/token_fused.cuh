// Fused token stage of the line-descriptor forward (sm_100a, tcgen05 + TMA).
//
// One persistent CTA per SM walks tiles of LPT = floor(128 / T) whole lines (LPT*T <= 128 token
// rows).  Per tile, entirely on chip:
//   P0  narrow positional MLP 3 -> 32 -> 64 (+ReLU) on CUDA cores            -> h64  (smem, split-bf16)
//   L3  64 -> 128 (+ReLU)   tcgen05, accumulator in TMEM, epilogue -> smem   -> h128
//   L4  128 -> 256 (+ReLU)  tcgen05                                          -> h256
//   L5  256 -> 256          tcgen05, epilogue adds the sampled descriptors   -> x = desc + wpe (fp32, smem)
//   CLS pooling: folded CLS-query scores x.u_h, softmax over the T tokens + CLS per head,
//       z_h = sum_n p_h[n] x[n]                                               -> z image (global)
// Replaces WordPositionalEncoder (models/line_transformer.py:61-73), `desc + pos` and the CLS
// concat (:117-121) and the attention part of MultiHeadAttention restricted to the CLS query row
// (models/line_attention.py:13-21,55-63), i.e. what a narrow-MLP kernel, two GEMM launches and a
// pooling kernel would do through global memory.  Only `desc` (the mandatory HBM read, one
// SWIZZLE_128B tensor-map TMA per 32-column block of the tile), the weights (L2 resident, streamed
// by 1-D TMA through a 2-slot ring) and the pooled z leave/enter the SM.
//
// Warp roles: warp 0 = TMA weight producer, warp 1 = MMA issuer (+TMEM alloc), warps 2-9 =
// 256 worker threads (thread pair per token row: lane = row within the warp's TMEM lane
// quarter, `half` = which half of the output columns).
#pragma once
#include <cuda.h>   // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)
#include "act_img.cuh"
#include "common.cuh"
#include "tc_weight.cuh"
#include "ptx_sm100.cuh"

namespace ltr {

struct TokenFusedArgs {
  const float* pnt;    // [R*T, 2]
  const float* score;  // [R*T]
  const float* desc;   // [R*T, 256]
  // narrow layers (fp32, BN folded): w1 [32,3], w2 [64,32]
  const float *w1, *b1, *w2, *b2;
  TcWeight W3, W4, W5;  // [128,64], [256,128], [256,256] packed split-bf16
  const float *b3, *b4, *b5;
  const float* U;      // [4,256] folded CLS query u_h = W_k,h^T q_h / 8 (ltr_create)
  const float* s_cls;  // [4]
  const float* cls;    // [256]
  ActImg z;            // out: image of [R, 1024]
  int R, T, lpt, n_tiles;
  float cx, cy, scale;
};

struct TokenFusedSmem {
  static constexpr int ACT = 128 * 1024;   // activation region: h64/h128/h256 images, then x fp32
  static constexpr int SLOT = 32 * 1024;   // one W tile [128 n x 64 k] hi + lo
  static constexpr int NSLOT = 2;
  static constexpr int OFF_RING = ACT;
  static constexpr int OFF_W1 = OFF_RING + NSLOT * SLOT;   // 96 floats
  static constexpr int OFF_B1 = OFF_W1 + 96 * 4;           // 32
  static constexpr int OFF_W2 = OFF_B1 + 32 * 4;           // 64*32
  static constexpr int OFF_B2 = OFF_W2 + 2048 * 4;         // 64
  static constexpr int OFF_B3 = OFF_B2 + 64 * 4;           // 128
  static constexpr int OFF_B4 = OFF_B3 + 128 * 4;          // 256
  static constexpr int OFF_B5 = OFF_B4 + 256 * 4;          // 256
  static constexpr int OFF_U = OFF_B5 + 256 * 4;           // 1024
  static constexpr int OFF_CLS = OFF_U + 1024 * 4;         // 256
  static constexpr int OFF_SC = OFF_CLS + 256 * 4;         // partial scores [2][128][4]
  static constexpr int OFF_P = OFF_SC + 2 * 128 * 4 * 4;   // softmax scratch (see below)
  static constexpr int OFF_BAR = OFF_P + (512 + 512 + 512) * 4;   // exps [128][4], 1/sum [<=512], CLS exp [<=512]
  static constexpr int TOTAL = OFF_BAR + 128 + 1024;       // + alignment slack
};

// x tile in the activation region: 8 column blocks of [128 rows x 32 fp32 (128 B)], each row's eight
// 16-byte chunks XOR-swizzled by row & 7 - the layout a SWIZZLE_128B tensor-map box load produces.
// Conflict-free both for a thread per row (L5 epilogue) and for a thread per channel quad (pooling).
__device__ __forceinline__ int xs_index(int row, int col) {
  return (col >> 5) * 4096 + row * 32 + (((((col >> 2) & 7) ^ (row & 7)) << 2) | (col & 3));
}

__global__ void __launch_bounds__(320, 1) token_fused_kernel(TokenFusedArgs p, const __grid_constant__ CUtensorMap desc_map) {
  using S = TokenFusedSmem;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint8_t* act = smem;
  float* sW1 = reinterpret_cast<float*>(smem + S::OFF_W1);
  float* sB1 = reinterpret_cast<float*>(smem + S::OFF_B1);
  float* sW2 = reinterpret_cast<float*>(smem + S::OFF_W2);
  float* sB2 = reinterpret_cast<float*>(smem + S::OFF_B2);
  float* sB3 = reinterpret_cast<float*>(smem + S::OFF_B3);
  float* sB4 = reinterpret_cast<float*>(smem + S::OFF_B4);
  float* sB5 = reinterpret_cast<float*>(smem + S::OFF_B5);
  float* sU = reinterpret_cast<float*>(smem + S::OFF_U);
  float* sCls = reinterpret_cast<float*>(smem + S::OFF_CLS);
  float* sSc = reinterpret_cast<float*>(smem + S::OFF_SC);
  float* sP = reinterpret_cast<float*>(smem + S::OFF_P);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);
  uint64_t* full = bars;            // [2] W slot filled (TMA tx)
  uint64_t* empty = bars + 2;       // [2] W slot consumed (tcgen05.commit)
  uint64_t* a_ready = bars + 4;     // workers -> MMA: A operand image complete (256 arrivals)
  uint64_t* acc_ready = bars + 5;   // MMA -> workers: accumulator complete (tcgen05.commit)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
  uint64_t* dbar = bars + 8;        // [4] descriptor columns {32 s .. 32 s + 32} u {128 + 32 s ..} of the tile have landed
  uint64_t* l5_done = bars + 12;    // MMA -> MMA thread: all L5 MMAs complete (last descriptor group may be fetched)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_launch_dependents();

  for (int i = tid; i < 96; i += blockDim.x) sW1[i] = p.w1[i];
  for (int i = tid; i < 32; i += blockDim.x) sB1[i] = p.b1[i];
  for (int i = tid; i < 2048; i += blockDim.x) sW2[i] = p.w2[i];
  for (int i = tid; i < 64; i += blockDim.x) sB2[i] = p.b2[i];
  for (int i = tid; i < 128; i += blockDim.x) sB3[i] = p.b3[i];
  for (int i = tid; i < 256; i += blockDim.x) { sB4[i] = p.b4[i]; sB5[i] = p.b5[i]; sCls[i] = p.cls[i]; }
  for (int i = tid; i < 1024; i += blockDim.x) sU[i] = p.U[i];
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < 2; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
    ptx::mbar_init(a_ready, 256);
    ptx::mbar_init(acc_ready, 1);
    for (int i = 0; i < 4; ++i) ptx::mbar_init(&dbar[i], 1);
    ptx::mbar_init(l5_done, 1);
    ptx::prefetch_tensormap(&desc_map);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 256);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // z (output image) may still be read by the previous launch sequence

  if (warp == 0) {
    // ---------------------------------------------------------------- W producer: 13 slots per tile
    {
      uint32_t it = 0;
      auto push = [&](const TcWeight& W, int nb, int kb) {
        if (lane != 0) return;
        const int s = it & 1;
        const uint32_t ph = (it >> 1) & 1;
        ptx::mbar_wait(&empty[s], ph ^ 1);
        uint8_t* dst = smem + S::OFF_RING + s * S::SLOT;
        const size_t off = ((size_t)kb * (W.N / 8) + (size_t)nb * 16) * 1024;
        ptx::mbar_arrive_expect_tx(&full[s], S::SLOT);
        ptx::bulk_g2s(dst, reinterpret_cast<const uint8_t*>(W.hi) + off, 16384, &full[s]);
        ptx::bulk_g2s(dst + 16384, reinterpret_cast<const uint8_t*>(W.lo) + off, 16384, &full[s]);
        ++it;
      };
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        push(p.W3, 0, 0);
        for (int nb = 0; nb < 2; ++nb)
          for (int kb = 0; kb < 2; ++kb) push(p.W4, nb, kb);
        for (int kb = 0; kb < 4; ++kb)       // L5: k-block outer (see the MMA issuer: frees the A tiles early)
          for (int nb = 0; nb < 2; ++nb) push(p.W5, nb, kb);
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(128, 128);
      const uint32_t act_u = ptx::smem_u32(act);
      uint32_t it = 0, na = 0;
      // one [128 x 128 x 64] block: A k-block kb of an activation with nkb k-blocks, W from the ring
      auto block = [&](int nkb, int kb, int nb, bool first) {
        const int s = it & 1;
        const uint32_t ph = (it >> 1) & 1;
        ptx::mbar_wait(&full[s], ph);
        ptx::tc_fence_after();
        const uint32_t a_hi = act_u + kb * 16384, a_lo = act_u + (nkb + kb) * 16384;
        const uint32_t w_hi = ptx::smem_u32(smem + S::OFF_RING + s * S::SLOT), w_lo = w_hi + 16384;
        const uint32_t d = tmem_base + nb * 128;
#pragma unroll
        for (int k16 = 0; k16 < 4; ++k16) {
          const uint32_t ko = k16 * 32;
          const uint64_t dah = ptx::make_sw128_kmajor_desc(a_hi + ko, 1024);
          const uint64_t dal = ptx::make_sw128_kmajor_desc(a_lo + ko, 1024);
          const uint64_t dwh = ptx::make_sw128_kmajor_desc(w_hi + ko, 1024);
          const uint64_t dwl = ptx::make_sw128_kmajor_desc(w_lo + ko, 1024);
          ptx::umma_bf16(d, dal, dwh, idesc, !(first && k16 == 0));
          ptx::umma_bf16(d, dah, dwl, idesc, 1);
          ptx::umma_bf16(d, dah, dwh, idesc, 1);
        }
        ptx::umma_commit(&empty[s]);
        ++it;
      };
      // The tile's sampled descriptors ([128 token rows x 256] fp32 - the one mandatory HBM stream of this stage) land
      // in the activation region, which holds the L5 A operand (h256) until L5's MMAs have read it.  Column group
      // s (blocks s and 4 + s of 32 columns) occupies exactly the memory of h256's k-block s (hi and lo tile), so L5
      // runs k-block outer / n-block inner and group s is fetched as soon as both MMA blocks of k-block s are
      // complete - which this thread knows without extra barriers: the 2-slot W ring made it wait for them before it
      // could issue the blocks of k-block s + 1.  Three of the four groups (96 KB) stream in under L5's own MMAs.
      float* xs_t = reinterpret_cast<float*>(act);
      uint32_t nl5 = 0;
      auto issue_desc = [&](int sgrp, int tok0) {
        ptx::mbar_arrive_expect_tx(&dbar[sgrp], 2 * 16384);
        ptx::tma_load_2d(xs_t + sgrp * 4096, &desc_map, sgrp * 32, tok0, &dbar[sgrp]);
        ptx::tma_load_2d(xs_t + (4 + sgrp) * 4096, &desc_map, (4 + sgrp) * 32, tok0, &dbar[sgrp]);
      };
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const int tok0 = (int)((long long)tile * p.lpt * p.T);
        {
          // The next tile's descriptors (one contiguous range of rows) start towards L2 now.  All CTAs run their
          // tiles in the same phase, so without this the whole chip asks HBM for 19 MB in the same few thousand
          // cycles at the end of L5 and then leaves it idle for the rest of the tile (trace: the last descriptor
          // group landed 6.1 k cycles after L5's MMAs were complete).
          const long long nt = (long long)(tile + (int)gridDim.x) * p.lpt * p.T, total = (long long)p.R * p.T;
          if (nt < total) {
            const long long rows = min((long long)p.lpt * p.T, total - nt);
            ptx::bulk_prefetch_l2(p.desc + nt * 256, (uint32_t)(rows * 1024));
          }
        }
        ptx::mbar_wait(a_ready, na++ & 1);   // h64
        ptx::tc_fence_after();
        block(1, 0, 0, true);
        ptx::umma_commit(acc_ready);
        ptx::mbar_wait(a_ready, na++ & 1);   // h128
        ptx::tc_fence_after();
        for (int nb = 0; nb < 2; ++nb)
          for (int kb = 0; kb < 2; ++kb) block(2, kb, nb, kb == 0);
        ptx::umma_commit(acc_ready);
        ptx::mbar_wait(a_ready, na++ & 1);   // h256
        ptx::tc_fence_after();
        for (int kb = 0; kb < 4; ++kb) {
          for (int nb = 0; nb < 2; ++nb) block(4, kb, nb, kb == 0);
          if (kb >= 1) issue_desc(kb - 1, tok0);   // both blocks of k-block kb-1 are complete: their A tiles are dead
        }
        ptx::umma_commit(acc_ready);
        ptx::umma_commit(l5_done);
        ptx::mbar_wait(l5_done, nl5++ & 1);
        issue_desc(3, tok0);
      }
    }
  } else {
    // ---------------------------------------------------------------- workers (256 threads)
    const int q = warp & 3;            // TMEM lane quarter
    const int half = (warp - 2) >> 2;  // column half
    const int r_in = q * 32 + lane;    // token row inside the tile
    const int wt = tid - 64;           // 0..255
    const int rows_used = p.lpt * p.T;
    uint32_t nacc = 0, dph = 0;   // phase parities: accumulator barrier, descriptor-group barriers
    auto worker_sync = [&]() { asm volatile("bar.sync 1, 256;" ::: "memory"); };
    // split-bf16 store of 8 consecutive columns into an activation image with nkb k-blocks
    auto act_store8 = [&](int nkb, int col, const float (&v)[8]) {
      uint4 h, l;
      ptx::split8_bf16(v, h, l);
      const uint32_t off = (col >> 6) * 16384 + ptx::sw128_offset(r_in, col & 63);
      *reinterpret_cast<uint4*>(act + off) = h;
      *reinterpret_cast<uint4*>(act + nkb * 16384 + off) = l;
    };
    // P0: narrow MLP 3 -> 32 -> 64 (+ReLU) for this thread's token row of tile `t`; the thread produces
    // outputs [32*half, 32*half+32) and keeps them as packed split-bf16 registers until p0_store().
    uint32_t p0_hi[16], p0_lo[16];
    auto p0_compute = [&](int t) {
      const long long tk0 = (long long)t * p.lpt * p.T;
      float x0 = 0.f, x1 = 0.f, x2 = 0.f;
      if (r_in < rows_used && tk0 + r_in < (long long)p.R * p.T) {
        const long long tk = tk0 + r_in;
        x0 = (p.pnt[2 * tk] - p.cx) / p.scale;
        x1 = (p.pnt[2 * tk + 1] - p.cy) / p.scale;
        x2 = p.score[tk];
      }
      float h1[32];
#pragma unroll
      for (int n = 0; n < 32; ++n)
        h1[n] = fmaxf(fmaf(sW1[n * 3 + 2], x2, fmaf(sW1[n * 3 + 1], x1, fmaf(sW1[n * 3], x0, sB1[n]))), 0.f);
#pragma unroll
      for (int n0 = 0; n0 < 32; n0 += 2) {
        float o[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int n = half * 32 + n0 + j;
          float a = sB2[n];
#pragma unroll
          for (int k = 0; k < 32; k += 4) {
            const float4 w = *reinterpret_cast<const float4*>(&sW2[n * 32 + k]);
            a = fmaf(w.x, h1[k], a); a = fmaf(w.y, h1[k + 1], a);
            a = fmaf(w.z, h1[k + 2], a); a = fmaf(w.w, h1[k + 3], a);
          }
          o[j] = fmaxf(a, 0.f);
        }
        ptx::split2_bf16(o[0], o[1], p0_hi[n0 >> 1], p0_lo[n0 >> 1]);
      }
    };
    auto p0_store = [&]() {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint32_t off = ptx::sw128_offset(r_in, half * 32 + c * 8);
        *reinterpret_cast<uint4*>(act + off) = make_uint4(p0_hi[c * 4], p0_hi[c * 4 + 1], p0_hi[c * 4 + 2], p0_hi[c * 4 + 3]);
        *reinterpret_cast<uint4*>(act + 16384 + off) = make_uint4(p0_lo[c * 4], p0_lo[c * 4 + 1], p0_lo[c * 4 + 2], p0_lo[c * 4 + 3]);
      }
    };
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      const int line0 = tile * p.lpt;
      const bool tr = (tile == blockIdx.x + gridDim.x) && warp == 2 && lane == 0;   // trace the CTA's 2nd tile
      if (tr) LTR_DBG_STAMP(0);
      // ---- P0 result of THIS tile (computed during the previous tile's L5 MMAs) -> h64 operand tile
      if (tile == (int)blockIdx.x) p0_compute(tile);
      p0_store();
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(a_ready);
      if (tr) LTR_DBG_STAMP(1);
      // ---- epilogue L3: 128 columns (64 per half) -> h128
      ptx::mbar_wait(acc_ready, nacc++ & 1);
      ptx::tc_fence_after();
      if (tr) LTR_DBG_STAMP(2);
#pragma unroll 1
      for (int c0 = half * 64; c0 < half * 64 + 64; c0 += 32) {
        float acc[32];
        ptx::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, acc);
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          float o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = fmaxf(acc[j + e] + sB3[c0 + j + e], 0.f);
          act_store8(2, c0 + j, o);
        }
      }
      ptx::tc_fence_before();
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(a_ready);
      if (tr) LTR_DBG_STAMP(3);
      // ---- epilogue L4: 256 columns (128 per half) -> h256
      ptx::mbar_wait(acc_ready, nacc++ & 1);
      ptx::tc_fence_after();
      if (tr) LTR_DBG_STAMP(4);
#pragma unroll 1
      for (int c0 = half * 128; c0 < half * 128 + 128; c0 += 32) {
        float acc[32];
        ptx::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, acc);
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          float o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = fmaxf(acc[j + e] + sB4[c0 + j + e], 0.f);
          act_store8(4, c0 + j, o);
        }
      }
      ptx::tc_fence_before();
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(a_ready);
      if (tr) LTR_DBG_STAMP(5);
      if (tile + (int)gridDim.x < p.n_tiles) p0_compute(tile + gridDim.x);   // overlaps the L5 MMAs
      // ---- epilogue L5: x = acc + b5 + desc -> fp32 tile in the (now dead) activation region,
      //      partial CLS scores over this thread's 128 columns
      ptx::mbar_wait(acc_ready, nacc++ & 1);
      ptx::tc_fence_after();
      if (tr) LTR_DBG_STAMP(6);
      float* xs = reinterpret_cast<float*>(act);
      // phase A: the tile's descriptors arrive in the (dead) activation region by TMA - eight SWIZZLE_128B
      // boxes of 32 columns in four groups (group s = column blocks s and 4 + s, one mbarrier per group),
      // issued by the MMA thread while L5 still runs (see there).  Rows past the end of the batch are
      // zero-filled by the TMA unit; rows of the next tile that ride along are never used.
      if (tr) LTR_DBG_STAMP(12);
      float sc0 = 0.f, sc1 = 0.f, sc2 = 0.f, sc3 = 0.f;
#pragma unroll 1
      for (int c0 = half * 128; c0 < half * 128 + 128; c0 += 32) {
        float acc[32];
        ptx::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, acc);
        ptx::mbar_wait(&dbar[(c0 >> 5) & 3], dph);   // this column group of every row has landed
        if (tr) LTR_DBG_STAMP(13 + ((c0 >> 5) & 3));
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const int n = c0 + j;
          float4* xp = reinterpret_cast<float4*>(&xs[xs_index(r_in, n)]);
          const float4 d = *xp;
          float4 x;
          x.x = acc[j] + sB5[n] + d.x;
          x.y = acc[j + 1] + sB5[n + 1] + d.y;
          x.z = acc[j + 2] + sB5[n + 2] + d.z;
          x.w = acc[j + 3] + sB5[n + 3] + d.w;
          *xp = x;
          const float4 u0 = *reinterpret_cast<const float4*>(&sU[n]);
          const float4 u1 = *reinterpret_cast<const float4*>(&sU[256 + n]);
          const float4 u2 = *reinterpret_cast<const float4*>(&sU[512 + n]);
          const float4 u3 = *reinterpret_cast<const float4*>(&sU[768 + n]);
          sc0 = fmaf(x.x, u0.x, fmaf(x.y, u0.y, fmaf(x.z, u0.z, fmaf(x.w, u0.w, sc0))));
          sc1 = fmaf(x.x, u1.x, fmaf(x.y, u1.y, fmaf(x.z, u1.z, fmaf(x.w, u1.w, sc1))));
          sc2 = fmaf(x.x, u2.x, fmaf(x.y, u2.y, fmaf(x.z, u2.z, fmaf(x.w, u2.w, sc2))));
          sc3 = fmaf(x.x, u3.x, fmaf(x.y, u3.y, fmaf(x.z, u3.z, fmaf(x.w, u3.w, sc3))));
        }
      }
      dph ^= 1;
      ptx::tc_fence_before();
      *reinterpret_cast<float4*>(&sSc[(half * 128 + r_in) * 4]) = make_float4(sc0, sc1, sc2, sc3);
      if (tr) LTR_DBG_STAMP(7);
      worker_sync();
      if (tr) LTR_DBG_STAMP(8);
      // ---- softmax over the T tokens + CLS of every (line, head), spread over the row threads:
      //      (1) combine the two column halves of the scores, (2) per (line, head) maximum,
      //      (3) one exp per (row, head), (4) per (line, head) sum -> 1/sum and the CLS weight.
      //      sP keeps UNNORMALISED exps; the pooling multiplies by 1/sum once per output.
      float* sM = sSc + 512;      // [4*lpt] maxima   (second half of the partial-score buffer, dead after (1))
      float* sInv = sP + 512;     // [4*lpt] 1/sum;   sP + 1024 .. : [4*lpt] CLS weight e0
      if (wt < 128) {
        const float4 a = *reinterpret_cast<const float4*>(&sSc[wt * 4]);
        const float4 b = *reinterpret_cast<const float4*>(&sSc[(128 + wt) * 4]);
        *reinterpret_cast<float4*>(&sP[wt * 4]) = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
      }
      worker_sync();
      for (int pi = wt; pi < 4 * p.lpt; pi += 256) {
        const int ln = pi >> 2, h = pi & 3, rb = ln * p.T;
        float m = p.s_cls[h];
        for (int n = 0; n < p.T; ++n) m = fmaxf(m, sP[(rb + n) * 4 + h]);
        sM[pi] = m;
      }
      worker_sync();
      if (wt < rows_used) {
        const int ln = wt / p.T;
        float4 v = *reinterpret_cast<const float4*>(&sP[wt * 4]);
        v.x = expf(v.x - sM[ln * 4 + 0]);
        v.y = expf(v.y - sM[ln * 4 + 1]);
        v.z = expf(v.z - sM[ln * 4 + 2]);
        v.w = expf(v.w - sM[ln * 4 + 3]);
        *reinterpret_cast<float4*>(&sP[wt * 4]) = v;
      }
      worker_sync();
      for (int pi = wt; pi < 4 * p.lpt; pi += 256) {
        const int ln = pi >> 2, h = pi & 3, rb = ln * p.T;
        const float e0 = expf(p.s_cls[h] - sM[pi]);
        float sum = e0;
        for (int n = 0; n < p.T; ++n) sum += sP[(rb + n) * 4 + h];
        sInv[pi] = 1.f / sum;
        sP[1024 + pi] = e0;
      }
      worker_sync();
      if (tr) LTR_DBG_STAMP(9);
      // ---- pooling: z_h[c] = (e0 * cls[c] + sum_n e[n] x[n][c]) / sum.  Thread = (4 channels, 2 of the
      //      4 heads, every second line): one 16-byte x read feeds 8 FMAs (a thread per channel was
      //      latency-bound at 2 loads per 4 FMAs).  All three indices are warp-uniform except the channels.
      {
        const int cq = wt & 63, lg = (wt >> 6) & 1, hp = wt >> 7;
        const float4 cv = *reinterpret_cast<const float4*>(&sCls[cq * 4]);
        for (int ln = lg; ln < p.lpt; ln += 2) {
          const int gl = line0 + ln;
          if (gl >= p.R) break;
          const float2 pc = *reinterpret_cast<const float2*>(&sP[1024 + ln * 4 + 2 * hp]);
          float za0 = pc.x * cv.x, za1 = pc.x * cv.y, za2 = pc.x * cv.z, za3 = pc.x * cv.w;
          float zb0 = pc.y * cv.x, zb1 = pc.y * cv.y, zb2 = pc.y * cv.z, zb3 = pc.y * cv.w;
          const int rb = ln * p.T;
#pragma unroll 3
          for (int n = 0; n < p.T; ++n) {
            const float4 xv = *reinterpret_cast<const float4*>(&xs[xs_index(rb + n, cq * 4)]);
            const float2 pr = *reinterpret_cast<const float2*>(&sP[(rb + n) * 4 + 2 * hp]);
            za0 = fmaf(pr.x, xv.x, za0); za1 = fmaf(pr.x, xv.y, za1); za2 = fmaf(pr.x, xv.z, za2); za3 = fmaf(pr.x, xv.w, za3);
            zb0 = fmaf(pr.y, xv.x, zb0); zb1 = fmaf(pr.y, xv.y, zb1); zb2 = fmaf(pr.y, xv.z, zb2); zb3 = fmaf(pr.y, xv.w, zb3);
          }
          const float2 iv = *reinterpret_cast<const float2*>(&sInv[ln * 4 + 2 * hp]);
          img_store4(p.z, gl, (2 * hp) * 256 + cq * 4, za0 * iv.x, za1 * iv.x, za2 * iv.x, za3 * iv.x);
          img_store4(p.z, gl, (2 * hp + 1) * 256 + cq * 4, zb0 * iv.y, zb1 * iv.y, zb2 * iv.y, zb3 * iv.y);
        }
      }
      if (tr) LTR_DBG_STAMP(10);
      worker_sync();   // x tile and probabilities are dead: the next tile may overwrite them
      if (tr) LTR_DBG_STAMP(11);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, 256);
}

// cuTensorMapEncodeTiled through the runtime (no link against libcuda)
typedef CUresult (*TensorMapEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                           const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline TensorMapEncodeTiledFn tensor_map_encoder() {
  static TensorMapEncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<TensorMapEncodeTiledFn>(p);
  }
  return fn;
}

inline int launch_token_fused(TokenFusedArgs a, cudaStream_t s) {
  if (a.R <= 0) return 0;
  if (a.T < 1 || a.T > 128) return set_error(-1, "token_fused: T must be in 1..128");
  if ((long long)a.R * a.T >= (1ll << 31)) return set_error(-1, "token_fused: more than 2^31 tokens in one call");
  if (reinterpret_cast<uintptr_t>(a.desc) % 16) return set_error(-1, "token_fused: desc_sublines must be 16-byte aligned");
  LTR_CUDA_TRY(ensure_dynamic_smem(token_fused_kernel, TokenFusedSmem::TOTAL));
  a.lpt = 128 / a.T;
  a.n_tiles = cdiv(a.R, a.lpt);
  // tensor map of the sampled descriptors viewed as [R*T rows, 256] fp32; box = 32 columns (128 B) x 128 rows
  TensorMapEncodeTiledFn enc = tensor_map_encoder();
  if (!enc) return set_error(-1, "token_fused: cuTensorMapEncodeTiled is not available from this driver");
  CUtensorMap map;
  const cuuint64_t gdim[2] = {256, (cuuint64_t)a.R * a.T};
  const cuuint64_t gstride[1] = {256 * sizeof(float)};
  const cuuint32_t box[2] = {32, 128};
  const cuuint32_t estride[2] = {1, 1};
  const CUresult cr = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(a.desc), gdim, gstride, box, estride,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return set_error(-1, "token_fused: cuTensorMapEncodeTiled failed (" + std::to_string((int)cr) + ")");
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (sms <= 0) sms = 148;
  const int grid = a.n_tiles < sms ? a.n_tiles : sms;
  LaunchScope ls(KC_TOKEN_FUSED, s);
  LTR_CUDA_TRY(launch_pdl(token_fused_kernel, dim3(grid), dim3(320), TokenFusedSmem::TOTAL, s, a, map));
  return 0;
}

}  // namespace ltr
