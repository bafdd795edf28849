// Fused token stage of the line-descriptor forward (sm_100a, tcgen05 + TMA).
//
// One persistent CTA per SM walks tiles of LPT = floor(128 / T) whole lines (LPT*T <= 128 token
// rows).  Per tile, entirely on chip:
//   P0  3 -> 32 (+ReLU) on CUDA cores                                        -> h32  (TMEM, split-bf16)
//   L2  32 -> 64 (+ReLU)    tcgen05 (N = 64, two K steps)                    -> h64
//   L3  64 -> 128 (+ReLU)   tcgen05, accumulator in TMEM, epilogue -> TMEM   -> h128
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
// The activations never touch shared memory: the A operand of every layer lives in TENSOR MEMORY
// (tcgen05.mma with A from TMEM) as split-bf16 - packed hi pairs and packed lo pairs, written by the
// previous layer's epilogue with tcgen05.st.  With A in shared memory a 128 x 128 x 16 MMA read
// 8 KB of operands per 64 cycles - the whole 128 B/clk of the SM's shared memory - while the weight
// ring and the descriptor stream were writing into the same memory (L5: 9.9 k cycles for 6.1 k of
// MMAs); and the 128 KB activation image shared its memory with the tile's descriptors, so the last
// descriptor group could only be fetched after L5 (it landed 6.1 k cycles into the L5 epilogue).
// Now the 128 KB are the descriptor / x tile alone: the next tile's descriptors are requested the
// moment the pooling of this tile has read x, and land under the next tile's P0 .. L5.
// TMEM columns (512): accumulators [0, 256); h128 hi [256, 320) lo [320, 384);
// h32 hi [256, 272) lo [272, 288); h64 hi [384, 416) lo [416, 448); h256: k < 128 hi [384, 448) lo [448, 512), k >= 128 hi [256, 320)
// lo [320, 384) - the first half of h256 is written while L4's second n-block still reads h128.
//
// Warp roles: warp 0 = TMA weight producer, warp 1 = MMA issuer (+TMEM alloc, descriptor TMA),
// warps 2-9 = 256 worker threads (thread pair per token row: lane = row within the warp's TMEM lane
// quarter; the pair splits every 128-column n-block of a layer in halves, so the first n-block's
// epilogue runs under the second n-block's MMAs).
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
  // first layer (fp32, BN folded): w1 [32,3]; b2 [64] is the bias of the second
  const float *w1, *b1, *b2;
  TcWeight W2, W3, W4, W5;  // [64,64 (K 32 zero-padded)], [128,64], [256,128], [256,256] packed split-bf16
  const float *b3, *b4, *b5;
  const float* U;      // [4,256] folded CLS query u_h = W_k,h^T q_h / 8 (ltr_create)
  const float* s_cls;  // [4]
  const float* cls;    // [256]
  ActImg z;            // out: image of [R, 1024]
  int R, T, lpt, n_tiles;
  float cx, cy, scale;
};

struct TokenFusedSmem {
  static constexpr int ACT = 128 * 1024;   // the descriptor / x tile (fp32); the activations live in tensor memory
  static constexpr int SLOT = 32 * 1024;   // one W tile [128 n x 64 k] hi + lo
  static constexpr int NSLOT = 2;
  static constexpr int OFF_RING = ACT;
  static constexpr int OFF_W1 = OFF_RING + NSLOT * SLOT;   // 96 floats
  static constexpr int OFF_B1 = OFF_W1 + 96 * 4;           // 32
  static constexpr int OFF_B2 = OFF_B1 + 32 * 4;           // 64
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

// x tile: 8 column blocks of [128 rows x 32 fp32 (128 B)], each row's eight
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
  uint64_t* a_ready = bars + 4;     // workers -> MMA: A operand complete in tensor memory (256 arrivals)
  uint64_t* acc_bar = bars + 5;     // [2] MMA -> workers: first / second 128-column n-block of the layer complete (tcgen05.commit)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);
  uint64_t* dbar = bars + 8;        // [4] descriptor columns {32 s .. 32 s + 32} u {128 + 32 s ..} of the tile have landed
  uint64_t* x_free = bars + 12;     // workers -> MMA thread: the pooling has read the x tile (256 arrivals)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_launch_dependents();

  for (int i = tid; i < 96; i += blockDim.x) sW1[i] = p.w1[i];
  for (int i = tid; i < 32; i += blockDim.x) sB1[i] = p.b1[i];
  for (int i = tid; i < 64; i += blockDim.x) sB2[i] = p.b2[i];
  for (int i = tid; i < 128; i += blockDim.x) sB3[i] = p.b3[i];
  for (int i = tid; i < 256; i += blockDim.x) { sB4[i] = p.b4[i]; sB5[i] = p.b5[i]; sCls[i] = p.cls[i]; }
  for (int i = tid; i < 1024; i += blockDim.x) sU[i] = p.U[i];
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < 2; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
    ptx::mbar_init(a_ready, 256);
    ptx::mbar_init(&acc_bar[0], 1);
    ptx::mbar_init(&acc_bar[1], 1);
    for (int i = 0; i < 4; ++i) ptx::mbar_init(&dbar[i], 1);
    ptx::mbar_init(x_free, 256);
    ptx::prefetch_tensormap(&desc_map);
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
  // tensor-memory columns (see the header)
  constexpr uint32_t T_H128_HI = 256, T_H128_LO = 320, T_H64_HI = 384, T_H64_LO = 416, T_H32_HI = 256, T_H32_LO = 272;
  constexpr uint32_t T_H256A_HI = 384, T_H256A_LO = 448, T_H256B_HI = 256, T_H256B_LO = 320;   // A: k < 128, B: k >= 128
  pdl_wait();   // z (output image) may still be read by the previous launch sequence

  if (warp == 0) {
    // ---------------------------------------------------------------- W producer: 14 slots per tile
    {
      uint32_t it = 0;
      auto push = [&](const TcWeight& W, int nb, int kb) {
        const int s = it & 1;
        const uint32_t ph = (it >> 1) & 1;
        ptx::mbar_wait(&empty[s], ph ^ 1);
        uint8_t* dst = smem + S::OFF_RING + s * S::SLOT;
        const size_t off = ((size_t)kb * (W.N / 8) + (size_t)nb * 16) * 1024;
        const uint32_t bytes = (uint32_t)min(128, W.N - nb * 128) * 128u;   // rows of this n-block x 128 B, per plane
        ptx::mbar_arrive_expect_tx(&full[s], 2 * bytes);
        ptx::bulk_g2s(dst, reinterpret_cast<const uint8_t*>(W.hi) + off, bytes, &full[s]);
        ptx::bulk_g2s(dst + 16384, reinterpret_cast<const uint8_t*>(W.lo) + off, bytes, &full[s]);
        ++it;
      };
      if (lane == 0) {
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
          push(p.W2, 0, 0);
          push(p.W3, 0, 0);
          for (int nb = 0; nb < 2; ++nb)
            for (int kb = 0; kb < 2; ++kb) push(p.W4, nb, kb);
          for (int nb = 0; nb < 2; ++nb)
            for (int kb = 0; kb < 4; ++kb) push(p.W5, nb, kb);
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc128 = ptx::make_idesc_bf16_f32(128, 128), idesc64 = ptx::make_idesc_bf16_f32(128, 64);
      uint32_t it = 0, na = 0;
      // one [128 x 128 x 64] block: K steps k0 .. k0 + 3 (16 wide) of an A operand whose packed hi / lo pairs start at
      // tensor-memory columns a_hi / a_lo (8 columns per step), W from the ring, accumulator columns nb * 128 ..
      auto block = [&](uint32_t a_hi, uint32_t a_lo, int nb, bool first, uint32_t idesc, int nk16) {
        const int s = it & 1;
        const uint32_t ph = (it >> 1) & 1;
        ptx::mbar_wait(&full[s], ph);
        ptx::tc_fence_after();
        const uint32_t w_hi = ptx::smem_u32(smem + S::OFF_RING + s * S::SLOT), w_lo = w_hi + 16384;
        const uint32_t d = tmem_base + nb * 128;
#pragma unroll
        for (int k16 = 0; k16 < 4; ++k16) {
          if (k16 >= nk16) break;
          const uint32_t ko = k16 * 32;
          const uint64_t dwh = ptx::make_sw128_kmajor_desc(w_hi + ko, 1024);
          const uint64_t dwl = ptx::make_sw128_kmajor_desc(w_lo + ko, 1024);
          ptx::umma_bf16_ts(d, tmem_base + a_lo + k16 * 8, dwh, idesc, !(first && k16 == 0));
          ptx::umma_bf16_ts(d, tmem_base + a_hi + k16 * 8, dwl, idesc, 1);
          ptx::umma_bf16_ts(d, tmem_base + a_hi + k16 * 8, dwh, idesc, 1);
        }
        ptx::umma_commit(&empty[s]);
        ++it;
      };
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        ptx::mbar_wait(a_ready, na++ & 1);   // h32
        ptx::tc_fence_after();
        block(T_H32_HI, T_H32_LO, 0, true, idesc64, 2);   // K = 32: the k-block's upper half is zero padding
        ptx::umma_commit(&acc_bar[0]);
        ptx::mbar_wait(a_ready, na++ & 1);   // h64
        ptx::tc_fence_after();
        block(T_H64_HI, T_H64_LO, 0, true, idesc128, 4);
        ptx::umma_commit(&acc_bar[0]);
        ptx::mbar_wait(a_ready, na++ & 1);   // h128
        ptx::tc_fence_after();
        for (int nb = 0; nb < 2; ++nb) {
          for (int kb = 0; kb < 2; ++kb) block(T_H128_HI + kb * 32, T_H128_LO + kb * 32, nb, kb == 0, idesc128, 4);
          ptx::umma_commit(&acc_bar[nb]);
        }
        ptx::mbar_wait(a_ready, na++ & 1);   // h256
        ptx::tc_fence_after();
        for (int nb = 0; nb < 2; ++nb) {
          for (int kb = 0; kb < 4; ++kb)
            block((kb < 2 ? T_H256A_HI : T_H256B_HI) + (kb & 1) * 32, (kb < 2 ? T_H256A_LO : T_H256B_LO) + (kb & 1) * 32, nb, kb == 0, idesc128, 4);
          ptx::umma_commit(&acc_bar[nb]);
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- workers (256 threads)
    const int q = warp & 3;            // TMEM lane quarter
    const int half = (warp - 2) >> 2;  // column half
    const int r_in = q * 32 + lane;    // token row inside the tile
    const int wt = tid - 64;           // 0..255
    const int rows_used = p.lpt * p.T;
    uint32_t nacc0 = 0, nacc1 = 0, dph = 0;   // phase parities: the two accumulator barriers, descriptor-group barriers
    auto worker_sync = [&]() { asm volatile("bar.sync 1, 256;" ::: "memory"); };
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    // relu(acc + bias) of 32 consecutive columns -> split-bf16 pairs -> 16 + 16 tensor-memory columns of the next
    // layer's A operand (hi pairs at t_hi, lo pairs at t_lo; the caller passes the column of the first pair)
    auto act_store32 = [&](const float (&acc)[32], const float* bias, uint32_t t_hi, uint32_t t_lo) {
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int j = 0; j < 32; j += 2) {   // packed fp32 pairs (add.f32x2 / fma.f32x2): half the issue slots of the scalar form
        const float2 v = __fadd2_rn(make_float2(acc[j], acc[j + 1]), *reinterpret_cast<const float2*>(bias + j));
        ptx::split2_bf16_x2(make_float2(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f)), hi[j >> 1], lo[j >> 1]);
      }
      ptx::tmem_st16(lane_addr + t_hi, hi);
      ptx::tmem_st16(lane_addr + t_lo, lo);
    };
    const float scls0 = p.s_cls[0], scls1 = p.s_cls[1], scls2 = p.s_cls[2], scls3 = p.s_cls[3];
    // token coordinates + score of this thread's row of tile `t` (raw; requested one tile ahead of their use)
    float in0 = 0.f, in1 = 0.f, in2 = 0.f;
    auto p0_fetch = [&](int t) {
      const long long tk0 = (long long)t * p.lpt * p.T;
      in0 = in1 = in2 = 0.f;
      if (t < p.n_tiles && r_in < rows_used && tk0 + r_in < (long long)p.R * p.T) {
        const long long tk = tk0 + r_in;
        in0 = p.pnt[2 * tk]; in1 = p.pnt[2 * tk + 1]; in2 = p.score[tk];
      }
    };
    p0_fetch(blockIdx.x);
    // P0: first layer 3 -> 32 (+ReLU) of this thread's token row; the pair of threads of a row splits the 32 outputs.
    // Result: 8 + 8 packed split-bf16 pairs -> tensor-memory columns of the h32 operand.
    auto p0 = [&](int t) {
      const long long tk0 = (long long)t * p.lpt * p.T;
      float x0 = 0.f, x1 = 0.f, x2 = 0.f;
      if (r_in < rows_used && tk0 + r_in < (long long)p.R * p.T) {
        x0 = (in0 - p.cx) / p.scale;
        x1 = (in1 - p.cy) / p.scale;
        x2 = in2;
      }
      p0_fetch(t + (int)gridDim.x);
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        const int n = half * 16 + j;
        const float a = fmaxf(fmaf(sW1[n * 3 + 2], x2, fmaf(sW1[n * 3 + 1], x1, fmaf(sW1[n * 3], x0, sB1[n]))), 0.f);
        const float b = fmaxf(fmaf(sW1[n * 3 + 5], x2, fmaf(sW1[n * 3 + 4], x1, fmaf(sW1[n * 3 + 3], x0, sB1[n + 1]))), 0.f);
        ptx::split2_bf16(a, b, hi[j >> 1], lo[j >> 1]);
      }
      ptx::tmem_st8(lane_addr + T_H32_HI + half * 8, hi);
      ptx::tmem_st8(lane_addr + T_H32_LO + half * 8, lo);
      ptx::tmem_wait_st();
    };
    // ---- Softmax and pooling of a tile are DEFERRED into the next tile's MMA phases: they need only shared memory
    //      (the tile's partial scores and its x tile), while the workers would otherwise idle behind L2, L3, the first
    //      n-blocks of L4 and L5 (8.8 k of a 26.6 k-cycle tile in the clock64 trace).  The x tile is handed back for the
    //      next descriptors (x_free) as soon as the last deferred line has been pooled, during L5's first n-block.
    float* xs = reinterpret_cast<float*>(act);
    float* sInv = sP + 512;     // [4*lpt] 1/sum;   sP + 1024 .. : [4*lpt] CLS weight e0
    int prev_line0 = 0;
    bool have_prev = false;
    // The tile's sampled descriptors ([128 token rows x 256] fp32 - the one mandatory HBM stream of this stage) land in
    // the x tile, four groups of two 32-column SWIZZLE_128B boxes (group s = blocks s and 4 + s, the order the L5 epilogue
    // walks them).  Requested by worker thread 0: for the first tile at once, later the moment all 256 workers have
    // handed the x tile back (x_free) - a wait its warp has to sit out anyway before it can touch the new descriptors.
    uint32_t nx = 0;
    auto request_desc = [&](int t) {
      float* xs_t = reinterpret_cast<float*>(act);
      const int tok0 = (int)((long long)t * p.lpt * p.T);
      for (int sgrp = 0; sgrp < 4; ++sgrp) {
        ptx::mbar_arrive_expect_tx(&dbar[sgrp], 2 * 16384);
        ptx::tma_load_2d(xs_t + sgrp * 4096, &desc_map, sgrp * 32, tok0, &dbar[sgrp]);
        ptx::tma_load_2d(xs_t + (4 + sgrp) * 4096, &desc_map, (4 + sgrp) * 32, tok0, &dbar[sgrp]);
      }
    };
    if (wt == 0 && (int)blockIdx.x < p.n_tiles) request_desc(blockIdx.x);
    // softmax over the T tokens + CLS of every (line, head).  Thread = (token row, head pair): it scans the scores of
    // its row's line (both column halves, 2 x 8 B per token) for the maximum - redundantly per row, which costs T
    // cheap iterations instead of a barrier and a pass by 4 * lpt threads - and stores its own exp; after one barrier
    // the first row of every line sums the line's exps -> 1/sum and the CLS weight.  sP keeps UNNORMALISED exps; the
    // pooling multiplies by 1/sum once per output.
    auto softmax_prev = [&]() {
      const int srow = wt & 127, h2 = (wt >> 7) * 2;
      const bool live = srow < rows_used;
      const int ln = live ? srow / p.T : 0, rb = ln * p.T;
      const float c0 = h2 ? scls2 : scls0, c1 = h2 ? scls3 : scls1;
      float m0 = c0, m1 = c1;
      if (live) {
        for (int n = 0; n < p.T; ++n) {
          const float2 a = *reinterpret_cast<const float2*>(&sSc[(rb + n) * 4 + h2]);
          const float2 b = *reinterpret_cast<const float2*>(&sSc[(128 + rb + n) * 4 + h2]);
          m0 = fmaxf(m0, a.x + b.x);
          m1 = fmaxf(m1, a.y + b.y);
        }
        const float2 a = *reinterpret_cast<const float2*>(&sSc[srow * 4 + h2]);
        const float2 b = *reinterpret_cast<const float2*>(&sSc[(128 + srow) * 4 + h2]);
        *reinterpret_cast<float2*>(&sP[srow * 4 + h2]) = make_float2(expf(a.x + b.x - m0), expf(a.y + b.y - m1));
      }
      worker_sync();
      if (live && srow == rb) {
        const float e0 = expf(c0 - m0), e1 = expf(c1 - m1);
        float s0 = e0, s1 = e1;
        for (int n = 0; n < p.T; ++n) {
          const float2 e = *reinterpret_cast<const float2*>(&sP[(rb + n) * 4 + h2]);
          s0 += e.x;
          s1 += e.y;
        }
        *reinterpret_cast<float2*>(&sInv[ln * 4 + h2]) = make_float2(1.f / s0, 1.f / s1);
        *reinterpret_cast<float2*>(&sP[1024 + ln * 4 + h2]) = make_float2(e0, e1);
      }
      worker_sync();
    };
    // pooling: z_h[c] = (e0 * cls[c] + sum_n e[n] x[n][c]) / sum.  Thread = (4 channels, every fourth line), all four
    // heads: one 16-byte x read and one 16-byte read of the four heads' weights feed 16 FMAs, and the x tile is read
    // ONCE per tile (a head pair per thread read it twice: 256 KB of the SM's shared-memory bandwidth per tile).
    // Lines [first, first + count) of this thread's line sequence lg, lg + 4, ...
    auto pool_prev = [&](int first, int count) {
      const int cq = wt & 63, lg = wt >> 6;
      const float4 cv = *reinterpret_cast<const float4*>(&sCls[cq * 4]);
      for (int k = first; k < first + count; ++k) {
        const int ln = lg + 4 * k;
        if (ln >= p.lpt) break;
        const int gl = prev_line0 + ln;
        if (gl >= p.R) break;
        const float4 pc = *reinterpret_cast<const float4*>(&sP[1024 + ln * 4]);
        float z[4][4];
        const float pcv[4] = {pc.x, pc.y, pc.z, pc.w};
#pragma unroll
        for (int h = 0; h < 4; ++h) { z[h][0] = pcv[h] * cv.x; z[h][1] = pcv[h] * cv.y; z[h][2] = pcv[h] * cv.z; z[h][3] = pcv[h] * cv.w; }
        const int rb = ln * p.T;
#pragma unroll 3
        for (int n = 0; n < p.T; ++n) {
          const float4 xv = *reinterpret_cast<const float4*>(&xs[xs_index(rb + n, cq * 4)]);
          const float4 pr = *reinterpret_cast<const float4*>(&sP[(rb + n) * 4]);
          const float prv[4] = {pr.x, pr.y, pr.z, pr.w};
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const float2 pp = make_float2(prv[h], prv[h]);
            const float2 za = __ffma2_rn(pp, make_float2(xv.x, xv.y), make_float2(z[h][0], z[h][1]));
            const float2 zb = __ffma2_rn(pp, make_float2(xv.z, xv.w), make_float2(z[h][2], z[h][3]));
            z[h][0] = za.x; z[h][1] = za.y; z[h][2] = zb.x; z[h][3] = zb.y;
          }
        }
        const float4 iv = *reinterpret_cast<const float4*>(&sInv[ln * 4]);
        const float ivv[4] = {iv.x, iv.y, iv.z, iv.w};
#pragma unroll
        for (int h = 0; h < 4; ++h)
          img_store4(p.z, gl, h * 256 + cq * 4, z[h][0] * ivv[h], z[h][1] * ivv[h], z[h][2] * ivv[h], z[h][3] * ivv[h]);
      }
    };
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      const int line0 = tile * p.lpt;
      const bool tr = (tile == blockIdx.x + gridDim.x) && warp == 2 && lane == 0;   // trace the CTA's 2nd tile
      if (tr) LTR_DBG_STAMP(0);
      // ---- P0 -> h32 operand
      p0(tile);
      ptx::tc_fence_before();
      ptx::mbar_arrive(a_ready);
      if (have_prev) softmax_prev();   // under the L2 MMAs
      // ---- epilogue L2: 64 columns (32 per thread of the pair) -> h64
      ptx::mbar_wait(&acc_bar[0], nacc0++ & 1);
      ptx::tc_fence_after();
      {
        float acc[32];
        ptx::tmem_ld32(lane_addr + (uint32_t)(half * 32), acc);
        act_store32(acc, sB2 + half * 32, T_H64_HI + half * 16, T_H64_LO + half * 16);
      }
      ptx::tmem_wait_st();
      ptx::tc_fence_before();
      ptx::mbar_arrive(a_ready);
      if (tr) LTR_DBG_STAMP(1);
      if (have_prev) pool_prev(0, 1);   // under the L3 MMAs
      // ---- epilogue L3: 128 columns (64 per half) -> h128
      ptx::mbar_wait(&acc_bar[0], nacc0++ & 1);
      ptx::tc_fence_after();
      if (tr) LTR_DBG_STAMP(2);
#pragma unroll 1
      for (int c0 = half * 64; c0 < half * 64 + 64; c0 += 32) {
        float acc[32];
        ptx::tmem_ld32(lane_addr + (uint32_t)c0, acc);
        act_store32(acc, sB3 + c0, T_H128_HI + (c0 >> 1), T_H128_LO + (c0 >> 1));
      }
      ptx::tmem_wait_st();
      ptx::tc_fence_before();
      ptx::mbar_arrive(a_ready);
      if (tr) LTR_DBG_STAMP(3);
      if (have_prev) pool_prev(1, 1);   // under L4's first n-block
      // ---- epilogue L4: 256 columns -> h256.  The pair of threads of a row halves EACH 128-column n-block (this
      //      thread: columns nb * 128 + half * 64 .. + 64), so the first n-block is converted while the second
      //      is still being multiplied; its pairs go to the h256 columns that do not alias h128 (see the header).
#pragma unroll 1
      for (int nb = 0; nb < 2; ++nb) {
        if (nb == 0) ptx::mbar_wait(&acc_bar[0], nacc0++ & 1);
        else ptx::mbar_wait(&acc_bar[1], nacc1++ & 1);
        ptx::tc_fence_after();
        if (tr && nb == 0) LTR_DBG_STAMP(4);
#pragma unroll 1
        for (int c0 = nb * 128 + half * 64; c0 < nb * 128 + half * 64 + 64; c0 += 32) {
          float acc[32];
          ptx::tmem_ld32(lane_addr + (uint32_t)c0, acc);
          const uint32_t pc = (uint32_t)(c0 & 127) >> 1;   // pair column inside the 64-column half of h256
          act_store32(acc, sB4 + c0, (nb == 0 ? T_H256A_HI : T_H256B_HI) + pc, (nb == 0 ? T_H256A_LO : T_H256B_LO) + pc);
        }
      }
      ptx::tmem_wait_st();
      ptx::tc_fence_before();
      ptx::mbar_arrive(a_ready);
      if (tr) LTR_DBG_STAMP(5);
      if (have_prev) {                  // under L5's first n-block: the rest, then the x tile is free for this tile's descriptors
        pool_prev(2, 1 << 30);
        ptx::fence_proxy_async_smem();  // this thread's generic accesses to the x tile before the TMA writes that follow
        ptx::mbar_arrive(x_free);
        if (wt == 0) {
          ptx::mbar_wait(x_free, nx++ & 1);
          request_desc(tile);
        }
      }
      // ---- epilogue L5: x = acc + b5 + desc -> fp32, in place in the x tile (where the TMA put desc);
      //      partial CLS scores over this thread's 128 columns (64 of each n-block, first n-block first)
      float2 s0 = make_float2(0.f, 0.f), s1 = s0, s2 = s0, s3 = s0;
#pragma unroll 1
      for (int nb = 0; nb < 2; ++nb) {
        if (nb == 0) ptx::mbar_wait(&acc_bar[0], nacc0++ & 1);
        else ptx::mbar_wait(&acc_bar[1], nacc1++ & 1);
        ptx::tc_fence_after();
        if (tr && nb == 0) LTR_DBG_STAMP(6);
        if (tr && nb == 1) LTR_DBG_STAMP(12);
#pragma unroll 1
        for (int c0 = nb * 128 + half * 64; c0 < nb * 128 + half * 64 + 64; c0 += 32) {
          float acc[32];
          ptx::tmem_ld32(lane_addr + (uint32_t)c0, acc);
          ptx::mbar_wait(&dbar[(c0 >> 5) & 3], dph);   // this column group of every row has landed (long ago, normally)
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const int n = c0 + j;
            float4* xp = reinterpret_cast<float4*>(&xs[xs_index(r_in, n)]);
            const float4 d = *xp;
            const float4 b = *reinterpret_cast<const float4*>(&sB5[n]);
            // packed fp32 pairs: per head an (even, odd) column pair of partial scores, summed after the loop
            const float2 xa = __fadd2_rn(__fadd2_rn(make_float2(acc[j], acc[j + 1]), make_float2(b.x, b.y)), make_float2(d.x, d.y));
            const float2 xb = __fadd2_rn(__fadd2_rn(make_float2(acc[j + 2], acc[j + 3]), make_float2(b.z, b.w)), make_float2(d.z, d.w));
            *xp = make_float4(xa.x, xa.y, xb.x, xb.y);
            const float4 u0 = *reinterpret_cast<const float4*>(&sU[n]);
            const float4 u1 = *reinterpret_cast<const float4*>(&sU[256 + n]);
            const float4 u2 = *reinterpret_cast<const float4*>(&sU[512 + n]);
            const float4 u3 = *reinterpret_cast<const float4*>(&sU[768 + n]);
            s0 = __ffma2_rn(xa, make_float2(u0.x, u0.y), __ffma2_rn(xb, make_float2(u0.z, u0.w), s0));
            s1 = __ffma2_rn(xa, make_float2(u1.x, u1.y), __ffma2_rn(xb, make_float2(u1.z, u1.w), s1));
            s2 = __ffma2_rn(xa, make_float2(u2.x, u2.y), __ffma2_rn(xb, make_float2(u2.z, u2.w), s2));
            s3 = __ffma2_rn(xa, make_float2(u3.x, u3.y), __ffma2_rn(xb, make_float2(u3.z, u3.w), s3));
          }
        }
      }
      dph ^= 1;
      ptx::tc_fence_before();
      *reinterpret_cast<float4*>(&sSc[(half * 128 + r_in) * 4]) = make_float4(s0.x + s0.y, s1.x + s1.y, s2.x + s2.y, s3.x + s3.y);
      if (tr) LTR_DBG_STAMP(7);
      worker_sync();
      if (tr) LTR_DBG_STAMP(8);
      // softmax + pooling of THIS tile run inside the next tile's MMA phases (or after the loop for the last one)
      prev_line0 = line0;
      have_prev = true;
      if (tr) LTR_DBG_STAMP(11);
    }
    if (have_prev) {   // drain: the CTA's last tile
      softmax_prev();
      pool_prev(0, 1 << 30);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, 512);
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
