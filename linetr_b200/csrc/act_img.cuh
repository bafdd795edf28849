// Split-bf16 activation image: the operand format of the tensor-core GEMM engine.
// See gemm_img.cuh for the layout description.
#pragma once
#include "ptx_sm100.cuh"

namespace ltr {

struct ActImg {
  __nv_bfloat16* hi = nullptr;
  __nv_bfloat16* lo = nullptr;
  int kblocks = 0;  // K/64 tiles per 128-row block
};
constexpr int IMG_TILE_ELEMS = 128 * 64;  // 16 KB of bf16 per tile and plane

// element index (in bf16 units, same for both planes) of activation element (row, k)
__device__ __forceinline__ size_t img_index(const ActImg& a, int row, int k) {
  return ((size_t)(row >> 7) * a.kblocks + (k >> 6)) * IMG_TILE_ELEMS + ptx::sw128_offset(row & 127, k & 63) / 2;
}
__device__ __forceinline__ void img_store1(const ActImg& a, int row, int k, float v) {
  __nv_bfloat16 h, l;
  ptx::split_bf16(v, h, l);
  const size_t i = img_index(a, row, k);
  a.hi[i] = h;
  a.lo[i] = l;
}
// 4 consecutive k (k % 4 == 0): one 8-byte store per plane
__device__ __forceinline__ void img_store4(const ActImg& a, int row, int k, float v0, float v1, float v2, float v3) {
  __nv_bfloat16 h[4], l[4];
  ptx::split_bf16(v0, h[0], l[0]);
  ptx::split_bf16(v1, h[1], l[1]);
  ptx::split_bf16(v2, h[2], l[2]);
  ptx::split_bf16(v3, h[3], l[3]);
  const size_t i = img_index(a, row, k);
  *reinterpret_cast<uint2*>(a.hi + i) = make_uint2(ptx::pack_bf16(h[0], h[1]), ptx::pack_bf16(h[2], h[3]));
  *reinterpret_cast<uint2*>(a.lo + i) = make_uint2(ptx::pack_bf16(l[0], l[1]), ptx::pack_bf16(l[2], l[3]));
}
// 8 consecutive k (k % 8 == 0): one 16-byte store per plane
__device__ __forceinline__ void img_store8(const ActImg& a, int row, int k, const float (&v)[8]) {
  __nv_bfloat16 h[8], l[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) ptx::split_bf16(v[e], h[e], l[e]);
  const size_t i = img_index(a, row, k);
  *reinterpret_cast<uint4*>(a.hi + i) =
      make_uint4(ptx::pack_bf16(h[0], h[1]), ptx::pack_bf16(h[2], h[3]), ptx::pack_bf16(h[4], h[5]), ptx::pack_bf16(h[6], h[7]));
  *reinterpret_cast<uint4*>(a.lo + i) =
      make_uint4(ptx::pack_bf16(l[0], l[1]), ptx::pack_bf16(l[2], l[3]), ptx::pack_bf16(l[4], l[5]), ptx::pack_bf16(l[6], l[7]));
}

}  // namespace ltr
