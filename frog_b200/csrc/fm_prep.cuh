// fm_prep.cuh -- per-image preparation, run once at upload (replaces the AoS Point vector the
// reference's readers build, match.cpp:39-48).
//
// Produces, all device-resident:
//   * keypoints ordered by (laplacian value, scale): `perm`, `scale_sorted`, class table.  In that
//     order both reference gates (match.cpp:270, :273-275) select, for any row, one contiguous
//     column interval -- so whole operand tiles outside the band are never scored.
//   * FP16 operand images `rowop` / `colop`, already in the 128-row x 128-byte SWIZZLE_128B K-major
//     tile layout tcgen05.mma reads, so one 16 KB cp.async.bulk lands a ready tile in shared memory.
//     K is padded 48 -> 64: slots 48,49 carry (1, 1) on the row side and the hi/lo FP16 halves of
//     -|b|^2/2 on the column side, so the MMA itself yields  t = a.b - |b|^2/2  = (|a|^2 - d^2)/2.
//   * per-image flags + max squared norm (certification inputs for the FP32 rescoring).
#pragma once
#include <cub/device/device_radix_sort.cuh>

#include "fm_common.cuh"

namespace fm {

constexpr float kMaxNorm2 = 16.0f;  // FP16 operands certified for |desc|^2 <= 16

__device__ __forceinline__ uint32_t float_sortable(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float float_unsortable(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// One thread per keypoint: validity flags, squared norm, 64-bit sort key (laplacian | scale).
__global__ void __launch_bounds__(256)
prep_keys_kernel(const float* __restrict__ desc, const float* __restrict__ scale, const float* __restrict__ lap,
                 uint32_t n, uint32_t d, ImageMeta* __restrict__ meta, unsigned long long* __restrict__ keys,
                 uint32_t* __restrict__ idx, float* __restrict__ norm2) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t flags = 0;
  float n2 = 0.f, dl2 = 0.f;
  if (i < n) {
    const float* p = desc + (size_t)i * d;
    for (uint32_t k = 0; k < d; k++) {
      float v = __ldg(p + k);
      if (!isfinite(v)) flags |= kImgNotFinite;
      n2 = fmaf(v, v, n2);
      const float r = __half2float(__float2half_rn(v)) - v;  // exact: the operand's FP16 rounding residual
      dl2 = fmaf(r, r, dl2);
    }
    float s = scale[i], l = lap[i];
    if (!isfinite(s) || !isfinite(l)) flags |= kImgNotFinite;
    if (!(s > 0.f)) flags |= kImgBadScale;
    if (!(n2 <= kMaxNorm2)) flags |= kImgBigNorm;
    if (l == 0.f) l = 0.f;  // -0.0 == +0.0 for the reference's float compare: one class
    keys[i] = ((unsigned long long)float_sortable(l) << 32) | __float_as_uint(s);
    idx[i] = i;
    norm2[i] = n2;
  }
  // warp-aggregate then one atomic per warp
  float m = isfinite(n2) ? n2 : 0.f;
  float md = isfinite(dl2) ? dl2 : 0.f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    md = fmaxf(md, __shfl_xor_sync(0xffffffffu, md, o));
    flags |= __shfl_xor_sync(0xffffffffu, flags, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (flags) atomicOr(&meta->flags, flags);
    atomicMax(reinterpret_cast<unsigned int*>(&meta->max_norm2), __float_as_uint(m));   // non-negative floats
    atomicMax(reinterpret_cast<unsigned int*>(&meta->max_delta2), __float_as_uint(md));  // order like uints
  }
}

// One CTA: un-zip the sorted (key, idx) pairs and build the laplacian class table.
__global__ void __launch_bounds__(1024)
prep_finish_kernel(const unsigned long long* __restrict__ keys_sorted, uint32_t n, uint32_t d,
                   ImageMeta* __restrict__ meta, float* __restrict__ scale_sorted) {
  __shared__ uint32_t s_pos[kMaxClasses + 1];
  __shared__ uint32_t s_cnt;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  for (uint32_t s = threadIdx.x; s < n; s += blockDim.x) {
    unsigned long long k = keys_sorted[s];
    scale_sorted[s] = __uint_as_float((uint32_t)k);
    if (s == 0 || (uint32_t)(keys_sorted[s - 1] >> 32) != (uint32_t)(k >> 32)) {
      uint32_t slot = atomicAdd(&s_cnt, 1u);
      if (slot <= kMaxClasses) s_pos[slot] = s;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t cnt = s_cnt;
    uint32_t flags = 0;
    if (d != (uint32_t)kD) flags |= kImgBadDim;
    if (cnt > kMaxClasses) { flags |= kImgManyClasses; cnt = kMaxClasses; }
    for (uint32_t a = 1; a < cnt; a++) {  // insertion sort of <= 8 boundaries
      uint32_t v = s_pos[a];
      int b = (int)a - 1;
      while (b >= 0 && s_pos[b] > v) { s_pos[b + 1] = s_pos[b]; b--; }
      s_pos[b + 1] = v;
    }
    meta->n_classes = cnt;
    for (uint32_t a = 0; a < cnt; a++) {
      meta->class_begin[a] = s_pos[a];
      meta->class_lap[a] = float_unsortable((uint32_t)(keys_sorted[s_pos[a]] >> 32));
    }
    for (uint32_t a = cnt; a <= kMaxClasses; a++) meta->class_begin[a] = n;
    if (flags) atomicOr(&meta->flags, flags);
  }
}

// Byte offset of (row r, 16-byte chunk q) inside a 128 x 128 B SWIZZLE_128B K-major tile:
// 8-row x 128 B atoms, chunk index XORed with (row mod 8)  [cute Swizzle<3,4,3>].
__host__ __device__ __forceinline__ uint32_t sw128_offset(uint32_t r, uint32_t q) {
  return r * 128u + ((q ^ (r & 7u)) << 4);
}

// One thread per (sorted keypoint, 16-byte chunk): FP16 operand tiles for both roles.
__global__ void __launch_bounds__(256)
prep_pack_kernel(const float* __restrict__ desc, const float* __restrict__ norm2, const uint32_t* __restrict__ perm,
                 uint32_t n, uint32_t n_pad, uint8_t* __restrict__ rowop, uint8_t* __restrict__ colop) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t s = g >> 3, q = g & 7u;
  if (s >= n_pad) return;
  __align__(16) __half h[8];
#pragma unroll
  for (int k = 0; k < 8; k++) h[k] = __float2half_rn(0.f);
  __align__(16) __half hc[8];
#pragma unroll
  for (int k = 0; k < 8; k++) hc[k] = h[k];
  if (s < n) {
    const uint32_t o = perm[s];
    if (q < 6) {
      const float4* p = reinterpret_cast<const float4*>(desc + (size_t)o * kD + q * 8);
      float4 v0 = __ldg(p), v1 = __ldg(p + 1);
      h[0] = __float2half_rn(v0.x); h[1] = __float2half_rn(v0.y); h[2] = __float2half_rn(v0.z); h[3] = __float2half_rn(v0.w);
      h[4] = __float2half_rn(v1.x); h[5] = __float2half_rn(v1.y); h[6] = __float2half_rn(v1.z); h[7] = __float2half_rn(v1.w);
#pragma unroll
      for (int k = 0; k < 8; k++) hc[k] = h[k];
    } else if (q == 6) {
      h[0] = __float2half_rn(1.f);
      h[1] = __float2half_rn(1.f);
      float x = -0.5f * norm2[o];
      __half hi = __float2half_rn(x);
      hc[0] = hi;
      hc[1] = __float2half_rn(x - __half2float(hi));
    }
  }
  const size_t off = (size_t)(s >> 7) * 16384u + sw128_offset(s & 127u, q);
  *reinterpret_cast<uint4*>(rowop + off) = *reinterpret_cast<const uint4*>(h);
  *reinterpret_cast<uint4*>(colop + off) = *reinterpret_cast<const uint4*>(hc);
}

}  // namespace fm
