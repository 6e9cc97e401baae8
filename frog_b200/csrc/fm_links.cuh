// fm_links.cuh -- consumer hand-off: the adjacency `bin/frog` builds when it reads pairs.bin.
//
// ImageGroup::readPairs (registration/imageGroup.cxx:1386-1411) walks the pair blocks in file order and, for every
// entry (p1 of image1, p2 of image2), appends {image2, p2} to image1's point p1 and {image1, p1} to image2's point p2
// (Point::links, registration/point.h:11-28).  The ORDER of a point's links is the order of those push_backs --
// frog's reservoir sampling and floating-point sums depend on it -- so the device build reproduces it:
//   q = 2 * (entries of the blocks before this one in file order + index of the entry in its block) + half
// numbers every half-link in push_back order, half-links are dropped into their point's slice of a CSR array in any
// order (atomics), and each point's slice is then sorted by q.  Input: the compacted match lists already resident on
// the device after fm_match -- a consumer in the same process never reads pairs.bin back.
#pragma once
#include "fm_common.cuh"
#include "fm_exact.cuh"  // find_segment

namespace fm {

struct HalfLink {
  unsigned long long q;  // position in the reference's push_back order
  uint32_t image, point; // Link{image, point}
};

struct LinkBlock {       // one pair block, in file order
  uint64_t src;          // first entry of its list in the result's pair array
  uint64_t q0;           // entries of all earlier blocks (file order)
  uint32_t count;
  uint32_t image1, image2;
  uint32_t base1, base2; // global point id of point 0 of image1 / image2
  uint32_t pad_;
};

// entry_off: exclusive prefix of block counts in file order (n_blocks + 1), for the flat entry index -> block search
template <bool kFill>
__global__ void __launch_bounds__(256)
links_scatter_kernel(const LinkBlock* __restrict__ blocks, const unsigned long long* __restrict__ entry_off, uint32_t n_blocks,
                     unsigned long long n_entries, const uint2* __restrict__ pairs, uint32_t* __restrict__ deg,
                     const unsigned long long* __restrict__ off, uint32_t* __restrict__ cursor, HalfLink* __restrict__ tmp) {
  const unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_entries) return;
  uint32_t lo = 0, hi = n_blocks;  // largest b with entry_off[b] <= e
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (entry_off[mid] <= e) lo = mid; else hi = mid;
  }
  const LinkBlock b = blocks[lo];
  const unsigned long long k = e - entry_off[lo];
  const uint2 m = pairs[b.src + k];
  const uint32_t g1 = b.base1 + m.x, g2 = b.base2 + m.y;
  if (!kFill) {
    atomicAdd(deg + g1, 1u);
    atomicAdd(deg + g2, 1u);
  } else {
    const unsigned long long q = 2ull * (b.q0 + k);
    tmp[off[g1] + atomicAdd(cursor + g1, 1u)] = HalfLink{q, b.image2, m.y};      // image1's point gets {image2, p2}
    tmp[off[g2] + atomicAdd(cursor + g2, 1u)] = HalfLink{q + 1ull, b.image1, m.x};  // image2's point gets {image1, p1}
  }
}

// One CTA: exclusive prefix of the degrees (n + 1 entries out).
__global__ void __launch_bounds__(1024)
links_scan_kernel(const uint32_t* __restrict__ deg, uint32_t n, unsigned long long* __restrict__ off) {
  __shared__ unsigned long long s_part[1024];
  const uint32_t per = (n + 1023u) / 1024u;
  const uint32_t i0 = min(n, threadIdx.x * per), i1 = min(n, i0 + per);
  unsigned long long sum = 0;
  for (uint32_t i = i0; i < i1; i++) sum += deg[i];
  s_part[threadIdx.x] = sum;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const unsigned long long v = threadIdx.x >= (uint32_t)o ? s_part[threadIdx.x - o] : 0ull;
    __syncthreads();
    s_part[threadIdx.x] += v;
    __syncthreads();
  }
  unsigned long long run = s_part[threadIdx.x] - sum;
  for (uint32_t i = i0; i < i1; i++) { off[i] = run; run += deg[i]; }
  if (threadIdx.x == 1023) off[n] = s_part[1023];
}

// One thread per point: order its slice by q (insertion sort: degrees are small), then emit (image, point).
__global__ void __launch_bounds__(128)
links_sort_kernel(const unsigned long long* __restrict__ off, uint32_t n_points, HalfLink* __restrict__ tmp,
                  uint2* __restrict__ out) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_points) return;
  const unsigned long long a = off[p], b = off[p + 1];
  for (unsigned long long i = a + 1; i < b; i++) {
    const HalfLink x = tmp[i];
    unsigned long long j = i;
    while (j > a && tmp[j - 1].q > x.q) { tmp[j] = tmp[j - 1]; j--; }
    tmp[j] = x;
  }
  for (unsigned long long i = a; i < b; i++) out[i] = make_uint2(tmp[i].image, tmp[i].point);
}

}  // namespace fm
