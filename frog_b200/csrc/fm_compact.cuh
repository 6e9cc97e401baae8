// fm_compact.cuh -- stable stream compaction of accepted matches (replaces push_back, match.cpp:323-327).
//
// Input : rowres[row] = matched column id or kNone, rows laid out task after task.
// Output: per task a dense list of (first, second) uint32 pairs ordered by row -- exactly the
//         order ComputeMatches appends them in -- written at the running output offset so that
//         consecutive tasks (and the two directions of a -sym pair) are contiguous.
// Three coalesced passes over 4 B/row: count per chunk, one-block exclusive scan over chunks,
// scatter with an in-block scan.  HBM-bound: 8 B read per row + 8 B written per match.
#pragma once
#include "fm_common.cuh"
#include "fm_exact.cuh"  // find_segment

namespace fm {

constexpr int kCompactThreads = 256;
constexpr int kCompactChunk = 2048;  // rows per CTA (8 per thread)

__device__ __forceinline__ uint32_t block_exclusive_scan_256(uint32_t v, uint32_t* total, uint32_t* s_warp) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= (uint32_t)o) inc += n;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < kCompactThreads / 32 ? s_warp[lane] : 0;
    uint32_t winc = w;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      uint32_t n = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= (uint32_t)o) winc += n;
    }
    if (lane < kCompactThreads / 32) s_warp[lane] = winc - w;
    if (lane == kCompactThreads / 32 - 1) s_warp[8] = winc;
  }
  __syncthreads();
  *total = s_warp[8];
  return inc - v + s_warp[warp];
}

// chunk_off[t] = first chunk of task t (exclusive prefix, n_tasks + 1 entries, host-built).
__global__ void __launch_bounds__(kCompactThreads)
compact_count_kernel(const ImageDev* __restrict__ images, const Task* __restrict__ tasks,
                     const uint32_t* __restrict__ chunk_off, uint32_t n_tasks,
                     const uint32_t* __restrict__ rowres, uint32_t* __restrict__ chunk_count) {
  __shared__ uint32_t s_warp[9];
  const uint32_t t = find_segment(chunk_off, n_tasks, blockIdx.x);
  const Task task = tasks[t];
  const uint32_t n_rows = images[task.row_img].n;
  const uint32_t base = (blockIdx.x - chunk_off[t]) * kCompactChunk;
  uint32_t c = 0;
#pragma unroll
  for (int i = 0; i < kCompactChunk / kCompactThreads; i++) {
    uint32_t row = base + i * kCompactThreads + threadIdx.x;
    if (row < n_rows) c += rowres[task.row_off + row] != kNone;
  }
  uint32_t total;
  block_exclusive_scan_256(c, &total, s_warp);
  if (threadIdx.x == 0) chunk_count[blockIdx.x] = total;
}

// One CTA: exclusive scan of chunk counts -> chunk_out (offsets relative to *running_total),
// per-pair counts, and the new running total.  pair_of_task maps a batch task to its global pair.
__global__ void __launch_bounds__(1024)
compact_scan_kernel(const uint32_t* __restrict__ chunk_count, uint32_t n_chunks,
                    const uint32_t* __restrict__ chunk_off, uint32_t n_tasks,
                    const uint32_t* __restrict__ pair_of_task, uint64_t* __restrict__ chunk_out,
                    uint32_t* __restrict__ pair_count, unsigned long long* __restrict__ running_total) {
  __shared__ unsigned long long s_part[1024];
  __shared__ unsigned long long s_base;
  const unsigned long long base0 = *running_total;
  // per-thread contiguous slice
  const uint32_t per = (n_chunks + 1023) / 1024;
  const uint32_t lo = min(n_chunks, threadIdx.x * per), hi = min(n_chunks, lo + per);
  unsigned long long sum = 0;
  for (uint32_t i = lo; i < hi; i++) sum += chunk_count[i];
  s_part[threadIdx.x] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long run = 0;
    for (int i = 0; i < 1024; i++) { unsigned long long v = s_part[i]; s_part[i] = run; run += v; }
    s_base = run;
  }
  __syncthreads();
  unsigned long long run = base0 + s_part[threadIdx.x];
  for (uint32_t i = lo; i < hi; i++) { chunk_out[i] = run; run += chunk_count[i]; }
  __syncthreads();
  // per-pair counts: task t owns chunks [chunk_off[t], chunk_off[t+1])
  for (uint32_t t = threadIdx.x; t < n_tasks; t += 1024) {
    uint32_t c0 = chunk_off[t], c1 = chunk_off[t + 1];
    unsigned long long end = (c1 < n_chunks) ? chunk_out[c1] : base0 + s_base;
    unsigned long long beg = (c0 < n_chunks) ? chunk_out[c0] : base0 + s_base;
    atomicAdd(&pair_count[pair_of_task[t]], (uint32_t)(end - beg));
  }
  __syncthreads();
  if (threadIdx.x == 0) *running_total = base0 + s_base;
}

__global__ void __launch_bounds__(kCompactThreads)
compact_scatter_kernel(const ImageDev* __restrict__ images, const Task* __restrict__ tasks,
                       const uint32_t* __restrict__ chunk_off, uint32_t n_tasks,
                       const uint32_t* __restrict__ rowres, const uint64_t* __restrict__ chunk_out,
                       uint2* __restrict__ out_pairs) {
  __shared__ uint32_t s_warp[9];
  const uint32_t t = find_segment(chunk_off, n_tasks, blockIdx.x);
  const Task task = tasks[t];
  const uint32_t n_rows = images[task.row_img].n;
  const uint32_t base = (blockIdx.x - chunk_off[t]) * kCompactChunk;
  uint64_t out = chunk_out[blockIdx.x];
  // each thread owns 8 consecutive rows so the output order is the row order
  const uint32_t row0 = base + threadIdx.x * (kCompactChunk / kCompactThreads);
  uint32_t m[kCompactChunk / kCompactThreads];
  uint32_t c = 0;
#pragma unroll
  for (int i = 0; i < kCompactChunk / kCompactThreads; i++) {
    uint32_t row = row0 + i;
    m[i] = row < n_rows ? rowres[task.row_off + row] : kNone;
    c += m[i] != kNone;
  }
  uint32_t total;
  uint32_t pos = block_exclusive_scan_256(c, &total, s_warp);
#pragma unroll
  for (int i = 0; i < kCompactChunk / kCompactThreads; i++) {
    if (m[i] != kNone) {
      uint32_t row = row0 + i;
      out_pairs[out + pos] = (task.flags & 1u) ? make_uint2(row, m[i]) : make_uint2(m[i], row);
      pos++;
    }
  }
}

}  // namespace fm
