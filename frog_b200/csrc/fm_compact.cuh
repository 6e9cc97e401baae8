// fm_compact.cuh -- stable stream compaction of accepted matches (replaces push_back, match.cpp:323-327).
//
// Input : rowres[row] = matched column id or kNone, rows laid out task after task.
// Output: per task a dense list of (first, second) uint32 pairs ordered by row -- exactly the
//         order ComputeMatches appends them in -- written at the running output offset so that
//         consecutive tasks (and the two directions of a -sym pair) are contiguous.
//
// ONE pass, HBM-bound: 4 B read per row, 8 B written per match (+ 4 B per match with distances).
// A CTA owns a chunk of 2048 rows of one task.  Chunks take their number from a ticket counter, so a
// chunk's predecessors are always already running; the exclusive prefix of the chunk totals is
// obtained by decoupled look-back over one 64-bit status word per chunk (launch epoch | state |
// value: no clearing between launches).  Rows are read with fully coalesced 4-byte loads (eight in
// flight per thread), ranked with warp ballots, staged in shared memory in output order and written
// with 16-byte stores.
#pragma once
#include "fm_common.cuh"
#include "fm_exact.cuh"  // find_segment

namespace fm {

constexpr int kCompactThreads = 256;
constexpr int kCompactPer = 8;                                  // rows per thread
constexpr int kCompactChunk = kCompactThreads * kCompactPer;    // rows per CTA
constexpr int kCompactSlices = kCompactPer * (kCompactThreads / 32);  // (iteration, warp) slices of 32 rows

// Status word of a chunk: [63:34] launch epoch, [33:32] state, [31:0] value.
constexpr unsigned long long kChunkAggregate = 1ull;  // value = the chunk's own total
constexpr unsigned long long kChunkPrefix = 2ull;     // value = inclusive prefix up to and including the chunk
__device__ __forceinline__ unsigned long long chunk_word(uint32_t epoch, unsigned long long state, uint32_t value) {
  return ((unsigned long long)epoch << 34) | (state << 32) | value;
}

struct CompactArgs {
  const ImageDev* images;
  const Task* tasks;              // this batch
  const uint32_t* chunk_off;      // first chunk of task t (exclusive prefix, n_tasks + 1 entries, host-built)
  uint32_t n_tasks;
  uint32_t n_chunks;
  const uint32_t* pair_of_task;   // batch task -> caller's pair index
  const uint32_t* rowres;
  const float* rowdist;           // kDist only: squared distance of the row's match
  unsigned long long* status;     // n_chunks words (never cleared: epoch-tagged)
  uint32_t* ticket;               // chunk numbering; reset to 0 by the CTA that draws the last ticket
  uint32_t epoch;                 // 30 bits, different from the previous launches that used `status`
  uint32_t* pair_count;           // per caller pair: += matches
  unsigned long long* running_total;  // matches written before this batch; += this batch's on exit
  uint2* out_pairs;
  float* out_dist;                // kDist only
};

template <bool kDist>
__global__ void __launch_bounds__(kCompactThreads)
compact_kernel(const CompactArgs a) {
  __shared__ uint32_t s_chunk;
  __shared__ unsigned long long s_base0;
  __shared__ uint32_t s_slice[kCompactSlices];  // slice totals, then their exclusive prefix
  __shared__ uint32_t s_total, s_excl;
  __shared__ __align__(16) uint2 s_stage[kCompactChunk + 2];
  __shared__ float s_dist[kDist ? kCompactChunk : 1];

  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    const uint32_t c = atomicAdd(a.ticket, 1u);
    if (c == a.n_chunks - 1) *a.ticket = 0u;  // every ticket of this launch has been drawn
    s_chunk = c;
    s_base0 = *reinterpret_cast<volatile unsigned long long*>(a.running_total);  // written again only by the last chunk
  }
  __syncthreads();
  const uint32_t chunk = s_chunk;
  const uint32_t t = find_segment(a.chunk_off, a.n_tasks, chunk);
  const Task task = a.tasks[t];
  const uint32_t n_rows = a.images[task.row_img].n;
  const uint32_t base = (chunk - a.chunk_off[t]) * kCompactChunk;
  const uint32_t* src = a.rowres + task.row_off;

  // ---- coalesced loads: iteration i covers rows base + i*256 .. +255, a warp 32 consecutive rows ----
  uint32_t m[kCompactPer], rank[kCompactPer];
#pragma unroll
  for (int i = 0; i < kCompactPer; i++) {
    const uint32_t row = base + i * kCompactThreads + tid;
    m[i] = row < n_rows ? __ldcs(src + row) : kNone;
  }
#pragma unroll
  for (int i = 0; i < kCompactPer; i++) {
    const uint32_t b = __ballot_sync(0xffffffffu, m[i] != kNone);
    rank[i] = __popc(b & ((1u << lane) - 1u));
    if (lane == 0) s_slice[i * (kCompactThreads / 32) + warp] = __popc(b);
  }
  __syncthreads();

  // ---- warp 0: exclusive scan of the 64 slice totals (row order = slice order), then the look-back ----
  if (warp == 0) {
    const uint32_t v0 = s_slice[lane], v1 = s_slice[32 + lane];
    uint32_t i0 = v0, i1 = v1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t n0 = __shfl_up_sync(0xffffffffu, i0, o), n1 = __shfl_up_sync(0xffffffffu, i1, o);
      if (lane >= (uint32_t)o) { i0 += n0; i1 += n1; }
    }
    const uint32_t tot0 = __shfl_sync(0xffffffffu, i0, 31);
    const uint32_t total = tot0 + __shfl_sync(0xffffffffu, i1, 31);
    s_slice[lane] = i0 - v0;
    s_slice[32 + lane] = tot0 + i1 - v1;

    volatile unsigned long long* status = a.status;
    if (lane == 0) {
      status[chunk] = chunk_word(a.epoch, chunk == 0 ? kChunkPrefix : kChunkAggregate, total);
      __threadfence();
    }
    // look back 32 chunks at a time: lane l inspects chunk - 1 - l (- 32 per round)
    uint32_t excl = 0;
    int64_t look = (int64_t)chunk - 1;
    while (look >= 0) {
      const int64_t idx = look - lane;
      unsigned long long w = 0;
      bool ready = true;
      if (idx >= 0) {
        w = status[idx];
        ready = (uint32_t)(w >> 34) == a.epoch && ((w >> 32) & 3ull) != 0ull;
      }
      if (!__all_sync(0xffffffffu, ready)) continue;  // a predecessor has not published yet: poll again
      const bool is_prefix = idx >= 0 && ((w >> 32) & 3ull) == kChunkPrefix;
      const uint32_t pmask = __ballot_sync(0xffffffffu, is_prefix);
      const uint32_t first = pmask ? (uint32_t)__ffs(pmask) - 1u : 32u;  // nearest chunk holding an inclusive prefix
      uint32_t v = (idx >= 0 && lane <= first) ? (uint32_t)w : 0u;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      excl += v;
      if (pmask) break;
      look -= 32;
    }
    if (lane == 0) {
      if (chunk != 0) {
        status[chunk] = chunk_word(a.epoch, kChunkPrefix, excl + total);
        __threadfence();
      }
      s_total = total;
      s_excl = excl;
      if (total) atomicAdd(a.pair_count + a.pair_of_task[t], total);
      if (chunk == a.n_chunks - 1) *a.running_total = s_base0 + excl + total;
    }
  }
  __syncthreads();

  // ---- stage in output order ----
  const uint32_t total = s_total;
  if (total == 0) return;  // CTA-uniform
  const unsigned long long dst0 = s_base0 + s_excl;  // index of this chunk's first output pair
  const uint32_t shift = (uint32_t)(dst0 & 1ull);    // staged one slot late when the destination is not 16-byte aligned
  const bool swap = task.flags & kTaskSwap;
#pragma unroll
  for (int i = 0; i < kCompactPer; i++) {
    if (m[i] != kNone) {
      const uint32_t row = base + i * kCompactThreads + tid;
      const uint32_t pos = s_slice[i * (kCompactThreads / 32) + warp] + rank[i];
      s_stage[pos + shift] = swap ? make_uint2(row, m[i]) : make_uint2(m[i], row);
      if (kDist) s_dist[pos] = a.rowdist[task.row_off + row];
    }
  }
  __syncthreads();

  // ---- 16-byte stores: [dst0 - shift, ...) is 16-byte aligned; the slot before the first and after the last pair is skipped
  uint2* dst = a.out_pairs + (dst0 - shift);
  const uint32_t n_slots = total + shift;
  const uint4* st4 = reinterpret_cast<const uint4*>(s_stage);
  for (uint32_t q = tid; q < (n_slots + 1) / 2; q += kCompactThreads) {
    const uint32_t s0 = 2 * q;
    const bool lo_ok = s0 >= shift, hi_ok = s0 + 1 < n_slots;
    if (lo_ok && hi_ok) __stcs(reinterpret_cast<uint4*>(dst) + q, st4[q]);
    else if (lo_ok) dst[s0] = s_stage[s0];
    else if (hi_ok) dst[s0 + 1] = s_stage[s0 + 1];
  }
  if (kDist) {
    float* dd = a.out_dist + dst0;
    for (uint32_t q = tid; q < total; q += kCompactThreads) dd[q] = s_dist[q];
  }
}

}  // namespace fm
