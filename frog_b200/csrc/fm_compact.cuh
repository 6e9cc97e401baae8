// fm_compact.cuh -- stable stream compaction of accepted matches (replaces push_back, match.cpp:323-327).
//
// Input : rowres[row] = matched column id or kNone, rows laid out task after task.
// Output: per task a dense list of (first, second) uint32 pairs ordered by row -- exactly the
//         order ComputeMatches appends them in -- written at the running output offset so that
//         consecutive tasks (and the two directions of a -sym pair) are contiguous.
//
// ONE pass, HBM-bound: 4 B read per row, 8 B written per match (+ 4 B per match with distances).
// The rows of a batch are cut into chunks of <= 4096 rows of one task, described by a host-built
// 16-byte record each (no per-chunk searches or dependent loads on the device).  A grid of
// PERSISTENT CTAs -- never more than fit on the chip at once, so every CTA is resident and the
// waits below cannot deadlock -- walks the chunks round-robin: chunk = blockIdx.x + k * gridDim.x.
// Per chunk: fully coalesced 4-byte loads (sixteen per thread, the NEXT chunk's issued before this
// one is processed), warp ballots for the ranks, the exclusive prefix of the chunk totals by
// decoupled look-back over one 64-bit status word per chunk (launch epoch | state | value: no
// clearing between launches; 256 predecessors inspected per round), staging in shared memory in
// output order, 16-byte stores.
#pragma once
#include "fm_common.cuh"

namespace fm {

constexpr int kCompactThreads = 256;
constexpr int kCompactPer = 16;                                 // rows per thread
constexpr int kCompactChunk = kCompactThreads * kCompactPer;    // rows per chunk
constexpr int kCompactSlices = kCompactPer * (kCompactThreads / 32);  // (iteration, warp) slices of 32 rows

// One chunk: rows [row_abs, row_abs + n) of the batch's rowres array = rows row_local.. of their task.
struct ChunkDesc {
  uint32_t row_abs;
  uint32_t n_swap;     // [30:0] rows in the chunk, [31] the task is a -sym reverse pass: emit (row, match)
  uint32_t row_local;  // index, inside its task, of the chunk's first row
  uint32_t pair;       // caller's pair index the task belongs to
};

// Status word of a chunk: [63:34] launch epoch, [33:32] state, [31:0] value.
constexpr unsigned long long kChunkAggregate = 1ull;  // value = the chunk's own total
constexpr unsigned long long kChunkPrefix = 2ull;     // value = inclusive prefix up to and including the chunk
__device__ __forceinline__ unsigned long long chunk_word(uint32_t epoch, unsigned long long state, uint32_t value) {
  return ((unsigned long long)epoch << 34) | (state << 32) | value;
}

struct CompactArgs {
  const ChunkDesc* chunks;
  uint32_t n_chunks;
  const uint32_t* rowres;
  const float* rowdist;           // kDist only: squared distance of the row's match
  unsigned long long* status;     // n_chunks words (never cleared: epoch-tagged)
  uint32_t epoch;                 // 30 bits, different from the previous launches that used `status`
  uint32_t* pair_count;           // per caller pair: += matches
  unsigned long long* running_total;  // matches written before this batch; += this batch's on exit
  uint2* out_pairs;
  float* out_dist;                // kDist only
};

template <bool kDist>
__global__ void __launch_bounds__(kCompactThreads)
compact_kernel(const CompactArgs a) {
  __shared__ unsigned long long s_base0;
  __shared__ uint32_t s_slice[kCompactSlices];  // slice totals, then their exclusive prefix
  __shared__ uint32_t s_total, s_excl;
  __shared__ uint32_t s_warpflag[kCompactThreads / 32], s_warpsum[kCompactThreads / 32];
  __shared__ __align__(16) uint2 s_stage[kCompactChunk + 2];

  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t G = gridDim.x;
  // written again only by the CTA that finishes the LAST chunk, after every other chunk has published -- i.e. after
  // every CTA has passed this load
  if (tid == 0) s_base0 = *reinterpret_cast<volatile unsigned long long*>(a.running_total);
  volatile unsigned long long* status = a.status;

  auto load_desc = [&](uint32_t c) {
    ChunkDesc d{0u, 0u, 0u, 0u};
    if (c < a.n_chunks) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(a.chunks) + c);
      d = ChunkDesc{v.x, v.y, v.z, v.w};
    }
    return d;
  };
  auto load_rows = [&](const ChunkDesc& d, uint32_t (&m)[kCompactPer]) {
    const uint32_t n = d.n_swap & 0x7FFFFFFFu;
    const uint32_t* src = a.rowres + d.row_abs;
#pragma unroll
    for (int i = 0; i < kCompactPer; i++) {
      const uint32_t r = i * kCompactThreads + tid;
      m[i] = r < n ? __ldcs(src + r) : kNone;
    }
  };

  uint32_t chunk = blockIdx.x;
  ChunkDesc d_cur = load_desc(chunk), d_next = load_desc(chunk + G);
  uint32_t m[kCompactPer], m_next[kCompactPer], rank[kCompactPer];
  load_rows(d_cur, m);

  for (; chunk < a.n_chunks; chunk += G) {
    const ChunkDesc d_next2 = load_desc(chunk + 2 * G);
    load_rows(d_next, m_next);  // in flight while this chunk is ranked, scanned and written

#pragma unroll
    for (int i = 0; i < kCompactPer; i++) {
      const uint32_t b = __ballot_sync(0xffffffffu, m[i] != kNone);
      rank[i] = __popc(b & ((1u << lane) - 1u));
      if (lane == 0) s_slice[i * (kCompactThreads / 32) + warp] = __popc(b);
    }
    __syncthreads();

    // ---- warp 0: exclusive scan of the 128 slice totals (row order = slice order) ----
    if (warp == 0) {
      uint32_t carry = 0;
#pragma unroll
      for (int k = 0; k < kCompactSlices / 32; k++) {
        const uint32_t v = s_slice[32 * k + lane];
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= (uint32_t)o) inc += n;
        }
        s_slice[32 * k + lane] = carry + inc - v;
        carry += __shfl_sync(0xffffffffu, inc, 31);
      }
      if (lane == 0) {
        s_total = carry;
        // publish right away: successors add this chunk's total while it is still looking back itself
        status[chunk] = chunk_word(a.epoch, chunk == 0 ? kChunkPrefix : kChunkAggregate, carry);
        __threadfence();
      }
    }
    __syncthreads();
    const uint32_t total = s_total;

    // ---- decoupled look-back: thread i inspects chunk - 1 - i (- 256 per round) ----
    {
      uint32_t excl = 0;  // meaningful in thread 0
      int64_t look = (int64_t)chunk - 1;
      while (look >= 0) {  // CTA-uniform
        const int64_t idx = look - tid;
        unsigned long long w = 0;
        bool ready = true;
        if (idx >= 0) {
          w = status[idx];
          ready = (uint32_t)(w >> 34) == a.epoch && ((w >> 32) & 3ull) != 0ull;
        }
        if (!__syncthreads_and(ready)) continue;  // a predecessor has not published yet: poll again
        const bool is_prefix = idx >= 0 && ((w >> 32) & 3ull) == kChunkPrefix;
        const uint32_t pm = __ballot_sync(0xffffffffu, is_prefix);
        if (lane == 0) s_warpflag[warp] = pm;
        __syncthreads();
        uint32_t first = kCompactThreads;  // thread index of the nearest chunk holding an inclusive prefix
#pragma unroll
        for (int k = kCompactThreads / 32 - 1; k >= 0; k--)
          if (s_warpflag[k]) first = 32 * k + (uint32_t)__ffs(s_warpflag[k]) - 1u;
        uint32_t val = (idx >= 0 && tid <= first) ? (uint32_t)w : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
        if (lane == 0) s_warpsum[warp] = val;
        __syncthreads();
        if (tid == 0) {
#pragma unroll
          for (int k = 0; k < kCompactThreads / 32; k++) excl += s_warpsum[k];
        }
        if (first < (uint32_t)kCompactThreads) break;
        look -= kCompactThreads;
      }
      if (tid == 0) {
        if (chunk != 0) {
          status[chunk] = chunk_word(a.epoch, kChunkPrefix, excl + total);
          __threadfence();
        }
        s_excl = excl;
        if (total) atomicAdd(a.pair_count + d_cur.pair, total);
        if (chunk == a.n_chunks - 1) *a.running_total = s_base0 + excl + total;
      }
    }
    __syncthreads();

    if (total != 0) {  // CTA-uniform
      // ---- stage in output order ----
      const unsigned long long dst0 = s_base0 + s_excl;  // index of this chunk's first output pair
      const uint32_t shift = (uint32_t)(dst0 & 1ull);    // staged one slot late when the destination is not 16-byte aligned
      const bool swap = d_cur.n_swap >> 31;
#pragma unroll
      for (int i = 0; i < kCompactPer; i++) {
        if (m[i] != kNone) {
          const uint32_t row = d_cur.row_local + i * kCompactThreads + tid;
          const uint32_t pos = s_slice[i * (kCompactThreads / 32) + warp] + rank[i];
          s_stage[pos + shift] = swap ? make_uint2(row, m[i]) : make_uint2(m[i], row);
        }
      }
      __syncthreads();
      // ---- 16-byte stores: [dst0 - shift, ...) is 16-byte aligned; the slot before the first and the one after the
      // last pair are not this chunk's
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
      if (kDist) {  // the distances take the same route through the (now free) staging area
        __syncthreads();
        float* s_dist = reinterpret_cast<float*>(s_stage);
#pragma unroll
        for (int i = 0; i < kCompactPer; i++) {
          if (m[i] != kNone)
            s_dist[s_slice[i * (kCompactThreads / 32) + warp] + rank[i]] = a.rowdist[d_cur.row_abs + i * kCompactThreads + tid];
        }
        __syncthreads();
        float* dd = a.out_dist + dst0;
        for (uint32_t q = tid; q < total; q += kCompactThreads) dd[q] = s_dist[q];
      }
      __syncthreads();  // the staging area and s_slice are rewritten by the next chunk
    }

    d_cur = d_next;
    d_next = d_next2;
#pragma unroll
    for (int i = 0; i < kCompactPer; i++) m[i] = m_next[i];
  }
}

}  // namespace fm
