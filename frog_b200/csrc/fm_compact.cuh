// fm_compact.cuh -- stable stream compaction of accepted matches (replaces push_back, match.cpp:323-327).
//
// Input : rowres[row] = matched column id or kNone, rows laid out task after task.
// Output: per task a dense list of (first, second) uint32 pairs ordered by row -- exactly the
//         order ComputeMatches appends them in -- written at the running output offset so that
//         consecutive tasks (and the two directions of a -sym pair) are contiguous.
//
// ONE pass, HBM-bound: 4 B read per row, 8 B written per match (+ 4 B per match with distances).
// The rows of a batch are cut into chunks of <= 2048 rows of one task, described by a host-built
// 16-byte record each (no per-chunk searches or dependent loads on the device).  One small CTA per
// chunk (128 threads, 16 KB of shared memory: a dozen CTAs per SM, ~100 KB of loads in flight per
// SM): fully coalesced 4-byte loads (sixteen per thread), warp ballots for the ranks, the exclusive
// prefix of the chunk totals by decoupled look-back over one 64-bit status word per chunk (launch
// epoch | state | value: no clearing between launches; 512 predecessors inspected per round, so
// even the first wave of a launch -- where nobody has a finished neighbour yet -- needs three
// rounds at most), staging in shared memory in output order, 16-byte stores.  CTAs of a 1-D grid
// are dispatched in index order, so a chunk's predecessors are always running or done (the same
// assumption every single-pass scan makes).
#pragma once
#include "fm_common.cuh"

namespace fm {

constexpr int kCompactThreads = 128;
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

constexpr int kCompactLook = 4;  // status words inspected per thread and look-back round

template <bool kDist>
__global__ void __launch_bounds__(kCompactThreads, 10)
compact_kernel(const CompactArgs a) {
  __shared__ uint32_t s_slice[kCompactSlices];  // slice totals, then their exclusive prefix
  __shared__ uint32_t s_total, s_first, s_unready, s_sum;
  __shared__ __align__(16) uint2 s_stage[kCompactChunk + 2];

  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t chunk = blockIdx.x;
  volatile unsigned long long* status = a.status;
  // written again only by the CTA of the LAST chunk, after every other chunk has published -- i.e. after every CTA has
  // passed this load
  const unsigned long long base0 = *reinterpret_cast<volatile unsigned long long*>(a.running_total);
  const uint4 dv = __ldg(reinterpret_cast<const uint4*>(a.chunks) + chunk);
  const ChunkDesc d{dv.x, dv.y, dv.z, dv.w};
  const uint32_t n = d.n_swap & 0x7FFFFFFFu;
  const uint32_t* src = a.rowres + d.row_abs;

  // ---- coalesced loads: iteration i covers rows i*128 .. +127 of the chunk, a warp 32 consecutive rows ----
  uint32_t m[kCompactPer], rank[kCompactPer];
#pragma unroll
  for (int i = 0; i < kCompactPer; i++) {
    const uint32_t r = i * kCompactThreads + tid;
    m[i] = r < n ? __ldcs(src + r) : kNone;
  }
  if (tid == 0) { s_first = 0xFFFFFFFFu; s_unready = 0xFFFFFFFFu; s_sum = 0u; }
#pragma unroll
  for (int i = 0; i < kCompactPer; i++) {
    const uint32_t b = __ballot_sync(0xffffffffu, m[i] != kNone);
    rank[i] = __popc(b & ((1u << lane) - 1u));
    if (lane == 0) s_slice[i * (kCompactThreads / 32) + warp] = __popc(b);
  }
  __syncthreads();

  // ---- warp 0: exclusive scan of the 64 slice totals (row order = slice order), publish the chunk total ----
  if (warp == 0) {
    uint32_t carry = 0;
#pragma unroll
    for (int k = 0; k < kCompactSlices / 32; k++) {
      const uint32_t v = s_slice[32 * k + lane];
      uint32_t inc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t nb = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o) inc += nb;
      }
      s_slice[32 * k + lane] = carry + inc - v;
      carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) {
      s_total = carry;
      // successors can add this chunk's total while it is still looking back itself
      // (the 64-bit word is all a successor reads from this chunk: no fence is needed around the store)
      status[chunk] = chunk_word(a.epoch, chunk == 0 ? kChunkPrefix : kChunkAggregate, carry);
    }
  }
  __syncthreads();
  const uint32_t total = s_total;

  // ---- decoupled look-back: thread t inspects the predecessors at distance t + 128 r, r = 0..3, per round ----
  uint32_t excl = 0;
  {
    int64_t look = (int64_t)chunk - 1;
    while (look >= 0) {  // CTA-uniform
      unsigned long long w[kCompactLook];
      uint32_t unready = 0xFFFFFFFFu, mine = 0xFFFFFFFFu;  // smallest distance not published yet / holding a prefix
#pragma unroll
      for (int r = kCompactLook - 1; r >= 0; r--) {
        const int64_t idx = look - (int64_t)(tid + r * kCompactThreads);
        w[r] = idx >= 0 ? status[idx] : chunk_word(a.epoch, kChunkAggregate, 0u);
        const bool ok = (uint32_t)(w[r] >> 34) == a.epoch && ((w[r] >> 32) & 3ull) != 0ull;
        if (!ok) unready = tid + r * kCompactThreads;
        else if (((w[r] >> 32) & 3ull) == kChunkPrefix) mine = tid + r * kCompactThreads;
      }
      if (unready != 0xFFFFFFFFu) atomicMin(&s_unready, unready);
      if (mine != 0xFFFFFFFFu) atomicMin(&s_first, mine);
      __syncthreads();
      const uint32_t first = s_first;
      // only the predecessors NEARER than the first inclusive prefix matter: poll again if one of them has not published
      if (s_unready < first) {  // CTA-uniform
        __syncthreads();
        if (tid == 0) { s_unready = 0xFFFFFFFFu; s_first = 0xFFFFFFFFu; }
        __syncthreads();
        continue;
      }
      uint32_t val = 0;
#pragma unroll
      for (int r = 0; r < kCompactLook; r++)
        if (tid + r * kCompactThreads <= first) val += (uint32_t)w[r];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
      if (lane == 0 && val) atomicAdd(&s_sum, val);
      __syncthreads();
      if (first != 0xFFFFFFFFu) break;
      look -= kCompactLook * kCompactThreads;  // (s_unready is 0xFFFFFFFF here: every word of the window was published)
    }
    excl = s_sum;
    if (tid == 0) {
      if (chunk != 0) status[chunk] = chunk_word(a.epoch, kChunkPrefix, excl + total);
      if (total) atomicAdd(a.pair_count + d.pair, total);
      if (chunk == a.n_chunks - 1) *a.running_total = base0 + excl + total;
    }
  }
  if (total == 0) return;  // CTA-uniform

  // ---- stage in output order ----
  const unsigned long long dst0 = base0 + excl;    // index of this chunk's first output pair
  const uint32_t shift = (uint32_t)(dst0 & 1ull);  // staged one slot late when the destination is not 16-byte aligned
  const bool swap = d.n_swap >> 31;
#pragma unroll
  for (int i = 0; i < kCompactPer; i++) {
    if (m[i] != kNone) {
      const uint32_t row = d.row_local + i * kCompactThreads + tid;
      const uint32_t pos = s_slice[i * (kCompactThreads / 32) + warp] + rank[i];
      s_stage[pos + shift] = swap ? make_uint2(row, m[i]) : make_uint2(m[i], row);
    }
  }
  __syncthreads();
  // ---- 16-byte stores: [dst0 - shift, ...) is 16-byte aligned; the slot before the first and the one after the last
  // pair are not this chunk's
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
        s_dist[s_slice[i * (kCompactThreads / 32) + warp] + rank[i]] = a.rowdist[d.row_abs + i * kCompactThreads + tid];
    }
    __syncthreads();
    float* dd = a.out_dist + dst0;
    for (uint32_t q = tid; q < total; q += kCompactThreads) dd[q] = s_dist[q];
  }
}

}  // namespace fm
