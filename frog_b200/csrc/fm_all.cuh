// fm_all.cuh -- the reference's `-all` mode (matchAll == true in ComputeMatches, match.cpp:295-300),
// bug-compatible.  In this mode every gated-in column whose distance is under the threshold emits a
// pair, but the pair does not name that column: it names `match`, the running nearest among the
// columns that were NOT under the threshold (the top-2 update, :303-313, is the `else` branch of
// the emission), and `match` is declared outside the row loop (:259), so a row that has not met
// such a column yet emits the value left behind by an earlier row.  Reproducing that needs
//   pass 1  (one thread per row, columns in ascending original order, the reference's arithmetic):
//           the row's emission count and its final `match` (or "none");
//   scan    (one thread per task, rows in order): the value of `match` each row starts with and
//           the row's offset in the task's list;
//   pass 2  the same walk again, now writing (match, row) -- or (row, match) for the -sym reverse
//           pass (:297) -- at the row's offset.
// Lists are as long as the data makes them (up to N_first * N_second pairs per image pair).
#pragma once
#include "fm_common.cuh"
#include "fm_exact.cuh"

namespace fm {

// kEmit == false: pass 1, writes row_count / row_final.  kEmit == true: pass 2, reads row_carry /
// row_off (offset inside the task's list) / task_base (offset of the task's list in `out`).
template <int D_T, bool kEmit>
__global__ void __launch_bounds__(kExactRows)
match_all_kernel(const ImageDev* __restrict__ images, const Task* __restrict__ tasks,
                 const uint32_t* __restrict__ task_blk_off, uint32_t n_tasks, float thr,
                 uint32_t* __restrict__ row_count, uint32_t* __restrict__ row_final,
                 const uint32_t* __restrict__ row_carry, const unsigned long long* __restrict__ row_off,
                 const unsigned long long* __restrict__ task_base, uint2* __restrict__ out) {
  extern __shared__ float smem[];
  const uint32_t t = find_segment(task_blk_off, n_tasks, blockIdx.x);
  const Task task = tasks[t];
  const ImageDev A = images[task.col_img];
  const ImageDev B = images[task.row_img];
  constexpr int d = D_T;  // other descriptor lengths: exact_generic_kernel<1|2> (fm_generic.cuh)
  float* s_desc = smem;                      // [kExactCols][d]
  float* s_scale = s_desc + kExactCols * d;  // [kExactCols]
  float* s_lap = s_scale + kExactCols;       // [kExactCols]

  const uint32_t row = (blockIdx.x - task_blk_off[t]) * kExactRows + threadIdx.x;
  const bool active = row < B.n;
  const bool swap = task.flags & kTaskSwap;

  float r[D_T];
  float sc = 1.f, lp = 0.f;
  if (active) {
    sc = B.scale[row];
    lp = B.lap[row];
    const float4* src = reinterpret_cast<const float4*>(B.desc + (size_t)row * D_T);
#pragma unroll
    for (int q = 0; q < D_T / 4; q++) {
      float4 v = __ldg(src + q);
      r[4 * q] = v.x; r[4 * q + 1] = v.y; r[4 * q + 2] = v.z; r[4 * q + 3] = v.w;
    }
  }

  float d1 = FLT_MAX;
  uint32_t match = (kEmit && active) ? row_carry[task.row_off + row] : 0u;
  uint32_t count = 0;
  uint2* o = (kEmit && active) ? out + task_base[t] + row_off[task.row_off + row] : nullptr;
  for (uint32_t c0 = 0; c0 < A.n; c0 += kExactCols) {
    const uint32_t nc = min((uint32_t)kExactCols, A.n - c0);
    __syncthreads();
    for (uint32_t idx = threadIdx.x; idx < nc * (uint32_t)d; idx += kExactRows)
      s_desc[idx] = __ldg(A.desc + (size_t)c0 * d + idx);
    if (threadIdx.x < nc) {
      s_scale[threadIdx.x] = A.scale[c0 + threadIdx.x];
      s_lap[threadIdx.x] = A.lap[c0 + threadIdx.x];
    }
    __syncthreads();
    if (!active) continue;
    for (uint32_t c = 0; c < nc; c++) {
      if (lp != s_lap[c]) continue;                    // match.cpp:270
      if (scale_gate_fails(sc, s_scale[c])) continue;  // match.cpp:273-275
      float acc = 0.f;
      const float4* col = reinterpret_cast<const float4*>(s_desc + c * D_T);
#pragma unroll
      for (int q = 0; q < D_T / 4; q++) {
        float4 v = col[q];
        float e;
        e = __fsub_rn(r[4 * q], v.x);     acc = __fadd_rn(acc, __fmul_rn(e, e));
        e = __fsub_rn(r[4 * q + 1], v.y); acc = __fadd_rn(acc, __fmul_rn(e, e));
        e = __fsub_rn(r[4 * q + 2], v.z); acc = __fadd_rn(acc, __fmul_rn(e, e));
        e = __fsub_rn(r[4 * q + 3], v.w); acc = __fadd_rn(acc, __fmul_rn(e, e));
      }
      if (__fsqrt_rn(acc) < thr) {  // match.cpp:295: emit, and do NOT take part in the top-2 update
        if (kEmit) o[count] = swap ? make_uint2(row, match) : make_uint2(match, row);
        count++;
      } else if (acc < d1) {  // match.cpp:303-307 (d2 never influences `match`)
        d1 = acc;
        match = c0 + c;
      }
    }
  }
  if (active && !kEmit) {
    row_count[task.row_off + row] = count;
    row_final[task.row_off + row] = d1 != FLT_MAX ? match : kNone;  // kNone: the row left `match` untouched
  }
}

// One thread per task walks its rows in order: `match` as each row finds it (0 before the first row,
// match.cpp:259) and the exclusive prefix sum of the emission counts.
__global__ void all_scan_kernel(const ImageDev* __restrict__ images, const Task* __restrict__ tasks, uint32_t n_tasks,
                                const uint32_t* __restrict__ row_count, const uint32_t* __restrict__ row_final,
                                uint32_t* __restrict__ row_carry, unsigned long long* __restrict__ row_off,
                                unsigned long long* __restrict__ task_total) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tasks) return;
  const Task task = tasks[t];
  const uint32_t n = images[task.row_img].n;
  uint32_t m = 0;
  unsigned long long run = 0;
  for (uint32_t i = 0; i < n; i++) {
    const uint32_t f = row_final[task.row_off + i];
    row_carry[task.row_off + i] = m;
    row_off[task.row_off + i] = run;
    run += row_count[task.row_off + i];
    if (f != kNone) m = f;
  }
  task_total[t] = run;
}

}  // namespace fm
