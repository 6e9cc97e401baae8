// Device side of libfrogsurf.so: the SURF3D producer's stages as CUDA kernels for sm_100a.
//
// Every kernel here is integer / gather work over the 64-bit integral volume, bound by L2 and HBM
// traffic, not by arithmetic.  None of it is GEMM-shaped, so there is no tensor-core code in this
// file: the rules that matter are coalesced x-fastest access, enough loads in flight per thread and
// grids that fill the 148 SMs.
//
// Exactness: the reference computes in float / double on x86-64 without FMA contraction.  Every
// floating-point operation below that takes part in a result is written as an explicit IEEE
// round-to-nearest intrinsic (__fmul_rn, __dadd_rn, ...) in the reference's evaluation order, so the
// compiler can neither fuse nor reassociate it, whatever -fmad says.  Integer box sums are exact in
// any order (unsigned 64-bit, wrapping like the reference's).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/frogsurf.h"

namespace fs {

typedef unsigned long long u64;

// ---------------------------------------------------------------------------------------------
// cast + shift (vtk3DSURF.cxx:158-176): vtkImageCast to int with ClampOverflow, then
// vtkImageShiftScale with shift = -range[0], both through double and static_cast<int>.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int clamp_to_int(double v) {
  if (v > 2147483647.0) v = 2147483647.0;
  if (v < -2147483648.0) v = -2147483648.0;
  return __double2int_rz(v);
}

template <typename T>
__device__ __forceinline__ int cast_shift(T v, double shift) {
  const int c = clamp_to_int((double)v);
  return clamp_to_int(__dmul_rn(__dadd_rn((double)c, shift), 1.0));
}

// per-block minimum of the volume, as double (exact for every supported voxel type)
template <typename T>
__global__ void volume_min_kernel(const T* __restrict__ in, size_t n, double* __restrict__ block_min) {
  double m = 1.0e300;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double v = (double)in[i];
    m = v < m ? v : m;
  }
  for (int o = 16; o; o >>= 1) {
    const double t = __shfl_xor_sync(0xffffffffu, m, o);
    m = t < m ? t : m;
  }
  __shared__ double wm[32];
  if ((threadIdx.x & 31) == 0) wm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (unsigned w = 1; w < (blockDim.x >> 5); w++) m = wm[w] < m ? wm[w] : m;
    block_min[blockIdx.x] = m;
  }
}

// ---------------------------------------------------------------------------------------------
// Integral volume (integral.cxx:11-121).  The reference runs three in-place passes (x, y, z); sums
// of unsigned 64-bit integers are exact in any order, so this does x and y in one pass per z slice
// (row scan + running column sums kept in shared memory) and z in a second, fully coalesced pass.
// HBM traffic: voxel read + 8 B write, then 8 B read + 8 B write = 28 B / voxel for 4-byte voxels.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(512) integral_xy_kernel(const T* __restrict__ in, int32_t* __restrict__ cast_out,
                                                           u64* __restrict__ out, int nx, int ny, double shift) {
  extern __shared__ u64 fs_sm[];
  u64* col = fs_sm;        // nx running column sums: the 2-D integral of the row above
  u64* wtot = fs_sm + nx;  // 32 warp totals
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const size_t slice = (size_t)nx * ny, base = (size_t)blockIdx.x * slice;
  for (int x = tid; x < nx; x += blockDim.x) col[x] = 0;
  __syncthreads();
  for (int y = 0; y < ny; y++) {
    u64 carry = 0;
    for (int x0 = 0; x0 < nx; x0 += blockDim.x) {
      const int x = x0 + tid;
      const size_t idx = base + (size_t)y * nx + x;
      u64 v = 0;
      if (x < nx) {
        const int c = cast_shift(in[idx], shift);
        if (cast_out) cast_out[idx] = c;
        v = (u64)(long long)c;  // int -> unsigned long long as the reference's assignment does (integral.cxx:113)
      }
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const u64 t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
      }
      if (lane == 31) wtot[warp] = v;
      __syncthreads();
      u64 before = carry, total = 0;
      for (int w = 0; w < nwarps; w++) {
        const u64 t = wtot[w];
        if (w < warp) before += t;
        total += t;
      }
      if (x < nx) {
        const u64 s = col[x] + v + before;
        col[x] = s;
        out[idx] = s;
      }
      carry += total;
      __syncthreads();
    }
  }
}

__global__ void integral_z_kernel(u64* __restrict__ vol, size_t slice, int nz) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= slice) return;
  u64 acc = vol[i];
  int z = 1;
  for (; z + 4 <= nz; z += 4) {
    u64* p = vol + (size_t)z * slice + i;
    const u64 a = p[0], b = p[slice], c = p[2 * slice], d = p[3 * slice];
    acc += a; p[0] = acc;
    acc += b; p[slice] = acc;
    acc += c; p[2 * slice] = acc;
    acc += d; p[3 * slice] = acc;
  }
  for (; z < nz; z++) {
    u64* p = vol + (size_t)z * slice + i;
    acc += *p;
    *p = acc;
  }
}

// ---------------------------------------------------------------------------------------------
// Box sums over the integral volume (integral.h:66-115, BoxIntegralOptim: no clamping).
// ---------------------------------------------------------------------------------------------
struct Integral {
  const u64* p;
  long long sy, sz;  // increments along y and z (x increment 1)
  int nx, ny, nz;
};

__device__ __forceinline__ u64 box_sum(const Integral& I, int x0, int y0, int z0, int sx, int sy, int sz) {
  const long long x1 = x0 - 1, y1 = (long long)(y0 - 1) * I.sy, z1 = (long long)(z0 - 1) * I.sz;
  const long long x2 = x0 + sx - 1, y2 = (long long)(y0 + sy - 1) * I.sy, z2 = (long long)(z0 + sz - 1) * I.sz;
  const u64* p = I.p;
  return __ldg(p + x2 + y2 + z2) - __ldg(p + x2 + y2 + z1) - __ldg(p + x2 + y1 + z2) - __ldg(p + x1 + y2 + z2) +
         __ldg(p + x1 + y1 + z2) + __ldg(p + x1 + y2 + z1) + __ldg(p + x2 + y1 + z1) - __ldg(p + x1 + y1 + z1);
}

__device__ __forceinline__ float boxf(const Integral& I, int x0, int y0, int z0, int sx, int sy, int sz) {
  return __ull2float_rn(box_sum(I, x0, y0, z0, sx, sy, sz));
}

// ---------------------------------------------------------------------------------------------
// One response layer (FastHessian::buildResponseLayer, fasthessian.cxx:343-481): box-filter
// approximations of the six second derivatives, determinant response, laplacian sign, blob flag.
// 144 eight-byte gathers per voxel; neighbouring threads read neighbouring x (stride `step`), so a
// warp's gathers fall into a handful of 128-byte lines that the following corner loads reuse from L1.
// ---------------------------------------------------------------------------------------------
struct LayerDev {
  float* responses;
  uint8_t* laplacian;
  uint8_t* isblob;
  int width, height, depth, step, filter;
  int limit;           // fasthessian.cxx:366
  float inv_volume9;   // fasthessian.cxx:355
};

__global__ void __launch_bounds__(256) response_layer_kernel(Integral I, LayerDev L) {
  const int iw = L.width - 2 * L.limit, ih = L.height - 2 * L.limit, id = L.depth - 2 * L.limit;
  const long long n = (long long)iw * ih * id;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int ax = L.limit + (int)(t % iw), ay = L.limit + (int)((t / iw) % ih), az = L.limit + (int)(t / ((long long)iw * ih));
  const int x = ax * L.step, y = ay * L.step, z = az * L.step;
  const int b = (L.filter - 1) / 2, l = L.filter / 3, w = L.filter;
  const int m = 2 * l - 1;

  const float Dxx = __fsub_rn(boxf(I, x - b, y - l + 1, z - l + 1, w, m, m),
                              __fmul_rn(boxf(I, x - l / 2, y - l + 1, z - l + 1, l, m, m), 3.0f));
  const float Dyy = __fsub_rn(boxf(I, x - l + 1, y - b, z - l + 1, m, w, m),
                              __fmul_rn(boxf(I, x - l + 1, y - l / 2, z - l + 1, m, l, m), 3.0f));
  const float Dzz = __fsub_rn(boxf(I, x - l + 1, y - l + 1, z - b, m, m, w),
                              __fmul_rn(boxf(I, x - l + 1, y - l + 1, z - l / 2, m, m, l), 3.0f));
  const float Dxy = __fsub_rn(__fsub_rn(__fadd_rn(boxf(I, x - l, y - l, z - l + 1, l, l, m), boxf(I, x + 1, y + 1, z - l + 1, l, l, m)),
                                        boxf(I, x - l, y + 1, z - l + 1, l, l, m)),
                              boxf(I, x + 1, y - l, z - l + 1, l, l, m));
  const float Dyz = __fsub_rn(__fsub_rn(__fadd_rn(boxf(I, x - l + 1, y - l, z - l, m, l, l), boxf(I, x - l + 1, y + 1, z + 1, m, l, l)),
                                        boxf(I, x - l + 1, y - l, z + 1, m, l, l)),
                              boxf(I, x - l + 1, y + 1, z - l, m, l, l));
  const float Dxz = __fsub_rn(__fsub_rn(__fadd_rn(boxf(I, x - l, y - l + 1, z - l, l, m, l), boxf(I, x + 1, y - l + 1, z + 1, l, m, l)),
                                        boxf(I, x - l, y - l + 1, z + 1, l, m, l)),
                              boxf(I, x + 1, y - l + 1, z - l, l, m, l));

  // fasthessian.cxx:428-430, in the reference's operand order and types (the 2.0 literal makes the second term,
  // and from there the running sum, double)
  const float sq = __fadd_rn(__fadd_rn(__fmul_rn(Dxy, Dxy), __fmul_rn(Dxz, Dxz)), __fmul_rn(Dyz, Dyz));
  const float Sdet2p = __fsub_rn(__fadd_rn(__fadd_rn(__fmul_rn(Dyy, Dzz), __fmul_rn(Dxx, Dyy)), __fmul_rn(Dxx, Dzz)),
                                 __fmul_rn(0.8330f, sq));
  const float Trace = __fadd_rn(__fadd_rn(Dxx, Dyy), Dzz);
  double det = (double)__fmul_rn(__fmul_rn(Dxx, Dyy), Dzz);
  det = __dadd_rn(det, __dmul_rn(__dmul_rn(__dmul_rn(__dmul_rn(2.0, (double)Dxy), (double)Dyz), (double)Dxz), (double)0.7603f));
  det = __dsub_rn(det, (double)__fmul_rn(__fmul_rn(__fmul_rn(Dxx, Dyz), Dyz), 0.8330f));
  det = __dsub_rn(det, (double)__fmul_rn(__fmul_rn(__fmul_rn(Dyy, Dxz), Dxz), 0.8330f));
  det = __dsub_rn(det, (double)__fmul_rn(__fmul_rn(__fmul_rn(Dzz, Dxy), Dxy), 0.8330f));
  const float Det = __double2float_rn(det);

  const size_t index = (size_t)ax + (size_t)ay * L.width + (size_t)az * L.width * L.height;
  L.isblob[index] = (Sdet2p > 0.0f) && (__fmul_rn(Trace, Det) > 0.0f);
  L.responses[index] = fabsf(__fmul_rn(Det, L.inv_volume9));
  L.laplacian[index] = Trace >= 0.0f ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------
// Extremum search over one (bottom, middle, top) layer triple (fasthessian.cxx:161-218, isExtremum
// :521-548) and, for each extremum, the derivative vector and Hessian the interpolation step solves
// with (deriv4D :666-696, hessian4D :938-1014).  Extrema are rare, so the record is written by the
// thread that found it; the host orders records by `key` = the reference's loop position.
// ---------------------------------------------------------------------------------------------
struct LayerView {
  const float* responses;
  const uint8_t* laplacian;
  const uint8_t* isblob;
  int width, height, depth;
};

__device__ __forceinline__ size_t lv_index(const LayerView& v, int scale, int r, int c, int d) {
  // responselayer.h:73-80: getResponse(row, column, layer, src) with scale = this->width / src->width
  return (size_t)(scale * c) + (size_t)(scale * r) * v.width + (size_t)(scale * d) * v.width * v.height;
}
__device__ __forceinline__ float lv_resp(const LayerView& v, int scale, int r, int c, int d) {
  return __ldg(v.responses + lv_index(v, scale, r, c, d));
}
__device__ __forceinline__ bool lv_blob(const LayerView& v, int scale, int r, int c, int d) {
  return __ldg(v.isblob + lv_index(v, scale, r, c, d)) != 0;
}

struct ExtremaPass {
  LayerView b, m, t;
  int scale_m, scale_b;        // m->width / t->width, b->width / t->width
  int limit;                   // fasthessian.cxx:181-189
  long long first_sup, first_down;  // loop positions from which `param` has lost its bits (fasthessian.cxx:202-210)
  int param0;                  // FIRST_SCALE 1, LAST_SCALE 2, NONE_SCALE 0
  int pass;
  float thresh;
};

struct Candidate {
  u64 key;
  int r, c, d, laplacian;
  float response;
  float pad_;
  double dD[4];
  double H[10];  // dxx dyy dzz dss dxy dxz dxs dyz dys dzs
};

__global__ void __launch_bounds__(256) extrema_kernel(ExtremaPass P, Candidate* __restrict__ out, unsigned* __restrict__ count, unsigned cap) {
  const int nc = P.t.width - 2 * P.limit, nr = P.t.height - 2 * P.limit, nd = P.t.depth - 2 * P.limit;
  const long long n = (long long)nc * nr * nd;
  const long long tix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (tix >= n) return;
  // x (column) fastest across threads for coalescing; the reference's loops nest r, c, d
  const int c = P.limit + (int)(tix % nc), r = P.limit + (int)((tix / nc) % nr), d = P.limit + (int)(tix / ((long long)nc * nr));
  const long long pos = ((long long)(r - P.limit) * nc + (c - P.limit)) * nd + (d - P.limit);
  int param = P.param0;
  if (pos >= P.first_sup) param &= 2;
  if (pos >= P.first_down) param &= 1;

  const float candidate = lv_resp(P.m, P.scale_m, r, c, d);
  if (candidate < P.thresh || !lv_blob(P.m, P.scale_m, r, c, d)) return;
  for (int rr = -1; rr <= 1; ++rr)
    for (int cc = -1; cc <= 1; ++cc)
      for (int dd = -1; dd <= 1; ++dd) {
        if ((param != 2 && lv_resp(P.t, 1, r + rr, c + cc, d + dd) >= candidate && lv_blob(P.t, 1, r + rr, c + cc, d + dd)) ||
            ((rr != 0 || cc != 0) && lv_resp(P.m, P.scale_m, r + rr, c + cc, d + dd) >= candidate &&
             lv_blob(P.m, P.scale_m, r + rr, c + cc, d + dd)) ||
            (param != 1 && lv_resp(P.b, P.scale_b, r + rr, c + cc, d + dd) >= candidate &&
             lv_blob(P.b, P.scale_b, r + rr, c + cc, d + dd)))
          return;
      }

  const unsigned slot = atomicAdd(count, 1u);
  if (slot >= cap) return;
  Candidate k;
  k.key = ((u64)P.pass << 48) | (u64)pos;
  k.r = r; k.c = c; k.d = d;
  k.laplacian = __ldg(P.m.laplacian + lv_index(P.m, P.scale_m, r, c, d));
  k.response = candidate;
  k.pad_ = 0.0f;
#define M_(rr, cc, dd) lv_resp(P.m, P.scale_m, r + (rr), c + (cc), d + (dd))
#define T_(rr, cc, dd) lv_resp(P.t, 1, r + (rr), c + (cc), d + (dd))
#define B_(rr, cc, dd) lv_resp(P.b, P.scale_b, r + (rr), c + (cc), d + (dd))
  // float differences, then a double division (the reference divides by the double literals 2.0 / 4.0)
  k.dD[0] = __ddiv_rn((double)__fsub_rn(M_(0, 1, 0), M_(0, -1, 0)), 2.0);
  k.dD[1] = __ddiv_rn((double)__fsub_rn(M_(1, 0, 0), M_(-1, 0, 0)), 2.0);
  k.dD[2] = __ddiv_rn((double)__fsub_rn(M_(0, 0, 1), M_(0, 0, -1)), 2.0);
  k.dD[3] = __ddiv_rn((double)__fsub_rn(T_(0, 0, 0), B_(0, 0, 0)), 2.0);
  const double v = (double)M_(0, 0, 0), v2 = __dmul_rn(2.0, v);
  k.H[0] = __dsub_rn((double)__fadd_rn(M_(0, 1, 0), M_(0, -1, 0)), v2);
  k.H[1] = __dsub_rn((double)__fadd_rn(M_(1, 0, 0), M_(-1, 0, 0)), v2);
  k.H[2] = __dsub_rn((double)__fadd_rn(M_(0, 0, 1), M_(0, 0, -1)), v2);
  k.H[3] = __dsub_rn((double)__fadd_rn(T_(0, 0, 0), B_(0, 0, 0)), v2);
#define CROSS_(a, b2, c2, d2) __ddiv_rn((double)__fadd_rn(__fsub_rn(__fsub_rn((a), (b2)), (c2)), (d2)), 4.0)
  k.H[4] = CROSS_(M_(1, 1, 0), M_(1, -1, 0), M_(-1, 1, 0), M_(-1, -1, 0));   // dxy
  k.H[5] = CROSS_(M_(1, 0, 1), M_(1, 0, -1), M_(-1, 0, 1), M_(-1, 0, -1));   // "dxz" (the reference steps r and d here)
  k.H[6] = CROSS_(T_(0, 1, 0), T_(0, -1, 0), B_(0, 1, 0), B_(0, -1, 0));     // dxs
  k.H[7] = CROSS_(M_(0, 1, 1), M_(0, 1, -1), M_(0, -1, 1), M_(0, -1, -1));   // "dyz" (steps c and d)
  k.H[8] = CROSS_(T_(1, 0, 0), T_(-1, 0, 0), B_(1, 0, 0), B_(-1, 0, 0));     // dys
  k.H[9] = CROSS_(T_(0, 0, 1), T_(0, 0, -1), B_(0, 0, 1), B_(0, 0, -1));     // dzs
#undef CROSS_
#undef M_
#undef T_
#undef B_
  out[slot] = k;
}

// ---------------------------------------------------------------------------------------------
// glibc 2.39's expf (sysdeps/ieee754/flt-32/e_expf.c -- Szabolcs Nagy's table-driven algorithm:
// x * 32/ln2 = k + r, 2^(k/32) from a 32-entry table, cubic in r, all in double, one final rounding
// to float), restated from its published description.  Surf::gaussian calls expf (surf.cxx:227),
// so bit-identical descriptors need bit-identical expf: this restatement equals this image's libm
// on every float in [-104, 88.7] (exhaustive host test, tests/test_surf_host.py) but two, which are
// patched below; the table is 2^(i/32) rounded to double with i << 47 subtracted.
// ---------------------------------------------------------------------------------------------
#define FS_EXP2_TABLE                                                                                   \
  {0x3ff0000000000000ULL, 0x3fefd9b0d3158574ULL, 0x3fefb5586cf9890fULL, 0x3fef9301d0125b51ULL,          \
   0x3fef72b83c7d517bULL, 0x3fef54873168b9aaULL, 0x3fef387a6e756238ULL, 0x3fef1e9df51fdee1ULL,          \
   0x3fef06fe0a31b715ULL, 0x3feef1a7373aa9cbULL, 0x3feedea64c123422ULL, 0x3feece086061892dULL,          \
   0x3feebfdad5362a27ULL, 0x3feeb42b569d4f82ULL, 0x3feeab07dd485429ULL, 0x3feea47eb03a5585ULL,          \
   0x3feea09e667f3bcdULL, 0x3fee9f75e8ec5f74ULL, 0x3feea11473eb0187ULL, 0x3feea589994cce13ULL,          \
   0x3feeace5422aa0dbULL, 0x3feeb737b0cdc5e5ULL, 0x3feec49182a3f090ULL, 0x3feed503b23e255dULL,          \
   0x3feee89f995ad3adULL, 0x3feeff76f2fb5e47ULL, 0x3fef199bdd85529cULL, 0x3fef3720dcef9069ULL,          \
   0x3fef5818dcfba487ULL, 0x3fef7c97337b9b5fULL, 0x3fefa4afa2a490daULL, 0x3fefd0765b6e4540ULL}
__constant__ u64 kExp2TabDev[32] = FS_EXP2_TABLE;
static const u64 kExp2TabHost[32] = FS_EXP2_TABLE;

// the same statement on both sides of the compiler so that the host tests can run it against libm without a GPU
__host__ __device__ __forceinline__ float glibc_expf(float x) {
#ifdef __CUDA_ARCH__
#define FS_DMUL(a, b) __dmul_rn((a), (b))
#define FS_DADD(a, b) __dadd_rn((a), (b))
#define FS_DSUB(a, b) __dsub_rn((a), (b))
  const u64* tab = kExp2TabDev;
#else
#define FS_DMUL(a, b) ((a) * (b))  // host build: x86-64 baseline, no FMA contraction
#define FS_DADD(a, b) ((a) + (b))
#define FS_DSUB(a, b) ((a) - (b))
  const u64* tab = kExp2TabHost;
#endif
  if (!(x >= -0x1.9fe368p6f)) return x != x ? x : 0.0f;  // underflow to +0 (and NaN in, NaN out)
  if (x > 0x1.62e42ep6f) return x + x > x ? x * 3.0e38f : x;  // overflow to +inf
  if (x == -0x1.f8cbb2p+5f) return 0x1.f45326p-92f;      // the two inputs where libm's last bit differs
  if (x == 0x1.04845ep+5f) return 0x1.f93e38p+46f;
  const double z = FS_DMUL(0x1.71547652b82fep+0 * 32.0, (double)x);
  double kd = FS_DADD(z, 0x1.8p52);
  u64 ki;
  memcpy(&ki, &kd, 8);
  kd = FS_DSUB(kd, 0x1.8p52);
  const double r = FS_DSUB(z, kd);
  const u64 sbits = tab[ki & 31] + (ki << 47);
  double s;
  memcpy(&s, &sbits, 8);
  const double p = FS_DADD(FS_DMUL(0x1.c6af84b912394p-5 / 32 / 32 / 32, r), 0x1.ebfce50fac4f3p-3 / 32 / 32);
  const double r2 = FS_DMUL(r, r);
  double y = FS_DADD(FS_DMUL(0x1.62e42ff0c52d6p-1 / 32, r), 1.0);
  y = FS_DADD(FS_DMUL(p, r2), y);
  return (float)FS_DMUL(y, s);
#undef FS_DMUL
#undef FS_DADD
#undef FS_DSUB
}

// ---------------------------------------------------------------------------------------------
// Descriptors (Surf::getDescriptor surf.cxx:63-156, getRawDescriptor :161-217).  One CTA per
// keypoint: the (2 radius)^3 Haar samples are computed in parallel (36 gathers each: the two boxes of
// a Haar wavelet share a face) and parked in shared memory as doubles; 48 threads then add up their
// sub-block's samples IN THE REFERENCE'S ORDER (u, v, w nested), because the sums are double
// accumulations whose rounding depends on it.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int f_round(float f) { return (int)floorf(__fadd_rn(f, 0.5f)); }  // surf.h:86-89

struct Clamp {
  bool hit;
};

// Integral value at (x, y, z); where the reference would read outside the volume (undefined behaviour in
// BoxIntegralOptim) this follows BoxIntegral (integral.h:30-62): negative index -> 0, index past the end -> last.
__device__ __forceinline__ u64 integ_at(const Integral& I, int x, int y, int z, Clamp& cl) {
  if (x < 0 || y < 0 || z < 0) { cl.hit = true; return 0; }
  if (x >= I.nx) { x = I.nx - 1; cl.hit = true; }
  if (y >= I.ny) { y = I.ny - 1; cl.hit = true; }
  if (z >= I.nz) { z = I.nz - 1; cl.hit = true; }
  return __ldg(I.p + x + (long long)y * I.sy + (long long)z * I.sz);
}

// sum over [xa+1..xb] x [ya+1..yb] x [za+1..zb] (inclusive corner indices as BoxIntegralOptim forms them)
__device__ __forceinline__ long long box_corners(const Integral& I, int xa, int xb, int ya, int yb, int za, int zb, Clamp& cl) {
  return (long long)(integ_at(I, xb, yb, zb, cl) - integ_at(I, xb, yb, za, cl) - integ_at(I, xb, ya, zb, cl) -
                     integ_at(I, xa, yb, zb, cl) + integ_at(I, xa, ya, zb, cl) + integ_at(I, xa, yb, za, cl) +
                     integ_at(I, xb, ya, za, cl) - integ_at(I, xa, ya, za, cl));
}

// haarXOptim / haarYOptim / haarZOptim (surf.cxx:256-281) at (x, y, z) with s = 2 * h
__device__ __forceinline__ void haar3(const Integral& I, int x, int y, int z, int h, float& hx, float& hy, float& hz, Clamp& cl) {
  const int s = 2 * h, hh = s / 2;
  const int xl = x - hh - 1, xm = x - 1, xr = x + hh - 1, xR = x - hh + s - 1;
  const int yl = y - hh - 1, ym = y - 1, yr = y + hh - 1, yR = y - hh + s - 1;
  const int zl = z - hh - 1, zm = z - 1, zr = z + hh - 1, zR = z - hh + s - 1;
  hx = __ll2float_rn(box_corners(I, xm, xr, yl, yR, zl, zR, cl) - box_corners(I, xl, xm, yl, yR, zl, zR, cl));
  hy = __ll2float_rn(box_corners(I, xl, xR, ym, yr, zl, zR, cl) - box_corners(I, xl, xR, yl, ym, zl, zR, cl));
  hz = __ll2float_rn(box_corners(I, xl, xR, yl, yR, zm, zr, cl) - box_corners(I, xl, xR, yl, yR, zl, zm, cl));
}

__global__ void __launch_bounds__(256) describe_kernel(Integral I, const fs_point* __restrict__ pts, unsigned n, int radius,
                                                        int type, int normalize, float* __restrict__ desc,
                                                        unsigned* __restrict__ n_clamped) {
  extern __shared__ double fs_smd[];
  const unsigned id = blockIdx.x;
  if (id >= n) return;
  const fs_point pt = pts[id];
  const int r3 = radius * radius * radius, S = 8 * r3;
  double* sx = fs_smd;  // [S] per component, sub-block major, (u, v, w) inside
  double* sy = fs_smd + S;
  double* sz = fs_smd + 2 * S;
  __shared__ double acc[48];
  __shared__ int any_clamped;
  if (threadIdx.x == 0) any_clamped = 0;
  __syncthreads();

  const double scale = (double)pt.scale;
  const int x = f_round(pt.x), y = f_round(pt.y), z = f_round(pt.z);
  const float halfRadius = __double2float_rn(__ddiv_rn((double)__double2float_rn((double)radius - 1.0), 2.0));
  const int h = f_round(pt.scale);                                           // s = 2 * fRound(scale)
  const float sig = __double2float_rn(__dmul_rn((double)2.5f, scale));       // gaussian(..., 2.5f * scale)
  const float sig2 = __fmul_rn(sig, sig);
  const float norm = __fdiv_rn(1.0f, __fmul_rn(sig2, sig));                  // 1.0f / (sig*sig*sig)
  const float den = __fmul_rn(__fmul_rn(2.0f, sig), sig);                    // 2.0f*sig*sig
  Clamp cl{false};
  const size_t dsize = type == 0 ? 48 : (size_t)3 * S;

  for (int sidx = threadIdx.x; sidx < S; sidx += blockDim.x) {
    const int blk = sidx / r3, within = sidx - blk * r3;
    const int i = -radius + (blk >> 2) * radius, j = -radius + ((blk >> 1) & 1) * radius, k = -radius + (blk & 1) * radius;
    const int u = i + within / (radius * radius), v = j + (within / radius) % radius, w = k + within % radius;
    const int sample_x = f_round(__double2float_rn(__dadd_rn((double)x, __dmul_rn((double)u, scale))));
    const int sample_y = f_round(__double2float_rn(__dadd_rn((double)y, __dmul_rn((double)v, scale))));
    const int sample_z = f_round(__double2float_rn(__dadd_rn((double)z, __dmul_rn((double)w, scale))));
    float hx, hy, hz;
    haar3(I, sample_x, sample_y, sample_z, h, hx, hy, hz, cl);
    if (type == 0) {
      const float ix = __fadd_rn((float)i, halfRadius), jx = __fadd_rn((float)j, halfRadius), kx = __fadd_rn((float)k, halfRadius);
      const int xs = f_round(__double2float_rn(__dadd_rn((double)pt.x, __dmul_rn((double)ix, scale))));
      const int ys = f_round(__double2float_rn(__dadd_rn((double)pt.y, __dmul_rn((double)jx, scale))));
      const int zs = f_round(__double2float_rn(__dadd_rn((double)pt.z, __dmul_rn((double)kx, scale))));
      const float gx = __fsub_rn((float)xs, (float)sample_x), gy = __fsub_rn((float)ys, (float)sample_y),
                  gz = __fsub_rn((float)zs, (float)sample_z);
      const float num = __fadd_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)), __fmul_rn(gz, gz));
      const double gauss = (double)__fmul_rn(norm, glibc_expf(__fdiv_rn(-num, den)));
      sx[sidx] = __dmul_rn(gauss, (double)hx);
      sy[sidx] = __dmul_rn(gauss, (double)hy);
      sz[sidx] = __dmul_rn(gauss, (double)hz);
    } else {
      float* o = desc + (size_t)id * dsize + (size_t)3 * sidx;
      o[0] = hx; o[1] = hy; o[2] = hz;
    }
  }
  if (cl.hit) any_clamped = 1;
  __syncthreads();
  if (threadIdx.x == 0 && any_clamped) atomicAdd(n_clamped, 1u);
  if (type != 0) return;

  if (threadIdx.x < 48) {
    const int blk = threadIdx.x / 6, comp = threadIdx.x % 6;
    const double* src = (comp % 3 == 0 ? sx : comp % 3 == 1 ? sy : sz) + blk * r3;
    double a = 0.0;  // dx = dy = ... = 0.f
    if (comp < 3) {
      for (int q = 0; q < r3; q++) a = __dadd_rn(a, src[q]);
    } else {
      for (int q = 0; q < r3; q++) a = __dadd_rn(a, fabs(src[q]));
    }
    acc[threadIdx.x] = a;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float* o = desc + (size_t)id * 48;
    double len = 0.0;
    for (int blk = 0; blk < 8; blk++) {
      const double* a = acc + blk * 6;
      double q = __dmul_rn(a[0], a[0]);
      for (int c = 1; c < 6; c++) q = __dadd_rn(q, __dmul_rn(a[c], a[c]));
      len = __dadd_rn(len, q);
    }
    len = __dsqrt_rn(len);
    for (int c = 0; c < 48; c++) {
      float f = __double2float_rn(acc[c]);  // desc[count++] = dx : double -> float
      if (normalize && len != 0.0) f = __double2float_rn(__ddiv_rn((double)f, len));
      o[c] = f;
    }
  }
}

}  // namespace fs
