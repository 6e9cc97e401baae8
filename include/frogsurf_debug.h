/*
 * frogsurf_debug.h -- host-side pieces of libfrogsurf.so exposed so that the CPU test suite can
 * check them without a GPU.  Not part of the drop-in surface.
 */
#ifndef FROGSURF_DEBUG_H_
#define FROGSURF_DEBUG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* the restatement of glibc's expf the descriptor kernel uses (surf.cxx:227 calls expf), host build */
float fs_debug_expf(float x);
void fs_debug_expf_many(const float* x, float* y, size_t n);
/* interpolation step fasthessian.cxx:614-661: X = -pinv(H) dD; H10 = dxx dyy dzz dss dxy dxz dxs dyz dys dzs */
void fs_debug_solve_offsets(const double* dD, const double* H10, double* X);
/* response-layer geometry (fasthessian.cxx:53-81, 287-341, 366): per layer width, height, depth, step, filter, limit */
int fs_debug_layers(int nx, int ny, int nz, int32_t* out6, int cap);
/* vtk3DSURF.cxx:209-226: order[] receives the original indices of the points kept, strongest first */
uint32_t fs_debug_select(const float* response, uint32_t n, int number_of_points, uint32_t* order);

#ifdef __cplusplus
}
#endif

#endif
