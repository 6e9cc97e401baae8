// TEST INFRASTRUCTURE ONLY -- see vtkShimCore.h
#include "vtkShimCore.h"
#ifndef ORACLE_VTK_SHIM_RESAMPLE_H
#define ORACLE_VTK_SHIM_RESAMPLE_H
class vtkImageResample : public vtkShimImageFilter {
 public:
  static vtkImageResample* New() { return new vtkImageResample; }
  void SetAxisOutputSpacing(int, double) {}
  void SetInterpolationModeToNearestNeighbor() {}
  void Update() { Unavailable("vtkImageResample"); }
};
#endif
