// TEST INFRASTRUCTURE ONLY -- see vtkShimCore.h
#include "vtkShimCore.h"
#ifndef ORACLE_VTK_SHIM_RESIZE_H
#define ORACLE_VTK_SHIM_RESIZE_H
class vtkImageResize : public vtkShimImageFilter {
 public:
  static vtkImageResize* New() { return new vtkImageResize; }
  void SetResizeMethodToOutputDimensions() {}
  void SetOutputDimensions(int, int, int) {}
  void SetCropping(int) {}
  void SetCroppingRegion(double, double, double, double, double, double) {}
  void Update() { Unavailable("vtkImageResize"); }
};
#endif
