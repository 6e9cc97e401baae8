// TEST INFRASTRUCTURE ONLY -- see vtkShimCore.h
#include "vtkShimCore.h"
#ifndef ORACLE_VTK_SHIM_CAST_H
#define ORACLE_VTK_SHIM_CAST_H
class vtkImageCast : public vtkShimImageFilter {
 public:
  static vtkImageCast* New() { return new vtkImageCast; }
  void SetClampOverflow(int) {}
  void SetOutputScalarTypeToInt() {}
  void Update() { vtkShimToInt(in_, out_, 0.0, 1.0); }
};
#endif
