// TEST INFRASTRUCTURE ONLY -- see vtkShimCore.h
#include "vtkShimCore.h"
#ifndef ORACLE_VTK_SHIM_SHIFTSCALE_H
#define ORACLE_VTK_SHIM_SHIFTSCALE_H
class vtkImageShiftScale : public vtkShimImageFilter {
 public:
  static vtkImageShiftScale* New() { return new vtkImageShiftScale; }
  void SetClampOverflow(int) {}
  void SetOutputScalarTypeToInt() {}
  void SetShift(double s) { shift_ = s; }
  void SetScale(double s) { scale_ = s; }
  void Update() { vtkShimToInt(in_, out_, shift_, scale_); }
 private:
  double shift_ = 0.0, scale_ = 1.0;
};
#endif
