// TEST INFRASTRUCTURE ONLY -- see vtkShimCore.h
#include "vtkShimCore.h"
#ifndef ORACLE_VTK_SHIM_LUMINANCE_H
#define ORACLE_VTK_SHIM_LUMINANCE_H
class vtkImageLuminance : public vtkShimImageFilter {
 public:
  static vtkImageLuminance* New() { return new vtkImageLuminance; }
  void Update() { Unavailable("vtkImageLuminance"); }
};
#endif
