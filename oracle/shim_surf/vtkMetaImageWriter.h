// TEST INFRASTRUCTURE ONLY -- see vtkShimCore.h.  Only FastHessian::WriteResponseMap (debug output, never
// called by the producer) uses a writer; it writes nothing here.
#include "vtkShimCore.h"
#ifndef ORACLE_VTK_SHIM_MIW_H
#define ORACLE_VTK_SHIM_MIW_H
class vtkMetaImageWriter : public vtkObject {
 public:
  static vtkMetaImageWriter* New() { return new vtkMetaImageWriter; }
  void SetInputData(vtkImageData*) {}
  void SetFileName(const char*) {}
  void Write() {}
};
#endif
