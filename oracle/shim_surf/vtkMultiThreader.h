// TEST INFRASTRUCTURE ONLY -- see vtkShimCore.h.  Runs the single method serially as thread 0 of 1.
#include "vtkShimCore.h"
#ifndef ORACLE_VTK_SHIM_MT_H
#define ORACLE_VTK_SHIM_MT_H
class vtkMultiThreader : public vtkObject {
 public:
  struct ThreadInfo { int ThreadID; int NumberOfThreads; void* UserData; };
  static vtkMultiThreader* New() { return new vtkMultiThreader; }
  void SetNumberOfThreads(int) {}
  void SetSingleMethod(void* (*f)(void*), void* data) { f_ = f; d_ = data; }
  void SingleMethodExecute() { ThreadInfo ti{0, 1, d_}; f_(&ti); }
 private:
  void* (*f_)(void*) = nullptr;
  void* d_ = nullptr;
};
#endif
