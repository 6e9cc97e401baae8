// TEST INFRASTRUCTURE ONLY -- stand-in for the parts of VTK that the reference's SURF3D producer
// (vtkOpenSURF3D/{integral,fasthessian,surf,vtk3DSURF}.cxx) touches, so that those files compile
// UNMODIFIED here, where VTK is absent.  Nothing in this header computes anything the producer's
// arithmetic depends on, with two stated exceptions that restate VTK's documented behaviour:
// vtkImageCast (static_cast with clamping) and vtkImageShiftScale ((v + shift) * scale, clamped,
// static_cast).  Filters that would need VTK's own algorithms (vtkImageResample, vtkImageResize,
// vtkImageLuminance) abort when run: the oracle is only valid from an already isotropic
// single-component volume onward (surf3d without -s / -d, descriptor types 0 and 1).
#ifndef ORACLE_VTK_SHIM_CORE_H
#define ORACLE_VTK_SHIM_CORE_H

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <string>
#include <vector>

typedef long long vtkIdType;

#define VTK_VOID 0
#define VTK_CHAR 2
#define VTK_UNSIGNED_CHAR 3
#define VTK_SHORT 4
#define VTK_UNSIGNED_SHORT 5
#define VTK_INT 6
#define VTK_UNSIGNED_INT 7
#define VTK_FLOAT 10
#define VTK_DOUBLE 11
#define VTK_UNSIGNED_LONG_LONG 17

#define VTK_FLOAT_MIN (-1.0e+38f)
#define VTK_FLOAT_MAX 1.0e+38f
#define VTK_UNSIGNED_SHORT_MIN 0
#define VTK_UNSIGNED_SHORT_MAX 65535

#define VTK_THREAD_RETURN_TYPE void*
#define VTK_THREAD_RETURN_VALUE nullptr

class vtkObject {
 public:
  static vtkObject* New() { return new vtkObject; }
  virtual const char* GetClassName() { return "vtkObject"; }
  void Register(vtkObject*) { ++refs_; }
  void UnRegister(vtkObject*) { if (--refs_ <= 0) delete this; }
  void Delete() { UnRegister(nullptr); }
  virtual void Modified() {}
 protected:
  vtkObject() : refs_(1) {}
  virtual ~vtkObject() {}
 private:
  int refs_;
};

#define vtkTypeMacro(cls, super) \
  typedef super Superclass;      \
  const char* GetClassName() override { return #cls; }
#define vtkStandardNewMacro(cls) \
  cls* cls::New() { return new cls; }
#define vtkSetMacro(name, type) \
  virtual void Set##name(type _arg) { this->name = _arg; this->Modified(); }
#define vtkGetMacro(name, type) \
  virtual type Get##name() { return this->name; }
#define vtkGetObjectMacro(name, type) \
  virtual type* Get##name() { return this->name; }

template <class T>
class vtkSmartPointer {
 public:
  vtkSmartPointer() : p_(nullptr) {}
  vtkSmartPointer(T* p) : p_(p) { if (p_) p_->Register(nullptr); }
  vtkSmartPointer(const vtkSmartPointer& o) : p_(o.p_) { if (p_) p_->Register(nullptr); }
  ~vtkSmartPointer() { if (p_) p_->UnRegister(nullptr); }
  vtkSmartPointer& operator=(T* p) {
    if (p) p->Register(nullptr);
    if (p_) p_->UnRegister(nullptr);
    p_ = p;
    return *this;
  }
  vtkSmartPointer& operator=(const vtkSmartPointer& o) { return *this = o.p_; }
  static vtkSmartPointer Take(T* p) { vtkSmartPointer s; s.p_ = p; return s; }
  static vtkSmartPointer New() { return Take(T::New()); }
  T* operator->() const { return p_; }
  T& operator*() const { return *p_; }
  operator T*() const { return p_; }
  T* Get() const { return p_; }
  T* GetPointer() const { return p_; }
 private:
  T* p_;
};

template <class T>
class vtkNew {
 public:
  vtkNew() : p_(T::New()) {}
  ~vtkNew() { p_->Delete(); }
  vtkNew(const vtkNew&) = delete;
  vtkNew& operator=(const vtkNew&) = delete;
  T* operator->() const { return p_; }
  operator T*() const { return p_; }
  T* Get() const { return p_; }
  T* GetPointer() const { return p_; }
 private:
  T* p_;
};

class vtkTimerLog : public vtkObject {
 public:
  static vtkTimerLog* New() { return new vtkTimerLog; }
  void StartTimer() { t0_ = std::chrono::steady_clock::now(); }
  void StopTimer() { t1_ = std::chrono::steady_clock::now(); }
  double GetElapsedTime() { return std::chrono::duration<double>(t1_ - t0_).count(); }
 private:
  std::chrono::steady_clock::time_point t0_, t1_;
};

class vtkBoundingBox {
 public:
  void SetBounds(const double* b) { for (int i = 0; i < 6; i++) b_[i] = b[i]; }
  bool ContainsPoint(double x, double y, double z) const {
    return x >= b_[0] && x <= b_[1] && y >= b_[2] && y <= b_[3] && z >= b_[4] && z <= b_[5];
  }
 private:
  double b_[6];
};

class vtkImageData : public vtkObject {
 public:
  static vtkImageData* New() { return new vtkImageData; }
  const char* GetClassName() override { return "vtkImageData"; }

  void SetDimensions(int x, int y, int z) { dims_[0] = x; dims_[1] = y; dims_[2] = z; }
  void SetDimensions(const int* d) { SetDimensions(d[0], d[1], d[2]); }
  int* GetDimensions() { return dims_; }
  void GetDimensions(int* d) { for (int i = 0; i < 3; i++) d[i] = dims_[i]; }
  void SetSpacing(double x, double y, double z) { sp_[0] = x; sp_[1] = y; sp_[2] = z; }
  void SetSpacing(const double* s) { SetSpacing(s[0], s[1], s[2]); }
  void GetSpacing(double* s) { for (int i = 0; i < 3; i++) s[i] = sp_[i]; }
  double* GetSpacing() { return sp_; }
  void SetOrigin(double x, double y, double z) { org_[0] = x; org_[1] = y; org_[2] = z; }
  void SetOrigin(const double* o) { SetOrigin(o[0], o[1], o[2]); }
  void GetOrigin(double* o) { for (int i = 0; i < 3; i++) o[i] = org_[i]; }
  double* GetOrigin() { return org_; }
  double* GetBounds() {
    for (int i = 0; i < 3; i++) {
      bounds_[2 * i] = org_[i];
      bounds_[2 * i + 1] = org_[i] + (dims_[i] - 1) * sp_[i];
    }
    return bounds_;
  }
  void GetBounds(double* b) { double* s = GetBounds(); for (int i = 0; i < 6; i++) b[i] = s[i]; }
  void CopyStructure(vtkImageData* o) {
    SetDimensions(o->dims_); SetSpacing(o->sp_); SetOrigin(o->org_);
  }
  int GetNumberOfScalarComponents() { return comps_; }
  int GetScalarType() { return type_; }
  static size_t TypeSize(int t) {
    switch (t) {
      case VTK_CHAR: case VTK_UNSIGNED_CHAR: return 1;
      case VTK_SHORT: case VTK_UNSIGNED_SHORT: return 2;
      case VTK_INT: case VTK_UNSIGNED_INT: case VTK_FLOAT: return 4;
      case VTK_DOUBLE: case VTK_UNSIGNED_LONG_LONG: return 8;
    }
    std::abort();
  }
  void AllocateScalars(int type, int comps) {
    type_ = type; comps_ = comps;
    incs_[0] = comps; incs_[1] = (vtkIdType)comps * dims_[0]; incs_[2] = incs_[1] * dims_[1];
    data_.assign((size_t)incs_[2] * dims_[2] * TypeSize(type), 0);
  }
  vtkIdType* GetIncrements() { return incs_; }
  void* GetScalarPointer() { return data_.data(); }
  void* GetScalarPointer(int x, int y, int z) {
    if (x < 0 || y < 0 || z < 0 || x >= dims_[0] || y >= dims_[1] || z >= dims_[2]) return nullptr;
    return data_.data() + ((size_t)x * incs_[0] + (size_t)y * incs_[1] + (size_t)z * incs_[2]) * TypeSize(type_);
  }
  size_t NumberOfValues() const { return (size_t)dims_[0] * dims_[1] * dims_[2] * comps_; }
  double ValueAsDouble(size_t i) const {
    const unsigned char* p = data_.data();
    switch (type_) {
      case VTK_CHAR: return ((const signed char*)p)[i];
      case VTK_UNSIGNED_CHAR: return ((const unsigned char*)p)[i];
      case VTK_SHORT: return ((const short*)p)[i];
      case VTK_UNSIGNED_SHORT: return ((const unsigned short*)p)[i];
      case VTK_INT: return ((const int*)p)[i];
      case VTK_UNSIGNED_INT: return ((const unsigned int*)p)[i];
      case VTK_FLOAT: return ((const float*)p)[i];
      case VTK_DOUBLE: return ((const double*)p)[i];
      case VTK_UNSIGNED_LONG_LONG: return (double)((const unsigned long long*)p)[i];
    }
    std::abort();
  }
  void GetScalarRange(double* r) {
    size_t n = NumberOfValues();
    r[0] = std::numeric_limits<double>::max(); r[1] = -std::numeric_limits<double>::max();
    for (size_t i = 0; i < n; i++) { double v = ValueAsDouble(i); if (v < r[0]) r[0] = v; if (v > r[1]) r[1] = v; }
  }
 protected:
  vtkImageData() : type_(VTK_VOID), comps_(1) {
    for (int i = 0; i < 3; i++) { dims_[i] = 0; sp_[i] = 1.0; org_[i] = 0.0; incs_[i] = 0; }
  }
 private:
  int dims_[3], type_, comps_;
  double sp_[3], org_[3], bounds_[6];
  vtkIdType incs_[3];
  std::vector<unsigned char> data_;
};

// base for the image filters the producer chains together
class vtkShimImageFilter : public vtkObject {
 public:
  void SetInputData(vtkImageData* in) { in_ = in; }
  void SetNumberOfThreads(int) {}
  vtkImageData* GetOutput() { return out_; }
 protected:
  vtkShimImageFilter() : out_(vtkSmartPointer<vtkImageData>::New()) {}
  [[noreturn]] static void Unavailable(const char* what) {
    std::fprintf(stderr, "oracle VTK shim: %s needs VTK's own algorithm, which is absent here\n", what);
    std::abort();
  }
  vtkSmartPointer<vtkImageData> in_, out_;
};

// (v + shift) * scale in double, clamped to the int range, static_cast<int> -- with shift 0 and
// scale 1 this is vtkImageCast with ClampOverflow on.
inline void vtkShimToInt(vtkImageData* in, vtkImageData* out, double shift, double scale) {
  out->CopyStructure(in);
  out->AllocateScalars(VTK_INT, 1);
  int* o = static_cast<int*>(out->GetScalarPointer());
  size_t n = in->NumberOfValues();
  const double lo = std::numeric_limits<int>::min(), hi = std::numeric_limits<int>::max();
  for (size_t i = 0; i < n; i++) {
    double v = (in->ValueAsDouble(i) + shift) * scale;
    if (v > hi) v = hi;
    if (v < lo) v = lo;
    o[i] = static_cast<int>(v);
  }
}

#endif
