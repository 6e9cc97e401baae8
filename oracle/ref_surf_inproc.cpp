// TEST INFRASTRUCTURE ONLY.  In-process face of the UNMODIFIED reference SURF3D producer
// (/root/reference/vtkOpenSURF3D/{integral,fasthessian,surf,vtk3DSURF}.cxx, compiled where they lie
// against oracle/shim_surf -- see oracle/Makefile, target `surfref`).  The driver below replays what
// surf3d.cxx:258-328 does after the image is loaded (set the options, Update(), write the points) on
// a volume handed over as a plain array, and exposes the intermediate products (shifted volume,
// integral volume, response layers) so that each device stage can be compared on its own.
// Never linked, loaded or executed by the product path.
#include <cstdint>
#include <cstring>
#include <vector>

#include "vtkShimCore.h"
#include "ipoint.h"
#include "integral.h"
#include "vtk3DSURF.h"
#define private public  // test access to FastHessian::buildResponseMap / responseMap; the layout is unchanged
#include "fasthessian.h"
#undef private
#include "responselayer.h"

namespace {

struct Probe : public vtk3DSURF {
  static Probe* Make() { return new Probe; }
  std::vector<Ipoint>& Pts() { return this->points; }
  vtkImageData* CastImage() { return this->Cast; }
};

struct Run {
  Probe* surf = nullptr;
  vtkSmartPointer<vtkImageData> input;
  std::vector<Ipoint> fh_pts;
  FastHessian* fh = nullptr;
  ~Run() {
    delete fh;
    if (surf) surf->Delete();
  }
};

}  // namespace

extern "C" {

// vtk_type: VTK_SHORT 4, VTK_INT 6, VTK_FLOAT 10 ...  data is copied.
void* sr_create(const void* data, int vtk_type, const int* dims, const double* spacing, const double* origin) {
  Run* r = new Run;
  r->input = vtkSmartPointer<vtkImageData>::New();
  r->input->SetDimensions(dims);
  r->input->SetSpacing(spacing);
  r->input->SetOrigin(origin);
  r->input->AllocateScalars(vtk_type, 1);
  std::memcpy(r->input->GetScalarPointer(), data, r->input->NumberOfValues() * vtkImageData::TypeSize(vtk_type));
  r->surf = Probe::Make();
  r->surf->SetInput(r->input);
  // surf3d.cxx's own defaults (surf3d.cxx:46-66): no resampling, threshold 0, type 0, radius 5
  r->surf->SetMaxSize(0);
  r->surf->SetSpacing(0);
  r->surf->SetThreshold(0);
  return r;
}

void sr_destroy(void* h) { delete static_cast<Run*>(h); }

// the option setters surf3d.cxx:259-267 calls; point_file may be null
void sr_update(void* h, double threshold, int number_of_points, int descriptor_type, int radius, int normalize,
               const char* point_file) {
  Run* r = static_cast<Run*>(h);
  r->surf->SetNormalize(normalize != 0);
  r->surf->SetThreshold(threshold);
  r->surf->SetDescriptorType(descriptor_type);
  r->surf->SetNumberOfPoints(number_of_points);
  r->surf->SetSubVolumeRadius(radius);
  r->surf->SetNbThread(-1);
  if (point_file) r->surf->SetPointFile(const_cast<char*>(point_file));
  r->surf->Update();
}

int64_t sr_num_points(void* h) { return (int64_t) static_cast<Run*>(h)->surf->Pts().size(); }
int64_t sr_descriptor_size(void* h) {
  auto& p = static_cast<Run*>(h)->surf->Pts();
  return p.empty() ? 0 : (int64_t)p[0].descriptor.size();
}
// xyzsr: n x 5 floats (x, y, z, scale, response) in voxel units as FastHessian leaves them; lap: n ints
void sr_points(void* h, float* xyzsr, int* lap, float* desc) {
  auto& p = static_cast<Run*>(h)->surf->Pts();
  for (size_t i = 0; i < p.size(); i++) {
    xyzsr[5 * i] = p[i].x; xyzsr[5 * i + 1] = p[i].y; xyzsr[5 * i + 2] = p[i].z;
    xyzsr[5 * i + 3] = p[i].scale; xyzsr[5 * i + 4] = p[i].response;
    lap[i] = p[i].laplacian;
    if (desc) std::memcpy(desc + i * p[i].descriptor.size(), p[i].descriptor.data(), p[i].descriptor.size() * sizeof(float));
  }
}
void sr_cast_volume(void* h, int* out) {
  vtkImageData* c = static_cast<Run*>(h)->surf->CastImage();
  std::memcpy(out, c->GetScalarPointer(), c->NumberOfValues() * sizeof(int));
}
void sr_integral_volume(void* h, unsigned long long* out) {
  vtkImageData* c = static_cast<Run*>(h)->surf->GetIntegral();
  std::memcpy(out, c->GetScalarPointer(), c->NumberOfValues() * sizeof(unsigned long long));
}
void sr_write_csv(void* h, const char* path) { static_cast<Run*>(h)->surf->WritePointsCSV(path); }
void sr_write_csvgz(void* h, const char* path, const char* gz_opts, int precision) {
  static_cast<Run*>(h)->surf->WritePointsCSVGZ(path, gz_opts, precision);
}
void sr_write_bin(void* h, const char* path) { static_cast<Run*>(h)->surf->WritePointsBinary(path); }
void sr_write_json(void* h, const char* path) { static_cast<Run*>(h)->surf->WritePoints(path); }

// Response layers of the detector as vtk3DSURF::Update builds it (vtk3DSURF.cxx:193, octaves 4, intervals 4,
// init_sample 2); needs sr_update first (for the integral volume).  Returns the number of layers.
int sr_build_response_map(void* h, double threshold) {
  Run* r = static_cast<Run*>(h);
  delete r->fh;
  r->fh_pts.clear();
  r->fh = new FastHessian(r->surf->GetIntegral(), r->fh_pts, 4, 4, 2, (float)threshold);
  r->fh->buildResponseMap();
  return (int)r->fh->responseMap.size();
}
// info: width, height, depth, step, filter
void sr_layer_info(void* h, int layer, int* info) {
  ResponseLayer* l = static_cast<Run*>(h)->fh->responseMap.at(layer);
  info[0] = l->width; info[1] = l->height; info[2] = l->depth; info[3] = l->step; info[4] = l->filter;
}
// isblob / laplacian are left uninitialised by the reference outside the computed interior (responselayer.h:47-49)
void sr_layer_data(void* h, int layer, float* responses, unsigned char* laplacian, unsigned char* isblob) {
  ResponseLayer* l = static_cast<Run*>(h)->fh->responseMap.at(layer);
  size_t n = (size_t)l->width * l->height * l->depth;
  std::memcpy(responses, l->responses, n * sizeof(float));
  std::memcpy(laplacian, l->laplacian, n);
  for (size_t i = 0; i < n; i++) isblob[i] = l->isblob[i] ? 1 : 0;
}
// detector alone on the integral volume: points in push_back order (before vtk3DSURF sorts them)
int64_t sr_detect(void* h, double threshold, float* xyzsr, int* lap, int64_t cap) {
  Run* r = static_cast<Run*>(h);
  std::vector<Ipoint> pts;
  FastHessian fh(r->surf->GetIntegral(), pts, 4, 4, 2, (float)threshold);
  fh.getIpoints();
  for (size_t i = 0; i < pts.size() && (int64_t)i < cap; i++) {
    xyzsr[5 * i] = pts[i].x; xyzsr[5 * i + 1] = pts[i].y; xyzsr[5 * i + 2] = pts[i].z;
    xyzsr[5 * i + 3] = pts[i].scale; xyzsr[5 * i + 4] = pts[i].response;
    lap[i] = pts[i].laplacian;
  }
  return (int64_t)pts.size();
}

}  // extern "C"
