// bin/surf3d -- the SURF3D producer's executable over libfrogsurf.so.
// Same command line as the reference's surf3d (vtkOpenSURF3D/surf3d.cxx:16-157): the option loop advances two
// tokens per key and ignores keys it does not know.  Options that need VTK's own image filters (-s / -d
// resampling, -m mask, -pad, -type 2) or picojson (-json 1) are rejected loudly instead of being approximated; the input is a
// MetaImage volume already at its final sampling.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include "frogsurf.h"
#include "surf_io.h"

using std::cout;
using std::endl;

namespace {
double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int unsupported(const char* what) {
  std::cerr << "surf3d (B200): " << what << " needs a VTK image filter that this build does not restate; terminating." << endl;
  return 7;
}

template <typename T>
void clamp_values(std::vector<unsigned char>& data, bool lo_on, float lo, bool hi_on, float hi) {
  // vtkImageThreshold with ReplaceOut (surf3d.cxx:232-255): values below -cmin become cmin, above -cmax become cmax
  T* p = reinterpret_cast<T*>(data.data());
  const size_t n = data.size() / sizeof(T);
  for (size_t i = 0; i < n; i++) {
    if (lo_on && (double)p[i] < (double)lo) p[i] = (T)lo;
    if (hi_on && (double)p[i] > (double)hi) p[i] = (T)hi;
  }
}
}  // namespace

int main(int argc, char* argv[]) {
  if (argc < 2) {
    cout << "Usage : surf3d file [options]" << endl;
    cout << "Available options:" << endl;
    cout << "-bin 0/1       : write points as bin file. Default : 0" << endl;
    cout << "-csv 0/1       : write points as csv file. Default : 0" << endl;
    cout << "-cmin value    : clamp values lower than specified value" << endl;
    cout << "-cmax value    : clamp values larger than specified value" << endl;
    cout << "-csvgz 0/1     : write points as csv.gz file. Default : 1" << endl;
    cout << "-gz opts       : set gz options such as compression level" << endl;
    cout << "-precision n   : set coefficients precision in csv.gz file" << endl;
    cout << "-n number      : maximum number of points" << endl;
    cout << "-normalize 0/1 : normalize descriptors (default : 1 )" << endl;
    cout << "-o basename    : set output file name. Default: \"points\"" << endl;
    cout << "-p file        : describe the points of a csv file (x,y,z,scale) instead of detecting" << endl;
    cout << "-r radius      : descriptor volume radius. Default : 5" << endl;
    cout << "-t threshold   : set detector threshold. Default: 0" << endl;
    cout << "-type 0/1      : set descriptor type :" << endl;
    cout << "       0 : SURF3D descriptor (default). Descriptor size : 48" << endl;
    cout << "       1 : subvolume HAAR coefficients. Descriptor size : 24 * radius^3" << endl;
    cout << "-gpu id        : CUDA device. Default : 0" << endl;
    exit(1);
  }

  double spacing = 0, threshold = 0;
  int maxSize = 0, writeJSON = 0, writeBIN = 0, writeCSV = 0, writeCSVGZ = 1, descriptorType = 0, numberOfPoints = -1;
  int pad = 0, subVolumeRadius = 5, precision = -1, gpu = 0;
  char *maskfilename = 0, *pointFile = 0, *gzOpts = 0;
  bool clampMinValues = false, clampMaxValues = false, normalize = true;
  float clampMinValue = 0, clampMaxValue = 0;
  std::string outfilename("points");

  int argumentsIndex = 2;
  while (argumentsIndex < argc) {
    char* key = argv[argumentsIndex];
    char* value = argumentsIndex + 1 < argc ? argv[argumentsIndex + 1] : (char*)"";
    if (strcmp(key, "-d") == 0) maxSize = atoi(value);
    if (strcmp(key, "-s") == 0) spacing = atof(value);
    if (strcmp(key, "-t") == 0) threshold = atof(value);
    if (strcmp(key, "-cmin") == 0) { clampMinValues = true; clampMinValue = atof(value); }
    if (strcmp(key, "-cmax") == 0) { clampMaxValues = true; clampMaxValue = atof(value); }
    if (strcmp(key, "-m") == 0) maskfilename = value;
    if (strcmp(key, "-o") == 0) outfilename = value;
    if (strcmp(key, "-json") == 0) writeJSON = atoi(value);
    if (strcmp(key, "-csv") == 0) writeCSV = atoi(value);
    if (strcmp(key, "-bin") == 0) writeBIN = atoi(value);
    if (strcmp(key, "-csvgz") == 0) writeCSVGZ = atoi(value);
    if (strcmp(key, "-type") == 0) descriptorType = atoi(value);
    if (strcmp(key, "-n") == 0) numberOfPoints = atoi(value);
    if (strcmp(key, "-p") == 0) pointFile = value;
    if (strcmp(key, "-r") == 0) subVolumeRadius = atoi(value);
    if (strcmp(key, "-normalize") == 0) normalize = atoi(value);
    if (strcmp(key, "-pad") == 0) pad = atoi(value);
    if (strcmp(key, "-gz") == 0) gzOpts = value;
    if (strcmp(key, "-precision") == 0) precision = atoi(value);
    if (strcmp(key, "-gpu") == 0) gpu = atoi(value);
    argumentsIndex += 2;
  }
  if (spacing != 0 || maxSize > 0) return unsupported("-s / -d (vtkImageResample)");
  if (maskfilename) return unsupported("-m (mask resampling)");
  if (pad) return unsupported("-pad (vtkImageMirrorPad)");
  if (descriptorType == 2) return unsupported("-type 2 (vtkImageResize)");
  if (writeJSON) return unsupported("-json 1 (picojson point dump)");

  cout << "load : " << argv[1] << endl;
  double t0 = now();
  fsio::Volume vol;
  std::string err;
  if (!fsio::read_metaimage(argv[1], vol, err)) {
    std::cerr << "Cannot load file " << argv[1] << " as an image file; terminating.\n";
    std::cerr << err << endl;
    return 5;
  }
  if (clampMinValues || clampMaxValues) {
    switch (vol.voxel_type) {
      case FS_U8: clamp_values<uint8_t>(vol.data, clampMinValues, clampMinValue, clampMaxValues, clampMaxValue); break;
      case FS_I16: clamp_values<int16_t>(vol.data, clampMinValues, clampMinValue, clampMaxValues, clampMaxValue); break;
      case FS_U16: clamp_values<uint16_t>(vol.data, clampMinValues, clampMinValue, clampMaxValues, clampMaxValue); break;
      case FS_I32: clamp_values<int32_t>(vol.data, clampMinValues, clampMinValue, clampMaxValues, clampMaxValue); break;
      default: clamp_values<float>(vol.data, clampMinValues, clampMinValue, clampMaxValues, clampMaxValue); break;
    }
  }
  cout << "Image loaded in " << now() - t0 << "s" << endl;
  fsio::write_bounds_json(outfilename + ".json", vol);

  fs_ctx* ctx = nullptr;
  if (fs_create(gpu, &ctx) != FS_OK) {
    std::cerr << "surf3d (B200): " << fs_last_error(nullptr) << endl;
    return 1;
  }
  auto die = [&](const char* what) {
    std::cerr << "surf3d (B200): " << what << ": " << fs_last_error(ctx) << endl;
    fs_destroy(ctx);
    return 1;
  };
  cout << "Image dimensions : " << vol.dims[0] << " " << vol.dims[1] << " " << vol.dims[2] << endl;
  cout << "Initial spacing  : " << vol.spacing[0] << " " << vol.spacing[1] << " " << vol.spacing[2] << endl;
  fs_stats st;
  t0 = now();
  if (fs_set_volume(ctx, vol.data.data(), vol.voxel_type, vol.dims[0], vol.dims[1], vol.dims[2]) != FS_OK) return die("volume");
  cout << "Integral computed in " << now() - t0 << "s" << endl;
  t0 = now();
  uint32_t n = 0;
  if (pointFile) {  // vtk3DSURF.cxx:183: ReadIPoints instead of the detector
    cout << "Use points in " << pointFile << endl;
    cout << "Read : " << pointFile << endl;
    std::vector<float> xyzs;
    size_t outside = 0;
    if (!fsio::read_points_file(pointFile, vol.spacing, vol.origin, vol.dims, xyzs, outside, err)) {
      std::cerr << "surf3d (B200): " << err << endl;
      fs_destroy(ctx);
      return 1;
    }
    if (outside) cout << "Error : " << outside << " points are outside image" << endl;
    n = (uint32_t)(xyzs.size() / 4);
    if (fs_set_points(ctx, xyzs.data(), n) != FS_OK) return die("points");
  } else {
    if (fs_detect(ctx, (float)threshold, &n) != FS_OK) return die("detector");
    cout << " Ipoints : " << n << endl;
    cout << "FastHessian computed in " << now() - t0 << "s" << endl;
  }
  t0 = now();
  if (fs_select(ctx, numberOfPoints) != FS_OK) return die("select");
  uint32_t dsize = 0;
  if (fs_describe(ctx, descriptorType, subVolumeRadius, normalize) != FS_OK) return die("descriptors");
  fs_num_points(ctx, &n, &dsize);
  cout << "Number of keypoints : " << n << endl;
  std::vector<fs_point> pts(n);
  std::vector<float> desc((size_t)n * dsize);
  if (fs_get_points(ctx, pts.data(), desc.data()) != FS_OK) return die("read back");
  cout << "Descriptors computed in " << now() - t0 << "s" << endl;
  fs_get_stats(ctx, &st);
  cout << "GPU ms : integral " << st.ms_integral << ", response map " << st.ms_response_map << ", extrema " << st.ms_extrema
       << ", descriptors " << st.ms_describe << endl;
  if (st.n_clamped) cout << "Warning : " << st.n_clamped << " keypoints have descriptor windows leaving the volume" << endl;
  fs_destroy(ctx);

  if (writeBIN) {
    t0 = now();
    fsio::write_points_bin(outfilename + ".bin", pts.data(), desc.data(), n, dsize, vol.spacing, vol.origin);
    cout << "bin written in " << now() - t0 << "s" << endl;
  }
  if (writeCSV) {
    t0 = now();
    fsio::write_points_csv(outfilename + ".csv", pts.data(), desc.data(), n, dsize, vol.spacing, vol.origin);
    cout << "csv written in " << now() - t0 << "s" << endl;
  }
  if (writeCSVGZ) {
    t0 = now();
    fsio::write_points_csvgz(outfilename + ".csv.gz", gzOpts, precision, pts.data(), desc.data(), n, dsize, vol.spacing, vol.origin);
    cout << "csvgz written in " << now() - t0 << "s" << endl;
  }
  return 0;
}
