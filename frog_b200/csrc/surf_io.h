// Host I/O of bin/surf3d: a MetaImage (.mhd / .mha) volume reader and the reference's keypoint writers.
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "frogsurf.h"

namespace fsio {

struct Volume {
  int dims[3] = {0, 0, 0};
  double spacing[3] = {1, 1, 1};
  double origin[3] = {0, 0, 0};
  int voxel_type = FS_I16;  // fs_voxel_type
  std::vector<unsigned char> data;
};

// Uncompressed little-endian MetaImage, one component, element types MET_UCHAR / MET_SHORT / MET_USHORT / MET_INT /
// MET_FLOAT, data in the same file (ElementDataFile = LOCAL) or in the named file next to the header.  Axis flips for
// negative TransformMatrix diagonals (vtkRobustImageReader.h:52-60, 97-113) are applied.
bool read_metaimage(const std::string& path, Volume& out, std::string& err);
bool write_metaimage(const std::string& path, const Volume& v, std::string& err);

// vtk3DSURF::WritePointsCSV / WritePointsCSVGZ / WritePointsBinary (vtk3DSURF.cxx:405-525): voxel-unit keypoints are
// written as origin + coordinate * spacing, scale * cbrt(spacing product), laplacian, response, descriptor.
bool write_points_csv(const std::string& path, const fs_point* pts, const float* desc, size_t n, size_t dsize,
                      const double spacing[3], const double origin[3]);
bool write_points_csvgz(const std::string& path, const char* gz_opts, int precision, const fs_point* pts, const float* desc,
                        size_t n, size_t dsize, const double spacing[3], const double origin[3]);
bool write_points_bin(const std::string& path, const fs_point* pts, const float* desc, size_t n, size_t dsize,
                      const double spacing[3], const double origin[3]);
// vtk3DSURF::ReadIPoints (vtk3DSURF.cxx:34-77, surf3d -p): one keypoint per line, "x,y,z,scale" in world units; returns
// x, y, z, scale per point in VOXEL units (as fs_set_points takes them).  A cell that is missing at the end of a line
// repeats the previous cell (the reference's getline leaves its string untouched), an empty line is skipped, a cell
// that is not a number is an error (std::stof throws in the reference).  n_outside counts the points the reference
// reports as "outside image" (they are kept).
bool read_points_file(const std::string& path, const double spacing[3], const double origin[3], const int dims[3],
                      std::vector<float>& xyzs, size_t& n_outside, std::string& err);

// surf3d.cxx:269-285: {"bounds":{"xmax":..,"xmin":..,...}} in picojson's number format
bool write_bounds_json(const std::string& path, const Volume& v);

}  // namespace fsio
