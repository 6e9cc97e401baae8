// fm_prep_types.h -- host/device table entry of the batched image preparation (fm_prep.cuh).
#pragma once
#include <cstdint>

namespace fm {

struct PrepSeg {
  uint32_t img;       // image index (ImageDev table / meta slot)
  uint32_t n;         // keypoints
  uint32_t off;       // first keypoint of this image in the batch-wide arrays
  uint32_t blk_keys;  // first 256-thread block of this image in prep_keys_kernel
  uint32_t blk_pack;  // first 256-thread block of this image in prep_pack_kernel
  uint32_t pad_[3];
};

}  // namespace fm
