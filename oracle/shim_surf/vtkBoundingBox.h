// TEST INFRASTRUCTURE ONLY -- see vtkShimCore.h
#include "vtkShimCore.h"
