// Test-infrastructure shim standing in for tools/transformIO.h (VTK is not installed).
// match.cpp only needs the type name and readTransform() for the -transformPrefix flag
// (match.cpp:516-525,550-553), which the oracle never exercises.
#pragma once
#include <cstring>
#include <algorithm>
#include <array>
struct vtkGeneralTransform { void TransformPoint(const float*, float*) {} };
inline vtkGeneralTransform* readTransform(const char*) { return nullptr; }
