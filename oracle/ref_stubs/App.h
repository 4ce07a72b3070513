// Build stub (test infrastructure): stands in for the reference's src/App.h when compiling its
// src/Filter.cuh standalone.  Only the one type Filter.cuh needs from it is declared, with the member
// layout of src/App.h:41-44.
#pragma once
#include <cstdint>
#include <glm/glm.hpp>
namespace gpupt {
struct cudaFramebuffer {
    unsigned long long PositionTexture, NormalTexture, UVTexture, MotionTexture;
};
}  // namespace gpupt
