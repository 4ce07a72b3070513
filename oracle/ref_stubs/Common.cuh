// Build stub (test infrastructure): stands in for the reference's src/Common.cuh (which does not compile
// on its own) when building src/Filter.cuh standalone.  Provides the five names Filter.cuh
// uses from it: FN_DECL, GLOBAL_ID, INOUT (src/Common.cuh:8-16), commonCu::IsFinite and saturate()
// (a CUDA <= 11 builtin removed from CUDA 12 headers; __saturatef is the same instruction).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <glm/glm.hpp>
#define FN_DECL __device__
#define INOUT(Type) Type &
#define GLOBAL_ID() glm::uvec2(blockIdx.x *blockDim.x + threadIdx.x, blockIdx.y * blockDim.y + threadIdx.y)
namespace commonCu {
__device__ inline bool IsFinite(float x) { return !isnan(x); }
__device__ inline bool IsFinite(glm::vec3 v) { return !(isnan(v.x) || isnan(v.y) || isnan(v.z)); }
__device__ inline bool IsFinite(glm::vec4 v) { return !(isnan(v.x) || isnan(v.y) || isnan(v.z) || isnan(v.w)); }
}  // namespace commonCu
__device__ inline float saturate(float x) { return __saturatef(x); }
