// svgf_tu_staged_f32.cu — packed kernel writing lattice planes, fp32 storage (see svgf_tu_staged.inl)
#define SVGF_TU_F32 true
#define SVGF_TU_STAGED_ENTRY atrous_packed_staged_f32
#include "svgf_tu_staged.inl"
