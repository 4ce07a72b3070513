// svgf_tu_staged_f16.cu — packed kernel writing lattice planes, fp16 storage (see svgf_tu_staged.inl)
#define SVGF_TU_F32 false
#define SVGF_TU_STAGED_ENTRY atrous_packed_staged_f16
#include "svgf_tu_staged.inl"
