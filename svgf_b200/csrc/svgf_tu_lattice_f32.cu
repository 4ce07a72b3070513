// svgf_tu_lattice_f32.cu — TMA-staged lattice levels, fp32 storage (see svgf_tu_lattice.inl)
#define SVGF_TU_F32 true
#define SVGF_TU_LATTICE_ENTRY atrous_lattice_f32
#include "svgf_tu_lattice.inl"
