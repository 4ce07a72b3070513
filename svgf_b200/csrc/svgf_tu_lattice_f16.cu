// svgf_tu_lattice_f16.cu — TMA-staged lattice levels, fp16 storage (see svgf_tu_lattice.inl)
#define SVGF_TU_F32 false
#define SVGF_TU_LATTICE_ENTRY atrous_lattice_f16
#include "svgf_tu_lattice.inl"
