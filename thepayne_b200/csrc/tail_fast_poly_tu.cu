// Translation unit of the fast fused tail with continuum polynomial and / or model output (tail_fast_tu.inl).
#define PAYNE_TU_POLY 1
#include "tail_fast_tu.inl"
