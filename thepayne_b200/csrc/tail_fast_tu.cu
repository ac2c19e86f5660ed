// Translation unit of the fast fused tail without continuum polynomial / model output (tail_fast_tu.inl).
#define PAYNE_TU_POLY 0
#include "tail_fast_tu.inl"
