// classic evaluation kernels, energy-only instantiations (see launch_classic.inl)
#define QAS_GRAD 0
#include "launch_classic.inl"
