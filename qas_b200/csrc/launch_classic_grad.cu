// classic evaluation kernels, gradient instantiations (see launch_classic.inl)
#define QAS_GRAD 1
#include "launch_classic.inl"
