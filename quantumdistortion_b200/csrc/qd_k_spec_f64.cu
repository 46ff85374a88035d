// float64 spectral pass (parity path), plain variants
#include "qd_spec_launch.inl"
QD_INSTANTIATE_SPEC(double, 256, 8, false, false, 1, false)
QD_INSTANTIATE_SPEC(double, 512, 8, false, false, 1, false)
QD_INSTANTIATE_SPEC(double, 1024, 8, false, false, 1, false)
QD_INSTANTIATE_SPEC(double, 2048, 4, false, false, 1, false)
QD_INSTANTIATE_SPEC(double, 4096, 2, false, false, 1, false)
