// float32 spectral pass, FX variants at the other n_fft
#include "qd_spec_launch.inl"
QD_INSTANTIATE_SPEC(float, 256, 8, false, true, 1, false)
QD_INSTANTIATE_SPEC(float, 512, 8, false, true, 1, false)
QD_INSTANTIATE_SPEC(float, 2048, 4, false, true, 1, false)
QD_INSTANTIATE_SPEC(float, 4096, 2, false, true, 1, false)
QD_INSTANTIATE_SPEC(float, 4096, 1, false, true, 1, false)
