// float32 spectral pass, FX variants (spectral FX / freeze / formant shift) at n_fft 2048
#include "qd_spec_launch.inl"
QD_INSTANTIATE_SPEC(float, 1024, 12, true, true, 1, false)
QD_INSTANTIATE_SPEC(float, 1024, 12, false, true, 1, false)
QD_INSTANTIATE_SPEC(float, 1024, 8, false, true, 1, false)
