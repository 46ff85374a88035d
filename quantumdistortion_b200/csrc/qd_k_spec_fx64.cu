// float64 spectral pass (parity path), FX variants; n_fft 8192 runs one warp per CTA, with the formant scratch the
// samples are staged inside it (SA)
#include "qd_spec_launch.inl"
QD_INSTANTIATE_SPEC(double, 256, 8, false, true, 1, false)
QD_INSTANTIATE_SPEC(double, 512, 8, false, true, 1, false)
QD_INSTANTIATE_SPEC(double, 1024, 4, false, true, 1, false)
QD_INSTANTIATE_SPEC(double, 1024, 6, false, true, 1, false)
QD_INSTANTIATE_SPEC(double, 2048, 2, false, true, 1, false)
QD_INSTANTIATE_SPEC(double, 4096, 1, false, true, 1, false)
QD_INSTANTIATE_SPEC(double, 4096, 1, false, true, 1, true)
