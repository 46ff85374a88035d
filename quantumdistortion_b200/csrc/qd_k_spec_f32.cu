// float32 spectral pass, plain (no FX) variants -- the headline kernel <float, 1024, 8, TS, noFX, NG = 2> among them
#include "qd_spec_launch.inl"
QD_INSTANTIATE_SPEC_EF(float, 1024, 8, true, false, 2, false, true)   // reference defaults: no float64 in the epilogue
QD_INSTANTIATE_SPEC(float, 1024, 8, true, false, 2, false)
QD_INSTANTIATE_SPEC(float, 1024, 8, false, false, 1, false)
QD_INSTANTIATE_SPEC(float, 256, 8, false, false, 1, false)
QD_INSTANTIATE_SPEC(float, 512, 8, false, false, 1, false)
QD_INSTANTIATE_SPEC(float, 2048, 4, false, false, 1, false)
QD_INSTANTIATE_SPEC(float, 4096, 4, false, false, 1, false)
