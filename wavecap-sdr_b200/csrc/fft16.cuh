// In-register radix-16 FFT building blocks on packed complex values (re = lo, im = hi of a 64-bit
// register pair) shared by the channelizer (FFT-256 = 16 x 16) and the spectrum kernel (65536 = 256 x 256).
#pragma once
#include "common.cuh"

namespace wc {
#include "fft16_body.inc"
}  // namespace wc
