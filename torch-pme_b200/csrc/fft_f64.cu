// Instantiation of the hand-written FFT . G . iFFT passes (fft3d.cuh) for double meshes with the
// Green's function evaluated in double.
#include "fft3d.cuh"

namespace tpme {
int filter_pow2_f64(const void* in, void* out, void* hat, int channels, int nx, int ny, int nz,
                    const GreenDev<double>& green, void* dc_out, cudaStream_t s) {
  return fft::filter_pow2<double, double, true>(in, out, hat, channels, nx, ny, nz, green, dc_out, s);
}
}  // namespace tpme
