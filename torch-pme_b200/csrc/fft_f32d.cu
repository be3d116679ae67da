// Instantiation of the hand-written FFT . G . iFFT passes (fft3d.cuh) for float meshes with the
// Green's function evaluated in double.
#include "fft3d.cuh"

namespace tpme {
int filter_pow2_f32d(const void* in, void* out, void* hat, int channels, int nx, int ny, int nz,
                    const GreenDev<double>& green, void* dc_out, cudaStream_t s) {
  return fft::filter_pow2<float, double, false>(in, out, hat, channels, nx, ny, nz, green, dc_out, s);
}
int slab_x_f32d(void* hat, int channels, int nx, int ny, int nz, int y0, int ny_local,
                const GreenDev<double>& green, void* dc_out, cudaStream_t s, const RemoteStore* rs) {
  return fft::x_pass_green<float, double, false>(hat, channels, nx, ny, nz, y0, ny_local, green, dc_out, s, rs);
}
}  // namespace tpme
