// Instantiation of the hand-written FFT . G . iFFT passes (fft3d.cuh) for float meshes with the
// Green's function evaluated in float.
#include "fft3d.cuh"

namespace tpme {
int filter_pow2_f32(const void* in, void* out, void* hat, int channels, int nx, int ny, int nz,
                    const GreenDev<float>& green, void* dc_out, cudaStream_t s) {
  return fft::filter_pow2<float, float, true>(in, out, hat, channels, nx, ny, nz, green, dc_out, s);
}
int slab_x_f32(void* hat, int channels, int nx, int ny, int nz, int y0, int ny_local,
               const GreenDev<float>& green, void* dc_out, cudaStream_t s, const RemoteStore* rs) {
  return fft::x_pass_green<float, float, true>(hat, channels, nx, ny, nz, y0, ny_local, green, dc_out, s, rs);
}
int slab_yz_f32(bool forward, void* real, void* hat, int planes, int ny, int nz, cudaStream_t s,
                const RemoteStore* rs) {
  return fft::yz_passes<float, float>(forward, real, hat, planes, ny, nz, s, rs);
}
}  // namespace tpme
