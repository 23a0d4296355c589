// tsff_data.cu -- data-side stage in front of the fit: lineout extraction from the CCD image,
// tsadar/utils/process/lineouts.py:85-165 (get_lineouts), for the electron or the ion spectrometer:
//
//   raw[l][y]    = sum_{x = a_l - dpixel}^{a_l + dpixel - 1} image[y][x]                     (:87-90, 108-111; 2 dpixel columns)
//   smooth[l][y] = (1 / span) sum_{k = -dpixel}^{dpixel} raw[l][y + k],  span = 2 dpixel + 1  (:91-93, 112-114; np.convolve 'same')
//   data[l][y]   = smooth[l][y] / gain                                                        (:125-127, 141-143)
//   amps[l]      = max over the fit windows of data[l][:]                                     (:128-137, 145-152)
//
// HBM-bound byte work: the image is read once per lineout (2 dpixel contiguous values per row), one CTA per lineout, the
// column sums staged in shared memory for the boxcar.  FP64 like the reference (the CCD counts are small integers: exact).
#include "tsff_common.cuh"

using namespace tsff;

namespace {
constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads) k_lineouts(const double* __restrict__ image, int NY, int NX, const int* __restrict__ pixels,
                                                       int dpixel, double gain, const unsigned char* __restrict__ window,
                                                       double* __restrict__ data, double* __restrict__ amps) {
  extern __shared__ double s_raw[];      // [NY]
  __shared__ double sred[kThreads / 32];
  const int l = blockIdx.x;
  const int a = pixels[l];
  for (int y = threadIdx.x; y < NY; y += kThreads) {
    const double* row = image + (long long)y * NX;
    double s = 0.0;
    for (int x = a - dpixel; x < a + dpixel; x++) s += row[x];      // the reference's slice [a - dpixel : a + dpixel]
    s_raw[y] = s;
  }
  __syncthreads();
  const double inv = 1.0 / (double)(2 * dpixel + 1);
  double mx = -1.0 / 0.0;
  for (int y = threadIdx.x; y < NY; y += kThreads) {
    double s = 0.0;
    for (int k = -dpixel; k <= dpixel; k++) {
      const int yy = y + k;
      if (yy >= 0 && yy < NY) s += s_raw[yy] * inv;                  // np.convolve(raw, ones(span) / span, "same"), in tap order
    }
    const double v = s / gain;
    data[(long long)l * NY + y] = v;
    if (!window || window[y]) mx = fmax(mx, v);
  }
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0 && amps) {
    double m = sred[0];
    for (int w = 1; w < kThreads / 32; w++) m = fmax(m, sred[w]);
    amps[l] = m;
  }
}
}  // namespace

extern "C" int tsff_lineouts_fwd(const double* image, int32_t NY, int32_t NX, const int32_t* pixels, const int32_t* pixels_host, int32_t L,
                                 int32_t dpixel, double gain, const unsigned char* window, double* data, double* amps, void* stream) {
  if (L == 0) return TSFF_OK;
  if (!image || !pixels || !pixels_host || !data || NY < 1 || NX < 1 || L < 0 || dpixel < 0 || !(gain != 0.0)) { set_error("tsff_lineouts_fwd: bad argument"); return TSFF_E_INVALID; }
  for (int l = 0; l < L; l++)
    if (pixels_host[l] - dpixel < 0 || pixels_host[l] + dpixel > NX) { set_error("lineout %d at pixel %d +- %d leaves the image of %d columns", l, pixels_host[l], dpixel, NX); return TSFF_E_INVALID; }
  if ((size_t)NY * 8 > 200 * 1024) { set_error("image too tall for shared-memory staging (%d rows)", NY); return TSFF_E_INVALID; }
  TSFF_SMEM_OPTIN(k_lineouts);
  k_lineouts<<<(unsigned)L, kThreads, (size_t)NY * 8, static_cast<cudaStream_t>(stream)>>>(image, NY, NX, pixels, dpixel, gain, window, data, amps);
  TSFF_LAUNCH_OK("k_lineouts");
  return TSFF_OK;
}
