// Host build of mixlab_b200/csrc/dsp_math.cuh for CPU-side verification (tests only).
#include "../../mixlab_b200/csrc/dsp_math.cuh"
#include <stddef.h>

extern "C" {
void mxl_host_sin(const double* x, double* y, size_t n) { for (size_t i = 0; i < n; i++) y[i] = mxl::sin_f64(x[i]); }
void mxl_host_sin4(const double* x, double* y, size_t n)
{
    size_t i = 0;
    for (; i + 4 <= n; i += 4) mxl::sin_f64x4(x + i, y + i);
    for (; i < n; i++) y[i] = mxl::sin_f64(x[i]);
}
void mxl_host_libm_sin(const double* x, double* y, size_t n) { for (size_t i = 0; i < n; i++) y[i] = sin(x[i]); }
void mxl_host_div(const double* a, double b, double* q, size_t n)
{
    double inv = 1.0 / b;
    for (size_t i = 0; i < n; i++) q[i] = mxl::div_by_const(a[i], b, inv);
}
// oscillator phase exactly as the kernels form it
void mxl_host_osc_phase(unsigned long long t, double sr, double freq, double* x, size_t n)
{
    for (size_t i = 0; i < n; i++) { double t0 = (double)(t + i) / sr; x[i] = t0 * freq * 2.0 * mxl::kPi; }
}
}
