// osc_core.cuh -- Oscillator sample arithmetic (src/module/oscillator.rs:15-37,65-92), shared by
// oscillator_kernel (audio_kernels.cu) and the fused voice kernel (fused_voice.cu) so that both produce the
// same bits.
#pragma once

#include "../../include/mixlab_b200.h"
#include "dsp_math.cuh"

namespace mxl {
namespace k {

// `(t + i as u64) as f64 / SAMPLE_RATE as f64` (correctly rounded quotient) `* freq`: oscillator.rs:74-75
__device__ __forceinline__ double osc_phase(double seq, double sr, double inv_sr, double freq)
{
    return div_by_const(seq, sr, inv_sr) * freq;
}

__device__ __forceinline__ float osc_wave(double n, int wf)
{
    switch (wf) {
    case MXL_WAVE_SINE: return (float)wave_sine(n);
    case MXL_WAVE_SQUARE: return (float)sign_bit_f64(wave_sine(n));
    case MXL_WAVE_SAW: return (float)wave_saw(n);
    case MXL_WAVE_TRIANGLE: return (float)wave_triangle(n);
    case MXL_WAVE_ON: return 1.0f;
    default: return 0.0f;
    }
}

// Four consecutive samples of one oscillator.  The waveform is the same for every sample of an instance:
// branch once, not per sample; the four sines go through the reduction and the Horner chain together.
__device__ __forceinline__ void osc_wave4(int wf, const double n[4], float s[4])
{
    if (wf == MXL_WAVE_SINE || wf == MXL_WAVE_SQUARE) {
        double x[4], y[4];
#pragma unroll
        for (int j = 0; j < 4; j++) x[j] = n[j] * kTwoPi;          // oscillator.rs:25-27
        sin_f64x4(x, y);
#pragma unroll
        for (int j = 0; j < 4; j++) s[j] = (float)(wf == MXL_WAVE_SINE ? y[j] : sign_bit_f64(y[j]));
    } else if (wf == MXL_WAVE_SAW) {
#pragma unroll
        for (int j = 0; j < 4; j++) s[j] = (float)wave_saw(n[j]);
    } else {
#pragma unroll
        for (int j = 0; j < 4; j++) s[j] = osc_wave(n[j], wf);
    }
}

// Eight consecutive samples: two groups of four, the sines of both groups in flight together (same operations per
// sample as osc_wave4, so the same bits).
__device__ __forceinline__ void osc_wave8(int wf, const double n[8], float s[8])
{
    if (wf == MXL_WAVE_SINE || wf == MXL_WAVE_SQUARE) {
        double x[8], y[8];
        bool fast = true;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            x[j] = n[j] * kTwoPi;                                  // oscillator.rs:25-27
            const uint32_t hi = high_word(x[j]) & 0x7fffffffu;
            fast = fast && hi < 0x42c00000u && (hi | low_word(x[j])) != 0u;
        }
        if (fast) {
            sin_reduced4(x, y);
            sin_reduced4(x + 4, y + 4);
        } else {
#pragma unroll 1
            for (int j = 0; j < 8; j++) y[j] = sin_f64(x[j]);
        }
#pragma unroll
        for (int j = 0; j < 8; j++) s[j] = (float)(wf == MXL_WAVE_SINE ? y[j] : sign_bit_f64(y[j]));
    } else if (wf == MXL_WAVE_SAW) {
#pragma unroll
        for (int j = 0; j < 8; j++) s[j] = (float)wave_saw(n[j]);
    } else {
#pragma unroll
        for (int j = 0; j < 8; j++) s[j] = osc_wave(n[j], wf);
    }
}

}  // namespace k
}  // namespace mxl
