// pipe_rates.cu -- measured issue rates of the FP64 pipe and the f32<->f64 converts on the device the
// EqThree / Mixer / Amplifier / Oscillator kernels run on.  Build: nvcc -gencode arch=compute_100a,code=sm_100a
// -O3 -fmad=false tools/pipe_rates.cu -o tools/pipe_rates.bin ; run on the GPU box; prints one JSON line per test.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITER = 4096;
constexpr int CH = 8;

template <int OP>
__global__ void __launch_bounds__(256) rate_kernel(double* out, double a, double b, float fa)
{
    double x[CH];
    float f[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) { x[i] = a * (threadIdx.x + i + 1); f[i] = fa * (threadIdx.x + i + 1); }
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) {
            if (OP == 0) x[i] = fma(x[i], b, a);                          // DFMA
            if (OP == 1) x[i] = x[i] + a;                                  // DADD
            if (OP == 2) x[i] = x[i] * b;                                  // DMUL
            if (OP == 3) { x[i] = x[i] + (double)f[i]; f[i] = f[i] + fa; } // DADD + F2F.F64.F32 + FADD
            if (OP == 4) { f[i] = (float)(x[i]) + f[i]; x[i] = x[i] + a; } // DADD + F2F.F32.F64 + FADD
            if (OP == 5) f[i] = fmaf(f[i], fa, fa);                        // FFMA
            if (OP == 6) { f[i] = f[i] + fa; }                             // FADD
            if (OP == 7) f[i] = (float)((double)f[i] * b);                 // cvt, DMUL, cvt (mixer/amplifier inner op)
        }
    }
    double s = 0; float g = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) { s += x[i]; g += f[i]; }
    if (s == 12345.678 || g == 3.25f) out[0] = s + g;
}

// one chain, one warp per SM: latency
template <int OP>
__global__ void lat_kernel(double* out, double a, double b, long long* cyc)
{
    double x = a * (threadIdx.x + 1);
    long long t0 = clock64();
    for (int it = 0; it < ITER; it++) {
        if (OP == 0) x = fma(x, b, a);
        if (OP == 1) x = x + a;
        if (OP == 2) x = x * b;
    }
    long long t1 = clock64();
    if (x == 12345.678) out[0] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int OP>
void run(const char* name, int ops_per_iter, double* d_out)
{
    int dev; cudaGetDevice(&dev);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    const int grid = p.multiProcessorCount * 8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    rate_kernel<OP><<<grid, 256>>>(d_out, 1.0000001, 0.9999999, 1.0000001f);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        rate_kernel<OP><<<grid, 256>>>(d_out, 1.0000001, 0.9999999, 1.0000001f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double lane_iters = (double)grid * 256 * ITER * CH;
    const double per_s = lane_iters / (best * 1e-3);
    printf("{\"test\": \"%s\", \"ms\": %.4f, \"G_lane_iters_per_s\": %.1f, \"lane_iters_per_clk_per_sm_at_max_clock\": %.2f, \"ops_per_iter\": %d}\n",
           name, best, per_s / 1e9, per_s / (p.multiProcessorCount * (double)khz * 1e3), ops_per_iter);
}

template <int OP>
void lat(const char* name, double* d_out, long long* d_cyc)
{
    lat_kernel<OP><<<1, 32>>>(d_out, 1.0000001, 0.9999999, d_cyc);
    lat_kernel<OP><<<1, 32>>>(d_out, 1.0000001, 0.9999999, d_cyc);
    long long c; cudaMemcpy(&c, d_cyc, sizeof c, cudaMemcpyDeviceToHost);
    printf("{\"test\": \"%s\", \"cycles_per_dependent_op\": %.2f}\n", name, (double)c / ITER);
}

int main()
{
    double* d_out; long long* d_cyc;
    cudaMalloc(&d_out, 64); cudaMalloc(&d_cyc, 64);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    printf("{\"device\": \"%s\", \"sms\": %d, \"max_clock_mhz\": %d}\n", p.name, p.multiProcessorCount, khz / 1000);
    run<0>("DFMA", 1, d_out);
    run<1>("DADD", 1, d_out);
    run<2>("DMUL", 1, d_out);
    run<3>("DADD+F2F.F64.F32+FADD", 3, d_out);
    run<4>("DADD+F2F.F32.F64+FADD", 3, d_out);
    run<5>("FFMA", 1, d_out);
    run<6>("FADD", 1, d_out);
    run<7>("F2F.F64.F32+DMUL+F2F.F32.F64", 3, d_out);
    lat<0>("DFMA latency", d_out, d_cyc);
    lat<1>("DADD latency", d_out, d_cyc);
    lat<2>("DMUL latency", d_out, d_cyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { fprintf(stderr, "cuda error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
