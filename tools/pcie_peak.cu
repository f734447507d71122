// pcie_peak.cu -- pinned-host copy bandwidth of the box (H2D alone, D2H alone, both at once): the
// ceiling of bench.py's e2e number, which moves 2 layers up and 1 frame down per tick.
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>

static float timed(cudaStream_t s0, cudaStream_t s1, void* d0, void* h0, void* d1, void* h1, size_t n, int up, int down, int reps, size_t piece)
{
    cudaEvent_t a, b, c; cudaEventCreate(&a); cudaEventCreate(&b); cudaEventCreate(&c);
    cudaDeviceSynchronize();
    cudaEventRecord(a, s0);
    cudaStreamWaitEvent(s1, a, 0);
    for (int r = 0; r < reps; r++) {
        for (size_t o = 0; o < n; o += piece) {
            size_t len = o + piece <= n ? piece : n - o;
            if (up) cudaMemcpyAsync((char*)d0 + o, (char*)h0 + o, len, cudaMemcpyHostToDevice, s0);
            if (down) cudaMemcpyAsync((char*)h1 + o, (char*)d1 + o, len, cudaMemcpyDeviceToHost, s1);
        }
    }
    cudaEventRecord(c, s1);
    cudaStreamWaitEvent(s0, c, 0);
    cudaEventRecord(b, s0);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main(int argc, char** argv)
{
    const size_t n = 512ull << 20;
    void *h0, *h1, *d0, *d1;
    // `wc`: the upload buffer write-combined (not snooped): does the H2D side gain?
    const bool wc = argc > 1 && !strcmp(argv[1], "wc");
    cudaHostAlloc(&h0, n, wc ? cudaHostAllocWriteCombined : cudaHostAllocDefault); cudaHostAlloc(&h1, n, cudaHostAllocDefault);
    if (wc) printf("{\"upload_buffer\": \"write-combined\"}\n");
    cudaMalloc(&d0, n); cudaMalloc(&d1, n);
    memset(h0, 1, n); memset(h1, 2, n);
    cudaStream_t s0, s1; cudaStreamCreate(&s0); cudaStreamCreate(&s1);
    const int reps = 4;
    for (size_t piece : {n, (size_t)3110400, (size_t)(64 << 10)}) {
        timed(s0, s1, d0, h0, d1, h1, n, 1, 1, 1, piece);
        float up = timed(s0, s1, d0, h0, d1, h1, n, 1, 0, reps, piece);
        float dn = timed(s0, s1, d0, h0, d1, h1, n, 0, 1, reps, piece);
        float both = timed(s0, s1, d0, h0, d1, h1, n, 1, 1, reps, piece);
        double gb = (double)n * reps / 1e9;
        printf("{\"piece_bytes\": %zu, \"h2d_gbs\": %.2f, \"d2h_gbs\": %.2f, \"both_each_gbs\": %.2f}\n", piece, gb / (up * 1e-3), gb / (dn * 1e-3), gb / (both * 1e-3));
    }
    // 2:1 mix as in the e2e step (two bytes up per byte down)
    {
        cudaEvent_t a, b, c; cudaEventCreate(&a); cudaEventCreate(&b); cudaEventCreate(&c);
        cudaDeviceSynchronize();
        cudaEventRecord(a, s0); cudaStreamWaitEvent(s1, a, 0);
        for (int r = 0; r < reps; r++) {
            cudaMemcpyAsync(d0, h0, n, cudaMemcpyHostToDevice, s0);
            cudaMemcpyAsync(h1, d1, n / 2, cudaMemcpyDeviceToHost, s1);
        }
        cudaEventRecord(c, s1); cudaStreamWaitEvent(s0, c, 0); cudaEventRecord(b, s0); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("{\"mix\": \"2 up : 1 down\", \"h2d_gbs\": %.2f, \"d2h_gbs\": %.2f}\n", (double)n * reps / 1e9 / (ms * 1e-3), (double)n * reps / 2e9 / (ms * 1e-3));
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { fprintf(stderr, "cuda error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
