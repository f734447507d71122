// pcie_multi.cu -- aggregate pinned-host copy bandwidth of the box with N GPUs pulling AT ONCE: the ceiling of
// bench.py's host-fed (e2e) number at N GPUs, which moves two 1080p layers up and one composite down per tick and
// rank.  One host thread + its own streams per GPU, the e2e leg's 2-up : 1-down mix, wall clock over all GPUs.
//
//   pcie_multi.bin <n_gpus> [affinity] [streams=K] [mb=M] [reps=R]
//
// affinity: before allocating, the thread of GPU i moves to the CPUs of the NUMA node the GPU hangs off
//           (/sys/bus/pci/devices/<bdf>/numa_node) and binds its memory policy there, so the pinned buffers are
//           first-touched on the near node.  Without it every buffer lands wherever the main thread ran.
// streams=K: K upload streams per GPU (each moves 1/K of the bytes).
#include <cuda_runtime.h>
#include <pthread.h>
#include <sched.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

static int g_n = 1, g_streams = 1, g_reps = 6;
static size_t g_bytes = 512ull << 20;
static bool g_affinity = false;
static pthread_barrier_t g_bar;

struct Result { int node = -1; int cpu = -1; double h2d = 0, d2h = 0, ms = 0; char bdf[32] = {0}; int ok = 1; };
static std::vector<Result> g_res;

static std::string read_file(const std::string& p)
{
    FILE* f = fopen(p.c_str(), "r");
    if (!f) return "";
    char buf[4096];
    size_t n = fread(buf, 1, sizeof buf - 1, f);
    fclose(f);
    buf[n] = 0;
    while (n && (buf[n - 1] == '\n' || buf[n - 1] == ' ')) buf[--n] = 0;
    return buf;
}

static std::vector<int> parse_cpulist(const std::string& s)
{
    std::vector<int> out;
    size_t i = 0;
    while (i < s.size()) {
        int a = atoi(s.c_str() + i), b = a;
        while (i < s.size() && s[i] != '-' && s[i] != ',') i++;
        if (i < s.size() && s[i] == '-') { b = atoi(s.c_str() + i + 1); while (i < s.size() && s[i] != ',') i++; }
        for (int c = a; c <= b; c++) out.push_back(c);
        if (i < s.size()) i++;
    }
    return out;
}

static void* worker(void* arg)
{
    const int dev = (int)(intptr_t)arg;
    Result& r = g_res[dev];
    cudaSetDevice(dev);
    char bdf[32] = {0};
    cudaDeviceGetPCIBusId(bdf, sizeof bdf, dev);
    for (char* p = bdf; *p; p++) *p = (char)tolower(*p);
    snprintf(r.bdf, sizeof r.bdf, "%s", bdf);
    const std::string node_s = read_file(std::string("/sys/bus/pci/devices/") + bdf + "/numa_node");
    r.node = node_s.empty() ? -1 : atoi(node_s.c_str());
    if (g_affinity && r.node >= 0) {
        const std::vector<int> cpus = parse_cpulist(read_file("/sys/devices/system/node/node" + std::to_string(r.node) + "/cpulist"));
        if (!cpus.empty()) {
            cpu_set_t set;
            CPU_ZERO(&set);
            for (int c : cpus) CPU_SET(c, &set);
            sched_setaffinity(0, sizeof set, &set);
        }
        unsigned long mask = 1ul << r.node;
        syscall(SYS_set_mempolicy, 2 /* MPOL_BIND */, &mask, sizeof(mask) * 8);
    }
    r.cpu = sched_getcpu();
    void *h_up, *h_dn, *d_up, *d_dn;
    if (cudaHostAlloc(&h_up, g_bytes, cudaHostAllocDefault) != cudaSuccess || cudaHostAlloc(&h_dn, g_bytes / 2, cudaHostAllocDefault) != cudaSuccess ||
        cudaMalloc(&d_up, g_bytes) != cudaSuccess || cudaMalloc(&d_dn, g_bytes / 2) != cudaSuccess) { r.ok = 0; }
    if (r.ok) { memset(h_up, 1, g_bytes); memset(h_dn, 2, g_bytes / 2); }
    std::vector<cudaStream_t> up(g_streams);
    cudaStream_t dn;
    for (auto& s : up) cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&dn, cudaStreamNonBlocking);
    const size_t part = g_bytes / g_streams;
    auto pass = [&](int reps) {
        for (int rep = 0; rep < reps; rep++) {
            for (int s = 0; s < g_streams; s++)
                cudaMemcpyAsync((char*)d_up + s * part, (char*)h_up + s * part, part, cudaMemcpyHostToDevice, up[s]);
            cudaMemcpyAsync(h_dn, d_dn, g_bytes / 2, cudaMemcpyDeviceToHost, dn);
        }
        cudaDeviceSynchronize();
    };
    if (r.ok) pass(1);
    pthread_barrier_wait(&g_bar);
    const auto t0 = std::chrono::steady_clock::now();
    if (r.ok) pass(g_reps);
    r.ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    pthread_barrier_wait(&g_bar);
    r.h2d = (double)g_bytes * g_reps / 1e9 / (r.ms * 1e-3);
    r.d2h = r.h2d / 2;
    if (cudaGetLastError() != cudaSuccess) r.ok = 0;
    return nullptr;
}

int main(int argc, char** argv)
{
    if (argc > 1) g_n = atoi(argv[1]);
    for (int i = 2; i < argc; i++) {
        if (!strcmp(argv[i], "affinity")) g_affinity = true;
        else if (!strncmp(argv[i], "streams=", 8)) g_streams = atoi(argv[i] + 8);
        else if (!strncmp(argv[i], "mb=", 3)) g_bytes = (size_t)atoi(argv[i] + 3) << 20;
        else if (!strncmp(argv[i], "reps=", 5)) g_reps = atoi(argv[i] + 5);
    }
    int have = 0;
    cudaGetDeviceCount(&have);
    if (g_n > have) g_n = have;
    if (g_n < 1) { fprintf(stderr, "no CUDA device\n"); return 1; }
    g_res.resize(g_n);
    pthread_barrier_init(&g_bar, nullptr, g_n + 1);
    std::vector<pthread_t> th(g_n);
    for (int i = 0; i < g_n; i++) pthread_create(&th[i], nullptr, worker, (void*)(intptr_t)i);
    pthread_barrier_wait(&g_bar);
    const auto t0 = std::chrono::steady_clock::now();
    pthread_barrier_wait(&g_bar);
    const double wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    for (auto& t : th) pthread_join(t, nullptr);
    double sum = 0, mn = 1e30;
    int ok = 1;
    for (auto& r : g_res) { sum += r.h2d; if (r.h2d < mn) mn = r.h2d; ok &= r.ok; }
    const double agg = (double)g_bytes * g_reps * g_n / 1e9 / (wall_ms * 1e-3);
    printf("{\"gpus\": %d, \"affinity\": %s, \"upload_streams\": %d, \"mix\": \"2 up : 1 down\", \"mb_up_per_pass\": %zu, \"reps\": %d, \"ok\": %s, "
           "\"aggregate_h2d_gbs\": %.2f, \"aggregate_d2h_gbs\": %.2f, \"per_gpu_h2d_min_gbs\": %.2f, \"per_gpu_h2d_mean_gbs\": %.2f, \"numa_nodes_online\": \"%s\", \"host_cpus\": %ld, \"per_gpu\": [",
           g_n, g_affinity ? "true" : "false", g_streams, g_bytes >> 20, g_reps, ok ? "true" : "false", agg, agg / 2, mn, sum / g_n,
           read_file("/sys/devices/system/node/online").c_str(), sysconf(_SC_NPROCESSORS_ONLN));
    for (int i = 0; i < g_n; i++)
        printf("%s{\"gpu\": %d, \"bdf\": \"%s\", \"numa_node\": %d, \"thread_cpu\": %d, \"h2d_gbs\": %.2f}", i ? ", " : "", i, g_res[i].bdf, g_res[i].node, g_res[i].cpu, g_res[i].h2d);
    printf("]}\n");
    return ok ? 0 : 1;
}
