// d2h_probe.cu -- what the BOX can do: pinned device->host (and host->device) copy bandwidth of one GPU per process, N
// processes at once, so that the aggregate ceiling of the host side (PCIe root complexes, memory controllers, IOMMU /
// virtualisation) is measured rather than assumed.  Companion of bench.py's `e2e` number.
//
//   d2h_probe <device> <seconds> <block MiB> <pieces> <numa: 0|1> <dir: d2h|h2d|both> [start_epoch_seconds]
//
// Each iteration moves one block of <block MiB> as <pieces> equal cudaMemcpyAsync calls on one stream (pieces = 6 is what
// fg_step_host_compact issued per slice in round 1, pieces = 1 is one staged block).  numa = 1 binds the calling thread
// to the CPUs of the GPU's NUMA node (/sys/bus/pci/devices/<bdf>/numa_node) BEFORE cudaHostAlloc, so that first-touch
// places the pinned pages on that node.  All processes spin until start_epoch_seconds so that they overlap.
// Prints one JSON line: {"device":..,"numa_node":..,"bound":..,"dir":..,"pieces":..,"block_mib":..,"gbps":..,"iters":..}
#include <cuda_runtime.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <unistd.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } } while (0)

static double now() { struct timeval t; gettimeofday(&t, 0); return t.tv_sec + 1e-6 * t.tv_usec; }

static int gpu_numa_node(int dev) {
    char bdf[64] = "", path[160], buf[32] = "";
    if (cudaDeviceGetPCIBusId(bdf, sizeof bdf, dev) != cudaSuccess) return -1;
    for (char *p = bdf; *p; p++) if (*p >= 'A' && *p <= 'Z') *p += 32;
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bdf);
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    if (!fgets(buf, sizeof buf, f)) buf[0] = 0;
    fclose(f);
    return atoi(buf);
}

// bind the calling thread to the CPUs listed in /sys/devices/system/node/node<N>/cpulist ("0-15,32-47")
static int bind_to_node(int node) {
    char path[96], buf[512] = "";
    snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    if (!fgets(buf, sizeof buf, f)) buf[0] = 0;
    fclose(f);
    cpu_set_t set;
    CPU_ZERO(&set);
    int n = 0;
    for (char *tok = strtok(buf, ",\n"); tok; tok = strtok(0, ",\n")) {
        int a, b;
        if (sscanf(tok, "%d-%d", &a, &b) == 2) { for (int c = a; c <= b; c++) { CPU_SET(c, &set); n++; } }
        else if (sscanf(tok, "%d", &a) == 1) { CPU_SET(a, &set); n++; }
    }
    if (!n) return -1;
    return sched_setaffinity(0, sizeof set, &set);
}

int main(int argc, char **argv) {
    if (argc < 7) { fprintf(stderr, "usage: %s device seconds block_mib pieces numa dir [start]\n", argv[0]); return 1; }
    const int dev = atoi(argv[1]);
    const double seconds = atof(argv[2]);
    const size_t block = (size_t)atoi(argv[3]) << 20;
    const int pieces = atoi(argv[4]), numa = atoi(argv[5]);
    const char *dir = argv[6];
    const double start = argc > 7 ? atof(argv[7]) : 0.0;
    CK(cudaSetDevice(dev));
    const int node = gpu_numa_node(dev);
    int bound = 0;
    if (numa && node >= 0) bound = bind_to_node(node) == 0;
    char *h = 0, *d = 0;
    CK(cudaHostAlloc((void **)&h, block, cudaHostAllocDefault));
    memset(h, 1, block);                                    // first touch from the (bound) thread
    CK(cudaMalloc((void **)&d, block));
    CK(cudaMemset(d, 2, block));
    cudaStream_t s, s2;
    CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    const bool d2h = !strcmp(dir, "d2h") || !strcmp(dir, "both"), h2d = !strcmp(dir, "h2d") || !strcmp(dir, "both");
    const size_t piece = block / (size_t)pieces;
    for (int w = 0; w < 3; w++) {                            // warm-up
        if (d2h) CK(cudaMemcpyAsync(h, d, block, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
    }
    while (start > 0.0 && now() < start) usleep(200);
    const double t0 = now();
    long iters = 0;
    while (now() - t0 < seconds) {
        for (int k = 0; k < pieces; k++) {
            if (d2h) CK(cudaMemcpyAsync(h + k * piece, d + k * piece, piece, cudaMemcpyDeviceToHost, s));
            if (h2d) CK(cudaMemcpyAsync(d + k * piece, h + k * piece, piece, cudaMemcpyHostToDevice, !strcmp(dir, "both") ? s2 : s));
        }
        CK(cudaStreamSynchronize(s));
        CK(cudaStreamSynchronize(s2));
        iters++;
    }
    const double dt = now() - t0;
    printf("{\"device\": %d, \"numa_node\": %d, \"bound\": %d, \"dir\": \"%s\", \"pieces\": %d, \"block_mib\": %zu, "
           "\"gbps\": %.2f, \"iters\": %ld, \"seconds\": %.3f}\n",
           dev, node, bound, dir, pieces, block >> 20, (double)iters * (double)piece * pieces * ((d2h && h2d) ? 2 : 1) / dt / 1e9, iters, dt);
    return 0;
}
