// ffma2_probe.cu -- issue rate of FFMA vs FFMA2 (packed fp32 pairs) per SM sub-partition on sm_100a (developer probe).
// Each thread runs CHAINS independent accumulator chains of ITER dependent fmas; one CTA per SM, W warps per CTA.
// Prints cycles per warp-instruction per scheduler: the fma pipe's reciprocal throughput once enough warps hide latency.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/ffma2_probe tools/probes/ffma2_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int CHAINS = 16, ITER = 4096;

template <int MODE>   // 0: scalar FFMA, 1: FFMA2 with packed operands, 2: FFMA2 with one operand a duplicated scalar
__global__ void probe(float *out, long long *cycles, float a, float b) {
    float2 acc[CHAINS];
    for (int c = 0; c < CHAINS; c++) acc[c] = make_float2(threadIdx.x + c, threadIdx.x - c);
    const float2 m = make_float2(a, MODE == 2 ? a : b), d = make_float2(b, a);
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < ITER; i++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) {
            if (MODE == 0) { acc[c].x = fmaf(acc[c].x, m.x, d.x); acc[c].y = fmaf(acc[c].y, m.y, d.y); }
            else acc[c] = __ffma2_rn(acc[c], m, d);
        }
    }
    const long long t1 = clock64();
    float s = 0.0f;
    for (int c = 0; c < CHAINS; c++) s += acc[c].x + acc[c].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, int warps) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float *out; long long *cyc;
    cudaMalloc(&out, sizeof(float) * sms * warps * 32);
    cudaMalloc(&cyc, sizeof(long long) * sms);
    probe<MODE><<<sms, warps * 32>>>(out, cyc, 0.999f, 0.001f);
    probe<MODE><<<sms, warps * 32>>>(out, cyc, 0.999f, 0.001f);
    cudaDeviceSynchronize();
    long long h[1024];
    cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < sms; i++) mean += (double)h[i];
    mean /= sms;
    const double warp_insts_per_smsp = (double)ITER * CHAINS * (MODE == 0 ? 2 : 1) * warps / 4.0;
    printf("%-28s warps/SM %2d: %.2f cycles per warp-instruction per scheduler, %.2f fma lanes-ops per cycle per SM\n", name, warps,
           mean / warp_insts_per_smsp, (double)ITER * CHAINS * 2 * warps * 32 / mean);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int w : {4, 8, 16, 32}) {
        run<0>("FFMA", w);
        run<1>("FFMA2 packed", w);
        run<2>("FFMA2 duplicated multiplier", w);
    }
    return 0;
}
