// mma_probe.cu -- what the warp-level tensor-core path (mma.sync, the only MMA a 4-warp CTA with battles in registers can
// issue without restructuring into tcgen05 + TMEM) delivers per SM on sm_100a, against FFMA / FFMA2 (ffma2_probe.cu).
// Question it answers: can the 64 x 64 layer of the rollout policy go to tensor cores at fp32-grade
// accuracy?  torch's fp32 result to 2e-5 in log-probability needs 3 x TF32 (a_hi*b_hi + a_lo*b_hi + a_hi*b_lo) or a
// 3-way bf16 split (6 products): the MMA path has to deliver >= 3x (TF32) / >= 6x (bf16) the FFMA rate to tie.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/mma_probe tools/probes/mma_probe.cu
// Each warp runs CHAINS independent accumulator tiles of ITER dependent MMAs; one CTA per SM, W warps per CTA.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int CHAINS = 8, ITER = 2048;

template <int MODE>   // 0: m16n8k8 tf32, 1: m16n8k16 bf16
__global__ void probe(float *out, long long *cycles, uint32_t seed) {
    float acc[CHAINS][4];
    for (int c = 0; c < CHAINS; c++) for (int k = 0; k < 4; k++) acc[c][k] = (float)(threadIdx.x + c + k);
    uint32_t a[4] = { seed, seed ^ 0x3c003c00u, seed + 1u, seed ^ 0x1234u }, b[2] = { seed ^ 0x3f800000u, seed + 7u };
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < ITER; i++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) {
            if (MODE == 0)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(acc[c][0]), "+f"(acc[c][1]), "+f"(acc[c][2]), "+f"(acc[c][3])
                             : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
            else
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(acc[c][0]), "+f"(acc[c][1]), "+f"(acc[c][2]), "+f"(acc[c][3])
                             : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
        }
    }
    const long long t1 = clock64();
    float s = 0.0f;
    for (int c = 0; c < CHAINS; c++) for (int k = 0; k < 4; k++) s += acc[c][k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, int warps, double macs_per_mma, double split_products) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float *out; long long *cyc;
    cudaMalloc(&out, sizeof(float) * sms * warps * 32);
    cudaMalloc(&cyc, sizeof(long long) * sms);
    probe<MODE><<<sms, warps * 32>>>(out, cyc, 0x3f000000u);
    probe<MODE><<<sms, warps * 32>>>(out, cyc, 0x3f000000u);
    cudaDeviceSynchronize();
    long long h[1024];
    cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < sms; i++) mean += (double)h[i];
    mean /= sms;
    const double macs = (double)ITER * CHAINS * warps * macs_per_mma;
    printf("%-22s warps/SM %2d: %7.1f MAC per cycle per SM  -> %6.1f fp32-grade MAC per cycle per SM after the %.0f-product split "
           "(FFMA / FFMA2: ~127, profiles/r01h_ffma2_probe.log)\n", name, warps, macs / mean, macs / mean / split_products, split_products);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int w : {4, 8, 16, 32}) {
        run<0>("mma.sync m16n8k8 tf32", w, 16.0 * 8 * 8, 3.0);
        run<1>("mma.sync m16n8k16 bf16", w, 16.0 * 8 * 16, 6.0);
    }
    return 0;
}
