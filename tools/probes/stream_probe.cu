// Developer probe: what does the step kernel's HBM traffic shape cost with NO compute?
// Reads 4 state planes (16 B/env each) + 1 action byte, writes 4 planes + obs 32 B + reward 4 + term 1 + info 8.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s: %s\n",#x,cudaGetErrorString(e)); return 1;}}while(0)
struct P { uint4 *pl[4]; const uint8_t* act; float4* obs; float* rew; uint8_t* term; int* frame; uchar4* misc; int n; };
template<int MODE> __global__ void __launch_bounds__(256) k_stream(P p) {
    for (int i = blockIdx.x*256 + threadIdx.x; i < p.n; i += gridDim.x*256) {
        uint4 a=p.pl[0][i], b=p.pl[1][i], c=p.pl[2][i], d=p.pl[3][i]; uint32_t act=p.act[i];
        a.x += act; b.y ^= a.z; c.z += d.w; d.x += c.y;   // trivial dependency so nothing is optimised away
        p.pl[0][i]=a; p.pl[1][i]=b; p.pl[2][i]=c; p.pl[3][i]=d;
        if (MODE >= 1) {
            p.obs[2*(size_t)i] = make_float4(a.x,b.x,c.x,d.x); p.obs[2*(size_t)i+1] = make_float4(a.y,b.y,c.y,d.y);
            p.rew[i] = __uint_as_float(a.w); p.term[i] = (uint8_t)b.w; p.frame[i] = c.w; p.misc[i] = make_uchar4(a.x,a.y,a.z,a.w);
        }
    }
}
int main(int argc, char** argv) {
    int n = 4*1024*1024; P p; p.n = n;
    for (int k=0;k<4;k++) { CK(cudaMalloc(&p.pl[k], (size_t)n*16)); CK(cudaMemset(p.pl[k], 1, (size_t)n*16)); }
    uint8_t* act; CK(cudaMalloc(&act, n)); CK(cudaMemset(act, 3, n)); p.act = act;
    CK(cudaMalloc(&p.obs, (size_t)n*32)); CK(cudaMalloc(&p.rew, (size_t)n*4)); CK(cudaMalloc(&p.term, n)); CK(cudaMalloc(&p.frame, (size_t)n*4)); CK(cudaMalloc(&p.misc, (size_t)n*4));
    cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode=0; mode<2; mode++) for (int bps : {2,4,8,16}) {
        int grid = 148*bps;
        for (int w=0; w<5; w++) { if (mode==0) k_stream<0><<<grid,256>>>(p); else k_stream<1><<<grid,256>>>(p); }
        cudaEventRecord(e0);
        const int reps=50;
        for (int r=0;r<reps;r++) { if (mode==0) k_stream<0><<<grid,256>>>(p); else k_stream<1><<<grid,256>>>(p); }
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms,e0,e1);
        double us = ms*1e3/reps; double bytes = (double)n*(mode==0 ? 129.0 : 174.0);
        printf("mode %d (%s) blocks/SM %2d: %.2f us  %.1f GB/s\n", mode, mode? "planes+outputs 174 B/env":"planes only 129 B/env", bps, us, bytes/us/1e3);
    }
    return 0;
}
