// microbench_lat.cu -- dependent-issue latencies of the instructions the lane-cooperative tracker chains together,
// for one warp alone on its scheduler (sm_100a).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_lat microbench_lat.cu
#include <cstdio>
#include <cuda_runtime.h>

// OP 0: FADD chain; 1: SHFL.IDX chain; 2: MUFU.RCP chain; 3: FADD+SHFL alternating; 4: LDS chain (pointer chase);
// 5: two warps, one bar.sync per iteration; 6: STS -> bar.sync -> LDS ping-pong between two warps;
// 7: FADD chain, 2 independent chains (issue rate of a lone warp); 8: 4 independent chains; 9: SEL->LOP3->FADD chain
template <int OP>
__global__ void k(float *out, long long *cyc, float seed, int iters) {
    __shared__ int s_next[64];
    __shared__ float s_val[2][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    s_next[threadIdx.x & 63] = (threadIdx.x + 1) & 31;
    __syncthreads();
    float f = seed + lane, g = seed * 2 + lane, h2 = seed * 3, h3 = seed * 4;
    int p = lane;
    const int src = (lane + 1) & 31;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            if (OP == 0) f = __fadd_rn(f, seed);
            if (OP == 1) f = __shfl_sync(0xffffffffu, f, src);
            if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(f));
            if (OP == 3) f = __fadd_rn(__shfl_sync(0xffffffffu, f, src), seed);
            if (OP == 4) p = ((volatile int *) s_next)[p];
            if (OP == 5) asm volatile("bar.sync 1, 64;" ::: "memory");
            if (OP == 6) {
                if (warp == (u & 1)) s_val[u & 1][lane] = f;
                asm volatile("bar.sync 1, 64;" ::: "memory");
                if (warp != (u & 1)) f = __fadd_rn(((volatile float *) s_val[u & 1])[src], seed);
            }
            if (OP == 7) { f = __fadd_rn(f, seed); g = __fadd_rn(g, seed); }
            if (OP == 8) { f = __fadd_rn(f, seed); g = __fadd_rn(g, seed); h2 = __fadd_rn(h2, seed); h3 = __fadd_rn(h3, seed); }
            if (OP == 9) {
                unsigned b = __float_as_uint(f);
                asm volatile("lop3.b32 %0, %0, %1, %2, 0xf8;" : "+r"(b) : "r"(0xffffffffu), "r"(p & 0));
                f = __fadd_rn(__uint_as_float(b), seed);
            }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[threadIdx.x] = f + g + h2 + h3 + p;
}

template <int OP>
static void run(const char *name, int threads, int per_iter) {
    float *out; long long *cyc, h;
    cudaMalloc(&out, 256 * 4); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    k<OP><<<1, threads>>>(out, cyc, 1.0009f, iters);
    k<OP><<<1, threads>>>(out, cyc, 1.0009f, iters);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-52s %7.2f cycles per op\n", name, (double) h / iters / 16 / per_iter);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("FADD dependent chain", 32, 1);
    run<7>("FADD, 2 independent chains (per FADD)", 32, 2);
    run<8>("FADD, 4 independent chains (per FADD)", 32, 4);
    run<1>("SHFL.IDX dependent chain", 32, 1);
    run<3>("SHFL.IDX + FADD dependent pair", 32, 1);
    run<2>("MUFU.RCP dependent chain", 32, 1);
    run<4>("LDS dependent chain", 32, 1);
    run<9>("LOP3 + FADD dependent pair", 32, 1);
    run<5>("bar.sync, 2 warps", 64, 1);
    run<6>("STS -> bar.sync -> LDS -> FADD hand-over, 2 warps", 64, 1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
