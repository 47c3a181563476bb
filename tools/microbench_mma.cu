// microbench_mma.cu -- issue-rate probe for warp-level mma.sync on B200 (sm_100a), alone and mixed with the
// scalar FP32 adds of the exact kernels: can a tensor-core "proposer" run in the shadow of the FP32 pipe?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_mma microbench_mma.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_bf16(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// OP 0: bf16 mma only (8 independent accumulators); 1: tf32 mma only; 2: 8 FADD per bf16 mma; 3: 16 FADD per mma;
// 4: FADD only (reference); 5: 32 FADD per mma; 6: SHFL only; 7: 8 FADD + 1 SHFL
template <int OP>
__global__ void k(float *out, long long *cyc, float seed, int iters) {
    float c[8][4], f[8];
    unsigned a[4], b[2];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        f[i] = seed + i + threadIdx.x;
#pragma unroll
        for (int j = 0; j < 4; j++) c[i][j] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < 4; j++) a[j] = 0x3F803F80u ^ ((threadIdx.x >> j) << 15);
    b[0] = 0x3F80BF80u; b[1] = 0xBF803F80u;
    const float inc = seed * 0.999f;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (OP == 0 || OP == 2 || OP == 3 || OP == 5) mma_bf16(c[u], a, b);
            if (OP == 1) mma_tf32(c[u], a, b);
            if (OP == 2 || OP == 4 || OP == 7) {
#pragma unroll
                for (int i = 0; i < 8; i++) f[i] = __fadd_rn(f[i], inc);
            }
            if (OP == 3) {
#pragma unroll
                for (int r = 0; r < 2; r++)
#pragma unroll
                    for (int i = 0; i < 8; i++) f[i] = __fadd_rn(f[i], inc);
            }
            if (OP == 5) {
#pragma unroll
                for (int r = 0; r < 4; r++)
#pragma unroll
                    for (int i = 0; i < 8; i++) f[i] = __fadd_rn(f[i], inc);
            }
            if (OP == 6 || OP == 7) f[u] = __shfl_xor_sync(0xffffffffu, f[u], 1);
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += f[i] + c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char *name, int thr, double mma_per_u, double fadd_per_u, double shfl_per_u) {
    int sms = 148, iters = 2000;
    float *out; long long *cyc;
    cudaMalloc(&out, sms * thr * 4); cudaMalloc(&cyc, sms * 8);
    k<OP><<<sms, thr>>>(out, cyc, 1.0001f, 10); cudaDeviceSynchronize();
    k<OP><<<sms, thr>>>(out, cyc, 1.0001f, iters); cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sms * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < sms; i++) avg += h[i]; avg /= sms;
    double units = (double) (thr / 32) * iters * 8.0;    // per SM
    printf("%-34s %2d warps/SM: %7.3f mma/clk/SM  %7.3f fadd/clk/SM  %6.3f shfl/clk/SM  (%.0f cyc)\n", name, thr / 32,
           units * mma_per_u / avg, units * fadd_per_u / avg, units * shfl_per_u / avg, avg);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int thr : {128, 256, 512, 1024}) {
        run<0>("mma m16n8k16 bf16", thr, 1, 0, 0);
        run<1>("mma m16n8k8 tf32", thr, 1, 0, 0);
    }
    run<4>("FADD only", 1024, 0, 8, 0);
    run<2>("1 mma + 8 FADD", 1024, 1, 8, 0);
    run<3>("1 mma + 16 FADD", 1024, 1, 16, 0);
    run<5>("1 mma + 32 FADD", 1024, 1, 32, 0);
    run<6>("SHFL only", 1024, 0, 0, 1);
    run<7>("8 FADD + 1 SHFL", 1024, 0, 8, 1);
    return 0;
}
