// microbench_lat2.cu -- issue intervals of independent SHFL / LDS for one warp alone, and the in-warp
// STS -> __syncwarp -> LDS hand-over (sm_100a).
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 lds64(const float *p) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(__cvta_generic_to_shared(p)));
    return v;
}
__device__ __forceinline__ void sts64(float *p, float a, float b) {
    asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"l"(__cvta_generic_to_shared(p)), "f"(a), "f"(b) : "memory");
}

template <int OP, int NW>
__global__ void k(float *out, long long *cyc, float seed, int iters) {
    __shared__ __align__(16) float s[NW][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = lane; i < 256; i += 32) s[warp][i] = seed * i;
    __syncthreads();
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; i++) f[i] = seed + lane + i;
    const int src = (lane + 1) & 31;
    float *my = s[warp];
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        if (OP == 0) {          // 8 independent SHFL, then 8 FADD consuming them
            float g[8];
#pragma unroll
            for (int i = 0; i < 8; i++) g[i] = __shfl_sync(0xffffffffu, f[i], src);
#pragma unroll
            for (int i = 0; i < 8; i++) f[i] = __fadd_rn(g[i], seed);
        }
        if (OP == 1) {          // 8 independent LDS.64 (broadcast-free, conflict-free), then FADDs
            float2 g[8];
#pragma unroll
            for (int i = 0; i < 8; i++) g[i] = lds64(my + 2 * ((lane + i) & 31) + 64 * (i & 1) + (__float_as_int(f[i]) & 0));
#pragma unroll
            for (int i = 0; i < 8; i++) f[i] = __fadd_rn(g[i].x, g[i].y);
        }
        if (OP == 2) {          // 4 independent LDS.128 (same bytes as OP 1), then FADDs
            float4 g[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const float4 *p = reinterpret_cast<const float4 *>(my + 4 * ((lane + i) & 31) + (__float_as_int(f[i]) & 0));
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(g[i].x), "=f"(g[i].y), "=f"(g[i].z), "=f"(g[i].w) : "l"(__cvta_generic_to_shared(p)));
            }
#pragma unroll
            for (int i = 0; i < 4; i++) { f[2 * i] = __fadd_rn(g[i].x, g[i].y); f[2 * i + 1] = __fadd_rn(g[i].z, g[i].w); }
        }
        if (OP == 3) {          // STS.64 -> __syncwarp -> LDS.64 (neighbour's value) -> FADD: dependent hand-over
            sts64(my + 2 * lane, f[0], f[1]);
            __syncwarp();
            float2 g = lds64(my + 2 * src);
            __syncwarp();
            f[0] = __fadd_rn(g.x, seed); f[1] = __fadd_rn(g.y, seed);
        }
        if (OP == 4) {          // same hand-over with two SHFLs
            float a = __shfl_sync(0xffffffffu, f[0], src), b = __shfl_sync(0xffffffffu, f[1], src);
            f[0] = __fadd_rn(a, seed); f[1] = __fadd_rn(b, seed);
        }
        if (OP == 5) {          // 16 independent SHFL
            float g[8], h[8];
#pragma unroll
            for (int i = 0; i < 8; i++) { g[i] = __shfl_sync(0xffffffffu, f[i], src); h[i] = __shfl_sync(0xffffffffu, f[i], src ^ 3); }
#pragma unroll
            for (int i = 0; i < 8; i++) f[i] = __fadd_rn(g[i], h[i]);
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    float r = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) r += f[i];
    out[threadIdx.x] = r;
}

template <int OP, int NW>
static void run(const char *name) {
    float *out; long long *cyc, h;
    cudaMalloc(&out, 1024 * 4); cudaMalloc(&cyc, 8);
    const int iters = 4000;
    k<OP, NW><<<1, 32 * NW>>>(out, cyc, 1.0009f, iters);
    k<OP, NW><<<1, 32 * NW>>>(out, cyc, 1.0009f, iters);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-64s %d warp(s): %7.1f cycles per iteration\n", name, NW, (double) h / iters);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0, 1>("8 independent SHFL + 8 FADD");
    run<5, 1>("16 independent SHFL + 8 FADD");
    run<0, 4>("8 independent SHFL + 8 FADD");
    run<0, 8>("8 independent SHFL + 8 FADD");
    run<1, 1>("8 independent LDS.64 + 8 FADD");
    run<1, 4>("8 independent LDS.64 + 8 FADD");
    run<2, 1>("4 independent LDS.128 + 8 FADD");
    run<3, 1>("STS.64 -> syncwarp -> LDS.64 -> FADD (dependent)");
    run<4, 1>("2 SHFL -> FADD (dependent)");
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
