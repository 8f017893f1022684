// f32x2_lat.cu -- dependent-issue latency of scalar and packed FP32 instructions on sm_100a (one warp, one chain).
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
template <int MODE> __global__ void lat(float *out, long long *cyc, int iters, const u64 *c) {
    u64 p = c[3] + threadIdx.x, k2 = c[2] + threadIdx.x; float s = 1.f + threadIdx.x, k = __uint_as_float((unsigned)c[2]);
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 32; r++) {
            if (MODE == 0) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(s) : "f"(k));
            if (MODE == 1) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p) : "l"(k2));
            if (MODE == 2) asm volatile("mul.rn.ftz.f32x2 %0, %0, %1;" : "+l"(p) : "l"(k2));
            if (MODE == 3) { asm volatile("mul.rn.ftz.f32x2 %0, %0, %1;" : "+l"(p) : "l"(k2)); asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p) : "l"(k2)); }
            if (MODE == 4) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(s) : "f"(k));
            if (MODE == 5) { float2 t = *(float2 *)&p; asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(t.x) : "f"(t.y)); p = *(u64 *)&t; asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p) : "l"(k2)); }   // packed -> scalar half -> packed
        }
    }
    long long t1 = clock64();
    float2 t = *(float2 *)&p; out[threadIdx.x] = s + t.x + t.y;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int MODE> void run(const char *name, int per, float *d, long long *dc, const u64 *c) {
    lat<MODE><<<1, 32>>>(d, dc, 100, c); cudaDeviceSynchronize();
    lat<MODE><<<1, 32>>>(d, dc, 2000, c); long long h; cudaMemcpy(&h, dc, 8, cudaMemcpyDeviceToHost);
    printf("%-40s %6.2f cycles per dependent instruction\n", name, (double)h / (2000.0 * 32 * per));
}
int main() {
    float *d; long long *dc; u64 *c; cudaMalloc(&d, 4096); cudaMalloc(&dc, 8); cudaMalloc(&c, 32);
    u64 hc[4]; float2 t; t = make_float2(-0.f, -0.f); hc[0] = *(u64 *)&t; t = make_float2(1.f, 1.f); hc[1] = *(u64 *)&t;
    t = make_float2(1.0000001f, 0.9999999f); hc[2] = *(u64 *)&t; t = make_float2(1.5f, 2.5f); hc[3] = *(u64 *)&t;
    cudaMemcpy(c, hc, 32, cudaMemcpyHostToDevice);
    run<0>("FADD", 1, d, dc, c); run<4>("FMUL", 1, d, dc, c); run<1>("FADD2", 1, d, dc, c); run<2>("FMUL2.FTZ", 1, d, dc, c);
    run<3>("FMUL2.FTZ -> FADD2", 2, d, dc, c); run<5>("FADD2 -> FADD (half) -> FADD2", 2, d, dc, c);
    return 0;
}
