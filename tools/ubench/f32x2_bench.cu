// f32x2_bench.cu -- microbenchmark: issue rate of scalar FADD/FMUL versus the packed FADD2/FMUL2 (add.rn.f32x2 /
// mul.rn.f32x2, sm_100a) and whether packed FP32 co-issues with integer ALU work; also checks that each half of a
// packed op is the same IEEE round-to-nearest result as the scalar op.   nvcc -arch=sm_100a -fmad=false
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 c; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(c) : "l"(a), "l"(b)); return c; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 c; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(c) : "l"(a), "l"(b)); return c; }
__device__ __forceinline__ float add1(float a, float b) { float c; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(c) : "f"(a), "f"(b)); return c; }
__device__ __forceinline__ float mul1(float a, float b) { float c; asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(c) : "f"(a), "f"(b)); return c; }
__device__ __forceinline__ unsigned lop(unsigned a, unsigned b) { unsigned c; asm volatile("xor.b32 %0, %1, %2;" : "=r"(c) : "r"(a), "r"(b)); return c; }
#define CH 8
template <int MODE> __global__ void __launch_bounds__(256) bench(float *out, int iters, float seed) {
    float s[CH]; u64 p[CH]; unsigned q[CH];
    for (int i = 0; i < CH; i++) { s[i] = seed + i + threadIdx.x; float2 t = make_float2(seed + i, seed - i); p[i] = *(u64 *)&t; q[i] = i * 77u + threadIdx.x; }
    float k = seed * 1e-3f + threadIdx.x * 1e-9f; float2 kk2 = make_float2(k, k * 1.5f); u64 k2 = *(u64 *)&kk2; 
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < CH; i++) {
                if (MODE == 0) s[i] = add1(s[i], k);
                if (MODE == 1) p[i] = add2(p[i], k2);
                if (MODE == 2) s[i] = mul1(s[i], k);
                if (MODE == 3) p[i] = mul2(p[i], k2);
                if (MODE == 4) { p[i] = add2(p[i], k2); q[i] = lop(q[i], q[(i + 1) & 7]); }              // packed add + ALU
                if (MODE == 5) { s[i] = add1(s[i], k); q[i] = lop(q[i], q[(i + 1) & 7]); }               // scalar add + ALU
                if (MODE == 6) { p[i] = add2(p[i], k2); q[i] = lop(q[i], q[(i + 1) & 7]); q[i] = lop(q[i], q[(i + 3) & 7]); }  // packed + 2 ALU
                if (MODE == 7) { s[i] = add1(s[i], k); s[i] = mul1(s[i], k); }                // dependent scalar add,mul
                if (MODE == 9) { asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(k2)); }
                if (MODE == 10) { asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(s[i]) : "f"(k)); }
                if (MODE == 8) { p[i] = add2(p[i], k2); p[i] = mul2(p[i], k2); }
            }
        }
    }
    float acc = 0; for (int i = 0; i < CH; i++) { float2 t = *(float2 *)&p[i]; acc += s[i] + t.x + t.y + (float)q[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
__global__ void exact(const float *a, const float *b, unsigned *bad, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x; if (2 * i + 1 >= n) return;
    float2 x = make_float2(a[2 * i], a[2 * i + 1]), y = make_float2(b[2 * i], b[2 * i + 1]);
    u64 s = add2(*(u64 *)&x, *(u64 *)&y), m = mul2(*(u64 *)&x, *(u64 *)&y);
    float2 sf = *(float2 *)&s, mf = *(float2 *)&m;
    auto same = [](float p, float q) { return (p != p && q != q) || __float_as_uint(p) == __float_as_uint(q); };
    if (!same(sf.x, __fadd_rn(x.x, y.x)) || !same(sf.y, __fadd_rn(x.y, y.y)) || !same(mf.x, __fmul_rn(x.x, y.x)) || !same(mf.y, __fmul_rn(x.y, y.y))) atomicAdd(bad, 1u);
    if (false && __float_as_uint(sf.x) != __float_as_uint(__fadd_rn(x.x, y.x)) || __float_as_uint(sf.y) != __float_as_uint(__fadd_rn(x.y, y.y)) ||
        __float_as_uint(mf.x) != __float_as_uint(__fmul_rn(x.x, y.x)) || __float_as_uint(mf.y) != __float_as_uint(__fmul_rn(x.y, y.y))) atomicAdd(bad, 1u);
}
template <int MODE> void run(const char *name, int per_iter_instr, float *d) {
    int dev_sms; cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, 0);
    const int iters = 4000; cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 3; w++) bench<MODE><<<dev_sms * 4, 256>>>(d, iters, 1.5f);
    cudaDeviceSynchronize();
    float ms = 1e9f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0); bench<MODE><<<dev_sms * 4, 256>>>(d, iters, 1.5f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float t; cudaEventElapsedTime(&t, e0, e1); if (t < ms) ms = t;
    }
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double warp_instr = (double)dev_sms * 4 * 8 * iters * 4.0 * CH * per_iter_instr;
    double per_clk_sm = warp_instr / (ms * 1e-3 * clk * 1e3) / dev_sms;
    printf("%-34s %8.3f ms  %6.3f warp-instr/clk/SM (at %d MHz nominal)\n", name, ms, per_clk_sm, clk / 1000);
}
int main() {
    float *d; cudaMalloc(&d, 148 * 8 * 256 * 4 * 4);
    for (int w = 0; w < 200; w++) bench<0><<<148 * 4, 256>>>(d, 4000, 1.5f);   // clock ramp-up
    cudaDeviceSynchronize();
    run<0>("scalar FADD", 1, d); run<1>("packed FADD2", 1, d); run<2>("scalar FMUL", 1, d); run<3>("packed FMUL2", 1, d);
    run<4>("FADD2 + LOP3", 2, d); run<5>("FADD + LOP3", 2, d); run<6>("FADD2 + 2 LOP3", 3, d); run<7>("FADD,FMUL dependent", 2, d); run<8>("FADD2,FMUL2 dependent", 2, d); run<9>("packed FFMA2", 1, d); run<10>("scalar FFMA", 1, d);
    const int n = 1 << 24; float *ha = (float *)malloc(n * 4), *hb = (float *)malloc(n * 4);
    srand(7); for (int i = 0; i < n; i++) { uint32_t u = ((uint32_t)rand() << 16) ^ rand(), v = ((uint32_t)rand() << 16) ^ rand();
        if (i % 3 == 0) { u = (u & 0x807fffffu) | ((100u + (u >> 23) % 60u) << 23); v = (v & 0x807fffffu) | ((100u + (v >> 23) % 60u) << 23); }   // similar magnitudes: cancellation, subnormal products
        ha[i] = *(float *)&u; hb[i] = *(float *)&v; }
    float *da, *db; unsigned *bad, hbad = 0; cudaMalloc(&da, n * 4); cudaMalloc(&db, n * 4); cudaMalloc(&bad, 4);
    cudaMemcpy(da, ha, n * 4, cudaMemcpyHostToDevice); cudaMemcpy(db, hb, n * 4, cudaMemcpyHostToDevice); cudaMemset(bad, 0, 4);
    exact<<<n / 2 / 256, 256>>>(da, db, bad, n); cudaMemcpy(&hbad, bad, 4, cudaMemcpyDeviceToHost);
    printf("packed vs scalar rn results: %u mismatching pairs of %d (random bit patterns incl. subnormal/inf/nan)\n", hbad, n / 2);
    return 0;
}
