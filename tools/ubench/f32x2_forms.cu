// f32x2_forms.cu -- does the scalar-broadcast operand form (FMUL2 R, R.F32x2, R.F32) keep the FMUL2 || FADD2 overlap?
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define CH 8
#define MULB(d, a, s) asm volatile("{\n\t.reg .b64 t;\n\tmov.b64 t, {%2, %2};\n\tmul.rn.ftz.f32x2 %0, %1, t;\n\t}" : "=l"(d) : "l"(a), "f"(s))
#define ADDB(d, a, s) asm volatile("{\n\t.reg .b64 t;\n\tmov.b64 t, {%2, %2};\n\tadd.rn.f32x2 %0, %1, t;\n\t}" : "=l"(d) : "l"(a), "f"(s))
#define RSUBB(d, a, s) asm volatile("{\n\t.reg .b64 t;\n\tmov.b64 t, {%2, %2};\n\tsub.rn.f32x2 %0, t, %1;\n\t}" : "=l"(d) : "l"(a), "f"(s))
#define MUL(d, a, b) asm volatile("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b))
#define ADD(d, a, b) asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b))
#define SUB(d, a, b) asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b))
template <int MODE> __global__ void __launch_bounds__(256) bench(float *out, int iters, const u64 *c) {
    u64 p[CH], q[CH]; const u64 k2 = c[2] + threadIdx.x; const float k = __uint_as_float((unsigned)c[2]) + 1e-9f * threadIdx.x;
    for (int i = 0; i < CH; i++) { p[i] = c[3] + i * 8 + threadIdx.x; q[i] = c[3] + i * 16 + threadIdx.x; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < CH; i++) {
                if (MODE == 0) { MUL(p[i], p[i], k2); ADD(q[i], q[i], k2); }            // independent chains, pair operands
                if (MODE == 1) { MULB(p[i], p[i], k); ADD(q[i], q[i], k2); }            // FMUL2 broadcast form
                if (MODE == 2) { MUL(p[i], p[i], k2); ADDB(q[i], q[i], k); }            // FADD2 broadcast form
                if (MODE == 3) { MULB(p[i], p[i], k); ADDB(q[i], q[i], k); }            // both
                if (MODE == 4) { MUL(p[i], p[i], k2); RSUBB(q[i], q[i], k); }           // FADD2 R.F32, -pair
                if (MODE == 5) { MUL(p[i], p[i], k2); SUB(q[i], k2, q[i]); }            // FADD2 pair, -pair
                if (MODE == 6) { MUL(p[i], p[i], p[i]); ADD(q[i], q[i], q[i]); }        // same-register operands (v*v)
                if (MODE == 8) { ADD(p[i], p[i], k2); MUL(p[i], p[i], k2); }            // one chain: add then mul (not contractable)
                if (MODE == 9) { ADD(p[i], p[i], k2); asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(p[i]) : "l"(p[i]), "l"(k2)); }   // same, IEEE multiply
                if (MODE == 10) { ADD(p[i], p[i], k2); }
                if (MODE == 11) { MUL(p[i], p[i], k2); }
                if (MODE == 12) { ADD(p[i], p[i], k2); MUL(q[i], q[i], k2); ADD(q[i], q[i], k2); MUL(p[i], p[i], k2); }   // two chains, alternating types
                if (MODE == 7) { asm volatile("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(p[i]) : "l"(p[i]), "l"(0x3f7ffffc3f7ffffcull)); ADD(q[i], q[i], k2); }   // immediate
            }
        }
    }
    float acc = 0; for (int i = 0; i < CH; i++) { float2 t = *(float2 *)&p[i], u = *(float2 *)&q[i]; acc += t.x + t.y + u.x + u.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int MODE> void run(const char *name, float *d, const u64 *c) {
    int sms, clk; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0); cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int iters = 4000; cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 3; w++) bench<MODE><<<sms * 4, 256>>>(d, iters, c);
    float ms = 1e9f;
    for (int rep = 0; rep < 5; rep++) { cudaEventRecord(e0); bench<MODE><<<sms * 4, 256>>>(d, iters, c); cudaEventRecord(e1); cudaEventSynchronize(e1); float t; cudaEventElapsedTime(&t, e0, e1); if (t < ms) ms = t; }
    printf("%-44s %7.3f ms  %5.2f cycles per (FMUL2, FADD2) pair per scheduler\n", name, ms, ms * 1e-3 * clk * 1e3 / (8.0 * iters * 4 * CH));
}
int main() {
    float *d; cudaMalloc(&d, 148 * 8 * 256 * 4 * 4);
    u64 hc[4]; float2 t; t = make_float2(-0.f, -0.f); hc[0] = *(u64 *)&t; t = make_float2(1.f, 1.f); hc[1] = *(u64 *)&t;
    t = make_float2(1.0000001f, 0.9999999f); hc[2] = *(u64 *)&t; t = make_float2(1.5f, 2.5f); hc[3] = *(u64 *)&t;
    u64 *c; cudaMalloc(&c, 32); cudaMemcpy(c, hc, 32, cudaMemcpyHostToDevice);
    for (int w = 0; w < 300; w++) bench<0><<<148 * 4, 256>>>(d, 4000, c);
    cudaDeviceSynchronize();
    run<0>("FMUL2 pair,pair | FADD2 pair,pair", d, c); run<1>("FMUL2 pair,R.F32 | FADD2 pair,pair", d, c); run<2>("FMUL2 pair,pair | FADD2 pair,R.F32", d, c);
    run<3>("FMUL2 pair,R.F32 | FADD2 pair,R.F32", d, c); run<4>("FMUL2 pair,pair | FADD2 R.F32,-pair", d, c); run<5>("FMUL2 pair,pair | FADD2 pair,-pair", d, c);
    run<8>("one chain: FADD2 -> FMUL2.FTZ", d, c); run<9>("one chain: FADD2 -> FMUL2 (IEEE)", d, c); run<10>("FADD2 only", d, c); run<11>("FMUL2.FTZ only", d, c);
    run<12>("two chains x (FADD2, FMUL2) [2 pairs]", d, c);
    run<6>("FMUL2 p,p (same reg) | FADD2 q,q", d, c); run<7>("FMUL2 pair,imm | FADD2 pair,pair", d, c);
    return 0;
}
