// loop_bench.cu -- the Whitted query loops in isolation: packed pair tests (FADD2/FMUL2) versus the scalar pair tests,
// all lanes active, on a synthetic scene whose primitives are never hit (so the exact stage stays voted off).
// Prints SM cycles per primitive test per scheduler.   nvcc -arch=sm_100a -fmad=false -I../../se-195-project-ray-tracer_b200/csrc
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#include "packed_loops.cuh"
using namespace rtb;
template <int MODE> __global__ void __launch_bounds__(128) k(const f4 *geom, const f2x2 *pairs, int n, int reps, float *out) {
    __shared__ f4 s_geom[64]; __shared__ f2x2 s_pairs[64];
    for (int i = threadIdx.x; i < n; i += blockDim.x) s_geom[i] = geom[i];
    for (int i = threadIdx.x; i < n; i += blockDim.x) s_pairs[i] = pairs[i];
    __syncthreads();
    WLane L; 
    L.qox = 0.1f * threadIdx.x; L.qoy = 0.25f; L.qoz = -7.f; L.qdx = 0.01f * (threadIdx.x & 31); L.qdy = 0.1f; L.qdz = 0.99f;
    L.cumu = 1e7f; L.qhit = -1; L.qkind = 0; L.phase = PH_NEAREST;
    for (int k2 = 0; k2 < 3; k2++) { L.sox[k2] = L.qox + k2; L.soy[k2] = 0.3f; L.soz[k2] = -6.f; L.slx[k2] = 0.02f * k2; L.sly[k2] = 0.2f; L.slz[k2] = 0.97f; L.sreach[k2] = 5.f; }
    L.ns = 3; L.sblk = 0;
    for (int r = 0; r < reps; r++) {
        if (MODE == 0) { for (int i = 0; i + 1 < n; i += 2) w_sphere2<false>(L, s_geom + i, i, true); }
        if (MODE == 1) {
#pragma unroll 1
            for (int i = 0; i < n; i += 2) w_sphere_pair(L, s_pairs[i], s_pairs[i + 1], i, 1); }
        if (MODE == 2) { for (int i = 0; i < n; i++) w_shadow_sphere<false>(L, s_geom[i], 7, true); }
        if (MODE == 3) {
#pragma unroll 1
            for (int i = 0; i < n; i += 2) w_shadow_sphere_pair(L, s_pairs[i], s_pairs[i + 1], 7); }
        if (MODE == 4) { for (int i = 0; i + 1 < n; i += 2) w_plane2<false>(L, s_geom + i, i, true); }
        if (MODE == 5) {
#pragma unroll 1
            for (int i = 0; i < n; i += 2) w_plane_pair(L, s_pairs[i], s_pairs[i + 1], i, 1); }
        L.qox += 1e-6f;      // keep the loop from being hoisted
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = L.cumu + L.qhit + L.sblk;
}
template <int MODE> void run(const char *name, int tests_per_step, int warps_per_sched, const f4 *g, const f2x2 *p, float *o) {
    int sms, clk; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0); cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int n = 16, reps = 4000; cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int ctas = sms * warps_per_sched;        // 128 threads = 4 warps = one per scheduler
    for (int w = 0; w < 3; w++) k<MODE><<<ctas, 128>>>(g, p, n, reps, o);
    float best = 1e9f;
    for (int rep = 0; rep < 5; rep++) { cudaEventRecord(e0); k<MODE><<<ctas, 128>>>(g, p, n, reps, o); cudaEventRecord(e1); cudaEventSynchronize(e1); float t; cudaEventElapsedTime(&t, e0, e1); if (t < best) best = t; }
    const double tests = (double)warps_per_sched * reps * n * tests_per_step;      // per scheduler
    printf("%-34s %d warps/scheduler  %7.3f ms  %6.2f cycles per primitive-ray test per scheduler\n", name, warps_per_sched, best, best * 1e-3 * clk * 1e3 / tests);
}
int main() {
    std::vector<f4> g(16); std::vector<f2x2> p(16);
    for (int i = 0; i < 16; i++) g[i] = { 100.f + i, 50.f, -30.f - i, 1.0f };       // far-away spheres; as planes: steep normals, never candidates? (they are candidates sometimes -- fine)
    for (int i = 0; i < 16; i += 2) { p[i].a = { g[i].x, g[i + 1].x }; p[i].b = { g[i].y, g[i + 1].y }; p[i + 1].a = { g[i].z, g[i + 1].z }; p[i + 1].b = { g[i].w, g[i + 1].w }; }
    f4 *dg; f2x2 *dp; float *o; cudaMalloc(&dg, 16 * 16); cudaMalloc(&dp, 16 * 16); cudaMalloc(&o, 148 * 8 * 128 * 4);
    cudaMemcpy(dg, g.data(), 256, cudaMemcpyHostToDevice); cudaMemcpy(dp, p.data(), 256, cudaMemcpyHostToDevice);
    for (int w = 0; w < 200; w++) k<0><<<148 * 4, 128>>>(dg, dp, 16, 2000, o);
    cudaDeviceSynchronize();
    for (int wps : {1, 2, 4, 8}) {
        run<0>("nearest sphere, scalar pairs", 1, wps, dg, dp, o); run<1>("nearest sphere, packed pairs", 1, wps, dg, dp, o);
        run<2>("shadow sphere x3 rays, scalar", 3, wps, dg, dp, o); run<3>("shadow sphere x3 rays, packed", 3, wps, dg, dp, o);
        run<4>("nearest plane, scalar pairs", 1, wps, dg, dp, o); run<5>("nearest plane, packed pairs", 1, wps, dg, dp, o);
    }
    return 0;
}
