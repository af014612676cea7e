// ubench.cu -- register-resident instruction-mix microbenchmarks for the pair-evaluation loop.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench tools/ubench.cu
// Prints, per variant, the FP32-pipe lane-op rate and (for pair-eval variants) pair-evals/s.
#include <cstdio>
#include <cuda_runtime.h>
#include "../pycpet_b200/csrc/common.cuh"
using namespace cpet;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

// A: FFMA2, invariant multiplicands (operand reuse possible)
__global__ void __launch_bounds__(256) kA(int iters, float seed, float* sink) {
    u64 v[8]; const u64 a = pk2(1.0000001f, 0.9999999f), b = pk2(seed, -seed);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = pk2(seed + i, seed - i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = fma2(v[i], a, b);
    }
    float s = 0; for (int i = 0; i < 8; ++i) { float lo, hi; upk2(v[i], lo, hi); s += lo + hi; }
    if (s == 123.456f) sink[0] = s;
}
// B: FFMA2 with three distinct register pairs per instruction (no reuse)
__global__ void __launch_bounds__(256) kB(int iters, float seed, float* sink) {
    u64 v[8], y[8], z[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[i] = pk2(seed + i, seed - i); y[i] = pk2(1.0f + 1e-7f * i, 1.0f - 1e-7f * i); z[i] = pk2(seed * i, -seed * i); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = fma2(y[i], z[(i + u) & 7], v[i]);
    }
    float s = 0; for (int i = 0; i < 8; ++i) { float lo, hi; upk2(v[i], lo, hi); s += lo + hi; }
    if (s == 123.456f) sink[0] = s;
}
// C: FMUL2 + FADD2 alternating, distinct operands
__global__ void __launch_bounds__(256) kC(int iters, float seed, float* sink) {
    u64 v[8], y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[i] = pk2(seed + i, seed - i); y[i] = pk2(1.0f + 1e-7f * i, 1.0f - 1e-7f * i); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) { if (i & 1) v[i] = mul2(v[i], y[(i + u) & 7]); else v[i] = add2(v[i], y[(i + u + 1) & 7]); }
    }
    float s = 0; for (int i = 0; i < 8; ++i) { float lo, hi; upk2(v[i], lo, hi); s += lo + hi; }
    if (s == 123.456f) sink[0] = s;
}

// pair-eval mixes on register-resident fake charges (8 pairs cycled), P points per thread
template <int P, int VARIANT>   // 0 packed+MUFU, 1 packed no MUFU, 2 scalar+MUFU, 3 packed+MUFU soft
__global__ void __launch_bounds__(256) kPair(int iters, float seed, float* sink) {
    u64 cx[4], cy[4], cz[4], cq[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        cx[j] = pk2(-10.f - j - seed, -11.f - j); cy[j] = pk2(-7.f + j, -3.f - j * seed);
        cz[j] = pk2(5.f + j, 6.f + seed * j); cq[j] = pk2(0.3f * (j + 1), -0.2f * (j + 1));
    }
    PointRegs<P> r;
#pragma unroll
    for (int p = 0; p < P; ++p) set_point<P>(r, p, 0.1f * p + seed * threadIdx.x * 1e-3f, 0.2f * p, 0.3f * p);
    clear_partials<P>(r);
    float sx[P], sy[P], sz[P];
#pragma unroll
    for (int p = 0; p < P; ++p) sx[p] = sy[p] = sz[p] = 0.f;
    const u64 dlt = pk2(seed * 1e-6f, seed * 1e-6f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int p = 0; p < P; ++p) { r.px[p] = add2(r.px[p], dlt); r.py[p] = add2(r.py[p], dlt); r.pz[p] = add2(r.pz[p], dlt); }  // keeps d = p - x loop-variant
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (VARIANT == 0 || VARIANT == 3) {
                PairA a; a.nx = cx[j]; a.ny = cy[j];
                PairB b; b.nz = cz[j]; b.q = cq[j];
                if (VARIANT == 0) eval_pair<MODE_FIELD_RAW, P>(a, b, r);
                else eval_pair<MODE_FIELD_SOFT, P>(a, b, r);
            } else if (VARIANT == 1) {
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    const u64 dx = add2(r.px[p], cx[j]), dy = add2(r.py[p], cy[j]), dz = add2(r.pz[p], cz[j]);
                    u64 r2 = mul2(dx, dx); r2 = fma2(dy, dy, r2); r2 = fma2(dz, dz, r2);
                    const u64 inv = r2;                       // no MUFU
                    const u64 t = mul2(inv, inv), u = mul2(inv, cq[j]), s = mul2(t, u);
                    r.ax[p] = fma2(s, dx, r.ax[p]); r.ay[p] = fma2(s, dy, r.ay[p]); r.az[p] = fma2(s, dz, r.az[p]);
                }
            } else {
                float x0, x1, y0, y1, z0, z1, q0, q1;
                upk2(cx[j], x0, x1); upk2(cy[j], y0, y1); upk2(cz[j], z0, z1); upk2(cq[j], q0, q1);
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    float px, py, pz, d;
                    upk2(r.px[p], px, d); upk2(r.py[p], py, d); upk2(r.pz[p], pz, d);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const float dx = px + (h ? x1 : x0), dy = py + (h ? y1 : y0), dz = pz + (h ? z1 : z0);
                        const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                        const float inv = rsqrt_approx(r2);
                        const float s = (inv * inv) * (inv * (h ? q1 : q0));
                        sx[p] = fmaf(s, dx, sx[p]); sy[p] = fmaf(s, dy, sy[p]); sz[p] = fmaf(s, dz, sz[p]);
                    }
                }
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int p = 0; p < P; ++p) { float lo, hi; upk2(r.ax[p], lo, hi); s += lo + hi + sx[p] + sy[p] + sz[p];
        upk2(r.ay[p], lo, hi); s += lo + hi; upk2(r.az[p], lo, hi); s += lo + hi; }
    if (s == 123.456f) sink[0] = s;
}

template <typename F>
static float time_kernel(F launch) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    float* sink; CK(cudaMalloc(&sink, 64));
    const int iters = 2048;
    for (int bps : {4, 8}) {
        const int blocks = prop.multiProcessorCount * bps, threads = 256;
        const double thr = (double)blocks * threads;
        printf("== %d blocks/SM x 256 threads\n", bps);
        float ms;
        ms = time_kernel([&] { kA<<<blocks, threads>>>(iters, 0.5f, sink); });
        printf("A FFMA2 reuse      : %.1f TFLOP/s\n", thr * iters * 32 * 2 * 2 / (ms * 1e-3) / 1e12);
        ms = time_kernel([&] { kB<<<blocks, threads>>>(iters, 0.5f, sink); });
        printf("B FFMA2 distinct   : %.1f TFLOP/s\n", thr * iters * 32 * 2 * 2 / (ms * 1e-3) / 1e12);
        ms = time_kernel([&] { kC<<<blocks, threads>>>(iters, 0.5f, sink); });
        printf("C FMUL2+FADD2      : %.1f T lane-op/s (peak = %.1f)\n", thr * iters * 16 * 2 / (ms * 1e-3) / 1e12,
               prop.multiProcessorCount * 128 * 1.965e9 / 1e12);
#define RUNP(P, V, name) ms = time_kernel([&] { kPair<P, V><<<blocks, threads>>>(iters, 0.5f, sink); }); \
        printf("pair P=%d %-18s: %.3e pair-evals/s (%.1f%% of 3.72e12)\n", P, name, thr * iters * 8.0 * P / (ms * 1e-3), \
               thr * iters * 8.0 * P / (ms * 1e-3) / 3.7225e12 * 100);
        RUNP(1, 0, "packed raw") RUNP(2, 0, "packed raw") RUNP(4, 0, "packed raw")
        RUNP(2, 3, "packed soft") RUNP(4, 3, "packed soft")
        RUNP(2, 1, "packed noMUFU") RUNP(4, 1, "packed noMUFU")
        RUNP(1, 2, "scalar raw") RUNP(2, 2, "scalar raw") RUNP(4, 2, "scalar raw")
    }
    return 0;
}
