// ubench3.cu -- lattice field kernel inner loops (K1 on box meshes): the shipped form (two CHARGES per packed register, a
// thread owns PZ z-nodes of one column, dx / dy / dx^2+dy^2 shared along z) against the same arithmetic with two
// z-NODES per packed register and the charge consumed one at a time through 32-bit broadcast operands.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench3 tools/ubench3.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../pycpet_b200/csrc/common.cuh"
using namespace cpet;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

// FORM 0: shipped (eval_pair_lattice<MODE_FIELD_RAW, PZ>).  FORM 1: node pairs packed, PZ even.
template <int FORM, int PZ, int U>
__global__ void __launch_bounds__(256) kLat(const ChargePair* __restrict__ g, int npairs, int passes, float seed, float* sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    ChargePair* tile = reinterpret_cast<ChargePair*>(smem);
    for (int i = threadIdx.x; i < npairs * 2; i += 256) reinterpret_cast<uint4*>(smem)[i] = reinterpret_cast<const uint4*>(g)[i];
    __syncthreads();
    const float x = seed * threadIdx.x, y = 0.1f + seed * blockIdx.x;
    double tot = 0.0;
    for (int pass = 0; pass < passes; ++pass) {
        if (FORM == 0) {
            LatticeRegs<PZ> r;
            r.px = pk2(x, x); r.py = pk2(y, y);
#pragma unroll
            for (int p = 0; p < PZ; ++p) { const float z = 0.05f * p + 1e-4f * pass + 1e-3f * x; r.pz[p] = pk2(z, z); r.ax[p] = r.ay[p] = r.az[p] = 0ull; }
#pragma unroll U
            for (int j = 0; j < npairs; ++j) eval_pair_lattice<MODE_FIELD_RAW, PZ>(tile[j].a, tile[j].b, r);
#pragma unroll
            for (int p = 0; p < PZ; ++p) { float lo, hi; upk2(r.ax[p], lo, hi); tot += lo + hi; upk2(r.ay[p], lo, hi); tot += lo + hi; upk2(r.az[p], lo, hi); tot += lo + hi; }
        } else {
            u64 pz[PZ / 2], ax[PZ / 2], ay[PZ / 2], az[PZ / 2];
#pragma unroll
            for (int p = 0; p < PZ / 2; ++p) { const float z = 0.1f * p + 1e-4f * pass + 1e-3f * x; pz[p] = pk2(z, z + 0.05f); ax[p] = ay[p] = az[p] = 0ull; }
            const float* f = reinterpret_cast<const float*>(tile);
#pragma unroll U
            for (int j = 0; j < npairs; ++j) {
                // the pair's 8 floats {-x0,-x1,-y0,-y1,-z0,-z1,q0,q1}: two charges, one after the other
                const float4 v0 = *reinterpret_cast<const float4*>(f + 8 * j), v1 = *reinterpret_cast<const float4*>(f + 8 * j + 4);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float nx = h ? v0.y : v0.x, ny = h ? v0.w : v0.z, nz = h ? v1.y : v1.x, q = h ? v1.w : v1.z;
                    const float dx = x + nx, dy = y + ny;
                    const float rxy = fmaf(dy, dy, dx * dx);
#pragma unroll
                    for (int p = 0; p < PZ / 2; ++p) {
                        const u64 dz = add2(pz[p], pk2(nz, nz));
                        const u64 r2 = fma2(dz, dz, pk2(rxy, rxy));
                        float a, b; upk2(r2, a, b);
                        const u64 inv = pk2(rsqrt_approx(a), rsqrt_approx(b));
                        const u64 s = mul2(mul2(inv, inv), mul2(inv, pk2(q, q)));
                        ax[p] = fma2(s, pk2(dx, dx), ax[p]);
                        ay[p] = fma2(s, pk2(dy, dy), ay[p]);
                        az[p] = fma2(s, dz, az[p]);
                    }
                }
            }
#pragma unroll
            for (int p = 0; p < PZ / 2; ++p) { float lo, hi; upk2(ax[p], lo, hi); tot += lo + hi; upk2(ay[p], lo, hi); tot += lo + hi; upk2(az[p], lo, hi); tot += lo + hi; }
        }
    }
    if (tot == 123.456) sink[0] = (float)tot;
}

template <int FORM, int PZ, int U>
static int run(const char* name, const ChargePair* g, int npairs, int sms, float* sink) {
    auto kern = kLat<FORM, PZ, U>;
    const size_t smem = (size_t)npairs * sizeof(ChargePair);
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, kern));
    const int passes = 40, blocks = sms * 2;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0); kern<<<blocks, 256, smem>>>(g, npairs, passes, 0.013f, sink); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
    }
    CK(cudaGetLastError());
    const double pe = (double)blocks * 256 * passes * (double)npairs * 2 * PZ;
    printf("%-28s PZ=%d U=%d regs=%3d : %.3e pair-evals/s (%.1f%% of 3.7225e12)\n", name, PZ, U, fa.numRegs, pe / (best * 1e-3), pe / (best * 1e-3) / 3.7225e12 * 100);
    return 0;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount, npairs = 1024;
    float* h = (float*)malloc(sizeof(ChargePair) * npairs);
    srand(1);
    for (int i = 0; i < npairs * 8; ++i) h[i] = 3.0f + 20.0f * rand() / (float)RAND_MAX;
    ChargePair* g; CK(cudaMalloc(&g, sizeof(ChargePair) * npairs));
    CK(cudaMemcpy(g, h, sizeof(ChargePair) * npairs, cudaMemcpyHostToDevice));
    float* sink; CK(cudaMalloc(&sink, 64));
    for (int rep = 0; rep < 2; ++rep) {
        if (run<0, 5, 4>("charge pairs packed", g, npairs, sms, sink)) return 1;
        if (run<0, 4, 4>("charge pairs packed", g, npairs, sms, sink)) return 1;
        if (run<0, 6, 4>("charge pairs packed", g, npairs, sms, sink)) return 1;
        if (run<1, 4, 4>("node pairs packed", g, npairs, sms, sink)) return 1;
        if (run<1, 6, 4>("node pairs packed", g, npairs, sms, sink)) return 1;
        if (run<1, 6, 2>("node pairs packed", g, npairs, sms, sink)) return 1;
        if (run<1, 8, 2>("node pairs packed", g, npairs, sms, sink)) return 1;
        if (run<1, 8, 4>("node pairs packed", g, npairs, sms, sink)) return 1;
        if (run<1, 10, 2>("node pairs packed", g, npairs, sms, sink)) return 1;
    }
    return 0;
}
