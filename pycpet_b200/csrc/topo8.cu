// topo8.cu -- K2, points-packed hybrid form: the default streamline integrator for frames whose queue gives
// every warp of the chip at least 4 lines.
//
// Same semantics and the same warp-wide organisation as topo.cu (CPET/utils/math_module.c:523-591, 489-503,
// 296-333; CPET/source/calculator.py:675-712; CPET/utils/gpu.py:25-399): a persistent kernel, one CTA per
// SM, the 32 lanes of a warp split the charges of the frame, lines are pulled from a global LPT queue, K + 2
// field evaluations per line, the per-line state machine in FP32 like the reference's own float arithmetic.
// What differs from k2x_topo_kernel:
//   * a packed FP32x2 register holds two POINTS (not two charges); a lane takes one charge per step and feeds
//     it to the arithmetic through 32-bit broadcast operands (common.cuh: evalp_far / evalp_near, PBlock);
//   * a warp owns up to 8 lines = 4 packed point pairs in registers (NP = 4, 2 or 1 pairs are evaluated per
//     pass, compacted, so the end of the queue costs short passes);
//   * the 32 FP32 partials of a lane (8 positions x {T.x, T.y, T.z, S}) go through a transposing warp
//     reduction that leaves lane l with the total of position l >> 2, component l & 3.
#include "cpet_internal.h"

namespace cpet {

#ifndef CPET_K2P_MAXT
#define CPET_K2P_MAXT 384       // 12 warps x <= 168 registers
#endif
// Blocks (= charges per lane) per FP32 accumulation chain.  128 -> 256 is +1.1 % (one warp reduction per pass on the
// 3A frame instead of two) for 8.9e-6 -> 1.3e-5 maximum curvature error against float64 on 512 sampled lines
// (tolerance 4e-5); 512 gives nothing more (profiles/round2_k2p_tuning.txt).
#ifndef CPET_K2P_CHUNK
#define CPET_K2P_CHUNK 256
#endif
#define K2P_SLOTS 8

struct K2PParams {
    const PBlock* blocks;
    const K2XMeta* meta;
    int tile_blocks;
    int stages;
    int resident;
    int cap;                // streamlines per warp (1, 2, 4 or 8)
    int tail4, tail2;       // lines per warp of the grid left in the queue from which a warp tops up to 4 / to 2 only
    const float* seeds;
    const int32_t* n_iter;
    const int32_t* order;
    int n_lines;
    float h;
    float dimx, dimy, dimz;
    float* out;
    int32_t* steps;
    unsigned int* queue;
    unsigned long long* evals;
};

// State of the (up to) 8 streamlines of one warp, in shared memory so that none of it occupies registers
// during the charge loop.  Arrays of 8 = one entry per line slot; c* are the slots' current points compacted
// to positions 0..na-1 (what the charge loop reads; position of a slot = its rank among the active slots).
struct __align__(16) WarpLinesP {
    float px[K2P_SLOTS], py[K2P_SLOTS], pz[K2P_SLOTS];         // current point p_k
    float sx[K2P_SLOTS], sy[K2P_SLOTS], sz[K2P_SLOTS];         // seed
    float ux[K2P_SLOTS], uy[K2P_SLOTS], uz[K2P_SLOTS];         // unit field direction at p_{k-1}
    float cx[K2P_SLOTS], cy[K2P_SLOTS], cz[K2P_SLOTS];         // current points by position
    double t[K2P_SLOTS][4];                                    // sums of the current pass by POSITION: T - E_near (3), S
    float dist[K2P_SLOTS], kinit[K2P_SLOTS];
    int line[K2P_SLOTS], n_it[K2P_SLOTS], k[K2P_SLOTS], k_end[K2P_SLOTS];
    float m1x[K2P_SLOTS], m1y[K2P_SLOTS], m1z[K2P_SLOTS], m2x[K2P_SLOTS], m2y[K2P_SLOTS], m2z[K2P_SLOTS];
};

__device__ __forceinline__ float xadd_f32(float a, float b, int m, bool upper) {
    const float keep = upper ? b : a;
    const float send = upper ? a : b;
    return keep + __shfl_xor_sync(0xffffffffu, send, m);
}
__device__ __forceinline__ double xadd_f64(double a, double b, int m, bool upper) {
    const double keep = upper ? b : a;
    const double send = upper ? a : b;
    return keep + shfl_xor_f64(send, m);
}

// FP32 partials of the warp's 2*NP positions -> FP64 running total v of (position (lane >> 2) mod 2NP,
// component lane & 3).  Every value is summed over the lanes in the same tree whatever NP is -- partner at
// lane distance 16, then 8 (FP32), then 4, 2, 1 (FP64) -- so a line's sums do not depend on how many lines its
// warp holds; where the number of values does not halve, the step is a plain butterfly.
template <int NP>
__device__ __forceinline__ void fold8(PRegs& r, double& v, int lane) {
    float d[2 * NP][4];
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        upk2(r.a0[p], d[2 * p][0], d[2 * p + 1][0]);
        upk2(r.a1[p], d[2 * p][1], d[2 * p + 1][1]);
        upk2(r.a2[p], d[2 * p][2], d[2 * p + 1][2]);
        upk2(r.a3[p], d[2 * p][3], d[2 * p + 1][3]);
        r.a0[p] = r.a1[p] = r.a2[p] = r.a3[p] = 0ull;
    }
    const bool up16 = (lane & 16) != 0, up8 = (lane & 8) != 0, up4 = (lane & 4) != 0;
    const bool up2 = (lane & 2) != 0, up1 = (lane & 1) != 0;
    float e[2][4];                     // after the steps at distance 16 and 8: two positions per lane
    if (NP == 4) {
        float w[4][4];                 // lanes < 16 keep positions 0..3, the others 4..7
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int c = 0; c < 4; ++c) w[q][c] = xadd_f32(d[q][c], d[(q + 4) % (2 * NP)][c], 16, up16);
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int c = 0; c < 4; ++c) e[q][c] = xadd_f32(w[q][c], w[q + 2][c], 8, up8);
    } else if (NP == 2) {
        float w[4][4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int c = 0; c < 4; ++c) w[q][c] = d[q % (2 * NP)][c] + __shfl_xor_sync(0xffffffffu, d[q % (2 * NP)][c], 16);
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int c = 0; c < 4; ++c) e[q][c] = xadd_f32(w[q][c], w[q + 2][c], 8, up8);
    } else {
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float w = d[q % (2 * NP)][c] + __shfl_xor_sync(0xffffffffu, d[q % (2 * NP)][c], 16);
                e[q][c] = w + __shfl_xor_sync(0xffffffffu, w, 8);
            }
    }
    double f[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) f[c] = xadd_f64((double)e[0][c], (double)e[1][c], 4, up4);
    const double g0 = xadd_f64(f[0], f[2], 2, up2);
    const double g1 = xadd_f64(f[1], f[3], 2, up2);
    v += xadd_f64(g0, g1, 1, up1);
}

// n blocks starting at `pa` (this lane's a-record of block 0; its b-record is 512 + 4*lane - 16*lane bytes
// further, passed as pb).  FP32 chains are cut every CPET_K2P_CHUNK blocks.
template <int NP, int U, bool NEAR>
__device__ __forceinline__ void evalp_run(const unsigned char* __restrict__ pa, const unsigned char* __restrict__ pb,
                                          int n, int lane, PRegs& r, double& v, int& run) {
    while (n > 0) {
        int m = CPET_K2P_CHUNK - run;
        if (m > n) m = n;
        if (NEAR) {
#pragma unroll 1
            for (int j = 0; j < m; ++j, pa += sizeof(PBlock))
                evalp_near<NP>(*reinterpret_cast<const float4*>(pa), r);
        } else {
#pragma unroll U
            for (int j = 0; j < m; ++j, pa += sizeof(PBlock), pb += sizeof(PBlock))
                evalp_far<NP>(*reinterpret_cast<const float4*>(pa), *reinterpret_cast<const float*>(pb), r);
        }
        n -= m;
        run += m;
        if (run >= CPET_K2P_CHUNK) { fold8<NP>(r, v, lane); run = 0; }
    }
}

// Blocks [g0, g1) of the array [near | far] (tile[0] is block gbase) against the warp's 2*NP positions.
template <int NP, int U>
__device__ __forceinline__ void evalp_blocks(const PBlock* __restrict__ tile, int gbase, int g0, int g1, int nb_near,
                                             int lane, PRegs& r, double& v, int& run) {
    const unsigned char* base = reinterpret_cast<const unsigned char*>(tile) - (size_t)gbase * sizeof(PBlock);
    const unsigned char* la = base + 16 * lane;
    const unsigned char* lb = base + 512 + 4 * lane;
    int b = g0;
    if (b < nb_near && b < g1) {
        const int e = min(g1, nb_near);
        evalp_run<NP, 1, true>(la + (size_t)b * sizeof(PBlock), lb, e - b, lane, r, v, run);
        b = e;
    }
    if (b < g1)
        evalp_run<NP, U, false>(la + (size_t)b * sizeof(PBlock), lb + (size_t)b * sizeof(PBlock), g1 - b, lane, r, v, run);
}

template <bool SD, int U>
__global__ void __launch_bounds__(CPET_K2P_MAXT, 1) k2p_topo_kernel(const K2PParams prm) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
    WarpLinesP* lines_all = reinterpret_cast<WarpLinesP*>(smem_raw + 128);
    const int n_warps = blockDim.x >> 5;
    PBlock* ring = reinterpret_cast<PBlock*>(smem_raw + 128 + sizeof(WarpLinesP) * n_warps);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    WarpLinesP& W = lines_all[tid >> 5];
    const bool owner = lane < K2P_SLOTS;   // lane q < 8 runs the state machine of line slot q
    const int S = prm.stages;
    const int TB = prm.tile_blocks;
    const int nb_total = prm.meta->nb_total;
    const int nb_near = prm.meta->nb_near;
    const int NT = (nb_total + TB - 1) / TB;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (owner) W.line[lane] = -1;
    __syncthreads();

    auto issue = [&](int it) {                 // streamed mode: tile it % NT into stage it % S
        const int stage = it % S;
        const int t = it % NT;
        const int n_t = min(TB, nb_total - t * TB);
        const uint32_t bytes = (uint32_t)n_t * (uint32_t)sizeof(PBlock);
        mbar_expect_tx(&full[stage], bytes);
        tma_load_1d(ring + (size_t)stage * TB, prm.blocks + (size_t)t * TB, bytes, &full[stage]);
    };
    int issued = 0;
    if (prm.resident) {
        if (tid == 0 && nb_total > 0) {        // the whole frame once: bulk copies of <= 32 blocks on one barrier
            mbar_expect_tx(&full[0], (uint32_t)nb_total * (uint32_t)sizeof(PBlock));
            for (int b = 0; b < nb_total; b += 32) {
                const int n = min(32, nb_total - b);
                tma_load_1d(ring + b, prm.blocks + b, (uint32_t)n * (uint32_t)sizeof(PBlock), &full[0]);
            }
        }
        if (nb_total > 0) mbar_wait(&full[0], 0u);
    } else if (tid == 0) {
        const int pre = min(S, NT);
        for (; issued < pre; ++issued) issue(issued);
    }

    bool exhausted = false;
    // 8 lines per warp while the queue is long; over its last stretch (4 lines per warp of the grid left) a
    // warp only tops up to 4, so the end of the queue is worked off by half-width passes on all warps
    int cap_now = prm.cap;
    const long long grid_warps = (long long)gridDim.x * (long long)n_warps;
    const long long tail_start = (long long)prm.n_lines - (long long)prm.tail4 * grid_warps;
    const long long tail2_start = (long long)prm.n_lines - (long long)prm.tail2 * grid_warps;
    unsigned long long my_evals = 0ull;
    const float hf = prm.h;
    const float inv_hf = 1.0f / prm.h;

    int it = 0;   // consumed-tile counter (streamed mode)
    while (true) {
        // ---- refill the warp's empty line slots from the queue (one atomic per warp) ------------
        {
            const bool empty = owner && W.line[lane] < 0;
            const unsigned em = __ballot_sync(0xffffffffu, empty);
            const int n_empty = __popc(em);
            int want = min(n_empty, cap_now - (K2P_SLOTS - n_empty));
            if (exhausted || want < 0) want = 0;
            if (want > 0) {                                    // warp-uniform
                unsigned base = 0;
                if (lane == 0) base = atomicAdd(prm.queue, (unsigned)want);
                base = __shfl_sync(0xffffffffu, base, 0);
                const int rank = __popc(em & ((1u << lane) - 1u));
                const unsigned slot = base + (unsigned)rank;
                if (empty && rank < want && slot < (unsigned)prm.n_lines) {
                    const int line = prm.order ? prm.order[slot] : (int)slot;
                    const float sx = prm.seeds[3 * (size_t)line];
                    const float sy = prm.seeds[3 * (size_t)line + 1];
                    const float sz = prm.seeds[3 * (size_t)line + 2];
                    const int n_it = prm.n_iter[line];
                    W.line[lane] = line;
                    W.sx[lane] = sx; W.sy[lane] = sy; W.sz[lane] = sz;
                    W.px[lane] = sx; W.py[lane] = sy; W.pz[lane] = sz;
                    if (SD) {
                        W.m1x[lane] = W.m2x[lane] = sx; W.m1y[lane] = W.m2y[lane] = sy;
                        W.m1z[lane] = W.m2z[lane] = sz;
                    }
                    W.n_it[lane] = n_it;
                    W.k[lane] = 0;
                    W.k_end[lane] = (n_it <= 0) ? 0 : -1;
                    W.dist[lane] = 0.f;
                    W.kinit[lane] = 0.f;
                    W.ux[lane] = W.uy[lane] = W.uz[lane] = 0.f;
                }
                if (base + (unsigned)want >= (unsigned)prm.n_lines) exhausted = true;
                if (cap_now > 4 && (long long)base + want >= tail_start) cap_now = 4;
                if (cap_now > 2 && (long long)base + want >= tail2_start) cap_now = 2;
            }
        }
        // ---- compact the active slots' points to positions 0..na-1 ----------------------------------------
        const bool active = owner && W.line[lane] >= 0;       // own slot: written by this lane
        const unsigned am = __ballot_sync(0xffffffffu, active);
        bool go;
        if (prm.resident) go = (am != 0u);
        else go = __syncthreads_or(am != 0u ? 1 : 0) != 0;
        if (!go) break;
        const int na = __popc(am);
        if (active) {
            const int pos = __popc(am & ((1u << lane) - 1u));
            W.cx[pos] = W.px[lane]; W.cy[pos] = W.py[lane]; W.cz[pos] = W.pz[lane];
        }
        __syncwarp();

        PRegs r;
        double v = 0.0;
        {
            const float4 x03 = *reinterpret_cast<const float4*>(&W.cx[0]), x47 = *reinterpret_cast<const float4*>(&W.cx[4]);
            const float4 y03 = *reinterpret_cast<const float4*>(&W.cy[0]), y47 = *reinterpret_cast<const float4*>(&W.cy[4]);
            const float4 z03 = *reinterpret_cast<const float4*>(&W.cz[0]), z47 = *reinterpret_cast<const float4*>(&W.cz[4]);
            float qx[8] = {x03.x, x03.y, x03.z, x03.w, x47.x, x47.y, x47.z, x47.w};
            float qy[8] = {y03.x, y03.y, y03.z, y03.w, y47.x, y47.y, y47.z, y47.w};
            float qz[8] = {z03.x, z03.y, z03.z, z03.w, z47.x, z47.y, z47.z, z47.w};
#pragma unroll
            for (int p = 0; p < 8; ++p)
                if (p >= na) { qx[p] = qx[0]; qy[p] = qy[0]; qz[p] = qz[0]; }   // unused positions repeat a valid point
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                r.c0[j] = pk2(-2.0f * qx[2 * j], -2.0f * qx[2 * j + 1]);
                r.c1[j] = pk2(-2.0f * qy[2 * j], -2.0f * qy[2 * j + 1]);
                r.c2[j] = pk2(-2.0f * qz[2 * j], -2.0f * qz[2 * j + 1]);
                r.c3[j] = pk2(fmaf(qz[2 * j], qz[2 * j], fmaf(qy[2 * j], qy[2 * j], qx[2 * j] * qx[2 * j])),
                              fmaf(qz[2 * j + 1], qz[2 * j + 1], fmaf(qy[2 * j + 1], qy[2 * j + 1], qx[2 * j + 1] * qx[2 * j + 1])));
                r.a0[j] = r.a1[j] = r.a2[j] = r.a3[j] = 0ull;
            }
        }

        // ---- field sums at those points: all charges, split over the 32 lanes -------------------------
        int run = 0;
        if (prm.resident) {
            if (na > 4) { evalp_blocks<4, U>(ring, 0, 0, nb_total, nb_near, lane, r, v, run); fold8<4>(r, v, lane); }
            else if (na > 2) { evalp_blocks<2, U>(ring, 0, 0, nb_total, nb_near, lane, r, v, run); fold8<2>(r, v, lane); }
            else { evalp_blocks<1, 8>(ring, 0, 0, nb_total, nb_near, lane, r, v, run); fold8<1>(r, v, lane); }
        } else {
            for (int t = 0; t < NT; ++t, ++it) {
                const int stage = it % S;
                mbar_wait(&full[stage], (uint32_t)((it / S) & 1));
                const int g0 = t * TB, g1 = min(nb_total, g0 + TB);
                const PBlock* tile = ring + (size_t)stage * TB;
                if (na > 4) evalp_blocks<4, U>(tile, g0, g0, g1, nb_near, lane, r, v, run);
                else if (na > 2) evalp_blocks<2, U>(tile, g0, g0, g1, nb_near, lane, r, v, run);
                else if (na >= 1) evalp_blocks<1, 8>(tile, g0, g0, g1, nb_near, lane, r, v, run);
                __syncthreads();                      // stage fully consumed by the CTA
                if (tid == 0) { issue(issued); ++issued; }   // speculative: next pass's tiles too
            }
            if (na > 4) fold8<4>(r, v, lane);
            else if (na > 2) fold8<2>(r, v, lane);
            else fold8<1>(r, v, lane);
        }

        // ---- lane l holds the total of (position (l >> 2) mod 2NP, component l & 3) -------------------------
        {
            const int npos = na > 4 ? 8 : (na > 2 ? 4 : 2);
            if (lane < 4 * npos && (lane >> 2) < na) W.t[lane >> 2][lane & 3] = v;
        }
        __syncwarp();

        // ---- state machine of line slot `lane` (FP32, like the reference's own float arithmetic, C:489-503;
        //      the sums arrive in FP64 and E = p*S - T is formed there) ------------------------------------
        if (active) {
            const int q = lane;
            const int pos = __popc(am & ((1u << lane) - 1u));
            const float px = W.px[q], py = W.py[q], pz = W.pz[q];
            // E = p*S - (T - E_near), up to the Coulomb constant (only the direction is used)
            const double2 t01 = *reinterpret_cast<const double2*>(&W.t[pos][0]);
            const double2 t23 = *reinterpret_cast<const double2*>(&W.t[pos][2]);
            const double ss = t23.y;
            const float ex = (float)((double)px * ss - t01.x);
            const float ey = (float)((double)py * ss - t01.y);
            const float ez = (float)((double)pz * ss - t23.x);
            int k = W.k[q];
            const int k_end = W.k_end[q];
            // unit direction (no zero guard: E = 0 gives NaN exactly like C:501)
            const float n2 = fmaf(ez, ez, fmaf(ey, ey, ex * ex));
            float inv_n;
            if (n2 > 1e-30f && n2 < 1e30f) {
                const float y = rsqrt_approx(n2);
                inv_n = y * fmaf(-0.5f * n2 * y, y, 1.5f);       // one Newton step: ~1 ulp
            } else {
                inv_n = (float)(1.0 / sqrt((double)ex * ex + (double)ey * ey + (double)ez * ez));
            }
            const float ux = ex * inv_n, uy = ey * inv_n, uz = ez * inv_n;
            const bool last = (k_end >= 0) && (k == k_end + 1);
            float kdir = 0.f;
            if (!SD && (k == 1 || last)) {   // curvature is needed at the first and last point pair
                // |u0 x u1| / h with u1 = u0 + d: u0 x u1 = u0 x d, and d = u1 - u0 is exact in FP32 for
                // neighbouring directions, so the cross product carries no cancellation
                const float pux = W.ux[q], puy = W.uy[q], puz = W.uz[q];
                const float dxu = ux - pux, dyu = uy - puy, dzu = uz - puz;
                const float cx = fmaf(puy, dzu, -puz * dyu);
                const float cy = fmaf(puz, dxu, -pux * dzu);
                const float cz = fmaf(pux, dyu, -puy * dxu);
                kdir = sqrtf(fmaf(cz, cz, fmaf(cy, cy, cx * cx))) * inv_hf;
                if (k == 1) W.kinit[q] = kdir;
            }
            const float nx = fmaf(hf, ux, px);
            const float ny = fmaf(hf, uy, py);
            const float nz = fmaf(hf, uz, pz);
            if (last) {
                if (SD) {
                    kdir = curv3_f32(make_float3(W.m1x[q], W.m1y[q], W.m1z[q]), make_float3(px, py, pz),
                                     make_float3(nx, ny, nz));
                    if (k == 1) W.kinit[q] = kdir;
                }
                const int line = W.line[q];
                reinterpret_cast<float2*>(prm.out)[line] = make_float2(W.dist[q], (W.kinit[q] + kdir) * 0.5f);
                if (prm.steps) prm.steps[line] = k_end;
                my_evals += (unsigned long long)(k_end + 2);
                W.line[q] = -1;
            } else {
                if (SD) {
                    W.m2x[q] = W.m1x[q]; W.m2y[q] = W.m1y[q]; W.m2z[q] = W.m1z[q];
                    W.m1x[q] = px; W.m1y[q] = py; W.m1z[q] = pz;
                }
                W.px[q] = nx; W.py[q] = ny; W.pz[q] = nz;
                ++k;
                W.k[q] = k;
                if (SD && k == 2)
                    W.kinit[q] = curv3_f32(make_float3(W.m2x[q], W.m2y[q], W.m2z[q]),
                                           make_float3(W.m1x[q], W.m1y[q], W.m1z[q]), make_float3(nx, ny, nz));
                if (k_end < 0) {
                    const bool outside = (nx < -prm.dimx) || (nx > prm.dimx) || (ny < -prm.dimy) ||
                                         (ny > prm.dimy) || (nz < -prm.dimz) || (nz > prm.dimz);
                    if (k >= W.n_it[q] || outside) {
                        W.k_end[q] = k;
                        const float ddx = W.sx[q] - nx, ddy = W.sy[q] - ny, ddz = W.sz[q] - nz;
                        W.dist[q] = sqrtf(fmaf(ddz, ddz, fmaf(ddy, ddy, ddx * ddx)));
                    }
                }
                W.ux[q] = ux; W.uy[q] = uy; W.uz[q] = uz;
            }
        }
        __syncwarp();
    }

    if (!prm.resident) {
        // drain the speculative loads before the CTA (and its shared memory) retires
        if (tid == 0) {
            for (; it < issued; ++it) mbar_wait(&full[it % S], (uint32_t)((it / S) & 1));
        }
    }
    if (my_evals) atomicAdd(prm.evals, my_evals);
}

template <bool SD, int U>
static int launch_k2p_inst(cpet_ctx* c, const K2PParams& prm, int grid, int threads, size_t smem) {
    auto kern = k2p_topo_kernel<SD, U>;
    CPET_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, threads, smem, c->stream>>>(prm);
    CPET_CUDA_TRY(cudaGetLastError());
    return CPET_OK;
}

static int k2p_threads(const cpet_ctx* c) {
    int threads = c->tune.k2_threads > 0 ? c->tune.k2_threads : CPET_K2P_MAXT;
    threads = (threads / 32) * 32;
    if (threads < 32) threads = 32;
    if (threads > CPET_K2P_MAXT) threads = CPET_K2P_MAXT;
    return threads;
}

static bool k2p_resident(const cpet_ctx* c, int warps_per_cta) {
    const size_t hdr = 128 + sizeof(WarpLinesP) * (size_t)warps_per_cta;
    const int max_blocks = (c->n_charges + 31) / 32 + 2;
    return hdr + (size_t)max_blocks * sizeof(PBlock) <= (size_t)c->max_smem_optin && c->tune.k2_stages <= 0 &&
           c->tune.k2_tile_pairs <= 0;
}

// The points-packed kernel holds 8 lines per warp, so a short queue leaves it a ragged last round: measured
// against the charge-pair-packed kernel (4 lines per warp) it wins from about 20 lines per warp of the chip when
// the frame is resident in shared memory and from about 28 when the charges are streamed (lock-step passes):
// 39,304 lines x 7,890 charges 0.710 against 0.687 of the FP32 peak, 27,000 lines 0.662 against 0.673; 54,872
// lines x 100,000 charges 0.710 against 0.694, 39,304 lines 0.688 against 0.696 (profiles/round2_k2_forms.md).
// An explicit k2_cap of 1 or 2 also keeps the charge-pair-packed kernel, which fills a warp with 1 or 2 lines.
bool topo8_wants(cpet_ctx* c, int n_lines) {
    if (c->tune.k2_cap == 1 || c->tune.k2_cap == 2) return false;
    const int warps = k2p_threads(c) / 32;
    const long long all_warps = (long long)c->sm_count * warps;
    return (long long)n_lines >= (k2p_resident(c, warps) ? 20 : 28) * all_warps;
}

int launch_topo_points_packed(cpet_ctx* c, int n_lines, const float* d_seeds, const int32_t* d_n_iter, float step,
                              const float dims[3], unsigned flags, float* d_out, int32_t* d_steps) {
    const Tuning& tu = c->tune;
    const int sms = c->sm_count;
    int launches = 0;

    int threads = k2p_threads(c);
    if (tu.k2_threads <= 0 && (long long)n_lines < (long long)sms * (threads / 32)) {
        // fewer lines than warps on the chip: one line per warp, spread over all SMs
        threads = 32 * ((n_lines + sms - 1) / sms);
    }
    const int warps_per_cta = threads / 32;
    const size_t hdr = 128 + sizeof(WarpLinesP) * (size_t)warps_per_cta;

    // --- charge staging plan; the class sizes are only known on the device, so plan for the worst case: every
    //     class ends in a partly filled block
    const int max_blocks = (c->n_charges + 31) / 32 + 2;
    K2PParams prm;
    size_t smem;
    if (k2p_resident(c, warps_per_cta)) {
        prm.resident = 1;
        prm.tile_blocks = max_blocks;
        prm.stages = 1;
        smem = hdr + (size_t)max_blocks * sizeof(PBlock);
    } else {
        prm.resident = 0;
        prm.stages = tu.k2_stages > 0 ? tu.k2_stages : 2;
        if (prm.stages > 8) prm.stages = 8;
        if (prm.stages < 2) prm.stages = 2;
        const int fit = (int)(((size_t)c->max_smem_optin - hdr) / ((size_t)prm.stages * sizeof(PBlock)));
        prm.tile_blocks = tu.k2_tile_pairs > 0 ? tu.k2_tile_pairs / 16 : fit;    // 16 charge pairs per block
        if (prm.tile_blocks > fit) prm.tile_blocks = fit;
        if (prm.tile_blocks < 1) prm.tile_blocks = 1;
        smem = hdr + (size_t)prm.stages * prm.tile_blocks * sizeof(PBlock);
    }

    // --- lines per warp ----------------------------------------------------------------------------------
    const long long all_warps = (long long)sms * warps_per_cta;
    int cap = tu.k2_cap;
    if (cap != 1 && cap != 2 && cap != 4 && cap != 8)
        cap = n_lines >= 8 * all_warps ? 8 : (n_lines >= 4 * all_warps ? 4 : (n_lines >= 3 * all_warps ? 2 : 1));
    prm.cap = cap;
    prm.tail4 = tu.k2_tail4 >= 0 ? tu.k2_tail4 : 4;
    prm.tail2 = tu.k2_tail2 >= 0 ? tu.k2_tail2 : 0;
    int grid = sms;
    const long long need_ctas = (n_lines + (long long)warps_per_cta * cap - 1) / ((long long)warps_per_cta * cap);
    if (need_ctas < grid) grid = (int)need_ctas;
    const long long slots = (long long)grid * warps_per_cta * cap;
    const bool do_sort = (tu.k2_sort < 0) ? (n_lines > slots) : (tu.k2_sort != 0);
    if (int rc = prepare_queue(c, n_lines, d_n_iter, do_sort, &prm.queue, &prm.evals, &prm.order, &launches))
        return rc;

    K2XMeta* meta = nullptr;
    if (int rc = pack_hybrid(c, n_lines, d_seeds, step, dims, 1, max_blocks, &meta, &launches)) return rc;

    prm.blocks = c->xblocks.as<PBlock>();
    prm.meta = meta;
    prm.seeds = d_seeds;
    prm.n_iter = d_n_iter;
    prm.n_lines = n_lines;
    prm.h = step;
    prm.dimx = dims[0]; prm.dimy = dims[1]; prm.dimz = dims[2];
    prm.out = d_out;
    prm.steps = d_steps;

    KernelTimer timer(c);   // brackets the integrator kernel only (the roofline's "dominant kernel")
    const bool sd = (flags & CPET_TOPO_CURV_SECOND_DIFF) != 0u;
    int rc;
    const int unroll = tu.k2_unroll > 0 ? tu.k2_unroll : 6;      // 3A frame: 0.760 / 0.755 / 0.752 of peak at 6 / 8 / 12 (profiles/round2_k2p_tuning.txt)
#define K2P_LAUNCH(UU) (sd ? launch_k2p_inst<true, UU>(c, prm, grid, threads, smem) : launch_k2p_inst<false, UU>(c, prm, grid, threads, smem))
    if (unroll <= 4) rc = K2P_LAUNCH(4);
    else if (unroll <= 6) rc = K2P_LAUNCH(6);
    else if (unroll <= 8) rc = K2P_LAUNCH(8);
    else rc = K2P_LAUNCH(12);
#undef K2P_LAUNCH
    if (rc) return rc;
    launches += 1;
    c->last_counters[0] = launches;
    c->last_counters[1] = -1;   // resolved lazily from the device counter (see capi.cu)
    c->last_counters[2] = -1;
    return CPET_OK;
}

}  // namespace cpet
