// topo.cu -- K2: batched fixed-step streamline integrator with fused distance / curvature.
//
// Replaces, for a whole frame in ONE launch,
//   * the CPU path: Pool.starmap over task_complete_thread -> thread_operation
//     (CPET/source/calculator.py:675-712, CPET/utils/math_module.c:523-591, 489-503, 296-333)
//   * the torch path: compute_topo_GPU_batch_filter + CPET/utils/gpu.py:25-399
//     (path matrix windows, Python filter loops, torch.cat dumps).
//
// This file holds what the streamline kernels share -- the LPT queue sort (k2_count / scan / scatter), the per-launch
// classification and packing of the charges against the sampling box (k2x_extent / count / scatter), the dispatch
// (launch_topo) -- and two of the three integrator kernels:
//   k2w_topo_kernel  round 1, direct form (12 packed FMA-pipe instructions per two pair-evaluations), kept for A/B
//   k2x_topo_kernel  hybrid near/far form (10), two CHARGES per packed register, 4 lines per warp: short queues
// The default for long queues, k2p_topo_kernel (hybrid form, two POINTS per packed register, 8 lines per warp), is in
// topo8.cu.  All three are persistent kernels, one CTA per SM: a warp owns a few streamlines at a time, its 32
// lanes split the charges of the frame and every lane evaluates all of the warp's current points against each charge
// it loads.  One *pass* = the field at the warp's current points over all M charges, a transposing warp reduction,
// then the owner lanes advance the state machines of the lines (state in shared memory: current point, seed,
// previous unit field direction, step counters).  A finished line writes {distance, curvature} and the warp pulls the
// next line index from a global queue ordered by n_iter descending (longest-processing-time-first).
// The field at p_k is evaluated exactly once: K+2 evaluations per line (K = steps taken), versus
// K+4 in the reference C (the first two are recomputed there) -- the two look-ahead points at
// the end are simply the next two steps of the same integration.
//
// Charges: if the packed set fits in shared memory it is loaded once per CTA by TMA bulk copies and warps then run
// fully independently (no CTA barriers); otherwise tiles stream continuously through an S-stage TMA/mbarrier ring.
#include "cpet_internal.h"

namespace cpet {

#define K2_KEYMAX 2048
#ifndef CPET_K2_MAXT
#define CPET_K2_MAXT 512        // round-1 direct-form kernel (k2w): 16 warps x <= 128 registers
#endif
#ifndef CPET_K2X_MAXT
#define CPET_K2X_MAXT 384       // hybrid kernel (k2x): 12 warps x <= 168 registers -- at 512 threads ptxas needs 25
#endif                          // register-rotation MOVs per 160 packed instructions of the loop (10 at 384), measured
                                // 2.49e12 against 2.59e12 pair-evals/s on the 3A frame (profiles/round2_k2x.md)

// 1/sqrt(a) and sqrt(a) in double from one MUFU.RSQ seed + two Newton steps (relative error
// ~1e-15 after the second step) instead of the ~60-instruction software DSQRT/DDIV sequences: the
// per-round state update is executed by every lane and sits on the critical path of each round.
// Anything outside the comfortable float range (0, inf, NaN, denormal) takes the exact path, so the
// reference's unguarded E/|E| semantics (NaN when E = 0, C:501) are unchanged.
__device__ __forceinline__ double rsqrt_f64_fast(double a) {
    const float af = (float)a;
    if (af > 1e-30f && af < 1e30f) {
        double y = (double)rsqrt_approx(af);
        const double ha = 0.5 * a;
        y = y * (1.5 - ha * y * y);
        y = y * (1.5 - ha * y * y);
        return y;
    }
    return 1.0 / sqrt(a);
}
__device__ __forceinline__ double sqrt_f64_fast(double a) {
    const float af = (float)a;
    if (af > 1e-30f && af < 1e30f) return a * rsqrt_f64_fast(a);
    return sqrt(a);
}

// ---------------------------------------------------------------------------------------------
// Why warp-wide (measured, profiles/round1_k2w_ncu.md): the first version of this kernel gave every
// streamline its own G lanes ("slots", G = 1..32 by a host heuristic).  With G >= 4 distinct
// addresses per LDS.128 it saturated the shared-memory pipe, with G <= 2 a short queue left most
// lanes idle for the length of the last lines, and every slot re-read each charge pair for a single
// point.  Here lane utilisation does not depend on how many lines a warp holds (PE = 4, 2 or 1
// points are evaluated per pass, compacted), one pair of LDS.128 feeds 8 pair-evaluations, and the
// instruction stream is 86 % packed FMA-pipe work (79 % before).
// ---------------------------------------------------------------------------------------------
struct K2WParams {
    const ChargeBlock* blocks;
    int n_blocks;
    int tile_blocks;
    int stages;
    int ntiles;
    int resident;
    int cap;                // streamlines per warp (1, 2 or 4)
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

template <int PE>
__device__ __forceinline__ void evalw_pair(const PairA a, const PairB b, PointRegs<4>& r) {
#pragma unroll
    for (int p = 0; p < PE; ++p) {
        const u64 dx = add2(r.px[p], a.nx);
        const u64 dy = add2(r.py[p], a.ny);
        const u64 dz = add2(r.pz[p], b.nz);
        u64 r2 = mul2(dx, dx);
        r2 = fma2(dy, dy, r2);
        r2 = fma2(dz, dz, r2);
        float r2a, r2b;
        upk2(r2, r2a, r2b);
        const u64 inv = pk2(rsqrt_approx(r2a), rsqrt_approx(r2b));
        const u64 t = mul2(inv, inv);
        const u64 u = mul2(inv, b.q);
        const u64 s = mul2(t, u);
        r.ax[p] = fma2(s, dx, r.ax[p]);
        r.ay[p] = fma2(s, dy, r.ay[p]);
        r.az[p] = fma2(s, dz, r.az[p]);
    }
}

// blocks [0, nblk): lane `lane` takes pair `lane` of every block; FP32 partials are folded into
// the FP64 sums every CPET_K2W_CHUNK blocks (one FP32 chain is at most that many additions).
#ifndef CPET_K2W_CHUNK
#define CPET_K2W_CHUNK 128
#endif
template <int PE, int U>
__device__ __forceinline__ void evalw_range(const ChargeBlock* __restrict__ blk, int nblk, int lane,
                                            PointRegs<4>& r, double (&acc)[4][3]) {
    for (int b0 = 0; b0 < nblk; b0 += CPET_K2W_CHUNK) {
        const int b1 = min(nblk, b0 + CPET_K2W_CHUNK);
#pragma unroll U
        for (int b = b0; b < b1; ++b) {
            const PairA a = blk[b].a[lane];
            const PairB bb = blk[b].b[lane];
            evalw_pair<PE>(a, bb, r);
        }
#pragma unroll
        for (int p = 0; p < PE; ++p) {
            float lo, hi;
            upk2(r.ax[p], lo, hi);
            acc[p][0] += (double)(lo + hi);
            upk2(r.ay[p], lo, hi);
            acc[p][1] += (double)(lo + hi);
            upk2(r.az[p], lo, hi);
            acc[p][2] += (double)(lo + hi);
            r.ax[p] = r.ay[p] = r.az[p] = 0ull;
        }
    }
}

// One halving step of a transposing warp reduction: lanes whose `upper` flag is clear keep `a` and
// hand `b` to their partner (lane ^ m); the others keep `b` and hand over `a`.
__device__ __forceinline__ double xchg_add(double a, double b, int m, bool upper) {
    const double keep = upper ? b : a;
    const double send = upper ? a : b;
    return keep + shfl_xor_f64(send, m);
}

// Sum acc[p][0..2] (p < PE) over the 32 lanes.  On return (ex,ey,ez) in lane l holds the total of
// position l / (32/PE), i.e. PE = 4: lanes 8p..8p+7 hold position p.  90 / 57 / 45 instructions
// for PE = 4 / 2 / 1 instead of 45 per position for a plain butterfly.
template <int PE>
__device__ __forceinline__ void reduce_positions(const double (&acc)[4][3], int lane, double& ex,
                                                 double& ey, double& ez) {
    double v[3];
    if (PE == 4) {
        const bool up16 = (lane & 16) != 0, up8 = (lane & 8) != 0;
        double w[2][3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            w[0][c] = xchg_add(acc[0][c], acc[2][c], 16, up16);   // low half: positions 0,1; high: 2,3
            w[1][c] = xchg_add(acc[1][c], acc[3][c], 16, up16);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] = xchg_add(w[0][c], w[1][c], 8, up8);
#pragma unroll
        for (int m = 4; m >= 1; m >>= 1)
#pragma unroll
            for (int c = 0; c < 3; ++c) v[c] += shfl_xor_f64(v[c], m);
    } else if (PE == 2) {
        const bool up16 = (lane & 16) != 0;
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] = xchg_add(acc[0][c], acc[1][c], 16, up16);
#pragma unroll
        for (int m = 8; m >= 1; m >>= 1)
#pragma unroll
            for (int c = 0; c < 3; ++c) v[c] += shfl_xor_f64(v[c], m);
    } else {
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] = acc[0][c];
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1)
#pragma unroll
            for (int c = 0; c < 3; ++c) v[c] += shfl_xor_f64(v[c], m);
    }
    ex = v[0]; ey = v[1]; ez = v[2];
}

// State of the (up to) 4 streamlines of one warp, in shared memory so that none of it occupies
// registers during the charge loop.  Arrays of 4 = one entry per line slot.
struct __align__(16) WarpLines {
    float px[4], py[4], pz[4];         // current point p_k
    float sx[4], sy[4], sz[4];         // seed
    double ux[4], uy[4], uz[4];        // unit field direction at p_{k-1}
    double ex[4], ey[4], ez[4];        // field sums of the current pass
    float dist[4], kinit[4];
    int line[4], n_it[4], k[4], k_end[4];
    float m1x[4], m1y[4], m1z[4], m2x[4], m2y[4], m2z[4];   // p_{k-1}, p_{k-2} (second-difference mode)
};

template <bool SD, int U4, int U2>
__global__ void __launch_bounds__(CPET_K2_MAXT, 1) k2w_topo_kernel(const K2WParams prm) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
    WarpLines* lines_all = reinterpret_cast<WarpLines*>(smem_raw + 128);
    const int n_warps = blockDim.x >> 5;
    ChargeBlock* ring = reinterpret_cast<ChargeBlock*>(smem_raw + 128 + sizeof(WarpLines) * n_warps);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    WarpLines& W = lines_all[tid >> 5];
    const bool owner = lane < 4;       // lane q < 4 runs the state machine of line slot q
    const int S = prm.stages;
    const int TB = prm.tile_blocks;
    const int NT = prm.ntiles;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (owner) W.line[lane] = -1;
    __syncthreads();

    auto issue = [&](int it) {
        const int stage = it % S;
        const int t = it % NT;
        const int n_t = min(TB, prm.n_blocks - t * TB);
        const uint32_t bytes = (uint32_t)n_t * (uint32_t)sizeof(ChargeBlock);
        mbar_expect_tx(&full[stage], bytes);
        tma_load_1d(ring + (size_t)stage * TB, prm.blocks + (size_t)t * TB, bytes, &full[stage]);
    };
    int issued = 0;
    if (tid == 0) {
        const int pre = prm.resident ? NT : min(S, NT);
        for (; issued < pre; ++issued) issue(issued);
    }
    if (prm.resident) {
        for (int t = 0; t < NT; ++t) mbar_wait(&full[t], 0u);
    }

    bool exhausted = false;
    unsigned long long my_evals = 0ull;
    const double h = (double)prm.h;
    const double inv_h = 1.0 / h;

    int it = 0;   // consumed-tile counter (streamed mode)
    while (true) {
        // ---- refill the warp's empty line slots from the queue (one atomic per warp) ------------
        {
            const bool empty = owner && W.line[lane] < 0;
            const unsigned em = __ballot_sync(0xffffffffu, empty);
            const int n_empty = __popc(em);
            int want = min(n_empty, prm.cap - (4 - n_empty));
            if (exhausted) want = 0;
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
                    W.ux[lane] = W.uy[lane] = W.uz[lane] = 0.0;
                }
                if (base + (unsigned)want >= (unsigned)prm.n_lines) exhausted = true;
            }
        }
        __syncwarp();     // the owners' stores above must be visible to every lane's gather below
        const unsigned am = __ballot_sync(0xffffffffu, owner && W.line[lane] >= 0);
        bool go;
        if (prm.resident) go = (am != 0u);
        else go = __syncthreads_or(am != 0u ? 1 : 0) != 0;
        if (!go) break;

        // ---- the warp's current points, compacted to positions 0..na-1 ----------------------------
        const int na = __popc(am);
        int src[4];
        PointRegs<4> r;
        double acc[4][3];
        {
            unsigned m = am;
            const int first = am ? (__ffs(am) - 1) : 0;
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                src[p] = m ? (__ffs(m) - 1) : first;           // unused positions repeat a valid point
                m &= m - 1u;
                set_point<4>(r, p, W.px[src[p]], W.py[src[p]], W.pz[src[p]]);
                acc[p][0] = acc[p][1] = acc[p][2] = 0.0;
            }
            clear_partials<4>(r);
        }

        // ---- field at those points: all charges, split over the 32 lanes -----------------------------
        if (prm.resident) {
            if (na > 2) evalw_range<4, U4>(ring, prm.n_blocks, lane, r, acc);
            else if (na == 2) evalw_range<2, U2>(ring, prm.n_blocks, lane, r, acc);
            else evalw_range<1, 8>(ring, prm.n_blocks, lane, r, acc);
        } else {
            for (int t = 0; t < NT; ++t, ++it) {
                const int stage = it % S;
                mbar_wait(&full[stage], (uint32_t)((it / S) & 1));
                const int n_t = min(TB, prm.n_blocks - t * TB);
                const ChargeBlock* tile = ring + (size_t)stage * TB;
                if (na > 2) evalw_range<4, U4>(tile, n_t, lane, r, acc);
                else if (na == 2) evalw_range<2, U2>(tile, n_t, lane, r, acc);
                else if (na == 1) evalw_range<1, 8>(tile, n_t, lane, r, acc);
                __syncthreads();                      // stage fully consumed by the CTA
                if (tid == 0) { issue(issued); ++issued; }   // speculative: next pass's tiles too
            }
        }

        // ---- warp reduction of the FP64 sums; the total of position p goes to line slot src[p] ---
        {
            double ex, ey, ez;
            int pos;                                           // position whose total this lane holds
            if (na > 2) { reduce_positions<4>(acc, lane, ex, ey, ez); pos = lane >> 3; }
            else if (na == 2) { reduce_positions<2>(acc, lane, ex, ey, ez); pos = lane >> 4; }
            else { reduce_positions<1>(acc, lane, ex, ey, ez); pos = 0; }
            const int slot = pos == 0 ? src[0] : (pos == 1 ? src[1] : (pos == 2 ? src[2] : src[3]));
            if ((lane & 7) == 0 && pos < na && (na > 2 || (lane & 15) == 0) && (na > 1 || lane == 0)) {
                W.ex[slot] = ex; W.ey[slot] = ey; W.ez[slot] = ez;
            }
        }
        __syncwarp();

        // ---- state machine of line slot `lane` ----------------------------------------------------------
        if (owner && ((am >> lane) & 1u)) {
            const int q = lane;
            const double ex = W.ex[q], ey = W.ey[q], ez = W.ez[q];
            const float px = W.px[q], py = W.py[q], pz = W.pz[q];
            int k = W.k[q];
            const int k_end = W.k_end[q];
            // unit direction (no zero guard: E = 0 gives NaN exactly like C:501)
            const double inv_n = rsqrt_f64_fast(ex * ex + ey * ey + ez * ez);
            const double ux = ex * inv_n, uy = ey * inv_n, uz = ez * inv_n;
            const bool last = (k_end >= 0) && (k == k_end + 1);
            float kdir = 0.f;
            if (!SD && (k == 1 || last)) {   // curvature is needed at the first and last point pair
                const double pux = W.ux[q], puy = W.uy[q], puz = W.uz[q];
                const double cx = puy * uz - puz * uy;
                const double cy = puz * ux - pux * uz;
                const double cz = pux * uy - puy * ux;
                kdir = (float)(sqrt_f64_fast(cx * cx + cy * cy + cz * cz) * inv_h);
                if (k == 1) W.kinit[q] = kdir;
            }
            const float nx = (float)((double)px + h * ux);
            const float ny = (float)((double)py + h * uy);
            const float nz = (float)((double)pz + h * uz);
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
                        const double ddx = (double)W.sx[q] - (double)nx;
                        const double ddy = (double)W.sy[q] - (double)ny;
                        const double ddz = (double)W.sz[q] - (double)nz;
                        W.dist[q] = (float)sqrt_f64_fast(ddx * ddx + ddy * ddy + ddz * ddz);
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

// =============================================================================================
// Round 2: hybrid near/far kernel (k2x).  Same warp-wide organisation as k2w above; what changes is
// the arithmetic of the charge loop (common.cuh: evalx_far / evalx_near) and, because that needs the
// charges classified against the sampling box, a per-launch packing step:
//   k2x_extent_kernel   max |seed| per axis (seeds may lie outside the box the caller names)
//   k2x_count_kernel    class of every charge (near / far), counts per chunk
//   k2x_scatter_kernel  stable compaction into XBlocks [near | far], zero-weight padding
// All three are O(M + L) and run on the launch stream; nothing is read back by the host (the
// integrator takes the block counts from device memory).
// =============================================================================================
#define K2X_CHUNK 1024               // charges per CTA of the packers

__global__ void __launch_bounds__(256) k2x_extent_kernel(const float* __restrict__ seeds, int n_lines,
                                                         K2XMeta* __restrict__ meta) {
    __shared__ unsigned s_m[8][3];
    unsigned m[3] = {0u, 0u, 0u};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_lines; i += gridDim.x * blockDim.x) {
#pragma unroll
        for (int c = 0; c < 3; ++c) m[c] = max(m[c], __float_as_uint(fabsf(seeds[3 * (size_t)i + c])));
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const unsigned w = __reduce_max_sync(0xffffffffu, m[c]);
        if ((threadIdx.x & 31) == 0) s_m[threadIdx.x >> 5][c] = w;
    }
    __syncthreads();
    if (threadIdx.x < 3) {                   // one atomic per block and axis (bit-wise max: |x| >= 0, NaN sorts last)
        unsigned w = 0u;
        for (int k = 0; k < 8; ++k) w = max(w, s_m[k][threadIdx.x]);
        if (w) atomicMax(&meta->ext[threadIdx.x], w);
    }
}

struct K2XBox { float bx, by, bz, pmax, amax; };

// region every point of every line stays in: max(box, seeds) inflated by three steps (the exiting
// point and the two look-ahead points, C:565-568).  NaN extents propagate (bit-wise max), which
// classifies every charge as near.
__device__ __forceinline__ K2XBox k2x_box(const K2XMeta* meta, float dx, float dy, float dz, float h, float amax) {
    K2XBox b;
    const float m = 3.0f * fabsf(h) * 1.0001f + 1e-6f;
    b.bx = __uint_as_float(max(__float_as_uint(fabsf(dx)), meta->ext[0])) + m;
    b.by = __uint_as_float(max(__float_as_uint(fabsf(dy)), meta->ext[1])) + m;
    b.bz = __uint_as_float(max(__float_as_uint(fabsf(dz)), meta->ext[2])) + m;
    b.pmax = sqrtf(b.bx * b.bx + b.by * b.by + b.bz * b.bz);
    b.amax = amax;
    return b;
}

// 0 = near (direct form), 1 = far (expanded form).  A charge is far when the rounding
// amplification of the expanded r^2, (|x| + |p|max)^2 / dist(x, region)^2, is at most amax.
__device__ __forceinline__ int k2x_class(float x, float y, float z, float q, const K2XBox& b) {
    const float ex = fmaxf(fabsf(x) - b.bx, 0.f), ey = fmaxf(fabsf(y) - b.by, 0.f), ez = fmaxf(fabsf(z) - b.bz, 0.f);
    const float r2min = ex * ex + ey * ey + ez * ez;
    const float xn = sqrtf(x * x + y * y + z * z) + b.pmax;
    // (r2min >= 1e-5: a far charge can never come within the 1e-3 A where the volume path's softening acts)
    const bool far = (xn * xn <= b.amax * r2min) && (xn < 1.0e6f) && (r2min >= 1.0e-5f);      // false for NaN
    if (!far) return 0;
    if (q == 0.f) return 1;                                            // zero-weight far record
    if (!(fabsf(q) >= CPET_X_MIN_ABS_Q)) return 0;                     // tiny or NaN charge
    return 1;
}

__device__ __forceinline__ void k2x_load_charge(const ChargePair* __restrict__ pairs, int i, float& x, float& y,
                                                float& z, float& q) {
    const float* f = reinterpret_cast<const float*>(pairs + (i >> 1));   // {-x0,-x1,-y0,-y1,-z0,-z1,q0,q1}
    const int h = i & 1;
    x = -f[h]; y = -f[2 + h]; z = -f[4 + h]; q = f[6 + h];
}

__global__ void __launch_bounds__(K2X_CHUNK) k2x_count_kernel(const ChargePair* __restrict__ pairs, int n_charges,
                                                              float dx, float dy, float dz, float h, float amax,
                                                              const K2XMeta* __restrict__ meta,
                                                              int* __restrict__ chunk_counts) {
    __shared__ int cnt[2];
    if (threadIdx.x < 2) cnt[threadIdx.x] = 0;
    __syncthreads();
    const K2XBox box = k2x_box(meta, dx, dy, dz, h, amax);
    const int i = blockIdx.x * K2X_CHUNK + threadIdx.x;
    int cls = -1;
    if (i < n_charges) {
        float x, y, z, q;
        k2x_load_charge(pairs, i, x, y, z, q);
        cls = k2x_class(x, y, z, q, box);
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const unsigned bal = __ballot_sync(0xffffffffu, cls == k);
        if ((threadIdx.x & 31) == 0 && bal) atomicAdd(&cnt[k], __popc(bal));
    }
    __syncthreads();
    if (threadIdx.x < 2) chunk_counts[2 * blockIdx.x + threadIdx.x] = cnt[threadIdx.x];
}

__device__ __forceinline__ void k2x_store(XBlock* __restrict__ out, int block, int slot, float f0, float f1,
                                          float f2, float f3, float f4) {
    const int l = (slot >> 1) & 31, hf = slot & 1;
    float* v0 = reinterpret_cast<float*>(&out[block].v0[l]);
    float* v1 = reinterpret_cast<float*>(&out[block].v1[l]);
    float* v2 = reinterpret_cast<float*>(&out[block].v2[l]);
    v0[hf] = f0; v0[2 + hf] = f1; v1[hf] = f2; v1[2 + hf] = f3; v2[hf] = f4;
}
__device__ __forceinline__ void k2x_store_class(XBlock* __restrict__ out, int cls, int first_block, int slot,
                                                float x, float y, float z, float q, bool pad) {
    const int block = first_block + (slot >> 6);
    if (cls == 0) {
        if (pad) k2x_store(out, block, slot, 2.0f * CPET_PAD_COORD, 2.0f * CPET_PAD_COORD, 2.0f * CPET_PAD_COORD, 0.f, 0.f);
        else k2x_store(out, block, slot, 2.0f * x, 2.0f * y, 2.0f * z, 4.0f * q, 0.f);
    } else if (pad || q == 0.f) {
        k2x_store(out, block, slot, 0.f, 0.f, 0.f, CPET_X_PAD_B, 0.f);
    } else {
        const double al = (q < 0.f ? -1.0 : 1.0) / ((double)q * (double)q);   // the record carries the sign of q
        const double x2 = (double)x * x + (double)y * y + (double)z * z;
        k2x_store(out, block, slot, (float)(al * x), (float)(al * y), (float)(al * z), (float)(al * x2), (float)al);
    }
}

// The same records in the layout of the points-packed kernel (topo8.cu): 32 charges per PBlock,
//   far : a = {alpha x, alpha y, alpha z, alpha}   b = alpha |x|^2
//   near: a = {2x, 2y, 2z, 4q}                     b unused
__device__ __forceinline__ void k2p_store_class(PBlock* __restrict__ out, int cls, int first_block, int slot,
                                                float x, float y, float z, float q, bool pad) {
    PBlock& blk = out[first_block + (slot >> 5)];
    const int l = slot & 31;
    if (cls == 0) {
        blk.a[l] = pad ? make_float4(2.0f * CPET_PAD_COORD, 2.0f * CPET_PAD_COORD, 2.0f * CPET_PAD_COORD, 0.f)
                       : make_float4(2.0f * x, 2.0f * y, 2.0f * z, 4.0f * q);
        blk.b[l] = 0.f;
    } else if (pad || q == 0.f) {
        blk.a[l] = make_float4(0.f, 0.f, 0.f, 0.f);
        blk.b[l] = CPET_X_PAD_B;
    } else {
        const double al = (q < 0.f ? -1.0 : 1.0) / ((double)q * (double)q);   // the record carries the sign of q
        const double x2 = (double)x * x + (double)y * y + (double)z * z;
        blk.a[l] = make_float4((float)(al * x), (float)(al * y), (float)(al * z), (float)al);
        blk.b[l] = (float)(al * x2);
    }
}

template <int LAYOUT>      // 0: XBlock (64 charges, charge pairs packed), 1: PBlock (32 charges, one per lane)
__global__ void __launch_bounds__(K2X_CHUNK) k2x_scatter_kernel(const ChargePair* __restrict__ pairs, int n_charges,
                                                                float dx, float dy, float dz, float h, float amax,
                                                                K2XMeta* __restrict__ meta,
                                                                const int* __restrict__ chunk_counts, int n_chunks,
                                                                void* __restrict__ out_raw) {
    constexpr int SH = LAYOUT == 0 ? 6 : 5;          // log2(charges per block)
    XBlock* out = reinterpret_cast<XBlock*>(out_raw);
    PBlock* outp = reinterpret_cast<PBlock*>(out_raw);
    __shared__ int s_red[32][4];
    __shared__ int s_pre[2], s_tot[2];
    __shared__ int s_warp[32][2];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    // charges of each class in the chunks before this one, and in all chunks
    int v[4] = {0, 0, 0, 0};
    for (int c = tid; c < n_chunks; c += K2X_CHUNK) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int n = chunk_counts[2 * c + k];
            v[2 + k] += n;
            if (c < (int)blockIdx.x) v[k] += n;
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        v[k] = __reduce_add_sync(0xffffffffu, v[k]);
        if (lane == 0) s_red[w][k] = v[k];
    }
    const K2XBox box = k2x_box(meta, dx, dy, dz, h, amax);
    const int i = blockIdx.x * K2X_CHUNK + tid;
    int cls = -1;
    float x = 0.f, y = 0.f, z = 0.f, q = 0.f;
    if (i < n_charges) {
        k2x_load_charge(pairs, i, x, y, z, q);
        cls = k2x_class(x, y, z, q, box);
    }
    int rank = 0;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const unsigned bal = __ballot_sync(0xffffffffu, cls == k);
        if (cls == k) rank = __popc(bal & ((1u << lane) - 1u));
        if (lane == 0) s_warp[w][k] = __popc(bal);
    }
    __syncthreads();
    if (tid < 4) {
        int t = 0;
        for (int j = 0; j < 32; ++j) t += s_red[j][tid];
        if (tid < 2) s_pre[tid] = t; else s_tot[tid - 2] = t;
    }
    if (tid >= 32 && tid < 34) {          // exclusive scan of the per-warp counts of class tid-32
        const int k = tid - 32;
        int run = 0;
        for (int j = 0; j < 32; ++j) { const int n = s_warp[j][k]; s_warp[j][k] = run; run += n; }
    }
    __syncthreads();
    const int nb0 = (s_tot[0] + (1 << SH) - 1) >> SH, nb1 = (s_tot[1] + (1 << SH) - 1) >> SH;
    const int first[2] = {0, nb0};
    if (cls >= 0) {
        const int slot = s_pre[cls] + s_warp[w][cls] + rank;
        if (LAYOUT == 0) k2x_store_class(out, cls, first[cls], slot, x, y, z, q, false);
        else k2p_store_class(outp, cls, first[cls], slot, x, y, z, q, false);
    }
    if (blockIdx.x == 0) {
        // zero-weight padding up to a whole block per class, and the counts the integrator reads
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int nb = k == 0 ? nb0 : nb1;
            for (int slot = s_tot[k] + tid; slot < (nb << SH); slot += K2X_CHUNK) {
                if (LAYOUT == 0) k2x_store_class(out, k, first[k], slot, 0.f, 0.f, 0.f, 0.f, true);
                else k2p_store_class(outp, k, first[k], slot, 0.f, 0.f, 0.f, 0.f, true);
            }
        }
        if (tid == 0) {
            meta->n_near = s_tot[0]; meta->n_far = s_tot[1];
            meta->nb_near = nb0; meta->nb_far = nb1;
            meta->nb_total = nb0 + nb1;
        }
    }
}

struct K2XParams {
    const XBlock* blocks;
    const K2XMeta* meta;
    int tile_blocks;
    int stages;
    int resident;
    int cap;                // streamlines per warp (1, 2 or 4)
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

struct __align__(16) WarpLinesX {
    float px[4], py[4], pz[4];         // current point p_k
    float sx[4], sy[4], sz[4];         // seed
    float ux[4], uy[4], uz[4];         // unit field direction at p_{k-1}
    double tx[4], ty[4], tz[4], ts[4]; // sums of the current pass: T - E_near (3) and S (signed by the charges)
    float dist[4], kinit[4];
    int line[4], n_it[4], k[4], k_end[4];
    float m1x[4], m1y[4], m1z[4], m2x[4], m2y[4], m2z[4];   // p_{k-1}, p_{k-2} (second-difference mode)
};

// FP32 partials of the warp's PE points, immediately through the halving steps of the transposing
// warp reduction (lane distance 16, then 8; FP32: at most four lane partials meet here), the survivor
// added in FP64 to v[0..3] = (T.x, T.y, T.z, S) of the position this lane ends up with (PE = 4:
// position lane >> 3).  Keeping only these 4 doubles alive across the charge loop (instead of 16
// per-position sums) is what leaves ptxas the registers to software-pipeline the loop.
__device__ __forceinline__ float xchg_add_f32(float a, float b, int m, bool upper) {
    const float keep = upper ? b : a;
    const float send = upper ? a : b;
    return keep + __shfl_xor_sync(0xffffffffu, send, m);
}
template <int PE>
__device__ __forceinline__ void fold_exchange(XRegs<4>& r, double (&v)[4], int lane) {
    float d[PE][4];
#pragma unroll
    for (int p = 0; p < PE; ++p) {
        float lo, hi;
        upk2(r.a0[p], lo, hi); d[p][0] = lo + hi;
        upk2(r.a1[p], lo, hi); d[p][1] = lo + hi;
        upk2(r.a2[p], lo, hi); d[p][2] = lo + hi;
        upk2(r.a3[p], lo, hi); d[p][3] = lo + hi;
        r.a0[p] = r.a1[p] = r.a2[p] = r.a3[p] = 0ull;
    }
    if (PE == 4) {
        const bool up16 = (lane & 16) != 0, up8 = (lane & 8) != 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float w0 = xchg_add_f32(d[0][c], d[2 % PE][c], 16, up16);   // low half: positions 0,1; high: 2,3
            const float w1 = xchg_add_f32(d[1 % PE][c], d[3 % PE][c], 16, up16);
            v[c] += (double)xchg_add_f32(w0, w1, 8, up8);
        }
    } else if (PE == 2) {
        // same additions as PE = 4 -- (l + l^16) + (l^8 + l^24) in FP32 -- so that a line's sums do not depend
        // on how many lines its warp holds (FP32 addition is commutative: who keeps which half is irrelevant)
        const bool up16 = (lane & 16) != 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float w = xchg_add_f32(d[0][c], d[1 % PE][c], 16, up16);
            v[c] += (double)(w + __shfl_xor_sync(0xffffffffu, w, 8));
        }
    } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float w = d[0][c] + __shfl_xor_sync(0xffffffffu, d[0][c], 16);
            v[c] += (double)(w + __shfl_xor_sync(0xffffffffu, w, 8));
        }
    }
}
// the remaining FP64 butterfly over lane distances 4, 2, 1 (for PE < 4 the lanes at distance 8 / 16 hold
// copies of the same sums)
template <int PE>
__device__ __forceinline__ void finish_exchange(double (&v)[4]) {
#pragma unroll
    for (int m = 4; m >= 1; m >>= 1)
#pragma unroll
        for (int c = 0; c < 4; ++c) v[c] += shfl_xor_f64(v[c], m);
}

// n blocks starting at shared-memory byte address of block 0's lane entries: p16 -> v0[lane]
// (v1[lane] is 512 bytes further), p8 -> v2[lane].  FP32 chains are cut every CPET_K2W_CHUNK blocks.
template <int PE, int U, bool NEAR>
__device__ __forceinline__ void evalx_run(const unsigned char* __restrict__ p16, const unsigned char* __restrict__ p8,
                                          int n, int lane, XRegs<4>& r, double (&v)[4], int& run) {
    while (n > 0) {
        int m = CPET_K2W_CHUNK - run;
        if (m > n) m = n;
        if (NEAR) {
#pragma unroll 1
            for (int j = 0; j < m; ++j, p16 += sizeof(XBlock))
                evalx_near<PE, 4>(*reinterpret_cast<const V16*>(p16), *reinterpret_cast<const V16*>(p16 + 512), r);
        } else {
#pragma unroll U
            for (int j = 0; j < m; ++j, p16 += sizeof(XBlock), p8 += sizeof(XBlock))
                evalx_far<PE, 4>(*reinterpret_cast<const V16*>(p16), *reinterpret_cast<const V16*>(p16 + 512),
                                 *reinterpret_cast<const u64*>(p8), r);
        }
        n -= m;
        run += m;
        if (run >= CPET_K2W_CHUNK) { fold_exchange<PE>(r, v, lane); run = 0; }
    }
}

// Blocks [g0, g1) of the array [near | far] (tile[0] is block gbase) against the warp's PE points.
template <int PE, int U>
__device__ __forceinline__ void evalx_blocks(const XBlock* __restrict__ tile, int gbase, int g0, int g1,
                                             int nb_near, int lane, XRegs<4>& r, double (&v)[4], int& run) {
    const unsigned char* base = reinterpret_cast<const unsigned char*>(tile) - (size_t)gbase * sizeof(XBlock);
    const unsigned char* l16 = base + 16 * lane;
    const unsigned char* l8 = base + 1024 + 8 * lane;
    int b = g0;
    if (b < nb_near && b < g1) {
        const int e = min(g1, nb_near);
        evalx_run<PE, 1, true>(l16 + (size_t)b * sizeof(XBlock), l8, e - b, lane, r, v, run);
        b = e;
    }
    if (b < g1)
        evalx_run<PE, U, false>(l16 + (size_t)b * sizeof(XBlock), l8 + (size_t)b * sizeof(XBlock), g1 - b, lane, r, v, run);
}

template <int PE>
__device__ __forceinline__ void finish_pass(XRegs<4>& r, double (&v)[4], int lane) {
    fold_exchange<PE>(r, v, lane);
    finish_exchange<PE>(v);
}

template <bool SD, int U4, int U2>
__global__ void __launch_bounds__(CPET_K2X_MAXT, 1) k2x_topo_kernel(const K2XParams prm) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
    WarpLinesX* lines_all = reinterpret_cast<WarpLinesX*>(smem_raw + 128);
    const int n_warps = blockDim.x >> 5;
    XBlock* ring = reinterpret_cast<XBlock*>(smem_raw + 128 + sizeof(WarpLinesX) * n_warps);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    WarpLinesX& W = lines_all[tid >> 5];
    const bool owner = lane < 4;       // lane q < 4 runs the state machine of line slot q
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
        const uint32_t bytes = (uint32_t)n_t * (uint32_t)sizeof(XBlock);
        mbar_expect_tx(&full[stage], bytes);
        tma_load_1d(ring + (size_t)stage * TB, prm.blocks + (size_t)t * TB, bytes, &full[stage]);
    };
    int issued = 0;
    if (prm.resident) {
        if (tid == 0 && nb_total > 0) {        // the whole frame once: bulk copies of <= 32 blocks on one barrier
            mbar_expect_tx(&full[0], (uint32_t)nb_total * (uint32_t)sizeof(XBlock));
            for (int b = 0; b < nb_total; b += 32) {
                const int n = min(32, nb_total - b);
                tma_load_1d(ring + b, prm.blocks + b, (uint32_t)n * (uint32_t)sizeof(XBlock), &full[0]);
            }
        }
        if (nb_total > 0) mbar_wait(&full[0], 0u);
    } else if (tid == 0) {
        const int pre = min(S, NT);
        for (; issued < pre; ++issued) issue(issued);
    }

    bool exhausted = false;
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
            int want = min(n_empty, prm.cap - (4 - n_empty));
            if (exhausted) want = 0;
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
            }
        }
        __syncwarp();     // the owners' stores above must be visible to every lane's gather below
        const unsigned am = __ballot_sync(0xffffffffu, owner && W.line[lane] >= 0);
        bool go;
        if (prm.resident) go = (am != 0u);
        else go = __syncthreads_or(am != 0u ? 1 : 0) != 0;
        if (!go) break;

        // ---- the warp's current points, compacted to positions 0..na-1 ----------------------------
        const int na = __popc(am);
        int src[4];
        XRegs<4> r;
        double v[4] = {0.0, 0.0, 0.0, 0.0};
        {
            unsigned m = am;
            const int first = am ? (__ffs(am) - 1) : 0;
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                src[p] = m ? (__ffs(m) - 1) : first;           // unused positions repeat a valid point
                m &= m - 1u;
                set_point_x<4>(r, p, W.px[src[p]], W.py[src[p]], W.pz[src[p]]);
                r.a0[p] = r.a1[p] = r.a2[p] = r.a3[p] = 0ull;
            }
        }

        // ---- field sums at those points: all charges, split over the 32 lanes -------------------------
        int run = 0;
        if (prm.resident) {
            if (na > 2) { evalx_blocks<4, U4>(ring, 0, 0, nb_total, nb_near, lane, r, v, run); finish_pass<4>(r, v, lane); }
            else if (na == 2) { evalx_blocks<2, U2>(ring, 0, 0, nb_total, nb_near, lane, r, v, run); finish_pass<2>(r, v, lane); }
            else { evalx_blocks<1, 8>(ring, 0, 0, nb_total, nb_near, lane, r, v, run); finish_pass<1>(r, v, lane); }
        } else {
            for (int t = 0; t < NT; ++t, ++it) {
                const int stage = it % S;
                mbar_wait(&full[stage], (uint32_t)((it / S) & 1));
                const int g0 = t * TB, g1 = min(nb_total, g0 + TB);
                const XBlock* tile = ring + (size_t)stage * TB;
                if (na > 2) evalx_blocks<4, U4>(tile, g0, g0, g1, nb_near, lane, r, v, run);
                else if (na == 2) evalx_blocks<2, U2>(tile, g0, g0, g1, nb_near, lane, r, v, run);
                else if (na == 1) evalx_blocks<1, 8>(tile, g0, g0, g1, nb_near, lane, r, v, run);
                __syncthreads();                      // stage fully consumed by the CTA
                if (tid == 0) { issue(issued); ++issued; }   // speculative: next pass's tiles too
            }
            if (na > 2) finish_pass<4>(r, v, lane);
            else if (na == 2) finish_pass<2>(r, v, lane);
            else finish_pass<1>(r, v, lane);
        }

        // ---- the totals of position p go to line slot src[p] ---------------------------------------------
        {
            const int pos = na > 2 ? (lane >> 3) : (na == 2 ? (lane >> 4) : 0);   // position whose total this lane holds
            const int slot = pos == 0 ? src[0] : (pos == 1 ? src[1] : (pos == 2 ? src[2] : src[3]));
            if ((lane & 7) == 0 && pos < na && (na > 2 || (lane & 15) == 0) && (na > 1 || lane == 0)) {
                W.tx[slot] = v[0]; W.ty[slot] = v[1]; W.tz[slot] = v[2]; W.ts[slot] = v[3];
            }
        }
        __syncwarp();

        // ---- state machine of line slot `lane` (FP32, like the reference's own float arithmetic, C:489-503;
        //      the sums arrive in FP64 and E = p*S - T is formed there) ------------------------------------
        if (owner && ((am >> lane) & 1u)) {
            const int q = lane;
            const float px = W.px[q], py = W.py[q], pz = W.pz[q];
            // E = p*S - (T - E_near), up to the Coulomb constant (only the direction is used)
            const double ss = W.ts[q];
            const float ex = (float)((double)px * ss - W.tx[q]);
            const float ey = (float)((double)py * ss - W.ty[q]);
            const float ez = (float)((double)pz * ss - W.tz[q]);
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

// ---------------------------------------------------------------------------------------------
// queue ordering: counting sort of line ids by n_iter, descending (LPT), block-aggregated.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int k2_key(int n_iter) {
    return n_iter < 0 ? 0 : (n_iter >= K2_KEYMAX ? K2_KEYMAX - 1 : n_iter);
}

__global__ void __launch_bounds__(256) k2_count_kernel(const int32_t* __restrict__ n_iter, int n,
                                                       unsigned* __restrict__ hist) {
    __shared__ unsigned cnt[K2_KEYMAX];
    for (int i = threadIdx.x; i < K2_KEYMAX; i += blockDim.x) cnt[i] = 0u;
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        atomicAdd(&cnt[k2_key(n_iter[i])], 1u);
    __syncthreads();
    for (int i = threadIdx.x; i < K2_KEYMAX; i += blockDim.x)
        if (cnt[i]) atomicAdd(&hist[i], cnt[i]);
}

// offsets[key] = number of lines with a LARGER key (descending order); one block of 1024 threads,
// two keys per thread, Hillis-Steele inclusive scan over the per-thread sums.
__global__ void __launch_bounds__(1024) k2_scan_kernel(const unsigned* __restrict__ hist,
                                                       unsigned* __restrict__ offsets) {
    static_assert(K2_KEYMAX == 2048, "scan assumes 2 keys per thread");
    __shared__ unsigned s[1024];
    const int t = threadIdx.x;
    // reversed order: position i holds key K2_KEYMAX-1-i
    const unsigned a = hist[K2_KEYMAX - 1 - 2 * t];
    const unsigned b = hist[K2_KEYMAX - 2 - 2 * t];
    s[t] = a + b;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        const unsigned v = (t >= d) ? s[t - d] : 0u;
        __syncthreads();
        s[t] += v;
        __syncthreads();
    }
    const unsigned excl = s[t] - (a + b);
    offsets[K2_KEYMAX - 1 - 2 * t] = excl;
    offsets[K2_KEYMAX - 2 - 2 * t] = excl + a;
}

__global__ void __launch_bounds__(256) k2_scatter_kernel(const int32_t* __restrict__ n_iter, int n,
                                                         const unsigned* __restrict__ offsets,
                                                         unsigned* __restrict__ cursor,
                                                         int32_t* __restrict__ order) {
    __shared__ unsigned cnt[K2_KEYMAX];
    __shared__ unsigned base[K2_KEYMAX];
    for (int start = blockIdx.x * blockDim.x; start < n; start += gridDim.x * blockDim.x) {
        for (int i = threadIdx.x; i < K2_KEYMAX; i += blockDim.x) cnt[i] = 0u;
        __syncthreads();
        const int i = start + threadIdx.x;
        int key = 0;
        unsigned rank = 0;
        if (i < n) { key = k2_key(n_iter[i]); rank = atomicAdd(&cnt[key], 1u); }
        __syncthreads();
        for (int b = threadIdx.x; b < K2_KEYMAX; b += blockDim.x)
            if (cnt[b]) base[b] = offsets[b] + atomicAdd(&cursor[b], cnt[b]);
        __syncthreads();
        if (i < n) order[base[key] + rank] = i;
        __syncthreads();
    }
}

// queue order (LPT): line ids sorted by n_iter descending into c->work1; leaves `order` null when
// the sort is skipped.  Also zeroes the queue cursor / evaluation counter block.
int prepare_queue(cpet_ctx* c, int n_lines, const int32_t* d_n_iter, bool do_sort,
                         unsigned int** queue, unsigned long long** evals, const int32_t** order,
                         int* launches) {
    const int sms = c->sm_count;
    if (int rc = c->counters.reserve(64 + sizeof(unsigned) * 3 * K2_KEYMAX)) return rc;
    unsigned char* cb = c->counters.as<unsigned char>();
    CPET_CUDA_TRY(cudaMemsetAsync(cb, 0, 64 + sizeof(unsigned) * 3 * K2_KEYMAX, c->stream));
    *queue = reinterpret_cast<unsigned int*>(cb);
    *evals = reinterpret_cast<unsigned long long*>(cb + 8);
    unsigned* hist = reinterpret_cast<unsigned*>(cb + 64);
    unsigned* offsets = hist + K2_KEYMAX;
    unsigned* cursor = offsets + K2_KEYMAX;
    *order = nullptr;
    if (do_sort) {
        if (int rc = c->work1.reserve(sizeof(int32_t) * (size_t)n_lines)) return rc;
        int blocks = (n_lines + 255) / 256;
        if (blocks > sms * 8) blocks = sms * 8;
        k2_count_kernel<<<blocks, 256, 0, c->stream>>>(d_n_iter, n_lines, hist);
        k2_scan_kernel<<<1, 1024, 0, c->stream>>>(hist, offsets);
        k2_scatter_kernel<<<blocks, 256, 0, c->stream>>>(d_n_iter, n_lines, offsets, cursor,
                                                         c->work1.as<int32_t>());
        CPET_CUDA_TRY(cudaGetLastError());
        *order = c->work1.as<int32_t>();
        *launches += 3;
    }
    return CPET_OK;
}

template <bool SD, int U4, int U2>
static int launch_k2w_inst(cpet_ctx* c, const K2WParams& prm, int grid, int threads, size_t smem) {
    auto kern = k2w_topo_kernel<SD, U4, U2>;
    CPET_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, threads, smem, c->stream>>>(prm);
    CPET_CUDA_TRY(cudaGetLastError());
    return CPET_OK;
}

static int launch_topo_warpwide(cpet_ctx* c, int n_lines, const float* d_seeds, const int32_t* d_n_iter,
                                float step, const float dims[3], unsigned flags, float* d_out,
                                int32_t* d_steps) {
    const Tuning& tu = c->tune;
    const int sms = c->sm_count;
    int launches = 0;

    // --- warps per CTA ---------------------------------------------------------------------------------
    int threads = tu.k2_threads > 0 ? tu.k2_threads : 512;
    threads = (threads / 32) * 32;
    if (threads < 32) threads = 32;
    if (threads > CPET_K2_MAXT) threads = CPET_K2_MAXT;
    if (tu.k2_threads <= 0 && (long long)n_lines < (long long)sms * (threads / 32)) {
        // fewer lines than warps on the chip: one line per warp, spread over all SMs
        threads = 32 * ((n_lines + sms - 1) / sms);
    }
    const int warps_per_cta = threads / 32;
    const size_t hdr = 128 + sizeof(WarpLines) * (size_t)warps_per_cta;   // barriers + per-warp line state

    // --- charge staging plan: whole frame resident in shared memory when it fits ------------------
    const int n_blocks = (c->n_pairs + 31) / 32;
    const size_t all_bytes = (size_t)n_blocks * sizeof(ChargeBlock);
    K2WParams prm;
    prm.blocks = c->charge_blocks.as<ChargeBlock>();
    prm.n_blocks = n_blocks;
    size_t smem;
    if (hdr + all_bytes <= (size_t)c->max_smem_optin && tu.k2_stages <= 0 && tu.k2_tile_pairs <= 0) {
        prm.resident = 1;
        prm.tile_blocks = 32;                        // 32 KB per bulk copy
        prm.ntiles = (n_blocks + prm.tile_blocks - 1) / prm.tile_blocks;
        if (prm.ntiles > 15) {                       // at most 16 mbarriers in the 128-byte header
            prm.tile_blocks = (n_blocks + 14) / 15;
            prm.ntiles = (n_blocks + prm.tile_blocks - 1) / prm.tile_blocks;
        }
        prm.stages = prm.ntiles > 0 ? prm.ntiles : 1;
        smem = hdr + all_bytes;
    } else {
        prm.resident = 0;
        // measured (tools/stream_exp.py, M = 100k): 512-pair tiles x 8 stages 2.20e12, 1024 x 6 2.36e12,
        // 2048 x 3 2.45e12, 3072 x 2 2.47e12 -- every tile costs a CTA barrier and an FP64 fold
        prm.tile_blocks = (tu.k2_tile_pairs > 0 ? tu.k2_tile_pairs : 3072) / 32;
        if (prm.tile_blocks < 1) prm.tile_blocks = 1;
        prm.stages = tu.k2_stages > 0 ? tu.k2_stages : 2;
        if (prm.stages > 8) prm.stages = 8;
        if (prm.stages < 2) prm.stages = 2;
        while (prm.tile_blocks > 1 &&
               hdr + (size_t)prm.stages * prm.tile_blocks * sizeof(ChargeBlock) > (size_t)c->max_smem_optin)
            prm.tile_blocks /= 2;
        prm.ntiles = (n_blocks + prm.tile_blocks - 1) / prm.tile_blocks;
        smem = hdr + (size_t)prm.stages * prm.tile_blocks * sizeof(ChargeBlock);
    }

    // --- lines per warp ----------------------------------------------------------------------------
    const long long all_warps = (long long)sms * warps_per_cta;
    int cap = tu.k2_cap;
    if (cap != 1 && cap != 2 && cap != 4) {
        // 4 lines per warp once every warp of the chip gets that many in the first wave; fewer for
        // short queues, so that the lines spread over all SMs instead of filling the first warps
        // (5,832 lines on 2,368 warps: 1 line per warp 1.72e12, 2 lines 1.57e12, 4 lines 0.94e12).
        cap = n_lines >= 4 * all_warps ? 4 : (n_lines >= 3 * all_warps ? 2 : 1);
    }
    prm.cap = cap;
    int grid = sms;
    const long long need_ctas = (n_lines + (long long)warps_per_cta * cap - 1) / ((long long)warps_per_cta * cap);
    if (need_ctas < grid) grid = (int)need_ctas;

    const long long slots = (long long)grid * warps_per_cta * cap;
    const bool do_sort = (tu.k2_sort < 0) ? (n_lines > slots) : (tu.k2_sort != 0);
    if (int rc = prepare_queue(c, n_lines, d_n_iter, do_sort, &prm.queue, &prm.evals, &prm.order, &launches))
        return rc;

    prm.seeds = d_seeds;
    prm.n_iter = d_n_iter;
    prm.n_lines = n_lines;
    prm.h = step;
    prm.dimx = dims[0]; prm.dimy = dims[1]; prm.dimz = dims[2];
    prm.out = d_out;
    prm.steps = d_steps;

    KernelTimer timer(c);   // brackets the integrator kernel only (the roofline's "dominant kernel")
    const bool sd = (flags & CPET_TOPO_CURV_SECOND_DIFF) != 0u;
    // inner-loop unroll: 4 blocks x 4 points (192 packed FMA-pipe instructions per iteration, no
    // register-rotation MOVs at the back edge; U = 2 leaves 24 MOVs per 96 -- profiles/round1_k2w.md)
    const int rc = sd ? launch_k2w_inst<true, 4, 4>(c, prm, grid, threads, smem)
                      : launch_k2w_inst<false, 4, 4>(c, prm, grid, threads, smem);
    if (rc) return rc;
    launches += 1;
    c->last_counters[0] = launches;
    c->last_counters[1] = -1;   // resolved lazily from the device counter (see capi.cu)
    c->last_counters[2] = -1;
    return CPET_OK;
}

// Seed extent, class of every charge (near / far against this launch's box) and stable compaction into the
// blocks of `layout` (0 = XBlock, 1 = PBlock) in c->xblocks; nothing is read back by the host, the integrator
// takes the block counts from *meta_out in device memory.  Must follow prepare_queue (which zeroes the meta).
int pack_hybrid(cpet_ctx* c, int n_lines, const float* d_seeds, float step, const float dims[3], int layout,
                int max_blocks, K2XMeta** meta_out, int* launches) {
    const int sms = c->sm_count;
    K2XMeta* meta = reinterpret_cast<K2XMeta*>(c->counters.as<unsigned char>() + 16);   // zeroed by prepare_queue
    static_assert(sizeof(K2XMeta) <= 48, "K2XMeta must fit the counter block header");
    const int n_chunks = (c->n_charges + K2X_CHUNK - 1) / K2X_CHUNK;
    const size_t block_bytes = layout == 0 ? sizeof(XBlock) : sizeof(PBlock);
    if (int rc = c->xblocks.reserve(block_bytes * (size_t)max_blocks)) return rc;
    if (int rc = c->xchunks.reserve(sizeof(int) * 2 * (size_t)(n_chunks > 0 ? n_chunks : 1))) return rc;
    const float amax = c->tune.k2_amax > 0 ? (float)c->tune.k2_amax : 8.0f;
    int eb = (n_lines + 255) / 256;
    if (eb > sms * 4) eb = sms * 4;
    k2x_extent_kernel<<<eb, 256, 0, c->stream>>>(d_seeds, n_lines, meta);
    *launches += 1;
    if (n_chunks > 0) {
        k2x_count_kernel<<<n_chunks, K2X_CHUNK, 0, c->stream>>>(c->charges.as<ChargePair>(), c->n_charges, dims[0],
                                                               dims[1], dims[2], step, amax, meta,
                                                               c->xchunks.as<int>());
        if (layout == 0)
            k2x_scatter_kernel<0><<<n_chunks, K2X_CHUNK, 0, c->stream>>>(c->charges.as<ChargePair>(), c->n_charges,
                                                                        dims[0], dims[1], dims[2], step, amax, meta,
                                                                        c->xchunks.as<int>(), n_chunks, c->xblocks.p);
        else
            k2x_scatter_kernel<1><<<n_chunks, K2X_CHUNK, 0, c->stream>>>(c->charges.as<ChargePair>(), c->n_charges,
                                                                        dims[0], dims[1], dims[2], step, amax, meta,
                                                                        c->xchunks.as<int>(), n_chunks, c->xblocks.p);
        *launches += 2;
    }
    CPET_CUDA_TRY(cudaGetLastError());
    *meta_out = meta;
    return CPET_OK;
}

template <bool SD, int U4, int U2>
static int launch_k2x_inst(cpet_ctx* c, const K2XParams& prm, int grid, int threads, size_t smem) {
    auto kern = k2x_topo_kernel<SD, U4, U2>;
    CPET_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, threads, smem, c->stream>>>(prm);
    CPET_CUDA_TRY(cudaGetLastError());
    return CPET_OK;
}

// Hybrid near/far integrator (default).  Launch sequence on the context's stream: queue sort (as
// before), seed extent, charge classification + stable compaction into XBlocks, integrator.
static int launch_topo_hybrid(cpet_ctx* c, int n_lines, const float* d_seeds, const int32_t* d_n_iter,
                              float step, const float dims[3], unsigned flags, float* d_out,
                              int32_t* d_steps) {
    const Tuning& tu = c->tune;
    const int sms = c->sm_count;
    int launches = 0;

    int threads = tu.k2_threads > 0 ? tu.k2_threads : CPET_K2X_MAXT;
    threads = (threads / 32) * 32;
    if (threads < 32) threads = 32;
    if (threads > CPET_K2X_MAXT) threads = CPET_K2X_MAXT;
    if (tu.k2_threads <= 0 && (long long)n_lines < (long long)sms * (threads / 32)) {
        // fewer lines than warps on the chip: one line per warp, spread over all SMs
        threads = 32 * ((n_lines + sms - 1) / sms);
    }
    const int warps_per_cta = threads / 32;
    const size_t hdr = 128 + sizeof(WarpLinesX) * (size_t)warps_per_cta;

    // --- charge staging plan; the class sizes are only known on the device, so plan for the worst
    //     case: every class ends in a partly filled block
    const int max_blocks = (c->n_charges + 63) / 64 + 2;
    K2XParams prm;
    size_t smem;
    if (hdr + (size_t)max_blocks * sizeof(XBlock) <= (size_t)c->max_smem_optin && tu.k2_stages <= 0 &&
        tu.k2_tile_pairs <= 0) {
        prm.resident = 1;
        prm.tile_blocks = max_blocks;
        prm.stages = 1;
        smem = hdr + (size_t)max_blocks * sizeof(XBlock);
    } else {
        prm.resident = 0;
        prm.stages = tu.k2_stages > 0 ? tu.k2_stages : 2;
        if (prm.stages > 8) prm.stages = 8;
        if (prm.stages < 2) prm.stages = 2;
        const int fit = (int)(((size_t)c->max_smem_optin - hdr) / ((size_t)prm.stages * sizeof(XBlock)));
        prm.tile_blocks = tu.k2_tile_pairs > 0 ? tu.k2_tile_pairs / 32 : fit;
        if (prm.tile_blocks > fit) prm.tile_blocks = fit;
        if (prm.tile_blocks < 1) prm.tile_blocks = 1;
        smem = hdr + (size_t)prm.stages * prm.tile_blocks * sizeof(XBlock);
    }

    // --- lines per warp (same heuristic as the direct-form kernel) ------------------------------------
    const long long all_warps = (long long)sms * warps_per_cta;
    int cap = tu.k2_cap;
    if (cap != 1 && cap != 2 && cap != 4) cap = n_lines >= 4 * all_warps ? 4 : (n_lines >= 3 * all_warps ? 2 : 1);
    prm.cap = cap;
    int grid = sms;
    const long long need_ctas = (n_lines + (long long)warps_per_cta * cap - 1) / ((long long)warps_per_cta * cap);
    if (need_ctas < grid) grid = (int)need_ctas;
    const long long slots = (long long)grid * warps_per_cta * cap;
    const bool do_sort = (tu.k2_sort < 0) ? (n_lines > slots) : (tu.k2_sort != 0);
    if (int rc = prepare_queue(c, n_lines, d_n_iter, do_sort, &prm.queue, &prm.evals, &prm.order, &launches))
        return rc;

    // --- classify and pack the charges against this launch's box ---------------------------------------
    K2XMeta* meta = nullptr;
    if (int rc = pack_hybrid(c, n_lines, d_seeds, step, dims, 0, max_blocks, &meta, &launches)) return rc;

    prm.blocks = c->xblocks.as<XBlock>();
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
    // far-loop unroll of the 4-point pass (k2_unroll; default by staging mode, see profiles/round2_k2x.md)
    int rc;
    const int unroll = tu.k2_unroll > 0 ? tu.k2_unroll : (prm.resident && c->n_charges >= 4000 ? 6 : 4);
#define K2X_LAUNCH(U) (sd ? launch_k2x_inst<true, U, 4>(c, prm, grid, threads, smem) : launch_k2x_inst<false, U, 4>(c, prm, grid, threads, smem))
    if (unroll <= 3) rc = K2X_LAUNCH(3);
    else if (unroll <= 4) rc = K2X_LAUNCH(4);
    else if (unroll <= 6) rc = K2X_LAUNCH(6);
    else if (unroll <= 8) rc = K2X_LAUNCH(8);
    else rc = K2X_LAUNCH(12);
#undef K2X_LAUNCH
    if (rc) return rc;
    launches += 1;
    c->last_counters[0] = launches;
    c->last_counters[1] = -1;   // resolved lazily from the device counter (see capi.cu)
    c->last_counters[2] = -1;
    return CPET_OK;
}

int launch_topo(cpet_ctx* c, int n_lines, const float* d_seeds, const int32_t* d_n_iter,
                float step, const float dims[3], unsigned flags, float* d_out, int32_t* d_steps) {
    CPET_REQUIRE(c->charges_set, CPET_ERR_STATE, "no charge set on this context: call cpet_set_charges first");
    c->last_counters[0] = c->last_counters[1] = c->last_counters[2] = 0;
    if (n_lines == 0) return CPET_OK;
    if (c->tune.k2_form == 1) { // the round-1 direct-form kernel, kept for A/B measurements
        c->last_path = 11;
        return launch_topo_warpwide(c, n_lines, d_seeds, d_n_iter, step, dims, flags, d_out, d_steps);
    }
    if (c->tune.k2_form == 2) { // the charge-pair-packed hybrid kernel
        c->last_path = 12;
        return launch_topo_hybrid(c, n_lines, d_seeds, d_n_iter, step, dims, flags, d_out, d_steps);
    }
    // default: the points-packed hybrid kernel (topo8.cu) once every warp of the chip gets at least 4 lines; shorter
    // queues keep the charge-pair-packed kernel, which fills a warp with one or two lines
    if (c->tune.k2_form == 3 || topo8_wants(c, n_lines)) {
        c->last_path = 13;
        return launch_topo_points_packed(c, n_lines, d_seeds, d_n_iter, step, dims, flags, d_out, d_steps);
    }
    c->last_path = 12;
    return launch_topo_hybrid(c, n_lines, d_seeds, d_n_iter, step, dims, flags, d_out, d_steps);
}

}  // namespace cpet
