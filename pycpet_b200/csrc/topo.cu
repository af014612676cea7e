// topo.cu -- K2: batched fixed-step streamline integrator with fused distance / curvature.
//
// Replaces, for a whole frame in ONE launch,
//   * the CPU path: Pool.starmap over task_complete_thread -> thread_operation
//     (CPET/source/calculator.py:675-712, CPET/utils/math_module.c:523-591, 489-503, 296-333)
//   * the torch path: compute_topo_GPU_batch_filter + CPET/utils/gpu.py:25-399
//     (path matrix windows, Python filter loops, torch.cat dumps).
//
// Design: a persistent kernel, one CTA per SM.  A *slot* (G lanes, G = 1..32) owns one
// streamline at a time and keeps its whole state in registers: current point, last two points,
// seed, end point, previous unit field direction.  One *round* = every slot evaluates the field
// at its current point by walking all M charges (G lanes split the pairs of each tile), then
// advances its own little state machine.  A slot whose line has finished writes
// {distance, curvature} and pulls the next line index from a global queue ordered by n_iter
// descending (longest-processing-time-first), so lanes stay busy until the queue runs dry.
// The field at p_k is evaluated exactly once: K+2 evaluations per line (K = steps taken), versus
// K+4 in the reference C (the first two are recomputed there) -- the two look-ahead points at
// the end are simply the next two steps of the same integration.
//
// Charges: if the packed set fits in shared memory (<= ~6.9k pairs = 13.8k charges) it is loaded
// once per CTA by TMA bulk copies and warps then run fully independently (no CTA barriers);
// otherwise tiles stream continuously through an S-stage TMA/mbarrier ring.
#include "cpet_internal.h"

namespace cpet {

#define K2_KEYMAX 2048
#ifndef CPET_K2_MAXT
#define CPET_K2_MAXT 512
#endif
#ifndef CPET_K2_UNROLL_P2
#define CPET_K2_UNROLL_P2 2
#endif
#ifndef CPET_K2_UNROLL
#define CPET_K2_UNROLL 8
#endif

struct K2Params {
    const ChargePair* charges;
    int n_pairs;
    int tile_pairs;
    int stages;
    int ntiles;
    int resident;
    const float* seeds;
    const int32_t* n_iter;
    const int32_t* order;   // queue -> line id (nullptr: identity)
    int n_lines;
    float h;
    float dimx, dimy, dimz;
    unsigned flags;
    float* out;
    int32_t* steps;
    unsigned int* queue;
    unsigned long long* evals;
};

// kappa = |v' x v''| / |v'|^3 from three consecutive FP32 positions, evaluated the way
// math_module.c does (C:575-580 differences in float; C:89-96, 108-121 norms through double).
__device__ __forceinline__ float curv3_f32(float3 a0, float3 a1, float3 a2) {
    const float v1x = a1.x - a0.x, v1y = a1.y - a0.y, v1z = a1.z - a0.z;
    const float v2x = a2.x - 2.0f * a1.x + a0.x;
    const float v2y = a2.y - 2.0f * a1.y + a0.y;
    const float v2z = a2.z - 2.0f * a1.z + a0.z;
    const float cx = v1y * v2z - v1z * v2y;
    const float cy = v1z * v2x - v1x * v2z;
    const float cz = v1x * v2y - v1y * v2x;
    const float nc = (float)sqrt((double)cx * cx + (double)cy * cy + (double)cz * cz);
    const float nd = (float)sqrt((double)v1x * v1x + (double)v1y * v1y + (double)v1z * v1z);
    const double d3 = (double)nd * (double)nd * (double)nd;
    return (float)((double)nc / d3);
}

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

// Per-streamline state, all in registers.  SD (second-difference curvature mode) additionally
// keeps the two previous FP32 positions.
template <bool SD>
struct LineState {
    int line, n_it, k, k_end;          // k = index of the current point; k_end = K once known
    float px, py, pz;                  // current point p_k
    float sx, sy, sz;                  // seed
    float dist;                        // |seed - p_K|, fixed when K becomes known
    double ux, uy, uz;                 // unit field direction at p_{k-1}
    float kinit;                       // curvature at the seed end
    float m1x, m1y, m1z, m2x, m2y, m2z;   // p_{k-1}, p_{k-2} (SD only)
};

template <int G, int P, bool SD>
__global__ void __launch_bounds__(CPET_K2_MAXT, 1) k2_topo_kernel(const K2Params prm) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
    ChargePair* ring = reinterpret_cast<ChargePair*>(smem_raw + 128);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int lane_g = tid % G;
    const bool leader = (lane_g == 0);
    const int S = prm.stages;
    const int TP = prm.tile_pairs;
    const int NT = prm.ntiles;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    auto issue = [&](int it) {   // it = running tile counter; tile index = it % NT
        const int stage = it % S;
        const int t = it % NT;
        const int n_t = min(TP, prm.n_pairs - t * TP);
        const uint32_t bytes = (uint32_t)n_t * (uint32_t)sizeof(ChargePair);
        mbar_expect_tx(&full[stage], bytes);
        tma_load_1d(ring + (size_t)stage * TP, prm.charges + (size_t)t * TP, bytes, &full[stage]);
    };
    int issued = 0;
    if (tid == 0) {
        const int pre = prm.resident ? NT : min(S, NT > 0 ? S : 0);
        for (; issued < pre; ++issued) issue(issued);
    }
    if (prm.resident) {
        for (int t = 0; t < NT; ++t) mbar_wait(&full[t], 0u);
    }

    // ---- slot state: P streamlines per thread (group of G lanes) -----------------------------
    LineState<SD> st[P];
#pragma unroll
    for (int q = 0; q < P; ++q) {
        st[q].line = -1; st[q].n_it = 0; st[q].k = 0; st[q].k_end = -1;
        st[q].px = st[q].py = st[q].pz = 0.f;
        st[q].sx = st[q].sy = st[q].sz = 0.f;
        st[q].dist = 0.f; st[q].kinit = 0.f;
        st[q].ux = st[q].uy = st[q].uz = 0.0;
        st[q].m1x = st[q].m1y = st[q].m1z = st[q].m2x = st[q].m2y = st[q].m2z = 0.f;
    }
    bool exhausted = false;
    unsigned long long my_evals = 0ull;
    const double h = (double)prm.h;
    const double inv_h = 1.0 / h;

    int it = 0;   // consumed-tile counter (streamed mode)
    while (true) {
        // ---- refill empty slots from the queue (warp-aggregated atomic) ------------------------
        bool any_active = false;
#pragma unroll
        for (int q = 0; q < P; ++q) {
            const bool want = leader && (st[q].line < 0) && !exhausted;
            const unsigned m = __ballot_sync(0xffffffffu, want);
            int slot = -1;
            if (m) {
                const int src = __ffs(m) - 1;
                unsigned base = 0;
                if (lane == src) base = atomicAdd(prm.queue, (unsigned)__popc(m));
                base = __shfl_sync(0xffffffffu, base, src);
                if (want) slot = (int)(base + (unsigned)__popc(m & ((1u << lane) - 1u)));
            }
            if (G > 1) {
                const int had = __shfl_sync(0xffffffffu, want ? 1 : 0, lane - lane_g);
                const int sl = __shfl_sync(0xffffffffu, slot, lane - lane_g);
                slot = had ? sl : -1;
            }
            if (slot >= 0) {
                if (slot < prm.n_lines) {
                    const int line = prm.order ? prm.order[slot] : slot;
                    st[q].line = line;
                    st[q].sx = prm.seeds[3 * (size_t)line];
                    st[q].sy = prm.seeds[3 * (size_t)line + 1];
                    st[q].sz = prm.seeds[3 * (size_t)line + 2];
                    st[q].n_it = prm.n_iter[line];
                    st[q].px = st[q].sx; st[q].py = st[q].sy; st[q].pz = st[q].sz;
                    if (SD) {
                        st[q].m1x = st[q].m2x = st[q].sx; st[q].m1y = st[q].m2y = st[q].sy;
                        st[q].m1z = st[q].m2z = st[q].sz;
                    }
                    st[q].k = 0;
                    st[q].k_end = (st[q].n_it <= 0) ? 0 : -1;
                    st[q].dist = 0.f;
                } else {
                    exhausted = true;
                }
            }
            any_active = any_active || (st[q].line >= 0);
        }
        bool go;
        if (prm.resident) go = __any_sync(0xffffffffu, any_active);
        else go = __syncthreads_or(any_active ? 1 : 0) != 0;
        if (!go) break;

        // ---- field at the P current points: all charges ------------------------------------------
        PointRegs<P> r;
        double acc[P][3];
#pragma unroll
        for (int q = 0; q < P; ++q) {
            set_point<P>(r, q, st[q].px, st[q].py, st[q].pz);
            acc[q][0] = acc[q][1] = acc[q][2] = 0.0;
        }
        clear_partials<P>(r);
        if (prm.resident) {
            for (int t = 0; t < NT; ++t) {
                const int n_t = min(TP, prm.n_pairs - t * TP);
                eval_tile_chunked<MODE_FIELD_RAW, P, (P >= 2 ? CPET_K2_UNROLL_P2 : CPET_K2_UNROLL), 64>(ring + (size_t)t * TP, lane_g,
                                                                           n_t, G, r, acc);
            }
        } else {
            const bool warp_on = __any_sync(0xffffffffu, any_active);
            for (int t = 0; t < NT; ++t, ++it) {
                const int stage = it % S;
                mbar_wait(&full[stage], (uint32_t)((it / S) & 1));
                if (warp_on) {
                    const int n_t = min(TP, prm.n_pairs - t * TP);
                    eval_tile_chunked<MODE_FIELD_RAW, P, (P >= 2 ? CPET_K2_UNROLL_P2 : CPET_K2_UNROLL), 64>(
                        ring + (size_t)stage * TP, lane_g, n_t, G, r, acc);
                }
                __syncthreads();                      // stage fully consumed by the CTA
                if (tid == 0) { issue(issued); ++issued; }   // speculative: next round's tiles too
            }
        }

        // ---- per-slot state machines ----------------------------------------------------------------
#pragma unroll
        for (int q = 0; q < P; ++q) {
            double ex = acc[q][0], ey = acc[q][1], ez = acc[q][2];
            if (G > 1) {
#pragma unroll
                for (int m = G / 2; m >= 1; m >>= 1) {
                    ex += shfl_xor_f64(ex, m);
                    ey += shfl_xor_f64(ey, m);
                    ez += shfl_xor_f64(ez, m);
                }
            }
            LineState<SD>& L = st[q];
            if (L.line < 0) continue;
            // unit direction (no zero guard: E = 0 gives NaN exactly like C:501)
            const double inv_n = rsqrt_f64_fast(ex * ex + ey * ey + ez * ez);
            const double ux = ex * inv_n, uy = ey * inv_n, uz = ez * inv_n;
            const bool last = (L.k_end >= 0) && (L.k == L.k_end + 1);
            float kdir = 0.f;
            if (!SD && (L.k == 1 || last)) {   // curvature is needed at the first and last point pair
                const double cx = L.uy * uz - L.uz * uy;
                const double cy = L.uz * ux - L.ux * uz;
                const double cz = L.ux * uy - L.uy * ux;
                kdir = (float)(sqrt_f64_fast(cx * cx + cy * cy + cz * cz) * inv_h);
                if (L.k == 1) L.kinit = kdir;
            }
            const float nx = (float)((double)L.px + h * ux);
            const float ny = (float)((double)L.py + h * uy);
            const float nz = (float)((double)L.pz + h * uz);
            if (last) {
                if (SD) {
                    kdir = curv3_f32(make_float3(L.m1x, L.m1y, L.m1z), make_float3(L.px, L.py, L.pz),
                                     make_float3(nx, ny, nz));
                    if (L.k == 1) L.kinit = kdir;
                }
                if (leader) {
                    reinterpret_cast<float2*>(prm.out)[L.line] = make_float2(L.dist, (L.kinit + kdir) * 0.5f);
                    if (prm.steps) prm.steps[L.line] = L.k_end;
                    my_evals += (unsigned long long)(L.k_end + 2);
                }
                L.line = -1;
            } else {
                if (SD) {
                    L.m2x = L.m1x; L.m2y = L.m1y; L.m2z = L.m1z;
                    L.m1x = L.px; L.m1y = L.py; L.m1z = L.pz;
                }
                L.px = nx; L.py = ny; L.pz = nz;
                ++L.k;
                if (SD && L.k == 2)
                    L.kinit = curv3_f32(make_float3(L.m2x, L.m2y, L.m2z), make_float3(L.m1x, L.m1y, L.m1z),
                                        make_float3(nx, ny, nz));
                if (L.k_end < 0) {
                    const bool outside = (nx < -prm.dimx) || (nx > prm.dimx) || (ny < -prm.dimy) ||
                                         (ny > prm.dimy) || (nz < -prm.dimz) || (nz > prm.dimz);
                    if (L.k >= L.n_it || outside) {
                        L.k_end = L.k;
                        const double ddx = (double)L.sx - (double)nx;
                        const double ddy = (double)L.sy - (double)ny;
                        const double ddz = (double)L.sz - (double)nz;
                        L.dist = (float)sqrt(ddx * ddx + ddy * ddy + ddz * ddz);
                    }
                }
                L.ux = ux; L.uy = uy; L.uz = uz;
            }
        }
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

template <int G, int P, bool SD>
static int launch_k2_inst(cpet_ctx* c, const K2Params& prm, int grid, int threads, size_t smem) {
    auto kern = k2_topo_kernel<G, P, SD>;
    CPET_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, threads, smem, c->stream>>>(prm);
    CPET_CUDA_TRY(cudaGetLastError());
    return CPET_OK;
}

int launch_topo(cpet_ctx* c, int n_lines, const float* d_seeds, const int32_t* d_n_iter,
                float step, const float dims[3], unsigned flags, float* d_out, int32_t* d_steps) {
    CPET_REQUIRE(c->charges_set, CPET_ERR_STATE, "no charge set on this context: call cpet_set_charges first");
    c->last_counters[0] = c->last_counters[1] = c->last_counters[2] = 0;
    if (n_lines == 0) return CPET_OK;
    const Tuning& tu = c->tune;
    const int sms = c->sm_count;
    int launches = 0;

    // --- charge staging plan -------------------------------------------------------------------
    const size_t smem_cap = (size_t)c->max_smem_optin - 128;
    const size_t all_bytes = (size_t)c->n_pairs * sizeof(ChargePair);
    K2Params prm;
    prm.charges = c->charges.as<ChargePair>();
    prm.n_pairs = c->n_pairs;
    size_t smem;
    if (all_bytes <= smem_cap && tu.k2_stages <= 0 && tu.k2_tile_pairs <= 0) {
        prm.resident = 1;
        prm.tile_pairs = 1024;                       // 32 KB per bulk copy
        prm.ntiles = (c->n_pairs + prm.tile_pairs - 1) / prm.tile_pairs;
        if (prm.ntiles > 15) {                       // at most 16 mbarriers in the 128-byte header
            prm.tile_pairs = ((c->n_pairs + 14) / 15 + 7) / 8 * 8;
            prm.ntiles = (c->n_pairs + prm.tile_pairs - 1) / prm.tile_pairs;
        }
        prm.stages = prm.ntiles > 0 ? prm.ntiles : 1;
        // tiles are contiguous; the last one may be partial: exactly n_pairs pairs live in smem
        smem = 128 + all_bytes;
    } else {
        prm.resident = 0;
        prm.tile_pairs = tu.k2_tile_pairs > 0 ? tu.k2_tile_pairs : 2048;
        prm.tile_pairs = ((prm.tile_pairs + 7) / 8) * 8;
        prm.stages = tu.k2_stages > 0 ? tu.k2_stages : 3;
        if (prm.stages > 8) prm.stages = 8;
        if (prm.stages < 2) prm.stages = 2;
        while (128 + (size_t)prm.stages * prm.tile_pairs * sizeof(ChargePair) > (size_t)c->max_smem_optin)
            prm.tile_pairs /= 2;
        prm.ntiles = (c->n_pairs + prm.tile_pairs - 1) / prm.tile_pairs;
        smem = 128 + (size_t)prm.stages * prm.tile_pairs * sizeof(ChargePair);
    }

    // --- slots ------------------------------------------------------------------------------------
    int threads = tu.k2_threads > 0 ? tu.k2_threads : 512;
    threads = (threads / 32) * 32;
    if (threads < 32) threads = 32;
    if (threads > CPET_K2_MAXT) threads = CPET_K2_MAXT;
    int P = tu.k2_points;
    if (P != 1 && P != 2) P = 1;
    int G = tu.k2_lanes;
    if (G <= 0) {
        // Measured on B200 (profiles/round1_sweep.md): 16 warps/SM are needed to hide latency, and
        // the longest-first queue evens out the tail once there are >= ~2.5 lines per slot.
        // Long lines (the reference's max_steps = round(2*|dims|/h), SC:272, in the hundreds) leave a
        // longer, more ragged tail: ask for >= 5 lines per slot there (h = 0.01: +12 %).
        const double diag = sqrt((double)dims[0] * dims[0] + (double)dims[1] * dims[1] + (double)dims[2] * dims[2]);
        const double max_steps = step > 0.f ? 2.0 * diag / (double)step : 0.0;
        const long long per_slot_x2 = max_steps > 32.0 ? 10 : 5;
        G = 1;
        while (G < 32 && 2LL * n_lines * G < per_slot_x2 * sms * threads * P) G *= 2;
    }
    if (G & (G - 1)) G = 1;
    if (G > 32) G = 32;
    int grid = sms;
    const long long slots_per_cta = (long long)(threads / G) * P;
    const long long need_ctas = (n_lines + slots_per_cta - 1) / slots_per_cta;
    if (need_ctas < grid) grid = (int)need_ctas;

    // --- queue order ---------------------------------------------------------------------------
    const bool do_sort = (tu.k2_sort < 0) ? (n_lines > (int)slots_per_cta * grid) : (tu.k2_sort != 0);
    if (int rc = c->counters.reserve(64 + sizeof(unsigned) * 3 * K2_KEYMAX)) return rc;
    unsigned char* cb = c->counters.as<unsigned char>();
    CPET_CUDA_TRY(cudaMemsetAsync(cb, 0, 64 + sizeof(unsigned) * 3 * K2_KEYMAX, c->stream));
    prm.queue = reinterpret_cast<unsigned int*>(cb);
    prm.evals = reinterpret_cast<unsigned long long*>(cb + 8);
    unsigned* hist = reinterpret_cast<unsigned*>(cb + 64);
    unsigned* offsets = hist + K2_KEYMAX;
    unsigned* cursor = offsets + K2_KEYMAX;
    prm.order = nullptr;

    if (do_sort) {
        if (int rc = c->work1.reserve(sizeof(int32_t) * (size_t)n_lines)) return rc;
        int blocks = (n_lines + 255) / 256;
        if (blocks > sms * 8) blocks = sms * 8;
        k2_count_kernel<<<blocks, 256, 0, c->stream>>>(d_n_iter, n_lines, hist);
        k2_scan_kernel<<<1, 1024, 0, c->stream>>>(hist, offsets);
        k2_scatter_kernel<<<blocks, 256, 0, c->stream>>>(d_n_iter, n_lines, offsets, cursor,
                                                         c->work1.as<int32_t>());
        CPET_CUDA_TRY(cudaGetLastError());
        prm.order = c->work1.as<int32_t>();
        launches += 3;
    }

    prm.seeds = d_seeds;
    prm.n_iter = d_n_iter;
    prm.n_lines = n_lines;
    prm.h = step;
    prm.dimx = dims[0]; prm.dimy = dims[1]; prm.dimz = dims[2];
    prm.flags = flags;
    prm.out = d_out;
    prm.steps = d_steps;

    KernelTimer timer(c);   // brackets the integrator kernel only (the roofline's "dominant kernel")
    int rc;
    const bool sd = (flags & CPET_TOPO_CURV_SECOND_DIFF) != 0u;
#define CPET_K2_CASE(GG)                                                                         \
    case GG:                                                                                     \
        if (sd) rc = (P == 2) ? launch_k2_inst<GG, 2, true>(c, prm, grid, threads, smem)          \
                              : launch_k2_inst<GG, 1, true>(c, prm, grid, threads, smem);         \
        else rc = (P == 2) ? launch_k2_inst<GG, 2, false>(c, prm, grid, threads, smem)            \
                           : launch_k2_inst<GG, 1, false>(c, prm, grid, threads, smem);           \
        break;
    switch (G) {
        CPET_K2_CASE(1) CPET_K2_CASE(2) CPET_K2_CASE(4) CPET_K2_CASE(8) CPET_K2_CASE(16)
        default:
        CPET_K2_CASE(32)
    }
#undef CPET_K2_CASE
    if (rc) return rc;
    launches += 1;
    c->last_counters[0] = launches;
    c->last_counters[1] = -1;   // resolved lazily from the device counter (see capi.cu)
    c->last_counters[2] = -1;
    return CPET_OK;
}

}  // namespace cpet
