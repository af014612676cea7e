// field.cu -- K1: Coulomb E-field / potential of M point charges on N points (sm_100a).
//
// Replaces compute_looped_field (CPET/utils/math_module.c:405-451), the per-point loops over
// calc_field / calc_field_base (C:255-333) and calc_esp_base (C:453-486; loop at
// CPET/utils/calculator.py:468-469).
//
// Shape of the kernel (N-body style, no tensor cores):
//   * charges are pre-packed in HBM as 32-byte pairs (common.cuh: ChargePair); a CTA streams its
//     charge range through a ring of shared-memory tiles filled by 1-D TMA bulk copies
//     (cp.async.bulk + mbarrier complete_tx), so loads of tile t+S-1 overlap the math on tile t;
//   * every thread keeps P points (G=1) in registers, or G lanes share one point and split the
//     pairs of each tile (small N); each broadcast LDS.128 pair feeds 2*P pair evaluations of
//     packed FADD2/FMUL2/FFMA2 + MUFU.RSQ;
//   * FP32 accumulation inside a tile (even/odd charge sums), FP64 across tiles;
//   * for small N the charge range is additionally split over gridDim.y and a finalize kernel
//     sums the FP64 partials, so that >= 2 CTAs/SM are in flight even for a 11^3 grid.
#include "cpet_internal.h"
#include <cuda_fp16.h>

#ifndef CPET_K1_UNROLL_P4
#define CPET_K1_UNROLL_P4 2      // inner-loop unroll of the 4-points-per-thread general kernel
#endif

namespace cpet {

// ---------------------------------------------------------------------------------------------
// charge packing: (M,3) f32 + (M,) f32  ->  ChargePair[ceil(M/2)]
// ---------------------------------------------------------------------------------------------
__global__ void pack_charges_kernel(const float* __restrict__ x, const float* __restrict__ q,
                                    int n_charges, int n_pairs, ChargePair* __restrict__ out,
                                    ChargeBlock* __restrict__ blocks) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ((n_pairs + 31) & ~31)) return;
    float cx[2], cy[2], cz[2], cq[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int i = 2 * j + h;
        if (i < n_charges) {
            cx[h] = -x[3 * (size_t)i + 0];
            cy[h] = -x[3 * (size_t)i + 1];
            cz[h] = -x[3 * (size_t)i + 2];
            cq[h] = q[i];
        } else {
            cx[h] = cy[h] = cz[h] = -CPET_PAD_COORD;
            cq[h] = 0.0f;
        }
    }
    ChargePair cp;
    cp.a.nx = pk2(cx[0], cx[1]);
    cp.a.ny = pk2(cy[0], cy[1]);
    cp.b.nz = pk2(cz[0], cz[1]);
    cp.b.q = pk2(cq[0], cq[1]);
    if (j < n_pairs) out[j] = cp;
    blocks[j >> 5].a[j & 31] = cp.a;      // the last block is padded with q = 0 pairs
    blocks[j >> 5].b[j & 31] = cp.b;
}

int launch_pack_charges(cpet_ctx* c, int n_charges, const float* d_x, const float* d_q) {
    const int n_pairs = (n_charges + 1) / 2;
    if (int rc = c->charges.reserve(sizeof(ChargePair) * (size_t)(n_pairs > 0 ? n_pairs : 1))) return rc;
    c->n_charges = n_charges;
    c->charges_set = true;
    c->n_pairs = n_pairs;
    const int n_blocks = (n_pairs + 31) / 32;
    if (int rc = c->charge_blocks.reserve(sizeof(ChargeBlock) * (size_t)(n_blocks > 0 ? n_blocks : 1))) return rc;
    if (n_pairs > 0) {
        pack_charges_kernel<<<(n_blocks * 32 + 255) / 256, 256, 0, c->stream>>>(
            d_x, d_q, n_charges, n_pairs, c->charges.as<ChargePair>(), c->charge_blocks.as<ChargeBlock>());
        CPET_CUDA_TRY(cudaGetLastError());
    }
    return CPET_OK;
}

// ---------------------------------------------------------------------------------------------
// K1
// ---------------------------------------------------------------------------------------------
struct K1Params {
    const ChargePair* charges;
    int n_pairs;
    int pairs_per_split;
    int tile_pairs;
    int stages;
    const float* x0;
    int n_points;
    int out_kind;     // 0 (N,3) f32 | 1 (N,6) f32 [x0|E] | 2 (N,) f32 | 3 (N,4) f16 [x0|phi]
                      // 4 (N,3) f32 p + h E/|E|  (propagate_topo, C:489-503)
    float step;
    void* out;
    double* partial;  // [splits][n_points][3 (field) or 1 (ESP)] when gridDim.y > 1
};

__device__ __forceinline__ void store_result(int out_kind, float step, void* out, int pt, float x,
                                             float y, float z, double s0, double s1, double s2) {
    const double k = (double)CPET_COULOMB_K;
    if (out_kind == 0) {
        float* o = (float*)out + 3 * (size_t)pt;
        o[0] = (float)(k * s0); o[1] = (float)(k * s1); o[2] = (float)(k * s2);
    } else if (out_kind == 1) {
        float2* o = (float2*)((float*)out + 6 * (size_t)pt);
        o[0] = make_float2(x, y);
        o[1] = make_float2(z, (float)(k * s0));
        o[2] = make_float2((float)(k * s1), (float)(k * s2));
    } else if (out_kind == 2) {
        ((float*)out)[pt] = (float)(k * s0);
    } else if (out_kind == 4) {
        // no zero guard, as in the reference: E = 0 -> NaN
        const double inv_n = 1.0 / sqrt(s0 * s0 + s1 * s1 + s2 * s2);
        float* o = (float*)out + 3 * (size_t)pt;
        o[0] = (float)((double)x + (double)step * s0 * inv_n);
        o[1] = (float)((double)y + (double)step * s1 * inv_n);
        o[2] = (float)((double)z + (double)step * s2 * inv_n);
    } else {
        // float64 -> float32 -> float16, the two casts the reference applies (UC:464-475)
        __half2 lo = __floats2half2_rn(x, y);
        __half2 hi = __floats2half2_rn(z, (float)(k * s0));
        uint2 v;
        v.x = *reinterpret_cast<unsigned*>(&lo);
        v.y = *reinterpret_cast<unsigned*>(&hi);
        ((uint2*)out)[pt] = v;
    }
}

template <int MODE, int P, int G>
__global__ void __launch_bounds__(256) k1_grid_kernel(const K1Params prm) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
    ChargePair* ring = reinterpret_cast<ChargePair*>(smem_raw + 128);

    const int tid = threadIdx.x;
    const int lane_g = tid % G;
    const int groups = blockDim.x / G;
    const int grp = tid / G;
    const int S = prm.stages;
    const int TP = prm.tile_pairs;

    const int pbeg = blockIdx.y * prm.pairs_per_split;
    const int pend = min(prm.n_pairs, pbeg + prm.pairs_per_split);
    const int npairs = max(0, pend - pbeg);
    const int ntiles = (npairs + TP - 1) / TP;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    auto issue = [&](int t) {
        const int stage = t % S;
        const int n_t = min(TP, npairs - t * TP);
        const uint32_t bytes = (uint32_t)n_t * (uint32_t)sizeof(ChargePair);
        mbar_expect_tx(&full[stage], bytes);
        tma_load_1d(ring + (size_t)stage * TP, prm.charges + pbeg + (size_t)t * TP, bytes,
                    &full[stage]);
    };
    if (tid == 0) {
        const int pre = min(S, ntiles);
        for (int t = 0; t < pre; ++t) issue(t);
    }

    PointRegs<P> r;
    float px[P], py[P], pz[P];
    int pt[P];
    double acc[P][3];
    const int base = blockIdx.x * groups * P;
#pragma unroll
    for (int p = 0; p < P; ++p) {
        pt[p] = base + p * groups + grp;
        const int ld = min(pt[p], prm.n_points - 1);
        px[p] = prm.x0[3 * (size_t)ld + 0];
        py[p] = prm.x0[3 * (size_t)ld + 1];
        pz[p] = prm.x0[3 * (size_t)ld + 2];
        set_point<P>(r, p, px[p], py[p], pz[p]);
        acc[p][0] = acc[p][1] = acc[p][2] = 0.0;
    }
    clear_partials<P>(r);

    for (int t = 0; t < ntiles; ++t) {
        const int stage = t % S;
        mbar_wait(&full[stage], (uint32_t)((t / S) & 1));
        const int n_t = min(TP, npairs - t * TP);
        eval_tile_chunked<MODE, P, (P >= 4 ? CPET_K1_UNROLL_P4 : 4), 64>(ring + (size_t)stage * TP, lane_g, n_t, G, r,
                                                          acc);
        if (t + S < ntiles) {
            __syncthreads();           // every warp is done reading this stage
            if (tid == 0) issue(t + S);
        }
    }

    if (G > 1) {
#pragma unroll
        for (int m = G / 2; m >= 1; m >>= 1) {
#pragma unroll
            for (int p = 0; p < P; ++p) {
                acc[p][0] += shfl_xor_f64(acc[p][0], m);
                if (MODE != MODE_ESP) {
                    acc[p][1] += shfl_xor_f64(acc[p][1], m);
                    acc[p][2] += shfl_xor_f64(acc[p][2], m);
                }
            }
        }
    }
    if (lane_g != 0) return;
#pragma unroll
    for (int p = 0; p < P; ++p) {
        if (pt[p] >= prm.n_points) continue;
        if (gridDim.y == 1) {
            store_result(prm.out_kind, prm.step, prm.out, pt[p], px[p], py[p], pz[p], acc[p][0],
                         acc[p][1], acc[p][2]);
        } else {
            // [split][point][1 (ESP) or 3 (field)] doubles
            constexpr int NC = (MODE == MODE_ESP) ? 1 : 3;
            double* o = prm.partial + ((size_t)blockIdx.y * prm.n_points + pt[p]) * NC;
            o[0] = acc[p][0];
            if (MODE != MODE_ESP) { o[1] = acc[p][1]; o[2] = acc[p][2]; }
        }
    }
}

__global__ void k1_finalize_kernel(const double* __restrict__ partial, int splits, int n_points, int ncomp,
                                   const float* __restrict__ x0, int out_kind, float step,
                                   void* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_points) return;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int s = 0; s < splits; ++s) {
        const double* p = partial + ((size_t)s * n_points + i) * ncomp;
        s0 += p[0];
        if (ncomp == 3) { s1 += p[1]; s2 += p[2]; }
    }
    store_result(out_kind, step, out, i, x0[3 * (size_t)i], x0[3 * (size_t)i + 1], x0[3 * (size_t)i + 2],
                 s0, s1, s2);
}

// ---------------------------------------------------------------------------------------------
// K1-lattice: the same kernel for tensor-product grids (see common.cuh: eval_pair_lattice).
// Work item = PZ consecutive z-nodes of one (x,y) column; point id = (ix*ny + iy)*nz + iz.
// ---------------------------------------------------------------------------------------------
struct K1LatParams {
    const ChargePair* charges;
    int n_pairs;
    int pairs_per_split;
    int tile_pairs;
    int stages;
    const float* xs;
    const float* ys;
    const float* zs;
    int nx, ny, nz, nzb;    // nzb = ceil(nz / PZ) z-blocks per column
    int n_items;            // nx * ny * nzb
    int n_points;
    int out_kind;
    float step;
    void* out;
    double* partial;
    const unsigned* soft_flag;   // optional: *soft_flag != 0 <=> some r^2 may fall below the softening
};

// The softening max(r^2, 1e-6) of the `volume` path (C:433-436) costs two FMNMX issue slots per two
// pair-evaluations (9 % of the lattice kernel).  It can only act when a charge lies within 1e-3 A of
// a grid node on all three axes at once; this scan sets *flag when such a charge exists.  When it
// does not, max(r^2, eps) == r^2 for every pair and the unsoftened instantiation returns the same
// bits, so the launcher enqueues both instantiations and each exits at once unless the flag names it.
__global__ void __launch_bounds__(256) soften_scan_kernel(const ChargePair* __restrict__ charges, int n_pairs,
                                                          const float* __restrict__ xs, int nx,
                                                          const float* __restrict__ ys, int ny,
                                                          const float* __restrict__ zs, int nz,
                                                          unsigned* __restrict__ flag) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    // a NaN node coordinate makes r^2 NaN, which fmaxf(r^2, eps) turns into eps: keep the softened kernel
    if (j < nx + ny + nz) {
        const float a = j < nx ? xs[j] : (j < nx + ny ? ys[j - nx] : zs[j - nx - ny]);
        if (a != a) *flag = 1u;
    }
    if (j >= n_pairs) return;
    // |d| >= 1.0005e-3 on one axis => fl(d)^2 > 1e-6 => r^2 > eps whatever the other axes are
    const float T = 1.0005e-3f;
    const ChargePair cp = charges[j];
    float cx[2], cy[2], cz[2];
    upk2(cp.a.nx, cx[0], cx[1]);
    upk2(cp.a.ny, cy[0], cy[1]);
    upk2(cp.b.nz, cz[0], cz[1]);
    bool hit = false;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const float sum = cx[h] + cy[h] + cz[h];
        if (sum != sum) { hit = true; continue; }          // NaN (or inf - inf) charge coordinate: see above
        bool near_x = false, near_y = false, near_z = false;
        for (int i = 0; i < nx; ++i) near_x = near_x || !(fabsf(__ldg(xs + i) + cx[h]) >= T);
        if (!near_x) continue;
        for (int i = 0; i < ny; ++i) near_y = near_y || !(fabsf(__ldg(ys + i) + cy[h]) >= T);
        if (!near_y) continue;
        for (int i = 0; i < nz; ++i) near_z = near_z || !(fabsf(__ldg(zs + i) + cz[h]) >= T);
        hit = hit || near_z;
    }
    if (hit) *flag = 1u;
}

template <int MODE, int PZ, int U, int NF = 0>    // NF: z-nodes per thread whose rsqrt runs on the FMA pipe (ESP only)
__global__ void __launch_bounds__(256) k1_lattice_kernel(const K1LatParams prm) {
    if (prm.soft_flag != nullptr) {
        const bool need_soft = (*prm.soft_flag != 0u);
        if (need_soft != (MODE == MODE_FIELD_SOFT)) return;     // the other instantiation serves this call
    }
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
    ChargePair* ring = reinterpret_cast<ChargePair*>(smem_raw + 128);

    const int tid = threadIdx.x;
    const int S = prm.stages;
    const int TP = prm.tile_pairs;
    const int pbeg = blockIdx.y * prm.pairs_per_split;
    const int pend = min(prm.n_pairs, pbeg + prm.pairs_per_split);
    const int npairs = max(0, pend - pbeg);
    const int ntiles = (npairs + TP - 1) / TP;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int t) {
        const int stage = t % S;
        const int n_t = min(TP, npairs - t * TP);
        const uint32_t bytes = (uint32_t)n_t * (uint32_t)sizeof(ChargePair);
        mbar_expect_tx(&full[stage], bytes);
        tma_load_1d(ring + (size_t)stage * TP, prm.charges + pbeg + (size_t)t * TP, bytes,
                    &full[stage]);
    };
    if (tid == 0) {
        const int pre = min(S, ntiles);
        for (int t = 0; t < pre; ++t) issue(t);
    }

    // this thread's column and z-block
    const int item = min(blockIdx.x * blockDim.x + tid, prm.n_items - 1);
    const bool live = (blockIdx.x * blockDim.x + tid) < prm.n_items;
    const int col = item / prm.nzb;
    const int zb = item - col * prm.nzb;
    const int ix = col / prm.ny;
    const int iy = col - ix * prm.ny;
    const float x = prm.xs[ix], y = prm.ys[iy];
    float z[PZ];
    LatticeRegs<PZ> r;
    double acc[PZ][3];
    r.px = pk2(x, x);
    r.py = pk2(y, y);
#pragma unroll
    for (int p = 0; p < PZ; ++p) {
        z[p] = prm.zs[min(zb * PZ + p, prm.nz - 1)];
        r.pz[p] = pk2(z[p], z[p]);
        r.ax[p] = r.ay[p] = r.az[p] = 0ull;
        acc[p][0] = acc[p][1] = acc[p][2] = 0.0;
    }

    for (int t = 0; t < ntiles; ++t) {
        const int stage = t % S;
        mbar_wait(&full[stage], (uint32_t)((t / S) & 1));
        const int n_t = min(TP, npairs - t * TP);
        eval_tile_lattice_chunked<MODE, PZ, U, 64, NF>(ring + (size_t)stage * TP, n_t, r, acc);
        if (t + S < ntiles) {
            __syncthreads();
            if (tid == 0) issue(t + S);
        }
    }
    if (!live) return;
#pragma unroll
    for (int p = 0; p < PZ; ++p) {
        const int iz = zb * PZ + p;
        if (iz >= prm.nz) continue;
        const int pt = col * prm.nz + iz;
        if (gridDim.y == 1) {
            store_result(prm.out_kind, prm.step, prm.out, pt, x, y, z[p], acc[p][0], acc[p][1], acc[p][2]);
        } else {
            constexpr int NC = (MODE == MODE_ESP) ? 1 : 3;
            double* o = prm.partial + ((size_t)blockIdx.y * prm.n_points + pt) * NC;
            o[0] = acc[p][0];
            if (MODE != MODE_ESP) { o[1] = acc[p][1]; o[2] = acc[p][2]; }
        }
    }
}

// The same kernel with a thread's PZ z-nodes packed in pairs (common.cuh: eval_pair_lattice_nodes); field modes only.
template <int MODE, int PZ, int U>
__global__ void __launch_bounds__(256) k1_lattice_nodes_kernel(const K1LatParams prm) {
    if (prm.soft_flag != nullptr) {
        const bool need_soft = (*prm.soft_flag != 0u);
        if (need_soft != (MODE == MODE_FIELD_SOFT)) return;     // the other instantiation serves this call
    }
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
    ChargePair* ring = reinterpret_cast<ChargePair*>(smem_raw + 128);

    const int tid = threadIdx.x;
    const int S = prm.stages;
    const int TP = prm.tile_pairs;
    const int pbeg = blockIdx.y * prm.pairs_per_split;
    const int pend = min(prm.n_pairs, pbeg + prm.pairs_per_split);
    const int npairs = max(0, pend - pbeg);
    const int ntiles = (npairs + TP - 1) / TP;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int t) {
        const int stage = t % S;
        const int n_t = min(TP, npairs - t * TP);
        const uint32_t bytes = (uint32_t)n_t * (uint32_t)sizeof(ChargePair);
        mbar_expect_tx(&full[stage], bytes);
        tma_load_1d(ring + (size_t)stage * TP, prm.charges + pbeg + (size_t)t * TP, bytes,
                    &full[stage]);
    };
    if (tid == 0) {
        const int pre = min(S, ntiles);
        for (int t = 0; t < pre; ++t) issue(t);
    }

    // this thread's column and z-block
    const int item = min(blockIdx.x * blockDim.x + tid, prm.n_items - 1);
    const bool live = (blockIdx.x * blockDim.x + tid) < prm.n_items;
    const int col = item / prm.nzb;
    const int zb = item - col * prm.nzb;
    const int ix = col / prm.ny;
    const int iy = col - ix * prm.ny;
    const float x = prm.xs[ix], y = prm.ys[iy];
    float z[PZ];
    LatticeNodeRegs<PZ> r;
    double acc[PZ][3];
    r.x = x;
    r.y = y;
#pragma unroll
    for (int p = 0; p < PZ; ++p) {
        z[p] = prm.zs[min(zb * PZ + p, prm.nz - 1)];
        acc[p][0] = acc[p][1] = acc[p][2] = 0.0;
    }
#pragma unroll
    for (int p = 0; p < PZ / 2; ++p) {
        r.pz[p] = pk2(z[2 * p], z[2 * p + 1]);
        r.ax[p] = r.ay[p] = r.az[p] = 0ull;
    }

    for (int t = 0; t < ntiles; ++t) {
        const int stage = t % S;
        mbar_wait(&full[stage], (uint32_t)((t / S) & 1));
        const int n_t = min(TP, npairs - t * TP);
        eval_tile_lattice_nodes_chunked<MODE, PZ, U, 64>(ring + (size_t)stage * TP, n_t, r, acc);
        if (t + S < ntiles) {
            __syncthreads();
            if (tid == 0) issue(t + S);
        }
    }
    if (!live) return;
#pragma unroll
    for (int p = 0; p < PZ; ++p) {
        const int iz = zb * PZ + p;
        if (iz >= prm.nz) continue;
        const int pt = col * prm.nz + iz;
        if (gridDim.y == 1) {
            store_result(prm.out_kind, prm.step, prm.out, pt, x, y, z[p], acc[p][0], acc[p][1], acc[p][2]);
        } else {
            constexpr int NC = (MODE == MODE_ESP) ? 1 : 3;
            double* o = prm.partial + ((size_t)blockIdx.y * prm.n_points + pt) * NC;
            o[0] = acc[p][0];
            if (MODE != MODE_ESP) { o[1] = acc[p][1]; o[2] = acc[p][2]; }
        }
    }
}

// finalize for the lattice path: coordinates come from the axis arrays
__global__ void k1_lattice_finalize_kernel(const double* __restrict__ partial, int splits, int n_points, int ncomp,
                                           const float* __restrict__ xs, const float* __restrict__ ys,
                                           const float* __restrict__ zs, int ny, int nz, int out_kind,
                                           float step, void* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_points) return;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int s = 0; s < splits; ++s) {
        const double* p = partial + ((size_t)s * n_points + i) * ncomp;
        s0 += p[0];
        if (ncomp == 3) { s1 += p[1]; s2 += p[2]; }
    }
    const int iz = i % nz;
    const int col = i / nz;
    store_result(out_kind, step, out, i, xs[col / ny], ys[col % ny], zs[iz], s0, s1, s2);
}

template <int MODE, int PZ, int U, int NF = 0>
static int launch_k1_lat_inst(cpet_ctx* c, const K1LatParams& prm, dim3 grid, int threads, size_t smem) {
    auto kern = k1_lattice_kernel<MODE, PZ, U, NF>;
    CPET_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, threads, smem, c->stream>>>(prm);
    CPET_CUDA_TRY(cudaGetLastError());
    return CPET_OK;
}

template <int MODE, int PZ, int U>
static int launch_k1_latn_inst(cpet_ctx* c, const K1LatParams& prm, dim3 grid, int threads, size_t smem) {
    auto kern = k1_lattice_nodes_kernel<MODE, PZ, U>;
    CPET_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, threads, smem, c->stream>>>(prm);
    CPET_CUDA_TRY(cudaGetLastError());
    return CPET_OK;
}

int launch_field_lattice(cpet_ctx* c, int mode, int nx, int ny, int nz, const float* d_xs,
                         const float* d_ys, const float* d_zs, int out_kind, void* d_out) {
    const long long n_points_ll = (long long)nx * ny * nz;
    CPET_REQUIRE(c->charges_set, CPET_ERR_STATE, "no charge set on this context: call cpet_set_charges first");
    c->last_counters[0] = 0;
    c->last_counters[1] = n_points_ll * (long long)c->n_charges;
    c->last_counters[2] = n_points_ll;
    if (n_points_ll == 0) return CPET_OK;
    CPET_REQUIRE(n_points_ll <= 0x7fffffffLL, CPET_ERR_INVALID, "lattice too large (%lld points)", n_points_ll);
    const int n_points = (int)n_points_ll;
    const Tuning& tu = c->tune;
    const int sms = c->sm_count;
    const int threads = 256;
    // measured (profiles/round1_lattice_sweep.txt): 5 z-nodes/thread + unroll 4 is fastest (3.05e12
    // pair-evals/s at 100^3 x 100k); take 4 nodes when that pads the z axis noticeably less
    int PZ = tu.k1_points;
    // ESP on large meshes: 6 z-nodes per thread, one of them with its rsqrt on the FMA pipe (the sum is bound by the
    // 16-lane special-function unit otherwise; common.cuh: rsqrt2_fma)
    const bool esp_mix = mode == MODE_ESP && (tu.k1_esp_mix > 0 || (tu.k1_esp_mix < 0 && n_points >= 100000 && PZ <= 0));
    if (esp_mix) PZ = 6;
    // field modes on large meshes: node pairs packed, 8 or 10 z-nodes per thread (whichever pads the z axis less):
    // 100^3 x 100k charges 0.879 against 0.870 of the FP32 peak, 48 x 48 x 464 0.902 against 0.877
    // (tools/lattice_nodes_ab.py, profiles/round2_lattice_nodes.txt)
    const bool nodes = mode != MODE_ESP && (tu.k1_lat_nodes > 0 || (tu.k1_lat_nodes < 0 && n_points >= 100000 && PZ <= 0));
    if (nodes) {
        if (PZ != 4 && PZ != 6 && PZ != 8 && PZ != 10) {
            const int pad8 = (nz + 7) / 8 * 8, pad10 = (nz + 9) / 10 * 10;
            PZ = pad8 <= pad10 ? 8 : 10;
        }
    } else if (PZ != 2 && PZ != 4 && PZ != 5 && PZ != 6) {
        const double pad5 = (double)((nz + 4) / 5 * 5) / nz, pad4 = (double)((nz + 3) / 4 * 4) / nz;
        PZ = (pad5 <= pad4 * 1.02) ? 5 : 4;
        // meshes below ~1e5 nodes (17^3 ... 41^3) need the threads more than the sharing
        // (tools/lattice_small.py: 17^3 34 vs 43 us, 41^3 240 vs 270-304 us for the E-field)
        if (n_points < 100000) PZ = 2;
    }
    const int nzb = (nz + PZ - 1) / PZ;
    const long long n_items_ll = (long long)nx * ny * nzb;
    const int n_items = (int)n_items_ll;
    const int gx = (n_items + threads - 1) / threads;

    int splits = tu.k1_splits;
    if (splits <= 0 && gx < 2 * sms) {
        // small mesh: fill the chip by splitting the charge range (>= 64 pairs per split)
        splits = (2 * sms + gx - 1) / gx;
        const int cap = c->n_pairs / 64 > 0 ? c->n_pairs / 64 : 1;
        if (splits > cap) splits = cap;
    } else if (splits <= 0) {
        splits = 1;
        const double slots = 2.0 * sms;
        const int smax = c->n_pairs / 1024 > 16 ? 16 : c->n_pairs / 1024;
        double best = 1e30;
        for (int s = 1; s <= (smax < 1 ? 1 : smax); ++s) {
            const double w = (double)gx * s / slots;
            const double ineff = ceil(w) / w;
            if (ineff < best - 0.004) { best = ineff; splits = s; }
            if (ineff <= 1.02) break;
        }
    }
    const int ncomp = (mode == MODE_ESP) ? 1 : 3;      // FP64 partial sums per point and split
    while (splits > 1 && (size_t)splits * (size_t)n_points * 8u * ncomp > ((size_t)1 << 30)) --splits;
    int pps = (c->n_pairs + splits - 1) / splits;
    pps = ((pps + 7) / 8) * 8;
    if (pps < 8) pps = 8;
    splits = c->n_pairs > 0 ? (c->n_pairs + pps - 1) / pps : 1;
    int tile_pairs = tu.k1_tile_pairs > 0 ? tu.k1_tile_pairs : 1024;
    if (tile_pairs > pps) tile_pairs = pps;
    tile_pairs = ((tile_pairs + 7) / 8) * 8;
    int stages = tu.k1_stages > 0 ? tu.k1_stages : 3;
    if (stages > 8) stages = 8;
    const int ntiles = (pps + tile_pairs - 1) / tile_pairs;
    if (stages > ntiles) stages = ntiles;
    if (stages < 1) stages = 1;
    const size_t smem = 128 + (size_t)stages * tile_pairs * sizeof(ChargePair);

    K1LatParams prm;
    prm.charges = c->charges.as<ChargePair>();
    prm.n_pairs = c->n_pairs;
    prm.pairs_per_split = pps;
    prm.tile_pairs = tile_pairs;
    prm.stages = stages;
    prm.xs = d_xs; prm.ys = d_ys; prm.zs = d_zs;
    prm.nx = nx; prm.ny = ny; prm.nz = nz; prm.nzb = nzb;
    prm.n_items = n_items;
    prm.n_points = n_points;
    prm.out_kind = out_kind;
    prm.step = 0.f;
    prm.out = d_out;
    prm.partial = nullptr;
    prm.soft_flag = nullptr;
    if (splits > 1) {
        if (int rc = c->work0.reserve(sizeof(double) * ncomp * (size_t)splits * (size_t)n_points)) return rc;
        prm.partial = c->work0.as<double>();
    }
    int launches = 0;
    // softened meshes with >= ~1 ms of work: prove on the device that the softening cannot act and
    // run the unsoftened instantiation (same bits, two issue slots fewer per two pair-evaluations)
    const bool scan = mode == MODE_FIELD_SOFT &&
                      (tu.k1_softscan > 0 ||
                       (tu.k1_softscan < 0 && (double)n_points * (double)c->n_charges >= 2.0e9));
    if (scan && c->n_pairs > 0) {
        if (int rc = c->flags.reserve(64)) return rc;
        CPET_CUDA_TRY(cudaMemsetAsync(c->flags.p, 0, 64, c->stream));
        soften_scan_kernel<<<((c->n_pairs > nx + ny + nz ? c->n_pairs : nx + ny + nz) + 255) / 256, 256, 0, c->stream>>>(
            c->charges.as<ChargePair>(), c->n_pairs, d_xs, nx, d_ys, ny, d_zs, nz, c->flags.as<unsigned>());
        CPET_CUDA_TRY(cudaGetLastError());
        prm.soft_flag = c->flags.as<unsigned>();
        launches += 1;
    }
    KernelTimer timer(c);
    dim3 grid((unsigned)gx, (unsigned)splits, 1);
    int rc;
    const int U = tu.k1_unroll == 1 ? 1 : (tu.k1_unroll == 2 ? 2 : 4);
#define CPET_LAT_PZ(M, UU)                                                                  \
    (PZ == 2 ? launch_k1_lat_inst<M, 2, UU>(c, prm, grid, threads, smem)                     \
             : (PZ == 5 ? launch_k1_lat_inst<M, 5, UU>(c, prm, grid, threads, smem)          \
                        : launch_k1_lat_inst<M, 4, UU>(c, prm, grid, threads, smem)))
#define CPET_LAT_CASE(M) (U == 1 ? CPET_LAT_PZ(M, 1) : (U == 4 ? CPET_LAT_PZ(M, 4) : CPET_LAT_PZ(M, 2)))
#define CPET_LATN_PZ(M, UU)                                                                           \
    (PZ == 4 ? launch_k1_latn_inst<M, 4, UU>(c, prm, grid, threads, smem)                                \
             : (PZ == 6 ? launch_k1_latn_inst<M, 6, UU>(c, prm, grid, threads, smem)                     \
                        : (PZ == 8 ? launch_k1_latn_inst<M, 8, UU>(c, prm, grid, threads, smem)          \
                                   : launch_k1_latn_inst<M, 10, UU>(c, prm, grid, threads, smem))))
#define CPET_LATN_CASE(M) (U == 2 ? CPET_LATN_PZ(M, 2) : CPET_LATN_PZ(M, 4))
    if (nodes) {
        if (mode == MODE_FIELD_SOFT) {
            if (prm.soft_flag) {                   // runs only when the scan found nothing
                rc = CPET_LATN_CASE(MODE_FIELD_RAW);
                if (rc) return rc;
                launches += 1;
            }
            rc = CPET_LATN_CASE(MODE_FIELD_SOFT);
        } else rc = CPET_LATN_CASE(MODE_FIELD_RAW);
    } else if (mode == MODE_FIELD_SOFT) {
        if (prm.soft_flag) {                       // runs only when the scan found nothing
            rc = CPET_LAT_CASE(MODE_FIELD_RAW);
            if (rc) return rc;
            launches += 1;
        }
        rc = CPET_LAT_CASE(MODE_FIELD_SOFT);       // runs only when it found something (or no scan)
    } else if (mode == MODE_FIELD_RAW) rc = CPET_LAT_CASE(MODE_FIELD_RAW);
    else if (PZ == 6) {
        rc = U == 1 ? launch_k1_lat_inst<MODE_ESP, 6, 1, 1>(c, prm, grid, threads, smem)
                    : (U == 4 ? launch_k1_lat_inst<MODE_ESP, 6, 4, 1>(c, prm, grid, threads, smem)
                              : launch_k1_lat_inst<MODE_ESP, 6, 2, 1>(c, prm, grid, threads, smem));
    } else rc = CPET_LAT_CASE(MODE_ESP);
#undef CPET_LAT_CASE
#undef CPET_LAT_PZ
#undef CPET_LATN_CASE
#undef CPET_LATN_PZ
    if (rc) return rc;
    launches += 1;
    c->last_path = 1;
    if (splits > 1) {
        k1_lattice_finalize_kernel<<<(n_points + 255) / 256, 256, 0, c->stream>>>(
            prm.partial, splits, n_points, ncomp, d_xs, d_ys, d_zs, ny, nz, out_kind, 0.f, d_out);
        CPET_CUDA_TRY(cudaGetLastError());
        launches += 1;
    }
    c->last_counters[0] = launches;
    return CPET_OK;
}

template <int MODE, int P, int G>
static int launch_k1_inst(cpet_ctx* c, const K1Params& prm, dim3 grid, int threads, size_t smem) {
    auto kern = k1_grid_kernel<MODE, P, G>;
    CPET_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, threads, smem, c->stream>>>(prm);
    CPET_CUDA_TRY(cudaGetLastError());
    return CPET_OK;
}

template <int MODE>
static int launch_k1_mode(cpet_ctx* c, const K1Params& prm, int P, int G, dim3 grid, int threads,
                          size_t smem) {
    if (G == 1) {
        if (P == 4) return launch_k1_inst<MODE, 4, 1>(c, prm, grid, threads, smem);
        if (P == 2) return launch_k1_inst<MODE, 2, 1>(c, prm, grid, threads, smem);
        return launch_k1_inst<MODE, 1, 1>(c, prm, grid, threads, smem);
    }
    if (G == 8) return launch_k1_inst<MODE, 1, 8>(c, prm, grid, threads, smem);
    return launch_k1_inst<MODE, 1, 32>(c, prm, grid, threads, smem);
}

// ---------------------------------------------------------------------------------------------
// Lattice detection on a flat (N,3) point list, so that the reference-shaped entry point
// (compute_looped_field / cpet_field_grid, which only ever sees mesh.reshape(-1,3)) can take the
// lattice kernel by itself.  Exact: every point is compared against the inferred axes.
//   info[0] = nz (first index where x or y changes), info[1] = ny*nz (first index where x changes),
//   info[2] = 1 if every point equals (x[(i/nynz)*nynz], y[((i/nz)%ny)*nz], z[i%nz]).
// ---------------------------------------------------------------------------------------------
__global__ void lattice_probe_kernel(const float* __restrict__ p, int n, unsigned* __restrict__ info) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= 0 || i >= n) return;
    const bool dx = p[3 * (size_t)i] != p[0];
    const bool dy = p[3 * (size_t)i + 1] != p[1];
    // only the first change matters: a point whose predecessor already differs cannot be the minimum
    if (dx || dy) {
        const bool pdx = p[3 * (size_t)(i - 1)] != p[0];
        const bool pdy = p[3 * (size_t)(i - 1) + 1] != p[1];
        if (!(pdx || pdy)) atomicMin(&info[0], (unsigned)i);
        if (dx && !pdx) atomicMin(&info[1], (unsigned)i);
    }
}

__global__ void lattice_verify_kernel(const float* __restrict__ p, int n, unsigned* __restrict__ info) {
    const unsigned nz = info[0], nynz = info[1] == 0xffffffffu ? (unsigned)n : info[1];
    if (nz == 0xffffffffu || nz == 0 || nynz % nz != 0 || (unsigned)n % nynz != 0) {
        if (blockIdx.x == 0 && threadIdx.x == 0) info[2] = 0;
        return;
    }
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned ix = (unsigned)i / nynz, rem = (unsigned)i % nynz;
    const unsigned iy = rem / nz, iz = rem % nz;
    const bool ok = p[3 * (size_t)i] == p[3 * (size_t)(ix * nynz)] &&
                    p[3 * (size_t)i + 1] == p[3 * (size_t)(iy * nz) + 1] &&
                    p[3 * (size_t)i + 2] == p[3 * (size_t)iz + 2];
    if (!ok) info[2] = 0;
}

__global__ void lattice_axes_kernel(const float* __restrict__ p, int nx, int ny, int nz,
                                    float* __restrict__ axes) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nx) axes[i] = p[3 * (size_t)i * ny * nz];
    else if (i < nx + ny) axes[i] = p[3 * (size_t)(i - nx) * nz + 1];
    else if (i < nx + ny + nz) axes[i] = p[3 * (size_t)(i - nx - ny) + 2];
}

// Returns 1 in *is_lattice (and the sizes + a device array xs|ys|zs in c->work2) when the point list
// is a z-fastest tensor-product mesh with nz >= 4.  Synchronises the stream once (12-byte readback).
int detect_lattice(cpet_ctx* c, int n_points, const float* d_x0, int* is_lattice, int* nx, int* ny,
                   int* nz, const float** d_axes) {
    *is_lattice = 0;
    if (n_points < 4096) return CPET_OK;
    if (int rc = c->counters.reserve(64 + sizeof(unsigned) * 3 * 2048)) return rc;
    unsigned* info = reinterpret_cast<unsigned*>(c->counters.as<unsigned char>() + 16);
    const unsigned init[3] = {0xffffffffu, 0xffffffffu, 1u};
    CPET_CUDA_TRY(cudaMemcpyAsync(info, init, sizeof(init), cudaMemcpyHostToDevice, c->stream));
    const int blocks = (n_points + 255) / 256;
    lattice_probe_kernel<<<blocks, 256, 0, c->stream>>>(d_x0, n_points, info);
    lattice_verify_kernel<<<blocks, 256, 0, c->stream>>>(d_x0, n_points, info);
    CPET_CUDA_TRY(cudaGetLastError());
    unsigned h[3] = {0, 0, 0};
    CPET_CUDA_TRY(cudaMemcpyAsync(h, info, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    CPET_CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (h[2] != 1u || h[0] == 0xffffffffu || h[0] < 4) return CPET_OK;
    const unsigned nynz = h[1] == 0xffffffffu ? (unsigned)n_points : h[1];
    *nz = (int)h[0];
    *ny = (int)(nynz / h[0]);
    *nx = (int)((unsigned)n_points / nynz);
    const int na = *nx + *ny + *nz;
    if (int rc = c->work2.reserve(sizeof(float) * (size_t)na)) return rc;
    lattice_axes_kernel<<<(na + 255) / 256, 256, 0, c->stream>>>(d_x0, *nx, *ny, *nz, c->work2.as<float>());
    CPET_CUDA_TRY(cudaGetLastError());
    *d_axes = c->work2.as<float>();
    *is_lattice = 1;
    return CPET_OK;
}

// ---------------------------------------------------------------------------------------------
// K1, hybrid near/far form for point LISTS (default from 2,048 points, tuning key "k1_hybrid"): the streamline kernel's
// arithmetic (common.cuh: evalp_far / evalp_near, 10 packed FMA-pipe instructions per two pair-evaluations for
// charges far from the points) applied to the general field sum.  The charges are classified per call against the
// origin-centred bounding box of the point list (topo.cu: pack_hybrid; PyCPET's box frame is centred at the
// origin -- a point cloud elsewhere simply classifies every charge as near and runs the direct form) and packed
// as PBlocks [near | far].  A thread owns 4 points (2 packed pairs); every charge record is a broadcast LDS.128 +
// LDS.32.  FP32 partials per point are folded into FP64 every 64 charges; E = k (p S - T) in FP64 at the end.
// The charge (block) range can be split over gridDim.y exactly like the general kernel (FP64 partials + finalize).
// ---------------------------------------------------------------------------------------------
struct K1XParams {
    const PBlock* blocks;
    const K2XMeta* meta;
    int tile_blocks;
    int stages;
    const float* x0;
    int n_points;
    int out_kind;
    float step;
    void* out;
    double* partial;
};

template <bool SOFT>
__device__ __forceinline__ void evalp_near2(const float4 a, PRegs& r) {
    const u64 x2 = pk2(a.x, a.x), y2 = pk2(a.y, a.y), z2 = pk2(a.z, a.z), q4 = pk2(a.w, a.w);
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        const u64 dx = add2(r.c0[p], x2);          // -2 (p - x), exact
        const u64 dy = add2(r.c1[p], y2);
        const u64 dz = add2(r.c2[p], z2);
        u64 r2 = mul2(dx, dx);
        r2 = fma2(dy, dy, r2);
        r2 = fma2(dz, dz, r2);
        if (SOFT) {                                // max(r^2, 1e-6) (C:433-436) on |D|^2 = 4 r^2: an exact scaling
            float lo, hi;
            upk2(r2, lo, hi);
            r2 = pk2(fmaxf(lo, 4.0f * CPET_SOFT_EPS), fmaxf(hi, 4.0f * CPET_SOFT_EPS));
        }
        const u64 inv = rsqrt2(r2);
        const u64 s = mul2(mul2(inv, inv), mul2(inv, q4));
        r.a0[p] = fma2(s, dx, r.a0[p]);
        r.a1[p] = fma2(s, dy, r.a1[p]);
        r.a2[p] = fma2(s, dz, r.a2[p]);
    }
}

template <bool SOFT>
__global__ void __launch_bounds__(256) k1x_grid_kernel(const K1XParams prm) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
    PBlock* ring = reinterpret_cast<PBlock*>(smem_raw + 128);
    const int tid = threadIdx.x;
    const int S = prm.stages;
    const int TB = prm.tile_blocks;
    const int nb_total = prm.meta->nb_total;
    const int nb_near = prm.meta->nb_near;
    const int bps = (nb_total + (int)gridDim.y - 1) / (int)gridDim.y;      // blocks per split
    const int bbeg = min(nb_total, (int)blockIdx.y * bps);
    const int bend = min(nb_total, bbeg + bps);
    const int nblk = bend - bbeg;
    const int ntiles = (nblk + TB - 1) / TB;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int t) {
        const int stage = t % S;
        const int n_t = min(TB, nblk - t * TB);
        const uint32_t bytes = (uint32_t)n_t * (uint32_t)sizeof(PBlock);
        mbar_expect_tx(&full[stage], bytes);
        tma_load_1d(ring + (size_t)stage * TB, prm.blocks + bbeg + (size_t)t * TB, bytes, &full[stage]);
    };
    if (tid == 0) {
        const int pre = min(S, ntiles);
        for (int t = 0; t < pre; ++t) issue(t);
    }

    // this thread's 4 points = 2 packed pairs (positions 0,1 and 2,3)
    float px[4], py[4], pz[4];
    int pt[4];
    const int base = blockIdx.x * 1024;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        pt[p] = base + p * 256 + tid;
        const int ld = min(pt[p], prm.n_points - 1);
        px[p] = prm.x0[3 * (size_t)ld + 0];
        py[p] = prm.x0[3 * (size_t)ld + 1];
        pz[p] = prm.x0[3 * (size_t)ld + 2];
    }
    PRegs r;
    double acc[4][4];                  // per point: T.x - E_near.x, T.y - .., T.z - .., S
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        r.c0[j] = pk2(-2.0f * px[2 * j], -2.0f * px[2 * j + 1]);
        r.c1[j] = pk2(-2.0f * py[2 * j], -2.0f * py[2 * j + 1]);
        r.c2[j] = pk2(-2.0f * pz[2 * j], -2.0f * pz[2 * j + 1]);
        r.c3[j] = pk2(fmaf(pz[2 * j], pz[2 * j], fmaf(py[2 * j], py[2 * j], px[2 * j] * px[2 * j])),
                      fmaf(pz[2 * j + 1], pz[2 * j + 1], fmaf(py[2 * j + 1], py[2 * j + 1], px[2 * j + 1] * px[2 * j + 1])));
        r.a0[j] = r.a1[j] = r.a2[j] = r.a3[j] = 0ull;
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) acc[p][0] = acc[p][1] = acc[p][2] = acc[p][3] = 0.0;

    auto flush = [&]() {               // FP32 partials of the last <= 64 charges -> FP64
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            float lo, hi;
            upk2(r.a0[j], lo, hi); acc[2 * j][0] += (double)lo; acc[2 * j + 1][0] += (double)hi;
            upk2(r.a1[j], lo, hi); acc[2 * j][1] += (double)lo; acc[2 * j + 1][1] += (double)hi;
            upk2(r.a2[j], lo, hi); acc[2 * j][2] += (double)lo; acc[2 * j + 1][2] += (double)hi;
            upk2(r.a3[j], lo, hi); acc[2 * j][3] += (double)lo; acc[2 * j + 1][3] += (double)hi;
            r.a0[j] = r.a1[j] = r.a2[j] = r.a3[j] = 0ull;
        }
    };

    for (int t = 0; t < ntiles; ++t) {
        const int stage = t % S;
        mbar_wait(&full[stage], (uint32_t)((t / S) & 1));
        const int n_t = min(TB, nblk - t * TB);
        const PBlock* tile = ring + (size_t)stage * TB;
        for (int b = 0; b < n_t; ++b) {
            const int g = bbeg + t * TB + b;               // global block index: near blocks come first
            const PBlock& blk = tile[b];
            if (g < nb_near) {
#pragma unroll 2
                for (int e = 0; e < 32; ++e) evalp_near2<SOFT>(blk.a[e], r);
            } else {
#pragma unroll 8
                for (int e = 0; e < 32; ++e) evalp_far<2>(blk.a[e], blk.b[e], r);
            }
            if (b & 1) flush();                            // chains of at most 64 charges
        }
        flush();
        if (t + S < ntiles) {
            __syncthreads();           // every warp is done reading this stage
            if (tid == 0) issue(t + S);
        }
    }

#pragma unroll
    for (int p = 0; p < 4; ++p) {
        if (pt[p] >= prm.n_points) continue;
        // E = p*S - (T - E_near), up to the Coulomb constant (store_result applies it)
        const double ex = (double)px[p] * acc[p][3] - acc[p][0];
        const double ey = (double)py[p] * acc[p][3] - acc[p][1];
        const double ez = (double)pz[p] * acc[p][3] - acc[p][2];
        if (gridDim.y == 1) {
            store_result(prm.out_kind, prm.step, prm.out, pt[p], px[p], py[p], pz[p], ex, ey, ez);
        } else {
            double* o = prm.partial + ((size_t)blockIdx.y * prm.n_points + pt[p]) * 3;
            o[0] = ex; o[1] = ey; o[2] = ez;
        }
    }
}

static int launch_field_grid_hybrid(cpet_ctx* c, int mode, int n_points, const float* d_x0, int out_kind,
                                    void* d_out, float step) {
    const Tuning& tu = c->tune;
    const int sms = c->sm_count;
    int launches = 0;
    unsigned int* queue = nullptr;
    unsigned long long* evals = nullptr;
    const int32_t* order = nullptr;
    if (int rc = prepare_queue(c, n_points, nullptr, false, &queue, &evals, &order, &launches)) return rc;   // zeroes the meta block
    const int max_blocks = (c->n_charges + 31) / 32 + 2;
    K2XMeta* meta = nullptr;
    const float zero3[3] = {0.f, 0.f, 0.f};
    if (int rc = pack_hybrid(c, n_points, d_x0, 0.0f, zero3, 1, max_blocks, &meta, &launches)) return rc;

    const int gx = (n_points + 1023) / 1024;
    int splits = tu.k1_splits;
    if (splits <= 0) {
        // ~2 CTAs are resident per SM.  Fewer CTAs than slots: split the block range so that ONE round is as full
        // as it gets (100,000 points: 98 CTAs x 3 splits = 294 of 296 slots, 0.72 of the peak against 0.65 with 4
        // splits, which need a second round); more: make the last round nearly full, as the direct kernel does.
        splits = 1;
        const int slots = 2 * sms;
        if (gx < slots) {
            const int smax = max_blocks / 4 > 0 ? max_blocks / 4 : 1;          // >= 128 charges per split
            splits = slots / gx;
            if (splits > smax) splits = smax;
        } else {
            const int smax = max_blocks / 64 > 16 ? 16 : max_blocks / 64;
            double best = 1e30;
            for (int s = 1; s <= (smax < 1 ? 1 : smax); ++s) {
                const double w = (double)gx * s / slots;
                const double ineff = ceil(w) / w;
                if (ineff < best - 0.004) { best = ineff; splits = s; }
                if (ineff <= 1.02) break;
            }
        }
    }
    while (splits > 1 && (size_t)splits * (size_t)n_points * 24u > ((size_t)1 << 30)) --splits;
    if (splits > max_blocks / 2) splits = max_blocks / 2 > 0 ? max_blocks / 2 : 1;
    if (splits > 65535) splits = 65535;

    K1XParams prm;
    prm.blocks = c->xblocks.as<PBlock>();
    prm.meta = meta;
    prm.tile_blocks = tu.k1_tile_pairs > 0 ? (tu.k1_tile_pairs / 16 > 0 ? tu.k1_tile_pairs / 16 : 1) : 48;   // 30 KB
    prm.stages = tu.k1_stages > 0 ? (tu.k1_stages > 8 ? 8 : tu.k1_stages) : 3;
    prm.x0 = d_x0;
    prm.n_points = n_points;
    prm.out_kind = out_kind;
    prm.step = step;
    prm.out = d_out;
    prm.partial = nullptr;
    if (splits > 1) {
        if (int rc = c->work0.reserve(sizeof(double) * 3 * (size_t)splits * (size_t)n_points)) return rc;
        prm.partial = c->work0.as<double>();
    }
    const size_t smem = 128 + (size_t)prm.stages * prm.tile_blocks * sizeof(PBlock);
    KernelTimer timer(c);
    dim3 grid((unsigned)gx, (unsigned)splits, 1);
    if (mode == MODE_FIELD_SOFT) {
        auto kern = k1x_grid_kernel<true>;
        CPET_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, 256, smem, c->stream>>>(prm);
    } else {
        auto kern = k1x_grid_kernel<false>;
        CPET_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, 256, smem, c->stream>>>(prm);
    }
    CPET_CUDA_TRY(cudaGetLastError());
    launches += 1;
    if (splits > 1) {
        k1_finalize_kernel<<<(n_points + 255) / 256, 256, 0, c->stream>>>(prm.partial, splits, n_points, 3, d_x0,
                                                                          out_kind, step, d_out);
        CPET_CUDA_TRY(cudaGetLastError());
        launches += 1;
    }
    c->last_counters[0] = launches;
    c->last_path = 3;
    return CPET_OK;
}

int launch_field_grid(cpet_ctx* c, int mode, int n_points, const float* d_x0, int out_kind,
                      void* d_out, float step) {
    CPET_REQUIRE(c->charges_set, CPET_ERR_STATE, "no charge set on this context: call cpet_set_charges first");
    c->last_counters[0] = 0;
    c->last_counters[1] = (int64_t)n_points * (int64_t)c->n_charges;
    c->last_counters[2] = n_points;
    if (n_points == 0) return CPET_OK;
    const Tuning& tu = c->tune;
    const int sms = c->sm_count;
    // Field sums over point lists of >= 2,048 points take the hybrid near/far kernel (1e6 random points x 100,000
    // charges: 0.78 of the FP32 peak against 0.64 softened / 0.68 raw for the direct form below; 5,000 points x 30,000
    // charges 0.53 against 0.30; tools/k1_hybrid_ab.py, profiles/round2_k1_hybrid.txt); ESP and short lists stay here.
    if ((tu.k1_hybrid > 0 || (tu.k1_hybrid < 0 && c->n_charges >= 256)) && mode != MODE_ESP && n_points >= 2048 &&
        c->n_charges > 0)
        return launch_field_grid_hybrid(c, mode, n_points, d_x0, out_kind, d_out, step);

    int threads = tu.k1_threads > 0 ? tu.k1_threads : 256;
    if (threads > 256) threads = 256;
    threads = (threads / 32) * 32;
    if (threads < 32) threads = 32;

    // Lanes per point.  Measured (tools/k1_midsize.py, profiles/round1_k1_midsize.md): from ~2,000
    // points up one thread per point with the charge range split over gridDim.y beats G lanes per
    // point by 1.1-2.6x (G distinct addresses per LDS.128 and a butterfly per point cost more than
    // the FP64 finalize); G = 32 only for tiny lists such as the 11^3 grid of example 2A.
    int G = tu.k1_lanes;
    if (G <= 0) G = (n_points >= 2048) ? 1 : 32;
    if (G != 1 && G != 8 && G != 32) G = (G < 8) ? 8 : 32;
    int P = tu.k1_points;
    if (G > 1) P = 1;
    else if (P <= 0) {
        // points per thread: register blocking pays once there are enough points to fill the chip
        const long long n = n_points;
        if (mode == MODE_ESP) P = (n >= 16384) ? 2 : 1;                 // MUFU-bound: 2 is enough
        else P = (n >= (long long)sms * 1024) ? 4 : ((n >= 16384) ? 2 : 1);
    }
    if (P != 1 && P != 2 && P != 4) P = 2;

    const int pts_per_cta = (threads / G) * P;
    const int gx = (n_points + pts_per_cta - 1) / pts_per_cta;

    // split the charge range when the point dimension alone cannot fill 2 CTAs per SM
    int splits = tu.k1_splits;
    if (splits <= 0) {
        splits = 1;
        if (gx < 2 * sms) {
            splits = (2 * sms + gx - 1) / gx;
        } else {
            // Wave quantisation: with ~2 resident CTAs/SM a grid of gx equal CTAs takes
            // ceil(gx / (2*sms)) rounds.  Splitting the charge range makes more, shorter CTAs so the
            // last round is nearly full (measured: 1007 CTAs 2.46e12 -> 16 splits 2.54e12
            // pair-evals/s, profiles/round1_tail_test.txt).  Keep >= 1024 pairs per split.
            const double slots = 2.0 * sms;
            const int smax = c->n_pairs / 1024 > 16 ? 16 : c->n_pairs / 1024;
            double best = 1e30;
            for (int s = 1; s <= (smax < 1 ? 1 : smax); ++s) {
                const double w = (double)gx * s / slots;
                const double ineff = ceil(w) / w;
                if (ineff < best - 0.004) { best = ineff; splits = s; }
                if (ineff <= 1.02) break;
            }
        }
    }
    // FP64 partials cost 24 B (ESP: 8 B) per point per split: keep that scratch under 1 GiB
    const int ncomp = (mode == MODE_ESP) ? 1 : 3;
    while (splits > 1 && (size_t)splits * (size_t)n_points * 8u * ncomp > ((size_t)1 << 30)) --splits;
    const int min_pairs_per_split = 64;
    int max_splits = c->n_pairs / min_pairs_per_split;
    if (max_splits < 1) max_splits = 1;
    if (splits > max_splits) splits = max_splits;
    if (splits > 65535) splits = 65535;
    int pps = (c->n_pairs + splits - 1) / splits;
    pps = ((pps + 7) / 8) * 8;
    if (pps < 8) pps = 8;
    splits = c->n_pairs > 0 ? (c->n_pairs + pps - 1) / pps : 1;

    int tile_pairs = tu.k1_tile_pairs > 0 ? tu.k1_tile_pairs : (mode == MODE_ESP ? 512 : 1024);
    if (tile_pairs > pps) tile_pairs = pps;
    tile_pairs = ((tile_pairs + 7) / 8) * 8;
    int stages = tu.k1_stages > 0 ? tu.k1_stages : (mode == MODE_ESP ? 2 : 3);
    if (stages > 8) stages = 8;
    const int ntiles = (pps + tile_pairs - 1) / tile_pairs;
    if (stages > ntiles) stages = ntiles;
    if (stages < 1) stages = 1;
    size_t smem = 128 + (size_t)stages * tile_pairs * sizeof(ChargePair);
    while (smem > (size_t)c->max_smem_optin && tile_pairs > 64) {
        tile_pairs /= 2;
        smem = 128 + (size_t)stages * tile_pairs * sizeof(ChargePair);
    }

    K1Params prm;
    prm.charges = c->charges.as<ChargePair>();
    prm.n_pairs = c->n_pairs;
    prm.pairs_per_split = pps;
    prm.tile_pairs = tile_pairs;
    prm.stages = stages;
    prm.x0 = d_x0;
    prm.n_points = n_points;
    prm.out_kind = out_kind;
    prm.step = step;
    prm.out = d_out;
    prm.partial = nullptr;
    if (splits > 1) {
        if (int rc = c->work0.reserve(sizeof(double) * ncomp * (size_t)splits * (size_t)n_points)) return rc;
        prm.partial = c->work0.as<double>();
    }

    KernelTimer timer(c);
    dim3 grid((unsigned)gx, (unsigned)splits, 1);
    int rc;
    if (mode == MODE_FIELD_SOFT) rc = launch_k1_mode<MODE_FIELD_SOFT>(c, prm, P, G, grid, threads, smem);
    else if (mode == MODE_FIELD_RAW) rc = launch_k1_mode<MODE_FIELD_RAW>(c, prm, P, G, grid, threads, smem);
    else rc = launch_k1_mode<MODE_ESP>(c, prm, P, G, grid, threads, smem);
    if (rc) return rc;
    c->last_counters[0] = 1;
    c->last_path = 0;
    if (splits > 1) {
        k1_finalize_kernel<<<(n_points + 255) / 256, 256, 0, c->stream>>>(
            prm.partial, splits, n_points, ncomp, d_x0, out_kind, step, d_out);
        CPET_CUDA_TRY(cudaGetLastError());
        c->last_counters[0] = 2;
    }
    return CPET_OK;
}

// ---------------------------------------------------------------------------------------------
// FP32 peak probe: register-resident FMA chains, no memory traffic.
// ---------------------------------------------------------------------------------------------
template <int PACKED>
__global__ void __launch_bounds__(256) fp32_probe_kernel(int iters, float seed, float* sink) {
    constexpr int CH = 8;
    if (PACKED) {
        u64 v[CH];
        const u64 a = pk2(1.0000001f, 0.9999999f), b = pk2(seed, -seed);
#pragma unroll
        for (int i = 0; i < CH; ++i) v[i] = pk2(seed + i, seed - i);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < CH; ++i) v[i] = fma2(v[i], a, b);
        }
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < CH; ++i) { float lo, hi; upk2(v[i], lo, hi); s += lo + hi; }
        if (s == 123.456f) sink[0] = s;
    } else {
        float v[2 * CH];
        const float a = 1.0000001f, b = seed;
#pragma unroll
        for (int i = 0; i < 2 * CH; ++i) v[i] = seed + i;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 2 * CH; ++i) v[i] = fmaf(v[i], a, b);
        }
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 2 * CH; ++i) s += v[i];
        if (s == 123.456f) sink[0] = s;
    }
}

int launch_fp32_probe(cpet_ctx* c, int packed, int iters, double* tflops) {
    if (iters <= 0) iters = 4096;
    if (int rc = c->work1.reserve(256)) return rc;
    const int blocks = c->sm_count * 8, threads = 256;
    cudaEvent_t e0, e1;
    CPET_CUDA_TRY(cudaEventCreate(&e0));
    CPET_CUDA_TRY(cudaEventCreate(&e1));
    float best_ms = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        CPET_CUDA_TRY(cudaEventRecord(e0, c->stream));
        if (packed) fp32_probe_kernel<1><<<blocks, threads, 0, c->stream>>>(iters, 0.5f, c->work1.as<float>());
        else fp32_probe_kernel<0><<<blocks, threads, 0, c->stream>>>(iters, 0.5f, c->work1.as<float>());
        CPET_CUDA_TRY(cudaEventRecord(e1, c->stream));
        CPET_CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0.f;
        CPET_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best_ms) best_ms = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    // per thread: iters * 4 * 16 FMAs (8 packed chains x 2 lanes, or 16 scalar chains)
    const double fmas = (double)blocks * threads * (double)iters * 4.0 * 16.0;
    *tflops = 2.0 * fmas / (best_ms * 1e-3) / 1e12;
    return CPET_OK;
}

}  // namespace cpet
