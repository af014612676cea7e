// common.cuh -- device-side building blocks shared by the sm_100a kernels.
//
//  * packed FP32x2 arithmetic (Blackwell FADD2/FMUL2/FFMA2): two point-charge pairs per
//    instruction, so the FP32 pipe is fed with half the issue slots;
//  * mbarrier + 1-D TMA bulk copy (cp.async.bulk -> UBLKCP) used to stage charge tiles;
//  * the pair-evaluation inner loops for the three flavours of the Coulomb sum.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cpet {

typedef unsigned long long u64;

// Coulomb factor and softening exactly as the reference stores them in `float`
// (CPET/utils/math_module.c:412 and :407).
#define CPET_COULOMB_K 14.3996451f
#define CPET_SOFT_EPS 1e-6f

// One charge pair (charges 2j and 2j+1) = 32 bytes = two 16-byte shared-memory vectors:
//   a = { (-x0,-x1), (-y0,-y1) }    b = { (-z0,-z1), (q0,q1) }
// Coordinates are stored negated so that d = p + (-x) is a single packed add.
struct __align__(16) PairA { u64 nx, ny; };
struct __align__(16) PairB { u64 nz, q; };
struct __align__(32) ChargePair { PairA a; PairB b; };
static_assert(sizeof(ChargePair) == 32, "charge pair must be 32 bytes");

// The same pairs in blocks of 32, structure-of-arrays inside a block (1 KB): lane l of a warp reads
// a[l] and b[l] with two conflict-free LDS.128 (32 x 16 contiguous bytes each).  Used by the
// warp-wide streamline kernel, where the 32 lanes of a warp split the charges of a frame.
struct __align__(16) ChargeBlock { PairA a[32]; PairB b[32]; };
static_assert(sizeof(ChargeBlock) == 1024, "charge block must be 1 KB");

// A padding charge: q = 0 at a far but finite position (no inf*0 in the raw-field mode).
#define CPET_PAD_COORD 1.0e8f

enum FieldMode : int { MODE_FIELD_SOFT = 0, MODE_FIELD_RAW = 1, MODE_ESP = 2 };

__device__ __forceinline__ u64 pk2(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(u64 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
    u64 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
    u64 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
// MUFU.RSQ.  The .ftz form is a single instruction; without it ptxas wraps every rsqrt in a
// denormal fix-up (FSETP + 2 predicated FMUL), which costs 3 extra issue slots per pair.  A
// denormal r^2 (< 1.2e-38, i.e. r < 1e-19 A) is a coincident point/charge either way: flushed to
// 0 it yields +inf, exactly what q*r^-3 overflows to in the reference's float arithmetic.
__device__ __forceinline__ float rsqrt_approx(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// ------------------------------------------------------------------------------------------
// mbarrier + TMA 1-D bulk copy
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy, completion signalled on `bar` (bytes multiple of 16, both
// addresses 16-byte aligned).
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                            uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ------------------------------------------------------------------------------------------
// Pair evaluation.  P evaluation points per thread (coordinates duplicated into both halves
// of a packed register), charge pairs read as broadcast LDS.128.  Accumulators are packed
// {even-charge sum, odd-charge sum}.
// ------------------------------------------------------------------------------------------
template <int P>
struct PointRegs {
    u64 px[P], py[P], pz[P];      // {p,p}
    u64 ax[P], ay[P], az[P];      // FP32 partial sums for the current tile (az unused for ESP)
};

template <int MODE, int P>
__device__ __forceinline__ void eval_pair(const PairA a, const PairB b, PointRegs<P>& r) {
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const u64 dx = add2(r.px[p], a.nx);
        const u64 dy = add2(r.py[p], a.ny);
        const u64 dz = add2(r.pz[p], b.nz);
        u64 r2 = mul2(dx, dx);
        r2 = fma2(dy, dy, r2);
        r2 = fma2(dz, dz, r2);
        float r2a, r2b;
        upk2(r2, r2a, r2b);
        if (MODE == MODE_FIELD_SOFT) {
            r2a = fmaxf(r2a, CPET_SOFT_EPS);
            r2b = fmaxf(r2b, CPET_SOFT_EPS);
        }
        const u64 inv = pk2(rsqrt_approx(r2a), rsqrt_approx(r2b));
        if (MODE == MODE_ESP) {
            r.ax[p] = fma2(inv, b.q, r.ax[p]);
        } else {
            const u64 t = mul2(inv, inv);
            const u64 u = mul2(inv, b.q);
            const u64 s = mul2(t, u);
            r.ax[p] = fma2(s, dx, r.ax[p]);
            r.ay[p] = fma2(s, dy, r.ay[p]);
            r.az[p] = fma2(s, dz, r.az[p]);
        }
    }
}

// Accumulate pairs [first, n) with stride `stride` of one shared-memory tile.
template <int MODE, int P, int UNROLL>
__device__ __forceinline__ void eval_tile(const ChargePair* __restrict__ tile, int first, int n,
                                          int stride, PointRegs<P>& r) {
#pragma unroll UNROLL
    for (int j = first; j < n; j += stride) {
        const PairA a = tile[j].a;
        const PairB b = tile[j].b;
        eval_pair<MODE, P>(a, b, r);
    }
}

template <int MODE, int P>
__device__ __forceinline__ void flush_partials(PointRegs<P>& r, double (&acc)[P][3]);

// Same, but the FP32 partial sums are folded into the FP64 accumulators every CHUNK pairs per
// lane, so no FP32 chain is longer than CHUNK additions (error ~ sqrt(CHUNK) * 2^-24 of the
// running partial sum; matters for the heavily cancelling ESP sum of a neutral system).
template <int MODE, int P, int UNROLL, int CHUNK>
__device__ __forceinline__ void eval_tile_chunked(const ChargePair* __restrict__ tile, int first,
                                                  int n, int stride, PointRegs<P>& r,
                                                  double (&acc)[P][3]) {
    for (int j0 = first; j0 < n; j0 += CHUNK * stride) {
        const int j1 = min(n, j0 + CHUNK * stride);
        eval_tile<MODE, P, UNROLL>(tile, j0, j1, stride, r);
        flush_partials<MODE, P>(r, acc);
    }
}

template <int P>
__device__ __forceinline__ void clear_partials(PointRegs<P>& r) {
#pragma unroll
    for (int p = 0; p < P; ++p) r.ax[p] = r.ay[p] = r.az[p] = 0ull;
}

// FP32 tile partials -> FP64 running sums (3 doubles per point; ESP uses [0] only).
template <int MODE, int P>
__device__ __forceinline__ void flush_partials(PointRegs<P>& r, double (&acc)[P][3]) {
#pragma unroll
    for (int p = 0; p < P; ++p) {
        float lo, hi;
        upk2(r.ax[p], lo, hi);
        acc[p][0] += (double)lo + (double)hi;
        if (MODE != MODE_ESP) {
            upk2(r.ay[p], lo, hi);
            acc[p][1] += (double)lo + (double)hi;
            upk2(r.az[p], lo, hi);
            acc[p][2] += (double)lo + (double)hi;
        }
    }
    clear_partials<P>(r);
}

template <int P>
__device__ __forceinline__ void set_point(PointRegs<P>& r, int p, float x, float y, float z) {
    r.px[p] = pk2(x, x);
    r.py[p] = pk2(y, y);
    r.pz[p] = pk2(z, z);
}

// ------------------------------------------------------------------------------------------
// Lattice variant (box grids are tensor products xs x ys x zs with z fastest,
// CPET/utils/calculator.py:218-233): a thread owns PZ consecutive z-nodes of one (x,y) column, so
// dx, dy and dx^2+dy^2 are computed once per charge pair and shared by the PZ points:
// 4 + 8*PZ packed FMA-pipe instructions per PZ points instead of 12*PZ (PZ=4: 9 vs 12 per point).
// The per-point arithmetic (order of the r^2 sum, s = (inv*inv)*(inv*q), accumulation) is the same
// as eval_pair, so results are bit-identical to the general kernel.
// ------------------------------------------------------------------------------------------
template <int PZ>
struct LatticeRegs {
    u64 px, py;                   // {x,x}, {y,y} of the column
    u64 pz[PZ];                   // {z,z} of the PZ nodes
    u64 ax[PZ], ay[PZ], az[PZ];
};

// 1/sqrt of two values WITHOUT the special-function unit: integer seed on the ALU pipe (0x5f3759df - (bits >> 1),
// 3.4 % off at most), three Newton steps y <- y (1.5 - x/2 y^2) on the FMA pipe (1.7e-3, 4.4e-6, then the FP32
// rounding level: measured against MUFU.RSQ and float64 in tests/test_gpu_parity.py).  10 packed instructions for
// two values against two MUFU.RSQ (8 cycles of the 16-lane XU pipe each): the ESP sum needs only 3.8 packed
// instructions per two pair-evaluations besides the rsqrt, so it is bound by the XU pipe (96 % busy, FMA pipe under
// half); giving every sixth z-node of a thread to this routine balances the two pipes (k1_lattice_kernel<ESP, 6, U, 1>).
__device__ __forceinline__ u64 rsqrt2_fma(u64 x) {
    float a, b;
    upk2(x, a, b);
    u64 y = pk2(__int_as_float(0x5f3759df - (__float_as_int(a) >> 1)),
                __int_as_float(0x5f3759df - (__float_as_int(b) >> 1)));
    const u64 hx = mul2(x, pk2(-0.5f, -0.5f));
    const u64 c15 = pk2(1.5f, 1.5f);
#pragma unroll
    for (int it = 0; it < 3; ++it) y = mul2(y, fma2(hx, mul2(y, y), c15));
    return y;
}

template <int MODE, int PZ, int NF = 0>
__device__ __forceinline__ void eval_pair_lattice(const PairA a, const PairB b, LatticeRegs<PZ>& r) {
    const u64 dx = add2(r.px, a.nx);
    const u64 dy = add2(r.py, a.ny);
    const u64 rxy = fma2(dy, dy, mul2(dx, dx));
#pragma unroll
    for (int p = 0; p < PZ; ++p) {
        const u64 dz = add2(r.pz[p], b.nz);
        const u64 r2 = fma2(dz, dz, rxy);
        float r2a, r2b;
        upk2(r2, r2a, r2b);
        if (MODE == MODE_FIELD_SOFT) {
            r2a = fmaxf(r2a, CPET_SOFT_EPS);
            r2b = fmaxf(r2b, CPET_SOFT_EPS);
        }
        const u64 inv = (MODE == MODE_ESP && p < NF) ? rsqrt2_fma(r2) : pk2(rsqrt_approx(r2a), rsqrt_approx(r2b));
        if (MODE == MODE_ESP) {
            r.ax[p] = fma2(inv, b.q, r.ax[p]);
        } else {
            const u64 t = mul2(inv, inv);
            const u64 u = mul2(inv, b.q);
            const u64 s = mul2(t, u);
            r.ax[p] = fma2(s, dx, r.ax[p]);
            r.ay[p] = fma2(s, dy, r.ay[p]);
            r.az[p] = fma2(s, dz, r.az[p]);
        }
    }
}

template <int MODE, int PZ, int UNROLL, int CHUNK, int NF = 0>
__device__ __forceinline__ void eval_tile_lattice_chunked(const ChargePair* __restrict__ tile, int n,
                                                          LatticeRegs<PZ>& r, double (&acc)[PZ][3]) {
    for (int j0 = 0; j0 < n; j0 += CHUNK) {
        const int j1 = min(n, j0 + CHUNK);
#pragma unroll UNROLL
        for (int j = j0; j < j1; ++j) {
            const PairA a = tile[j].a;
            const PairB b = tile[j].b;
            eval_pair_lattice<MODE, PZ, NF>(a, b, r);
        }
#pragma unroll
        for (int p = 0; p < PZ; ++p) {
            float lo, hi;
            upk2(r.ax[p], lo, hi);
            acc[p][0] += (double)lo + (double)hi;
            if (MODE != MODE_ESP) {
                upk2(r.ay[p], lo, hi);
                acc[p][1] += (double)lo + (double)hi;
                upk2(r.az[p], lo, hi);
                acc[p][2] += (double)lo + (double)hi;
            }
            r.ax[p] = r.ay[p] = r.az[p] = 0ull;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Lattice variant with two z-NODES per packed register (k1_lattice_nodes_kernel): the same arithmetic per point
// and charge -- r^2 = dz^2 + (dy^2 + dx^2), s = (inv*inv)*(inv*q), E += s*d -- but a thread's PZ nodes are packed
// in pairs and the charges of a pair are consumed one after the other through 32-bit broadcast operands: dx, dy and
// dx^2+dy^2 are scalars per charge and column, and only the z accumulation still reads three 64-bit registers
// (the streamline kernel's lesson, profiles/round2_ubench2.txt; lattice loops in tools/ubench3.cu).  A node's sum
// runs over the charges in index order (one FP32 chain instead of an even and an odd one), so the results differ
// from the charge-pair form in the last bits.
// ------------------------------------------------------------------------------------------
template <int PZ>
struct LatticeNodeRegs {
    float x, y;                                   // the column
    u64 pz[PZ / 2];                               // {z_2j, z_2j+1}
    u64 ax[PZ / 2], ay[PZ / 2], az[PZ / 2];       // FP32 partials of nodes 2j (low half) and 2j+1 (high half)
};

template <int MODE, int PZ>
__device__ __forceinline__ void eval_pair_lattice_nodes(const PairA a, const PairB b, LatticeNodeRegs<PZ>& r) {
    float nx[2], ny[2], nz[2], q[2];
    upk2(a.nx, nx[0], nx[1]); upk2(a.ny, ny[0], ny[1]); upk2(b.nz, nz[0], nz[1]); upk2(b.q, q[0], q[1]);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const float dx = r.x + nx[h], dy = r.y + ny[h];
        const float rxy = fmaf(dy, dy, dx * dx);
#pragma unroll
        for (int p = 0; p < PZ / 2; ++p) {
            const u64 dz = add2(r.pz[p], pk2(nz[h], nz[h]));
            const u64 r2 = fma2(dz, dz, pk2(rxy, rxy));
            float r2a, r2b;
            upk2(r2, r2a, r2b);
            if (MODE == MODE_FIELD_SOFT) {
                r2a = fmaxf(r2a, CPET_SOFT_EPS);
                r2b = fmaxf(r2b, CPET_SOFT_EPS);
            }
            const u64 inv = pk2(rsqrt_approx(r2a), rsqrt_approx(r2b));
            const u64 s = mul2(mul2(inv, inv), mul2(inv, pk2(q[h], q[h])));
            r.ax[p] = fma2(s, pk2(dx, dx), r.ax[p]);
            r.ay[p] = fma2(s, pk2(dy, dy), r.ay[p]);
            r.az[p] = fma2(s, dz, r.az[p]);
        }
    }
}

template <int MODE, int PZ, int UNROLL, int CHUNK>
__device__ __forceinline__ void eval_tile_lattice_nodes_chunked(const ChargePair* __restrict__ tile, int n,
                                                                LatticeNodeRegs<PZ>& r, double (&acc)[PZ][3]) {
    for (int j0 = 0; j0 < n; j0 += CHUNK) {
        const int j1 = min(n, j0 + CHUNK);
#pragma unroll UNROLL
        for (int j = j0; j < j1; ++j) {
            const PairA a = tile[j].a;
            const PairB b = tile[j].b;
            eval_pair_lattice_nodes<MODE, PZ>(a, b, r);
        }
#pragma unroll
        for (int p = 0; p < PZ / 2; ++p) {
            float lo, hi;
            upk2(r.ax[p], lo, hi); acc[2 * p][0] += (double)lo; acc[2 * p + 1][0] += (double)hi;
            upk2(r.ay[p], lo, hi); acc[2 * p][1] += (double)lo; acc[2 * p + 1][1] += (double)hi;
            upk2(r.az[p], lo, hi); acc[2 * p][2] += (double)lo; acc[2 * p + 1][2] += (double)hi;
            r.ax[p] = r.ay[p] = r.az[p] = 0ull;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Streamline kernel, round 2: hybrid pair evaluation (topo.cu, k2x_topo_kernel).
//
// Every point of a streamline lies inside the sampling box inflated by three steps, so a charge
// far from that region can be evaluated in an *expanded, charge-scaled* form that needs 10 packed
// FMA-pipe instructions per two pair-evaluations instead of 12:
//     alpha = s/q^2,  a = alpha*x,  b = alpha*|x|^2,  s = sgn(q)   (per charge, packed once per launch)
//     t = alpha*|p-x|^2 = alpha*|p|^2 + b - 2 p.a                  (4 FFMA2; t carries the sign of q)
//     u = |t|^(-3/2) = |q|^3 / |p-x|^3                             (MUFU.RSQ of |t| -- the absolute value is
//                                                                   an operand modifier --, 2 FMUL2)
//     S += u*alpha,  T += u*a                                      (4 FFMA2: sums of q/r^3 and q x/r^3)
//     E = p*S - T                                                  (FP64, once per pass)
// The sign of the charge rides on the whole record, so negative and positive charges run through
// the same loop and the same accumulators.
// The expansion loses about (|x|+|p|)^2 / |p-x|^2 ulps in t, so only charges for which that
// amplification is bounded (<= 8 by default) take this form; the few charges close to the box keep
// the direct form d = p - x, r^2 = d.d, E += q r^-3 d (12 instructions), sharing the point
// registers (-2p): the near record stores (2x, 2y, 2z, 4q) and evaluates D = -2p + 2x = -2d exactly
// (powers of two), q' D / |D|^3 = -q d / |d|^3, i.e. the near field enters T negatively.
// ------------------------------------------------------------------------------------------
struct __align__(16) V16 { u64 a, b; };
// 32 charge pairs, structure-of-arrays: lane l reads v0[l], v1[l] (LDS.128) and v2[l] (LDS.64).
//   far  block: v0 = {ax, ay}   v1 = {az, b}    v2 = alpha        (each entry {even, odd charge})
//   near block: v0 = {2x, 2y}   v1 = {2z, 4q}   v2 unused
struct __align__(16) XBlock { V16 v0[32]; V16 v1[32]; u64 v2[32]; };
static_assert(sizeof(XBlock) == 1280, "hybrid charge block must be 1280 bytes");

#define CPET_X_PAD_B 1.0e30f          // far pad: a = 0, b = 1e30, alpha = 0  ->  u ~ 1e-45, u*alpha = u*a = 0
#define CPET_X_MIN_ABS_Q 1.0e-12f     // smaller |q| (and q = 0 near the box) keep the direct form

template <int P>
struct XRegs {
    float c0[P], c1[P], c2[P], c3[P];   // -2p.x, -2p.y, -2p.z, |p|^2 (ptxas feeds them as broadcast .F32 operands)
    u64 a0[P], a1[P], a2[P], a3[P];     // FP32 partials {even, odd}: T.x - E_near.x, T.y - .., T.z - .., S
};

template <int P>
__device__ __forceinline__ void set_point_x(XRegs<P>& r, int p, float x, float y, float z) {
    r.c0[p] = -2.0f * x;
    r.c1[p] = -2.0f * y;
    r.c2[p] = -2.0f * z;
    r.c3[p] = fmaf(z, z, fmaf(y, y, x * x));
}
__device__ __forceinline__ u64 rsqrt2(u64 t) {
    float lo, hi;
    upk2(t, lo, hi);
    return pk2(rsqrt_approx(lo), rsqrt_approx(hi));
}
__device__ __forceinline__ u64 rsqrt2_abs(u64 t) {       // MUFU.RSQ |x|: the sign of the charge rides on t
    float lo, hi;
    upk2(t, lo, hi);
    return pk2(rsqrt_approx(fabsf(lo)), rsqrt_approx(fabsf(hi)));
}

template <int PE, int P>
__device__ __forceinline__ void evalx_far(const V16 v0, const V16 v1, const u64 al, XRegs<P>& r) {
#pragma unroll
    for (int p = 0; p < PE; ++p) {
        u64 t = fma2(al, pk2(r.c3[p], r.c3[p]), v1.b);
        t = fma2(pk2(r.c0[p], r.c0[p]), v0.a, t);
        t = fma2(pk2(r.c1[p], r.c1[p]), v0.b, t);
        t = fma2(pk2(r.c2[p], r.c2[p]), v1.a, t);
        const u64 inv = rsqrt2_abs(t);
        const u64 u = mul2(mul2(inv, inv), inv);
        r.a3[p] = fma2(u, al, r.a3[p]);
        r.a0[p] = fma2(u, v0.a, r.a0[p]);
        r.a1[p] = fma2(u, v0.b, r.a1[p]);
        r.a2[p] = fma2(u, v1.a, r.a2[p]);
    }
}

template <int PE, int P>
__device__ __forceinline__ void evalx_near(const V16 v0, const V16 v1, XRegs<P>& r) {
#pragma unroll
    for (int p = 0; p < PE; ++p) {
        const u64 dx = add2(pk2(r.c0[p], r.c0[p]), v0.a);
        const u64 dy = add2(pk2(r.c1[p], r.c1[p]), v0.b);
        const u64 dz = add2(pk2(r.c2[p], r.c2[p]), v1.a);
        u64 r2 = mul2(dx, dx);
        r2 = fma2(dy, dy, r2);
        r2 = fma2(dz, dz, r2);
        const u64 inv = rsqrt2(r2);
        const u64 s = mul2(mul2(inv, inv), mul2(inv, v1.b));
        r.a0[p] = fma2(s, dx, r.a0[p]);
        r.a1[p] = fma2(s, dy, r.a1[p]);
        r.a2[p] = fma2(s, dz, r.a2[p]);
    }
}

// ------------------------------------------------------------------------------------------
// Streamline kernel, points-packed hybrid form (topo8.cu, k2p_topo_kernel).
//
// The same two arithmetic forms, with the roles of the packed halves swapped: a packed register holds TWO
// POINTS of the warp and a lane takes ONE charge per step, whose record is consumed through 32-bit broadcast
// operands (FFMA2 R, R.F32x2, R.F32, R.F32x2).  The instruction count is unchanged (10 per two
// pair-evaluations in the far form); what changes is the register-file traffic per instruction: no FFMA2 of the
// far loop reads three distinct 64-bit registers any more (4 of 10 did), which is what held the charge-pair
// form at 73.7 % of the FP32 peak in the loop microbenchmark against 79.0 % for this one
// (tools/ubench2.cu X10 / XP10, profiles/round2_ubench2.txt).  A warp carries 8 points (4 packed pairs).
//   far  record (20 B): a = {alpha x, alpha y, alpha z, alpha} (LDS.128), b = alpha |x|^2 (LDS.32)
//   near record       : a = {2x, 2y, 2z, 4q}
// ------------------------------------------------------------------------------------------
struct __align__(16) PBlock { float4 a[32]; float b[32]; };
static_assert(sizeof(PBlock) == 640, "points-packed charge block must be 640 bytes");

struct PRegs {
    u64 c0[4], c1[4], c2[4], c3[4];     // {-2p.x}, {-2p.y}, {-2p.z}, {|p|^2} of the point pair (positions 2j, 2j+1)
    u64 a0[4], a1[4], a2[4], a3[4];     // FP32 partials per position: T.x - E_near.x, T.y - .., T.z - .., S
};

template <int NP>
__device__ __forceinline__ void evalp_far(const float4 a, const float b, PRegs& r) {
    const u64 ax = pk2(a.x, a.x), ay = pk2(a.y, a.y), az = pk2(a.z, a.z), al = pk2(a.w, a.w);   // .F32 operands
    const u64 bb = pk2(b, b);          // a broadcast ADDEND next to a broadcast multiplicand: FFMA2 R, R.F32x2, R.F32, R.F32
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        u64 t = fma2(r.c3[p], al, bb);
        t = fma2(ax, r.c0[p], t);
        t = fma2(ay, r.c1[p], t);
        t = fma2(az, r.c2[p], t);
        const u64 inv = rsqrt2_abs(t);
        const u64 u = mul2(mul2(inv, inv), inv);
        r.a3[p] = fma2(u, al, r.a3[p]);
        r.a0[p] = fma2(u, ax, r.a0[p]);
        r.a1[p] = fma2(u, ay, r.a1[p]);
        r.a2[p] = fma2(u, az, r.a2[p]);
    }
}

template <int NP>
__device__ __forceinline__ void evalp_near(const float4 a, PRegs& r) {
    const u64 x2 = pk2(a.x, a.x), y2 = pk2(a.y, a.y), z2 = pk2(a.z, a.z), q4 = pk2(a.w, a.w);
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        const u64 dx = add2(r.c0[p], x2);          // -2 (p - x), exact
        const u64 dy = add2(r.c1[p], y2);
        const u64 dz = add2(r.c2[p], z2);
        u64 r2 = mul2(dx, dx);
        r2 = fma2(dy, dy, r2);
        r2 = fma2(dz, dz, r2);
        const u64 inv = rsqrt2(r2);
        const u64 s = mul2(mul2(inv, inv), mul2(inv, q4));
        r.a0[p] = fma2(s, dx, r.a0[p]);
        r.a1[p] = fma2(s, dy, r.a1[p]);
        r.a2[p] = fma2(s, dz, r.a2[p]);
    }
}

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

__device__ __forceinline__ double shfl_xor_f64(double v, int lane_mask) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_xor_sync(0xffffffffu, lo, lane_mask);
    hi = __shfl_xor_sync(0xffffffffu, hi, lane_mask);
    return __hiloint2double(hi, lo);
}

}  // namespace cpet
