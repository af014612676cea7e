// legacy.cu -- the reference's math_module.c symbol names on top of the CUDA kernels, so that
// CPET/utils/c_ops.py `Math_ops(shared_loc=<libcpetb200.so>)` binds and runs unchanged
// (include/cpet_b200.h section (A)).  Ownership and accumulate-into-output quirks follow the
// reference: callers allocate outputs; calc_field_base / calc_esp_base ADD into their output.
//
// These calls use one lazily created process-wide context.  They are one-point / one-line per
// call by construction; the batched cpet_* entry points are the fast path.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "cpet_internal.h"

namespace cpet {

static std::mutex g_mu;
static cpet_ctx* g_ctx = nullptr;
static std::vector<float> g_last_x, g_last_q;   // content of the charge set currently on the device

static cpet_ctx* legacy_ctx() {
    if (g_ctx) return g_ctx;
    int dev = 0;
    if (const char* e = getenv("CPET_B200_DEVICE")) dev = atoi(e);
    if (cpet_create(dev, &g_ctx) != CPET_OK) {
        fprintf(stderr, "libcpetb200: %s\n", cpet_last_error());
        g_ctx = nullptr;
    }
    return g_ctx;
}

static void fail_fill(float* out, int n) {
    fprintf(stderr, "libcpetb200: %s\n", cpet_last_error());
    for (int i = 0; i < n; ++i) out[i] = NAN;
}

// Upload x/Q unless the device already holds exactly this content.
static int legacy_charges(cpet_ctx* c, int n, const float* x, const float* Q) {
    const size_t m = (size_t)(n > 0 ? n : 0);
    if (c->n_charges == n && g_last_x.size() == 3 * m && g_last_q.size() == m &&
        (m == 0 || (memcmp(g_last_x.data(), x, sizeof(float) * 3 * m) == 0 &&
                    memcmp(g_last_q.data(), Q, sizeof(float) * m) == 0)))
        return CPET_OK;
    int rc = cpet_set_charges(c, n, x, Q);
    if (rc == CPET_OK) {
        g_last_x.assign(x, x + 3 * m);
        g_last_q.assign(Q, Q + m);
    } else {
        g_last_x.clear();
        g_last_q.clear();
    }
    return rc;
}

// ---- tiny helper kernels for the non-hot-path symbols ------------------------------------------
__global__ void rowsum_kernel(const float* __restrict__ A, long long rows, int cols, float* __restrict__ ret) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    float s = 0.f;
    for (int j = 0; j < cols; ++j) s += A[i * cols + j];
    ret[i] = s;
}

// result[b][c] = sum_i k * Q[i] * r_mag[b][i] * R[b][i][c]      (C:135-211)
__global__ void __launch_bounds__(256) einsum_op_kernel(int rows, const float* __restrict__ r_mag,
                                                        const float* __restrict__ Q,
                                                        const float* __restrict__ R,
                                                        float* __restrict__ result) {
    const int b = blockIdx.x;
    const float* rm = r_mag + (size_t)b * rows;
    const float* Rb = R + (size_t)b * rows * 3;
    double s0 = 0, s1 = 0, s2 = 0;
    for (int i = threadIdx.x; i < rows; i += blockDim.x) {
        const float w = CPET_COULOMB_K * Q[i] * rm[i];
        s0 += (double)(w * Rb[3 * i]); s1 += (double)(w * Rb[3 * i + 1]); s2 += (double)(w * Rb[3 * i + 2]);
    }
    __shared__ double red[3][256];
    red[0][threadIdx.x] = s0; red[1][threadIdx.x] = s1; red[2][threadIdx.x] = s2;
    __syncthreads();
    for (int w = 128; w >= 1; w >>= 1) {
        if (threadIdx.x < w)
            for (int c = 0; c < 3; ++c) red[c][threadIdx.x] += red[c][threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x < 3) result[3 * (size_t)b + threadIdx.x] = (float)red[threadIdx.x][0];
}

__global__ void vecadd_kernel(const float* A, const float* B, int n, float* ret) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ret[i] = A[i] + B[i];
}

__global__ void dot_kernel(const double* A, const double* B, int rows, int cols, double* ret) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    double s = 0;
    for (int j = 0; j < cols; ++j) s += A[(size_t)i * cols + j] * B[j];
    ret[i] = s;
}

__global__ void spmv_kernel(const int* indptr, int rows, const int* ind, const double* A,
                            const double* B, double* ret) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    double s = 0;
    for (int j = indptr[i]; j < indptr[i + 1]; ++j) s += A[j] * B[ind[j]];
    ret[i] = s;
}

// One dipole streamline, one CTA (development-only symbol of the reference, C:335-371, 506-520,
// 593-660).  The field expression is the reference's, term for term.
__global__ void __launch_bounds__(256) dipole_line_kernel(int n, int n_iter, float h,
                                                          const float* __restrict__ seed,
                                                          const float* __restrict__ dims,
                                                          const float* __restrict__ x,
                                                          const float* __restrict__ mu,
                                                          float* __restrict__ ret) {
    __shared__ double red[3][256];
    __shared__ float cur[3];
    __shared__ float pts[6][3];
    __shared__ int stop;
    auto step_from = [&](const float* p_in, float* p_out) {
        const float px = p_in[0], py = p_in[1], pz = p_in[2];
        double e0 = 0, e1 = 0, e2 = 0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const float r0 = px - x[3 * i], r1 = py - x[3 * i + 1], r2 = pz - x[3 * i + 2];
            const float rn = sqrtf(r0 * r0 + r1 * r1 + r2 * r2);
            const float rsq = rn * rn;
            const float r5 = 1.0f / (rsq * rsq * rn);
            const float m0 = mu[3 * i], m1 = mu[3 * i + 1], m2 = mu[3 * i + 2];
            const float proj = 3.0f * m0 * r0 + m1 * r1 + m2 * r2;
            e0 += (double)(CPET_COULOMB_K * r5 * (proj * r0 - m0 * rsq));
            e1 += (double)(CPET_COULOMB_K * r5 * (proj * r1 - m1 * rsq));
            e2 += (double)(CPET_COULOMB_K * r5 * (proj * r2 - m2 * rsq));
        }
        red[0][threadIdx.x] = e0; red[1][threadIdx.x] = e1; red[2][threadIdx.x] = e2;
        __syncthreads();
        for (int w = 128; w >= 1; w >>= 1) {
            if (threadIdx.x < w)
                for (int c = 0; c < 3; ++c) red[c][threadIdx.x] += red[c][threadIdx.x + w];
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            const double nrm = sqrt(red[0][0] * red[0][0] + red[1][0] * red[1][0] + red[2][0] * red[2][0]);
            p_out[0] = (float)((double)px + (double)h * red[0][0] / nrm);
            p_out[1] = (float)((double)py + (double)h * red[1][0] / nrm);
            p_out[2] = (float)((double)pz + (double)h * red[2][0] / nrm);
        }
        __syncthreads();
    };
    if (threadIdx.x == 0) { for (int c = 0; c < 3; ++c) cur[c] = seed[c]; stop = 0; }
    __syncthreads();
    for (int i = 0; i < n_iter; ++i) {
        step_from(cur, cur);
        if (threadIdx.x == 0) {
            if (cur[0] < -dims[0] || cur[0] > dims[0] || cur[1] < -dims[1] || cur[1] > dims[1] ||
                cur[2] < -dims[2] || cur[2] > dims[2]) stop = 1;
        }
        __syncthreads();
        if (stop) break;
    }
    if (threadIdx.x == 0) for (int c = 0; c < 3; ++c) { pts[0][c] = seed[c]; pts[3][c] = cur[c]; }
    __syncthreads();
    step_from(pts[0], pts[1]);
    step_from(pts[1], pts[2]);
    step_from(pts[3], pts[4]);
    step_from(pts[4], pts[5]);
    if (threadIdx.x == 0) {
        float kap[2];
        for (int e = 0; e < 2; ++e) {
            const float* a0 = pts[3 * e]; const float* a1 = pts[3 * e + 1]; const float* a2 = pts[3 * e + 2];
            float v1[3], v2[3];
            for (int c = 0; c < 3; ++c) { v1[c] = a1[c] - a0[c]; v2[c] = a2[c] - 2.0f * a1[c] + a0[c]; }
            const float cx = v1[1] * v2[2] - v1[2] * v2[1], cy = v1[2] * v2[0] - v1[0] * v2[2],
                        cz = v1[0] * v2[1] - v1[1] * v2[0];
            const float nc = (float)sqrt((double)cx * cx + (double)cy * cy + (double)cz * cz);
            const float nd = (float)sqrt((double)v1[0] * v1[0] + (double)v1[1] * v1[1] + (double)v1[2] * v1[2]);
            kap[e] = (float)((double)nc / ((double)nd * nd * nd));
        }
        const double dx = (double)seed[0] - cur[0], dy = (double)seed[1] - cur[1], dz = (double)seed[2] - cur[2];
        ret[0] = (float)sqrt(dx * dx + dy * dy + dz * dz);
        ret[1] = (kap[0] + kap[1]) / 2;
    }
}

// Generic "upload inputs, run, download outputs" plumbing for the helper symbols.
struct Stage {
    cpet_ctx* c;
    std::vector<void*> bufs;
    bool ok = true;
    explicit Stage(cpet_ctx* ctx) : c(ctx) {}
    ~Stage() { for (void* p : bufs) cudaFree(p); }
    template <typename T>
    T* up(const T* host, size_t n) {
        void* d = nullptr;
        if (cudaMalloc(&d, sizeof(T) * (n ? n : 1)) != cudaSuccess) { ok = false; return nullptr; }
        bufs.push_back(d);
        if (host && n && cudaMemcpyAsync(d, host, sizeof(T) * n, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) ok = false;
        return (T*)d;
    }
    template <typename T>
    bool down(T* host, const T* dev, size_t n) {
        if (!ok) return false;
        if (cudaMemcpyAsync(host, dev, sizeof(T) * n, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) return false;
        return cudaStreamSynchronize(c->stream) == cudaSuccess && cudaGetLastError() == cudaSuccess;
    }
};

}  // namespace cpet

using namespace cpet;

#define LEGACY_CTX_OR_FAIL(out, n)                          \
    cpet_clear_error(); /* status = this call's outcome */  \
    std::lock_guard<std::mutex> lock(g_mu);                 \
    cpet_ctx* c = legacy_ctx();                             \
    if (!c) { fail_fill((out), (n)); return; }              \
    cudaSetDevice(c->device)

extern "C" {

void compute_looped_field(int total_points, int n_charges, float* x_0, float* x, float* Q, float* E) {
    LEGACY_CTX_OR_FAIL(E, 3 * total_points);
    if (legacy_charges(c, n_charges, x, Q) || cpet_field_grid(c, total_points, x_0, CPET_FIELD_SOFTEN, E))
        fail_fill(E, 3 * total_points);
}

void compute_batched_field(int total_points, int batch_size, int n_charges, float* x_0, float* x,
                           float* Q, float* E) {
    (void)batch_size;   // the batching of C:374-403 is a host-memory blocking detail
    LEGACY_CTX_OR_FAIL(E, 3 * total_points);
    if (legacy_charges(c, n_charges, x, Q) || cpet_field_grid(c, total_points, x_0, 0u, E))
        fail_fill(E, 3 * total_points);
}

void calc_field(float* E, float* x_init, int n_charges, float* x, float* Q) {
    LEGACY_CTX_OR_FAIL(E, 3);
    if (legacy_charges(c, n_charges, x, Q) || cpet_field_grid(c, 1, x_init, 0u, E)) fail_fill(E, 3);
}

void calc_field_base(float* E, float* x_init, int n_charges, float* x, float* Q) {
    LEGACY_CTX_OR_FAIL(E, 3);
    float tmp[3];
    if (legacy_charges(c, n_charges, x, Q) || cpet_field_grid(c, 1, x_init, 0u, tmp)) { fail_fill(E, 3); return; }
    E[0] += tmp[0]; E[1] += tmp[1]; E[2] += tmp[2];          // C:327-332 accumulates into E
}

void calc_esp_base(float* ESP, float* x_init, int n_charges, float* x, float* Q) {
    LEGACY_CTX_OR_FAIL(ESP, 1);
    float tmp = 0.f;
    if (legacy_charges(c, n_charges, x, Q) || cpet_esp_grid(c, 1, x_init, 0u, &tmp)) { fail_fill(ESP, 1); return; }
    ESP[0] += tmp;                                             // C:482-485 accumulates into ESP[0]
}

void thread_operation(int n_charges, int n_iter, float step_size, float* x_0, float* dimensions,
                      float* x, float* Q, float* ret) {
    LEGACY_CTX_OR_FAIL(ret, 2);
    const int32_t it = n_iter;
    if (legacy_charges(c, n_charges, x, Q) ||
        cpet_topo_batch(c, 1, x_0, &it, step_size, dimensions, 0u, ret, nullptr))
        fail_fill(ret, 2);
}

void thread_operation_dipole(int n_dipoles, int n_iter, float step_size, float* x_0,
                             float* dimensions, float* x, float* mu, float* ret) {
    LEGACY_CTX_OR_FAIL(ret, 2);
    Stage s(c);
    float* dx = s.up(x, 3 * (size_t)n_dipoles);
    float* dm = s.up(mu, 3 * (size_t)n_dipoles);
    float* ds = s.up(x_0, 3);
    float* dd = s.up(dimensions, 3);
    float* dr = s.up<float>(nullptr, 2);
    if (s.ok) dipole_line_kernel<<<1, 256, 0, c->stream>>>(n_dipoles, n_iter, step_size, ds, dd, dx, dm, dr);
    if (!s.down(ret, dr, 2)) { set_error(CPET_ERR_CUDA, "thread_operation_dipole failed"); fail_fill(ret, 2); }
}

void einsum_ij_i(int rows, int cols, float* A, float* ret) {
    LEGACY_CTX_OR_FAIL(ret, rows);
    Stage s(c);
    float* dA = s.up(A, (size_t)rows * cols);
    float* dr = s.up<float>(nullptr, rows);
    if (s.ok && rows) rowsum_kernel<<<(rows + 255) / 256, 256, 0, c->stream>>>(dA, rows, cols, dr);
    if (rows && !s.down(ret, dr, rows)) { set_error(CPET_ERR_CUDA, "einsum_ij_i failed"); fail_fill(ret, rows); }
}

void einsum_ij_i_batch(int batch, int rows, int cols, float* A, float* ret) {
    const long long total = (long long)batch * rows;
    LEGACY_CTX_OR_FAIL(ret, (int)total);
    Stage s(c);
    float* dA = s.up(A, (size_t)total * cols);
    float* dr = s.up<float>(nullptr, (size_t)total);
    if (s.ok && total) rowsum_kernel<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(dA, total, cols, dr);
    if (total && !s.down(ret, dr, (size_t)total)) { set_error(CPET_ERR_CUDA, "einsum_ij_i_batch failed"); fail_fill(ret, (int)total); }
}

void einsum_operation_batch(int batch, int rows, float* r_mag, float* Q, float* R, float* result) {
    LEGACY_CTX_OR_FAIL(result, 3 * batch);
    Stage s(c);
    float* drm = s.up(r_mag, (size_t)batch * rows);
    float* dq = s.up(Q, rows);
    float* dR = s.up(R, (size_t)batch * rows * 3);
    float* dres = s.up<float>(nullptr, 3 * (size_t)batch);
    if (s.ok && batch) einsum_op_kernel<<<batch, 256, 0, c->stream>>>(rows, drm, dq, dR, dres);
    if (batch && !s.down(result, dres, 3 * (size_t)batch)) { set_error(CPET_ERR_CUDA, "einsum_operation failed"); fail_fill(result, 3 * batch); }
}

void einsum_operation(int rows, float* r_mag, float* Q, float* R, float* result) {
    einsum_operation_batch(1, rows, r_mag, Q, R, result);
}

void vecaddn(float* ret, float* A, float* B, int lenA) {
    LEGACY_CTX_OR_FAIL(ret, lenA);
    Stage s(c);
    float* dA = s.up(A, lenA); float* dB = s.up(B, lenA); float* dr = s.up<float>(nullptr, lenA);
    if (s.ok && lenA) vecadd_kernel<<<(lenA + 255) / 256, 256, 0, c->stream>>>(dA, dB, lenA, dr);
    if (lenA && !s.down(ret, dr, lenA)) { set_error(CPET_ERR_CUDA, "vecaddn failed"); fail_fill(ret, lenA); }
}

void dot(double* ret, double* A, double* B, int rows, int cols) {
    cpet_clear_error();
    std::lock_guard<std::mutex> lock(g_mu);
    cpet_ctx* c = legacy_ctx();
    if (!c) { for (int i = 0; i < rows; ++i) ret[i] = NAN; return; }
    cudaSetDevice(c->device);
    Stage s(c);
    double* dA = s.up(A, (size_t)rows * cols); double* dB = s.up(B, cols); double* dr = s.up<double>(nullptr, rows);
    if (s.ok && rows) dot_kernel<<<(rows + 255) / 256, 256, 0, c->stream>>>(dA, dB, rows, cols, dr);
    if (rows && !s.down(ret, dr, rows)) { set_error(CPET_ERR_CUDA, "dot failed"); for (int i = 0; i < rows; ++i) ret[i] = NAN; }
}

void sparse_dot(double* ret, int* indptr, int indptrlen, int* indA, int lenindA, double* A, int lenA,
                double* B, int size_array) {
    const int rows = indptrlen - 1;
    cpet_clear_error();
    std::lock_guard<std::mutex> lock(g_mu);
    cpet_ctx* c = legacy_ctx();
    if (!c || rows <= 0) { for (int i = 0; i < rows; ++i) ret[i] = NAN; return; }
    cudaSetDevice(c->device);
    Stage s(c);
    int* dp = s.up(indptr, indptrlen); int* di = s.up(indA, lenindA);
    double* dA = s.up(A, lenA); double* dB = s.up(B, size_array); double* dr = s.up<double>(nullptr, rows);
    if (s.ok) spmv_kernel<<<(rows + 255) / 256, 256, 0, c->stream>>>(dp, rows, di, dA, dB, dr);
    if (!s.down(ret, dr, rows)) { set_error(CPET_ERR_CUDA, "sparse_dot failed"); for (int i = 0; i < rows; ++i) ret[i] = NAN; }
}

}  // extern "C"
