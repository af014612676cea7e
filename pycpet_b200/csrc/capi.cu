// capi.cu -- the extern "C" boundary of libcpetb200.so (see include/cpet_b200.h).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <cmath>
#include <mutex>
#include <vector>

#include "cpet_internal.h"

namespace cpet {

static thread_local std::string g_err;
static thread_local int g_status = 0;

void set_error(int status, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    g_status = status;
}

int DevBuf::reserve(size_t bytes) {
    if (bytes <= cap) return CPET_OK;
    if (p) { cudaFree(p); p = nullptr; cap = 0; }
    size_t want = bytes + bytes / 4;            // grow-only with slack
    want = (want + 255) & ~(size_t)255;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
        e = cudaMalloc(&p, (bytes + 255) & ~(size_t)255);
        want = (bytes + 255) & ~(size_t)255;
    }
    if (e != cudaSuccess) {
        p = nullptr;
        set_error(CPET_ERR_CUDA, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
        return CPET_ERR_CUDA;
    }
    cap = want;
    return CPET_OK;
}

void DevBuf::release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
}

static bool timer_slot(cpet_ctx* c, int i) {
    if (!c->ev0[i]) {
        if (cudaEventCreate(&c->ev0[i]) != cudaSuccess || cudaEventCreate(&c->ev1[i]) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
    }
    return true;
}
KernelTimer::KernelTimer(cpet_ctx* ctx) : c(ctx) {
    if (!c->tune.timing) return;
    const int i = c->timer_count % cpet_ctx::kTimerRing;
    if (timer_slot(c, i)) cudaEventRecord(c->ev0[i], c->stream);
}
KernelTimer::~KernelTimer() {
    if (!c->tune.timing) return;
    const int i = c->timer_count % cpet_ctx::kTimerRing;
    if (c->ev1[i]) cudaEventRecord(c->ev1[i], c->stream);
    ++c->timer_count;
}

static int resolve_timer(cpet_ctx* c) {
    if (!c->tune.timing || c->timer_count == 0) { c->last_kernel_ms = 0.0; return CPET_OK; }
    const int i = (c->timer_count - 1) % cpet_ctx::kTimerRing;
    CPET_CUDA_TRY(cudaEventSynchronize(c->ev1[i]));
    float ms = 0.f;
    CPET_CUDA_TRY(cudaEventElapsedTime(&ms, c->ev0[i], c->ev1[i]));
    c->last_kernel_ms = ms;
    return CPET_OK;
}

static int make_ctx(int device, void* stream, bool borrow, cpet_ctx** out) {
    CPET_REQUIRE(out != nullptr, CPET_ERR_INVALID, "cpet_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        set_error(CPET_ERR_NO_DEVICE, "no CUDA device available (%s); libcpetb200 has no CPU fallback",
                  e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return CPET_ERR_NO_DEVICE;
    }
    CPET_REQUIRE(device >= 0 && device < n, CPET_ERR_INVALID, "device %d out of range (0..%d)", device, n - 1);
    cudaDeviceProp prop;
    CPET_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error(CPET_ERR_NO_DEVICE,
                  "device %d (%s) is compute capability %d.%d; this library carries sm_100a code only",
                  device, prop.name, prop.major, prop.minor);
        return CPET_ERR_NO_DEVICE;
    }
    CPET_CUDA_TRY(cudaSetDevice(device));
    cpet_ctx* c = new cpet_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    if (borrow) {
        c->stream = (cudaStream_t)stream;
        c->own_stream = false;
    } else {
        e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            delete c;
            set_error(CPET_ERR_CUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(e));
            return CPET_ERR_CUDA;
        }
        c->own_stream = true;
    }
    *out = c;
    return CPET_OK;
}

// Host-pointer entry points enqueue copies that reference the caller's buffers.  Whatever path such
// a call leaves on, its stream is drained first: a caller may free (pinned) buffers right after an
// error return.  The success paths synchronise themselves and disarm the guard.
struct DrainOnExit {
    cudaStream_t s;
    bool armed = true;
    explicit DrainOnExit(cudaStream_t st) : s(st) {}
    ~DrainOnExit() {
        if (armed) cudaStreamSynchronize(s);
    }
};

#define CTX_GUARD(c)                                                                     \
    CPET_REQUIRE((c) != nullptr, CPET_ERR_INVALID, "context is NULL");                   \
    CPET_CUDA_TRY(cudaSetDevice((c)->device))

static int resolve_topo_counters(cpet_ctx* c) {
    if (c->last_counters[1] >= 0) return CPET_OK;
    unsigned long long ev = 0;
    CPET_CUDA_TRY(cudaMemcpyAsync(&ev, c->counters.as<unsigned char>() + 8, sizeof(ev),
                                  cudaMemcpyDeviceToHost, c->stream));
    CPET_CUDA_TRY(cudaStreamSynchronize(c->stream));
    c->last_counters[2] = (int64_t)ev;
    c->last_counters[1] = (int64_t)ev * (int64_t)c->n_charges;
    return CPET_OK;
}

// ---- MD-frame batch helpers (cpet_topo_hist_frames) ------------------------------------------
// Two child contexts (own streams, own staging buffers) take the frames alternately: everything a
// frame needs is enqueued on its child's stream -- H2D of its charges (and n_iter), pack, queue sort,
// integrator, histogram, D2H of its counts (and rows) -- so the copies of frame f+1 / f-1 overlap
// the kernels of frame f, and the CTAs of frame f+1 start on each SM as soon as frame f leaves it.
__global__ void add_evals_kernel(const unsigned long long* __restrict__ evals, int n_charges,
                                 unsigned long long* __restrict__ totals) {
    totals[0] += *evals;
    totals[1] += *evals * (unsigned long long)n_charges;
}

static int frames_child(cpet_ctx* c, int i, cpet_ctx** out) {
    if (!c->pipe[i]) {
        if (int rc = make_ctx(c->device, nullptr, false, &c->pipe[i])) return rc;
    }
    c->pipe[i]->tune = c->tune;
    c->pipe[i]->tune.timing = 0;
    *out = c->pipe[i];
    return CPET_OK;
}

}  // namespace cpet

using namespace cpet;

extern "C" {

int cpet_abi_version(void) { return CPET_ABI_VERSION; }
const char* cpet_last_error(void) { return g_err.c_str(); }
int cpet_last_status(void) { return g_status; }
void cpet_clear_error(void) { g_err.clear(); g_status = 0; }

int cpet_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int cpet_create(int device, cpet_ctx** out) { return make_ctx(device, nullptr, false, out); }
int cpet_create_on_stream(int device, void* cuda_stream, cpet_ctx** out) {
    return make_ctx(device, cuda_stream, true, out);
}

int cpet_destroy(cpet_ctx* c) {
    if (!c) return CPET_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    c->charges.release(); c->charge_blocks.release(); c->raw_x.release(); c->raw_q.release();
    c->xblocks.release(); c->xchunks.release();
    c->in0.release(); c->in1.release(); c->out0.release(); c->out1.release();
    c->work0.release(); c->work1.release(); c->work2.release(); c->counters.release(); c->flags.release(); c->totals.release();
    for (int i = 0; i < cpet_ctx::kTimerRing; ++i) {
        if (c->ev0[i]) cudaEventDestroy(c->ev0[i]);
        if (c->ev1[i]) cudaEventDestroy(c->ev1[i]);
    }
    for (int i = 0; i < 2; ++i)
        if (c->pipe[i]) { cpet_destroy(c->pipe[i]); c->pipe[i] = nullptr; }
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return CPET_OK;
}

int cpet_sync(cpet_ctx* c) {
    CTX_GUARD(c);
    CPET_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return CPET_OK;
}

int cpet_device_of(cpet_ctx* c) { return c ? c->device : -1; }
int cpet_last_path(cpet_ctx* c) { return c ? c->last_path : -1; }

int cpet_set_tuning(cpet_ctx* c, const char* key, int value) {
    CPET_REQUIRE(c && key, CPET_ERR_INVALID, "cpet_set_tuning: NULL argument");
    Tuning& t = c->tune;
    struct { const char* k; int* v; } tab[] = {
        {"k1_threads", &t.k1_threads}, {"k1_points", &t.k1_points}, {"k1_lanes", &t.k1_lanes},
        {"k1_tile_pairs", &t.k1_tile_pairs}, {"k1_stages", &t.k1_stages}, {"k1_splits", &t.k1_splits}, {"k1_lattice", &t.k1_lattice}, {"k1_unroll", &t.k1_unroll}, {"k1_softscan", &t.k1_softscan},
        {"k2_threads", &t.k2_threads}, {"k2_tile_pairs", &t.k2_tile_pairs},
        {"k2_stages", &t.k2_stages}, {"k2_sort", &t.k2_sort}, {"k2_cap", &t.k2_cap},
        {"k2_form", &t.k2_form}, {"k2_amax", &t.k2_amax}, {"k2_unroll", &t.k2_unroll},
        {"frames_pin", &t.frames_pin}, {"k1_esp_mix", &t.k1_esp_mix}, {"k1_lat_nodes", &t.k1_lat_nodes}, {"k1_hybrid", &t.k1_hybrid}, {"k2_tail4", &t.k2_tail4}, {"k2_tail2", &t.k2_tail2},
        {"timing", &t.timing},
    };
    for (auto& e : tab) {
        if (strcmp(e.k, key) == 0) {
            if (e.v == &t.k2_sort || e.v == &t.k1_lattice || e.v == &t.k1_softscan || e.v == &t.k1_esp_mix || e.v == &t.k1_lat_nodes || e.v == &t.k1_hybrid || e.v == &t.k2_tail4 || e.v == &t.k2_tail2) *e.v = value;
            else *e.v = value > 0 ? value : 0;
            return CPET_OK;
        }
    }
    set_error(CPET_ERR_INVALID, "unknown tuning key '%s'", key);
    return CPET_ERR_INVALID;
}

int cpet_last_counters(cpet_ctx* c, int64_t out[3]) {
    CTX_GUARD(c);
    CPET_REQUIRE(out != nullptr, CPET_ERR_INVALID, "out is NULL");
    if (int rc = resolve_topo_counters(c)) return rc;
    out[0] = c->last_counters[0]; out[1] = c->last_counters[1]; out[2] = c->last_counters[2];
    return CPET_OK;
}

int cpet_kernel_times(cpet_ctx* c, double* ms, int max_n, int* n_out) {
    CTX_GUARD(c);
    CPET_REQUIRE(ms != nullptr && n_out != nullptr && max_n >= 0, CPET_ERR_INVALID, "bad arguments");
    int n = c->timer_count;
    if (n > cpet_ctx::kTimerRing) n = cpet_ctx::kTimerRing;
    if (n > max_n) n = max_n;
    const int first = c->timer_count - n;
    for (int j = 0; j < n; ++j) {
        const int i = (first + j) % cpet_ctx::kTimerRing;
        CPET_CUDA_TRY(cudaEventSynchronize(c->ev1[i]));
        float t = 0.f;
        CPET_CUDA_TRY(cudaEventElapsedTime(&t, c->ev0[i], c->ev1[i]));
        ms[j] = t;
    }
    *n_out = n;
    c->timer_count = 0;
    return CPET_OK;
}

int cpet_last_kernel_ms(cpet_ctx* c, double* ms) {
    CTX_GUARD(c);
    CPET_REQUIRE(ms != nullptr, CPET_ERR_INVALID, "ms is NULL");
    if (int rc = resolve_timer(c)) return rc;
    *ms = c->last_kernel_ms;
    return CPET_OK;
}

// ---------------------------------------------------------------- charges -------------------
int cpet_set_charges_dev(cpet_ctx* c, int n_charges, const float* d_x, const float* d_Q) {
    CTX_GUARD(c);
    CPET_REQUIRE(n_charges >= 0, CPET_ERR_INVALID, "n_charges < 0");
    CPET_REQUIRE(n_charges == 0 || (d_x && d_Q), CPET_ERR_INVALID, "NULL charge arrays");
    return launch_pack_charges(c, n_charges, d_x, d_Q);
}

int cpet_set_charges(cpet_ctx* c, int n_charges, const float* x, const float* Q) {
    CTX_GUARD(c);
    CPET_REQUIRE(n_charges >= 0, CPET_ERR_INVALID, "n_charges < 0");
    CPET_REQUIRE(n_charges == 0 || (x && Q), CPET_ERR_INVALID, "NULL charge arrays");
    const size_t m = (size_t)(n_charges > 0 ? n_charges : 1);
    if (int rc = c->raw_x.reserve(sizeof(float) * 3 * m)) return rc;
    if (int rc = c->raw_q.reserve(sizeof(float) * m)) return rc;
    if (n_charges > 0) {
        CPET_CUDA_TRY(cudaMemcpyAsync(c->raw_x.p, x, sizeof(float) * 3 * m, cudaMemcpyHostToDevice, c->stream));
        CPET_CUDA_TRY(cudaMemcpyAsync(c->raw_q.p, Q, sizeof(float) * m, cudaMemcpyHostToDevice, c->stream));
    }
    return launch_pack_charges(c, n_charges, c->raw_x.as<float>(), c->raw_q.as<float>());
}

// ---------------------------------------------------------------- K1 -------------------------
static int field_mode_of(unsigned flags) { return (flags & CPET_FIELD_SOFTEN) ? MODE_FIELD_SOFT : MODE_FIELD_RAW; }

int cpet_field_grid_dev(cpet_ctx* c, int n_points, const float* d_x0, unsigned flags, float* d_out) {
    CTX_GUARD(c);
    CPET_REQUIRE(n_points >= 0, CPET_ERR_INVALID, "n_points < 0");
    CPET_REQUIRE((flags & ~(CPET_FIELD_SOFTEN | CPET_OUT_CONCAT)) == 0, CPET_ERR_INVALID, "unknown flag bits 0x%x", flags);
    CPET_REQUIRE(n_points == 0 || (d_x0 && d_out), CPET_ERR_INVALID, "NULL point/output arrays");
    return launch_field_grid(c, field_mode_of(flags), n_points, d_x0, (flags & CPET_OUT_CONCAT) ? 1 : 0, d_out);
}

int cpet_field_grid(cpet_ctx* c, int n_points, const float* x0, unsigned flags, float* out) {
    CTX_GUARD(c);
    CPET_REQUIRE(n_points >= 0, CPET_ERR_INVALID, "n_points < 0");
    CPET_REQUIRE(n_points == 0 || (x0 && out), CPET_ERR_INVALID, "NULL point/output arrays");
    if (n_points == 0) return CPET_OK;
    const size_t n = (size_t)n_points;
    const size_t out_floats = (flags & CPET_OUT_CONCAT) ? 6 * n : 3 * n;
    if (int rc = c->in0.reserve(sizeof(float) * 3 * n)) return rc;
    if (int rc = c->out0.reserve(sizeof(float) * out_floats)) return rc;
    DrainOnExit drain(c->stream);
    CPET_CUDA_TRY(cudaMemcpyAsync(c->in0.p, x0, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, c->stream));
    // The reference only ever hands over mesh.reshape(-1,3) (UC:442): recognise box meshes here and
    // give them the lattice kernel (bit-identical results, a quarter fewer FP32 instructions).
    int is_lat = 0, nx = 0, ny = 0, nz = 0;
    const float* d_axes = nullptr;
    const bool try_lattice = c->tune.k1_lattice > 0 || (c->tune.k1_lattice < 0 && c->tune.k1_lanes == 0);
    if (try_lattice) {
        if (int rc = detect_lattice(c, n_points, c->in0.as<float>(), &is_lat, &nx, &ny, &nz, &d_axes)) return rc;
    }
    if (is_lat) {
        CPET_REQUIRE((flags & ~(CPET_FIELD_SOFTEN | CPET_OUT_CONCAT)) == 0, CPET_ERR_INVALID, "unknown flag bits 0x%x", flags);
        if (int rc = launch_field_lattice(c, field_mode_of(flags), nx, ny, nz, d_axes, d_axes + nx, d_axes + nx + ny,
                                          (flags & CPET_OUT_CONCAT) ? 1 : 0, c->out0.p))
            return rc;
    } else {
        if (int rc = cpet_field_grid_dev(c, n_points, c->in0.as<float>(), flags, c->out0.as<float>())) return rc;
    }
    CPET_CUDA_TRY(cudaMemcpyAsync(out, c->out0.p, sizeof(float) * out_floats, cudaMemcpyDeviceToHost, c->stream));
    CPET_CUDA_TRY(cudaStreamSynchronize(c->stream));
    drain.armed = false;
    return CPET_OK;
}

int cpet_esp_grid_dev(cpet_ctx* c, int n_points, const float* d_x0, unsigned flags, void* d_out) {
    CTX_GUARD(c);
    CPET_REQUIRE(n_points >= 0, CPET_ERR_INVALID, "n_points < 0");
    CPET_REQUIRE((flags & ~CPET_OUT_CONCAT) == 0, CPET_ERR_INVALID, "unknown flag bits 0x%x", flags);
    CPET_REQUIRE(n_points == 0 || (d_x0 && d_out), CPET_ERR_INVALID, "NULL point/output arrays");
    return launch_field_grid(c, MODE_ESP, n_points, d_x0, (flags & CPET_OUT_CONCAT) ? 3 : 2, d_out);
}

int cpet_esp_grid(cpet_ctx* c, int n_points, const float* x0, unsigned flags, void* out) {
    CTX_GUARD(c);
    CPET_REQUIRE(n_points >= 0, CPET_ERR_INVALID, "n_points < 0");
    CPET_REQUIRE(n_points == 0 || (x0 && out), CPET_ERR_INVALID, "NULL point/output arrays");
    if (n_points == 0) return CPET_OK;
    const size_t n = (size_t)n_points;
    const size_t out_bytes = (flags & CPET_OUT_CONCAT) ? 8 * n : 4 * n;
    if (int rc = c->in0.reserve(sizeof(float) * 3 * n)) return rc;
    if (int rc = c->out0.reserve(out_bytes)) return rc;
    DrainOnExit drain(c->stream);
    CPET_CUDA_TRY(cudaMemcpyAsync(c->in0.p, x0, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, c->stream));
    // box meshes (UC:450-475 hands over mesh.reshape(-1,3)) get the lattice kernel, as in cpet_field_grid
    int is_lat = 0, nx = 0, ny = 0, nz = 0;
    const float* d_axes = nullptr;
    const bool try_lattice = c->tune.k1_lattice > 0 || (c->tune.k1_lattice < 0 && c->tune.k1_lanes == 0);
    if (try_lattice) {
        if (int rc = detect_lattice(c, n_points, c->in0.as<float>(), &is_lat, &nx, &ny, &nz, &d_axes)) return rc;
    }
    if (is_lat) {
        CPET_REQUIRE((flags & ~CPET_OUT_CONCAT) == 0, CPET_ERR_INVALID, "unknown flag bits 0x%x", flags);
        if (int rc = launch_field_lattice(c, MODE_ESP, nx, ny, nz, d_axes, d_axes + nx, d_axes + nx + ny,
                                          (flags & CPET_OUT_CONCAT) ? 3 : 2, c->out0.p))
            return rc;
    } else {
        if (int rc = cpet_esp_grid_dev(c, n_points, c->in0.as<float>(), flags, c->out0.p)) return rc;
    }
    CPET_CUDA_TRY(cudaMemcpyAsync(out, c->out0.p, out_bytes, cudaMemcpyDeviceToHost, c->stream));
    CPET_CUDA_TRY(cudaStreamSynchronize(c->stream));
    drain.armed = false;
    return CPET_OK;
}

// ---------------------------------------------------------------- K1 on lattices -------------
static int lattice_out_kind(int is_esp, unsigned flags) {
    if (is_esp) return (flags & CPET_OUT_CONCAT) ? 3 : 2;
    return (flags & CPET_OUT_CONCAT) ? 1 : 0;
}
static size_t lattice_out_bytes(int is_esp, unsigned flags, size_t n) {
    if (is_esp) return (flags & CPET_OUT_CONCAT) ? 8 * n : 4 * n;
    return sizeof(float) * ((flags & CPET_OUT_CONCAT) ? 6 * n : 3 * n);
}

static int lattice_dev(cpet_ctx* c, int is_esp, int nx, int ny, int nz, const float* d_xs, const float* d_ys,
                       const float* d_zs, unsigned flags, void* d_out) {
    CTX_GUARD(c);
    CPET_REQUIRE(nx >= 0 && ny >= 0 && nz >= 0, CPET_ERR_INVALID, "negative lattice size");
    const unsigned allowed = is_esp ? CPET_OUT_CONCAT : (CPET_FIELD_SOFTEN | CPET_OUT_CONCAT);
    CPET_REQUIRE((flags & ~allowed) == 0, CPET_ERR_INVALID, "unknown flag bits 0x%x", flags);
    if ((long long)nx * ny * nz == 0) return CPET_OK;
    CPET_REQUIRE(d_xs && d_ys && d_zs && d_out, CPET_ERR_INVALID, "NULL axis/output arrays");
    const int mode = is_esp ? MODE_ESP : field_mode_of(flags);
    return launch_field_lattice(c, mode, nx, ny, nz, d_xs, d_ys, d_zs, lattice_out_kind(is_esp, flags), d_out);
}

static int lattice_host(cpet_ctx* c, int is_esp, int nx, int ny, int nz, const float* xs, const float* ys,
                        const float* zs, unsigned flags, void* out) {
    CTX_GUARD(c);
    CPET_REQUIRE(nx >= 0 && ny >= 0 && nz >= 0, CPET_ERR_INVALID, "negative lattice size");
    const size_t n = (size_t)nx * ny * nz;
    if (n == 0) return CPET_OK;
    CPET_REQUIRE(xs && ys && zs && out, CPET_ERR_INVALID, "NULL axis/output arrays");
    const size_t ob = lattice_out_bytes(is_esp, flags, n);
    if (int rc = c->in0.reserve(sizeof(float) * ((size_t)nx + ny + nz))) return rc;
    if (int rc = c->out0.reserve(ob)) return rc;
    float* d_xs = c->in0.as<float>();
    float* d_ys = d_xs + nx;
    float* d_zs = d_ys + ny;
    DrainOnExit drain(c->stream);
    CPET_CUDA_TRY(cudaMemcpyAsync(d_xs, xs, sizeof(float) * nx, cudaMemcpyHostToDevice, c->stream));
    CPET_CUDA_TRY(cudaMemcpyAsync(d_ys, ys, sizeof(float) * ny, cudaMemcpyHostToDevice, c->stream));
    CPET_CUDA_TRY(cudaMemcpyAsync(d_zs, zs, sizeof(float) * nz, cudaMemcpyHostToDevice, c->stream));
    if (int rc = lattice_dev(c, is_esp, nx, ny, nz, d_xs, d_ys, d_zs, flags, c->out0.p)) return rc;
    CPET_CUDA_TRY(cudaMemcpyAsync(out, c->out0.p, ob, cudaMemcpyDeviceToHost, c->stream));
    CPET_CUDA_TRY(cudaStreamSynchronize(c->stream));
    drain.armed = false;
    return CPET_OK;
}

int cpet_field_lattice(cpet_ctx* c, int nx, int ny, int nz, const float* xs, const float* ys, const float* zs,
                       unsigned flags, float* out) {
    return lattice_host(c, 0, nx, ny, nz, xs, ys, zs, flags, out);
}
int cpet_field_lattice_dev(cpet_ctx* c, int nx, int ny, int nz, const float* d_xs, const float* d_ys,
                           const float* d_zs, unsigned flags, float* d_out) {
    return lattice_dev(c, 0, nx, ny, nz, d_xs, d_ys, d_zs, flags, d_out);
}
int cpet_esp_lattice(cpet_ctx* c, int nx, int ny, int nz, const float* xs, const float* ys, const float* zs,
                     unsigned flags, void* out) {
    return lattice_host(c, 1, nx, ny, nz, xs, ys, zs, flags, out);
}
int cpet_esp_lattice_dev(cpet_ctx* c, int nx, int ny, int nz, const float* d_xs, const float* d_ys,
                         const float* d_zs, unsigned flags, void* d_out) {
    return lattice_dev(c, 1, nx, ny, nz, d_xs, d_ys, d_zs, flags, d_out);
}

int cpet_propagate_dev(cpet_ctx* c, int n_points, const float* d_x0, float step_size, float* d_out) {
    CTX_GUARD(c);
    CPET_REQUIRE(n_points >= 0, CPET_ERR_INVALID, "n_points < 0");
    CPET_REQUIRE(n_points == 0 || (d_x0 && d_out), CPET_ERR_INVALID, "NULL point/output arrays");
    return launch_field_grid(c, MODE_FIELD_RAW, n_points, d_x0, 4, d_out, step_size);
}

int cpet_propagate(cpet_ctx* c, int n_points, const float* x0, float step_size, float* out) {
    CTX_GUARD(c);
    CPET_REQUIRE(n_points >= 0, CPET_ERR_INVALID, "n_points < 0");
    CPET_REQUIRE(n_points == 0 || (x0 && out), CPET_ERR_INVALID, "NULL point/output arrays");
    if (n_points == 0) return CPET_OK;
    const size_t n = (size_t)n_points;
    if (int rc = c->in0.reserve(sizeof(float) * 3 * n)) return rc;
    if (int rc = c->out0.reserve(sizeof(float) * 3 * n)) return rc;
    DrainOnExit drain(c->stream);
    CPET_CUDA_TRY(cudaMemcpyAsync(c->in0.p, x0, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, c->stream));
    if (int rc = cpet_propagate_dev(c, n_points, c->in0.as<float>(), step_size, c->out0.as<float>())) return rc;
    CPET_CUDA_TRY(cudaMemcpyAsync(out, c->out0.p, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, c->stream));
    CPET_CUDA_TRY(cudaStreamSynchronize(c->stream));
    drain.armed = false;
    return CPET_OK;
}

// ---------------------------------------------------------------- K2 -------------------------
int cpet_topo_batch_dev(cpet_ctx* c, int n_lines, const float* d_seeds, const int32_t* d_n_iter,
                        float step_size, const float dims[3], unsigned flags, float* d_out,
                        int32_t* d_steps) {
    CTX_GUARD(c);
    CPET_REQUIRE(n_lines >= 0, CPET_ERR_INVALID, "n_lines < 0");
    CPET_REQUIRE((flags & ~CPET_TOPO_CURV_SECOND_DIFF) == 0, CPET_ERR_INVALID, "unknown flag bits 0x%x", flags);
    CPET_REQUIRE(dims != nullptr, CPET_ERR_INVALID, "dims is NULL");
    CPET_REQUIRE(n_lines == 0 || (d_seeds && d_n_iter && d_out), CPET_ERR_INVALID, "NULL seed/n_iter/output arrays");
    return launch_topo(c, n_lines, d_seeds, d_n_iter, step_size, dims, flags, d_out, d_steps);
}

int cpet_topo_batch(cpet_ctx* c, int n_lines, const float* seeds, const int32_t* n_iter,
                    float step_size, const float dims[3], unsigned flags, float* out,
                    int32_t* steps) {
    CTX_GUARD(c);
    CPET_REQUIRE(n_lines >= 0, CPET_ERR_INVALID, "n_lines < 0");
    CPET_REQUIRE(n_lines == 0 || (seeds && n_iter && out), CPET_ERR_INVALID, "NULL seed/n_iter/output arrays");
    if (n_lines == 0) return CPET_OK;
    const size_t n = (size_t)n_lines;
    if (int rc = c->in0.reserve(sizeof(float) * 3 * n)) return rc;
    if (int rc = c->in1.reserve(sizeof(int32_t) * n)) return rc;
    if (int rc = c->out0.reserve(sizeof(float) * 2 * n)) return rc;
    if (steps) { if (int rc = c->out1.reserve(sizeof(int32_t) * n)) return rc; }
    DrainOnExit drain(c->stream);
    CPET_CUDA_TRY(cudaMemcpyAsync(c->in0.p, seeds, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, c->stream));
    CPET_CUDA_TRY(cudaMemcpyAsync(c->in1.p, n_iter, sizeof(int32_t) * n, cudaMemcpyHostToDevice, c->stream));
    if (int rc = cpet_topo_batch_dev(c, n_lines, c->in0.as<float>(), c->in1.as<int32_t>(), step_size, dims,
                                     flags, c->out0.as<float>(), steps ? c->out1.as<int32_t>() : nullptr))
        return rc;
    CPET_CUDA_TRY(cudaMemcpyAsync(out, c->out0.p, sizeof(float) * 2 * n, cudaMemcpyDeviceToHost, c->stream));
    if (steps)
        CPET_CUDA_TRY(cudaMemcpyAsync(steps, c->out1.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, c->stream));
    CPET_CUDA_TRY(cudaStreamSynchronize(c->stream));
    drain.armed = false;
    return CPET_OK;
}

// ---------------------------------------------------------------- K3 -------------------------
static int upload_edges(cpet_ctx* c, int nd, const double* d_edges, int nc, const double* c_edges,
                        const double** dev_d, const double** dev_c) {
    CPET_REQUIRE(nd >= 1 && nc >= 1, CPET_ERR_INVALID, "histogram needs nd >= 1 and nc >= 1");
    CPET_REQUIRE(d_edges && c_edges, CPET_ERR_INVALID, "NULL edge arrays");
    const size_t n = (size_t)nd + nc + 2;
    if (int rc = c->work2.reserve(sizeof(double) * n)) return rc;
    CPET_CUDA_TRY(cudaMemcpyAsync(c->work2.p, d_edges, sizeof(double) * (nd + 1), cudaMemcpyHostToDevice, c->stream));
    CPET_CUDA_TRY(cudaMemcpyAsync(c->work2.as<double>() + nd + 1, c_edges, sizeof(double) * (nc + 1),
                                  cudaMemcpyHostToDevice, c->stream));
    *dev_d = c->work2.as<double>();
    *dev_c = c->work2.as<double>() + nd + 1;
    return CPET_OK;
}

int cpet_hist2d_dev(cpet_ctx* c, int n_frames, int64_t n_per_frame, const float* d_values, int nd,
                    const double* d_edges_host, int nc, const double* c_edges_host, int64_t* d_counts) {
    CTX_GUARD(c);
    CPET_REQUIRE(n_frames >= 0 && n_per_frame >= 0, CPET_ERR_INVALID, "negative sizes");
    CPET_REQUIRE(d_counts != nullptr, CPET_ERR_INVALID, "counts is NULL");
    const double *dd, *dc;
    if (int rc = upload_edges(c, nd, d_edges_host, nc, c_edges_host, &dd, &dc)) return rc;
    return launch_hist2d(c, n_frames, n_per_frame, d_values, false, nd, dd, nc, dc,
                         reinterpret_cast<unsigned long long*>(d_counts));
}

static int hist2d_host(cpet_ctx* c, int n_frames, int64_t n_per_frame, const void* values, bool f64,
                       int nd, const double* d_edges, int nc, const double* c_edges, int64_t* counts) {
    CTX_GUARD(c);
    CPET_REQUIRE(n_frames >= 0 && n_per_frame >= 0, CPET_ERR_INVALID, "negative sizes");
    CPET_REQUIRE(counts != nullptr, CPET_ERR_INVALID, "counts is NULL");
    CPET_REQUIRE((int64_t)n_frames * n_per_frame == 0 || values, CPET_ERR_INVALID, "values is NULL");
    DrainOnExit drain(c->stream);
    const double *dd, *dc;
    if (int rc = upload_edges(c, nd, d_edges, nc, c_edges, &dd, &dc)) return rc;
    const size_t nval = (size_t)n_frames * (size_t)n_per_frame * 2;
    const size_t vbytes = nval * (f64 ? sizeof(double) : sizeof(float));
    const size_t cbytes = sizeof(int64_t) * (size_t)n_frames * nd * nc;
    if (int rc = c->in0.reserve(vbytes ? vbytes : 8)) return rc;
    if (int rc = c->out0.reserve(cbytes ? cbytes : 8)) return rc;
    if (vbytes) CPET_CUDA_TRY(cudaMemcpyAsync(c->in0.p, values, vbytes, cudaMemcpyHostToDevice, c->stream));
    if (int rc = launch_hist2d(c, n_frames, n_per_frame, c->in0.p, f64, nd, dd, nc, dc,
                               c->out0.as<unsigned long long>()))
        return rc;
    if (cbytes) CPET_CUDA_TRY(cudaMemcpyAsync(counts, c->out0.p, cbytes, cudaMemcpyDeviceToHost, c->stream));
    CPET_CUDA_TRY(cudaStreamSynchronize(c->stream));
    drain.armed = false;
    return CPET_OK;
}

int cpet_hist2d(cpet_ctx* c, int n_frames, int64_t n_per_frame, const double* values, int nd,
                const double* d_edges, int nc, const double* c_edges, int64_t* counts) {
    return hist2d_host(c, n_frames, n_per_frame, values, true, nd, d_edges, nc, c_edges, counts);
}

int cpet_hist2d_f32(cpet_ctx* c, int n_frames, int64_t n_per_frame, const float* values, int nd,
                    const double* d_edges, int nc, const double* c_edges, int64_t* counts) {
    return hist2d_host(c, n_frames, n_per_frame, values, false, nd, d_edges, nc, c_edges, counts);
}

int cpet_topo_hist(cpet_ctx* c, int n_lines, const float* seeds, const int32_t* n_iter, float step_size,
                   const float dims[3], unsigned flags, float* out_rows, int32_t* steps, int nd,
                   const double* d_edges, int nc, const double* c_edges, int64_t* counts) {
    CTX_GUARD(c);
    CPET_REQUIRE(n_lines >= 0, CPET_ERR_INVALID, "n_lines < 0");
    CPET_REQUIRE(n_lines == 0 || (seeds && n_iter), CPET_ERR_INVALID, "NULL seed/n_iter arrays");
    CPET_REQUIRE(counts != nullptr, CPET_ERR_INVALID, "counts is NULL");
    DrainOnExit drain(c->stream);
    const double *dd, *dc;
    if (int rc = upload_edges(c, nd, d_edges, nc, c_edges, &dd, &dc)) return rc;
    const size_t n = (size_t)(n_lines > 0 ? n_lines : 1);
    const size_t cbytes = sizeof(int64_t) * (size_t)nd * nc;
    if (int rc = c->in0.reserve(sizeof(float) * 3 * n)) return rc;
    if (int rc = c->in1.reserve(sizeof(int32_t) * n)) return rc;
    if (int rc = c->out0.reserve(sizeof(float) * 2 * n)) return rc;
    if (int rc = c->out1.reserve(sizeof(int32_t) * n)) return rc;
    if (int rc = c->work0.reserve(cbytes)) return rc;
    if (n_lines > 0) {
        CPET_CUDA_TRY(cudaMemcpyAsync(c->in0.p, seeds, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, c->stream));
        CPET_CUDA_TRY(cudaMemcpyAsync(c->in1.p, n_iter, sizeof(int32_t) * n, cudaMemcpyHostToDevice, c->stream));
        if (int rc = cpet_topo_batch_dev(c, n_lines, c->in0.as<float>(), c->in1.as<int32_t>(), step_size, dims,
                                         flags, c->out0.as<float>(), steps ? c->out1.as<int32_t>() : nullptr))
            return rc;
    }
    const int64_t saved[3] = {c->last_counters[0], c->last_counters[1], c->last_counters[2]};
    // the (L,2) rows never leave the device between the integrator and the binning kernel
    if (int rc = launch_hist2d(c, 1, n_lines, c->out0.p, false, nd, dd, nc, dc, c->work0.as<unsigned long long>()))
        return rc;
    c->last_counters[0] = saved[0] + c->last_counters[0];
    c->last_counters[1] = saved[1];
    c->last_counters[2] = saved[2];
    if (n_lines > 0 && out_rows)
        CPET_CUDA_TRY(cudaMemcpyAsync(out_rows, c->out0.p, sizeof(float) * 2 * n, cudaMemcpyDeviceToHost, c->stream));
    if (n_lines > 0 && steps)
        CPET_CUDA_TRY(cudaMemcpyAsync(steps, c->out1.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, c->stream));
    CPET_CUDA_TRY(cudaMemcpyAsync(counts, c->work0.p, cbytes, cudaMemcpyDeviceToHost, c->stream));
    CPET_CUDA_TRY(cudaStreamSynchronize(c->stream));
    drain.armed = false;
    return CPET_OK;
}

// ---------------------------------------------------------------- MD-frame batch -----------
// Page-locks a caller's pageable host range for the lifetime of the object.  A device-to-host cudaMemcpyAsync into
// pageable memory blocks the calling thread until the copy is done, i.e. until the frame's kernels have finished,
// so frame f + 1 would not even be enqueued before frame f has completed and the second stream would buy nothing
// for callers that pass plain NumPy arrays (the reference's own calling convention: np.zeros outputs).
struct ScopedHostPin {
    void* p = nullptr;
    ScopedHostPin(void* ptr, size_t bytes, bool enable) {
        if (!enable || !ptr || bytes < (1u << 16)) return;
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess) { cudaGetLastError(); return; }
        if (a.type != cudaMemoryTypeUnregistered) return;              // already page-locked (or not host memory)
        if (cudaHostRegister(ptr, bytes, cudaHostRegisterDefault) == cudaSuccess) p = ptr;
        else cudaGetLastError();                                        // not registrable: stay pageable
    }
    ~ScopedHostPin() { if (p) cudaHostUnregister(p); }
    ScopedHostPin(const ScopedHostPin&) = delete;
    ScopedHostPin& operator=(const ScopedHostPin&) = delete;
};

int cpet_topo_hist_frames(cpet_ctx* c, int n_frames, const int* n_charges, const float* const* x,
                          const float* const* Q, int n_lines, const float* seeds, const int32_t* n_iter,
                          int64_t n_iter_frame_stride, float step_size, const float dims[3], unsigned flags,
                          float* out_rows, int nd, const double* d_edges, int nc, const double* c_edges,
                          int64_t* counts) {
    CTX_GUARD(c);
    CPET_REQUIRE(n_frames >= 0 && n_lines >= 0, CPET_ERR_INVALID, "negative sizes");
    CPET_REQUIRE((flags & ~CPET_TOPO_CURV_SECOND_DIFF) == 0, CPET_ERR_INVALID, "unknown flag bits 0x%x", flags);
    CPET_REQUIRE(dims != nullptr, CPET_ERR_INVALID, "dims is NULL");
    CPET_REQUIRE(nd >= 1 && nc >= 1 && d_edges && c_edges, CPET_ERR_INVALID, "histogram needs edges, nd >= 1 and nc >= 1");
    CPET_REQUIRE(n_frames == 0 || (n_charges && x && Q && counts), CPET_ERR_INVALID, "NULL frame arrays");
    CPET_REQUIRE(n_lines == 0 || (seeds && n_iter), CPET_ERR_INVALID, "NULL seed/n_iter arrays");
    CPET_REQUIRE(n_iter_frame_stride == 0 || n_iter_frame_stride >= n_lines, CPET_ERR_INVALID,
                 "n_iter_frame_stride must be 0 (shared) or >= n_lines");
    c->last_counters[0] = c->last_counters[1] = c->last_counters[2] = 0;
    if (n_frames == 0) return CPET_OK;
    for (int f = 0; f < n_frames; ++f)
        CPET_REQUIRE(n_charges[f] >= 0 && (n_charges[f] == 0 || (x[f] && Q[f])), CPET_ERR_INVALID,
                     "frame %d: bad charge arrays", f);

    const size_t n = (size_t)(n_lines > 0 ? n_lines : 1);
    const size_t cbytes = sizeof(int64_t) * (size_t)nd * nc;
    const int n_children = n_frames > 1 ? 2 : 1;
    cpet_ctx* ch[2] = {nullptr, nullptr};
    const double* dd[2]; const double* dc[2];
    auto setup = [&]() -> int {
        for (int i = 0; i < n_children; ++i) {
            if (int rc = frames_child(c, i, &ch[i])) return rc;
            cpet_ctx* k = ch[i];
            if (int rc = upload_edges(k, nd, d_edges, nc, c_edges, &dd[i], &dc[i])) return rc;
            if (int rc = k->in0.reserve(sizeof(float) * 3 * n)) return rc;
            if (int rc = k->in1.reserve(sizeof(int32_t) * n)) return rc;
            if (int rc = k->out0.reserve(sizeof(float) * 2 * n)) return rc;
            if (int rc = k->work0.reserve(cbytes)) return rc;
            if (int rc = k->totals.reserve(64)) return rc;
            CPET_CUDA_TRY(cudaMemsetAsync(k->totals.p, 0, 64, k->stream));
            if (n_lines > 0) {
                CPET_CUDA_TRY(cudaMemcpyAsync(k->in0.p, seeds, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, k->stream));
                if (n_iter_frame_stride == 0)
                    CPET_CUDA_TRY(cudaMemcpyAsync(k->in1.p, n_iter, sizeof(int32_t) * n, cudaMemcpyHostToDevice, k->stream));
            }
        }
        return CPET_OK;
    };
    if (int rc = setup()) {
        for (int i = 0; i < n_children; ++i)
            if (ch[i]) cudaStreamSynchronize(ch[i]->stream);     // seeds / edges copies in flight
        return rc;
    }
    int64_t launches = 0;
    // frames_pin = 1: result buffers the caller left pageable are page-locked for the call (both streams are drained
    // before the destructors run).  Off by default: measured, the registration costs more than the blocking copies
    const bool pin = c->tune.frames_pin != 0 && n_frames > 1;
    ScopedHostPin pin_rows(out_rows, out_rows ? sizeof(float) * 2 * n * (size_t)n_frames : 0, pin && n_lines > 0);
    ScopedHostPin pin_counts(counts, cbytes * (size_t)n_frames, pin);
    ScopedHostPin pin_iter(const_cast<int32_t*>(n_iter),
                           n_iter_frame_stride ? sizeof(int32_t) * (size_t)n_iter_frame_stride * (size_t)n_frames : 0,
                           pin && n_lines > 0);
    // every frame's work is enqueued asynchronously; whatever happens, both streams are drained before
    // this call returns, because the copies reference the caller's host buffers
    auto enqueue = [&]() -> int {
        for (int f = 0; f < n_frames; ++f) {
            cpet_ctx* k = ch[f % n_children];
            const double* kd = dd[f % n_children];
            const double* kc = dc[f % n_children];
            if (int rc = cpet_set_charges(k, n_charges[f], x[f], Q[f])) return rc;
            launches += 1;
            if (n_lines > 0) {
                if (n_iter_frame_stride != 0)
                    CPET_CUDA_TRY(cudaMemcpyAsync(k->in1.p, n_iter + (size_t)f * (size_t)n_iter_frame_stride,
                                                  sizeof(int32_t) * n, cudaMemcpyHostToDevice, k->stream));
                if (int rc = launch_topo(k, n_lines, k->in0.as<float>(), k->in1.as<int32_t>(), step_size, dims, flags,
                                         k->out0.as<float>(), nullptr))
                    return rc;
                launches += k->last_counters[0];
                add_evals_kernel<<<1, 1, 0, k->stream>>>(
                    reinterpret_cast<const unsigned long long*>(k->counters.as<unsigned char>() + 8), k->n_charges,
                    k->totals.as<unsigned long long>());
                CPET_CUDA_TRY(cudaGetLastError());
            }
            if (int rc = launch_hist2d(k, 1, n_lines, k->out0.p, false, nd, kd, nc, kc, k->work0.as<unsigned long long>()))
                return rc;
            launches += k->last_counters[0];
            if (n_lines > 0 && out_rows)
                CPET_CUDA_TRY(cudaMemcpyAsync(out_rows + (size_t)f * 2 * n, k->out0.p, sizeof(float) * 2 * n,
                                              cudaMemcpyDeviceToHost, k->stream));
            CPET_CUDA_TRY(cudaMemcpyAsync(counts + (size_t)f * nd * nc, k->work0.p, cbytes, cudaMemcpyDeviceToHost, k->stream));
        }
        return CPET_OK;
    };
    const int rc_enqueue = enqueue();
    unsigned long long tot[2][2] = {{0, 0}, {0, 0}};
    int rc_sync = CPET_OK;
    for (int i = 0; i < n_children; ++i) {
        if (rc_enqueue == CPET_OK &&
            cudaMemcpyAsync(tot[i], ch[i]->totals.p, sizeof(tot[i]), cudaMemcpyDeviceToHost, ch[i]->stream) != cudaSuccess)
            rc_sync = CPET_ERR_CUDA;
        const cudaError_t e = cudaStreamSynchronize(ch[i]->stream);
        if (e != cudaSuccess && rc_enqueue == CPET_OK && rc_sync == CPET_OK) {
            set_error(CPET_ERR_CUDA, "cpet_topo_hist_frames: stream %d failed: %s", i, cudaGetErrorString(e));
            rc_sync = CPET_ERR_CUDA;
        }
    }
    if (rc_enqueue != CPET_OK) return rc_enqueue;
    if (rc_sync != CPET_OK) return rc_sync;
    c->last_counters[0] = launches;
    c->last_counters[2] = (int64_t)(tot[0][0] + tot[1][0]);
    c->last_counters[1] = (int64_t)(tot[0][1] + tot[1][1]);
    return CPET_OK;
}

// ---------------------------------------------------------------- order statistics ----------
int cpet_radix_hist_dev(cpet_ctx* c, int64_t n, const float* d_values, int stride, int offset, int n_targets,
                        const uint32_t* prefixes, int prefix_bits, uint64_t* hist) {
    CTX_GUARD(c);
    CPET_REQUIRE(n >= 0 && stride >= 1 && offset >= 0 && offset < stride, CPET_ERR_INVALID, "bad value layout");
    CPET_REQUIRE(prefixes && hist && (n == 0 || d_values), CPET_ERR_INVALID, "NULL arrays");
    CPET_REQUIRE(n_targets >= 1 && n_targets <= 16, CPET_ERR_INVALID, "1..16 targets per pass");
    const size_t hb = sizeof(unsigned long long) * 256 * (size_t)n_targets;
    if (int rc = c->work2.reserve(hb + 64)) return rc;
    unsigned* d_pre = c->work2.as<unsigned>();
    unsigned long long* d_hist = reinterpret_cast<unsigned long long*>(c->work2.as<unsigned char>() + 64);
    CPET_CUDA_TRY(cudaMemcpyAsync(d_pre, prefixes, sizeof(unsigned) * n_targets, cudaMemcpyHostToDevice, c->stream));
    if (int rc = launch_radix_hist(c, n, d_values, stride, offset, n_targets, d_pre, prefix_bits, d_hist)) return rc;
    CPET_CUDA_TRY(cudaMemcpyAsync(hist, d_hist, hb, cudaMemcpyDeviceToHost, c->stream));
    CPET_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return CPET_OK;
}

static int order_stats_impl(cpet_ctx* c, int64_t n, const float* d_values, int stride, int offset, int n_ranks,
                            const int64_t* ranks, float* out) {
    CPET_REQUIRE(n_ranks >= 1 && n_ranks <= 16, CPET_ERR_INVALID, "1..16 ranks per call");
    CPET_REQUIRE(n >= 1, CPET_ERR_INVALID, "order statistics of an empty array");
    int64_t k[16];
    uint32_t prefix[16];
    std::vector<uint64_t> hist(256 * 16);
    for (int t = 0; t < n_ranks; ++t) {
        CPET_REQUIRE(ranks[t] >= 0 && ranks[t] < n, CPET_ERR_INVALID, "rank %lld out of range", (long long)ranks[t]);
        k[t] = ranks[t];
        prefix[t] = 0;
    }
    int launches = 0;
    for (int pass = 0; pass < 4; ++pass) {
        // pass 0 has no prefix yet: one shared histogram serves every target
        const int nt = pass == 0 ? 1 : n_ranks;
        if (int rc = cpet_radix_hist_dev(c, n, d_values, stride, offset, nt, prefix, 8 * pass, hist.data())) return rc;
        ++launches;
        for (int t = 0; t < n_ranks; ++t) {
            const uint64_t* h = hist.data() + 256 * (pass == 0 ? 0 : t);
            uint64_t cum = 0;
            int b = 0;
            for (; b < 256; ++b) {
                if ((uint64_t)k[t] < cum + h[b]) break;
                cum += h[b];
            }
            CPET_REQUIRE(b < 256, CPET_ERR_STATE, "radix select lost rank %d (values changed between passes?)", t);
            k[t] -= (int64_t)cum;
            prefix[t] = (prefix[t] << 8) | (uint32_t)b;
        }
    }
    for (int t = 0; t < n_ranks; ++t) {
        const uint32_t key = prefix[t];
        const uint32_t u = (key & 0x80000000u) ? (key & 0x7fffffffu) : ~key;
        memcpy(&out[t], &u, sizeof(float));
    }
    c->last_counters[0] = launches;
    c->last_counters[1] = 0;
    c->last_counters[2] = 0;
    return CPET_OK;
}

int cpet_order_stats_dev(cpet_ctx* c, int64_t n, const float* d_values, int stride, int offset, int n_ranks,
                         const int64_t* ranks, float* out) {
    CTX_GUARD(c);
    CPET_REQUIRE(d_values && ranks && out, CPET_ERR_INVALID, "NULL arrays");
    CPET_REQUIRE(stride >= 1 && offset >= 0 && offset < stride, CPET_ERR_INVALID, "bad value layout");
    return order_stats_impl(c, n, d_values, stride, offset, n_ranks, ranks, out);
}

int cpet_order_stats(cpet_ctx* c, int64_t n, const float* values, int stride, int offset, int n_ranks,
                     const int64_t* ranks, float* out) {
    CTX_GUARD(c);
    CPET_REQUIRE(values && ranks && out, CPET_ERR_INVALID, "NULL arrays");
    CPET_REQUIRE(n >= 1 && stride >= 1 && offset >= 0 && offset < stride, CPET_ERR_INVALID, "bad value layout");
    const size_t bytes = sizeof(float) * (size_t)n * stride;
    if (int rc = c->in0.reserve(bytes)) return rc;
    CPET_CUDA_TRY(cudaMemcpyAsync(c->in0.p, values, bytes, cudaMemcpyHostToDevice, c->stream));
    return order_stats_impl(c, n, c->in0.as<float>(), stride, offset, n_ranks, ranks, out);
}

int cpet_chi2_matrix(cpet_ctx* c, int n_hists, int64_t n_bins, const double* H, double* out) {
    CTX_GUARD(c);
    CPET_REQUIRE(n_hists >= 0 && n_bins >= 0, CPET_ERR_INVALID, "negative sizes");
    if (n_hists == 0) return CPET_OK;
    CPET_REQUIRE(H && out, CPET_ERR_INVALID, "NULL arrays");
    const size_t hb = sizeof(double) * (size_t)n_hists * (size_t)n_bins;
    const size_t ob = sizeof(double) * (size_t)n_hists * n_hists;
    if (int rc = c->in0.reserve(hb ? hb : 8)) return rc;
    if (int rc = c->out0.reserve(ob)) return rc;
    DrainOnExit drain(c->stream);
    if (hb) CPET_CUDA_TRY(cudaMemcpyAsync(c->in0.p, H, hb, cudaMemcpyHostToDevice, c->stream));
    if (int rc = launch_chi2(c, n_hists, n_bins, c->in0.as<double>(), c->out0.as<double>())) return rc;
    CPET_CUDA_TRY(cudaMemcpyAsync(out, c->out0.p, ob, cudaMemcpyDeviceToHost, c->stream));
    CPET_CUDA_TRY(cudaStreamSynchronize(c->stream));
    drain.armed = false;
    return CPET_OK;
}

int cpet_chi2_rows_dev(cpet_ctx* c, int n_hists, int64_t n_bins, const double* d_H, int row0, int n_rows,
                       double* d_out) {
    CTX_GUARD(c);
    CPET_REQUIRE(n_hists >= 0 && n_bins >= 0 && n_rows >= 0, CPET_ERR_INVALID, "negative sizes");
    if (n_hists == 0 || n_rows == 0) return CPET_OK;
    CPET_REQUIRE(d_H && d_out, CPET_ERR_INVALID, "NULL arrays");
    return launch_chi2_rows(c, n_hists, n_bins, d_H, row0, n_rows, d_out);
}

int cpet_fp32_peak_probe(cpet_ctx* c, int packed, int iters, double* tflops) {
    CTX_GUARD(c);
    CPET_REQUIRE(tflops != nullptr, CPET_ERR_INVALID, "tflops is NULL");
    return launch_fp32_probe(c, packed, iters, tflops);
}

}  // extern "C"
