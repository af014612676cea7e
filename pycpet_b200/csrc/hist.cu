// hist.cu -- K3: batched distance x curvature 2-D histogram with NumPy's binning rule, and the
// pairwise chi^2 histogram distance.
//
// Replaces np.histogram2d as called by make_histograms (CPET/utils/calculator.py:702-707) and
// distance_numpy / construct_distance_matrix (UC:975-978, 1003-1015).
//
// Binning is bit-exact with numpy/lib/_histograms_impl.py::histogramdd: the float64 edge arrays
// are supplied by the caller (np.linspace), each value is located with a `side='right'` binary
// search in shared memory, a value equal to the last edge falls in the last bin, outliers (and
// NaNs) are dropped.  Counting uses warp-aggregated (match.any) shared-memory atomics, flushed
// once per CTA to 64-bit global counters.
#include "cpet_internal.h"

namespace cpet {

__device__ __forceinline__ int searchsorted_right(const double* __restrict__ e, int n_edges,
                                                  double v) {
    int lo = 0, hi = n_edges;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (e[mid] <= v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

template <typename T, bool SMEM_COUNTS>
__global__ void __launch_bounds__(256) k3_hist2d_kernel(const T* __restrict__ values,
                                                        long long n_per_frame, int nd, int nc,
                                                        const double* __restrict__ d_edges,
                                                        const double* __restrict__ c_edges,
                                                        unsigned long long* __restrict__ counts) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* ed = reinterpret_cast<double*>(smem_raw);
    double* ec = ed + (nd + 1);
    unsigned* cnt = reinterpret_cast<unsigned*>(ec + (nc + 1));
    const int nbins = nd * nc;
    for (int i = threadIdx.x; i <= nd; i += blockDim.x) ed[i] = d_edges[i];
    for (int i = threadIdx.x; i <= nc; i += blockDim.x) ec[i] = c_edges[i];
    if (SMEM_COUNTS)
        for (int i = threadIdx.x; i < nbins; i += blockDim.x) cnt[i] = 0u;
    __syncthreads();

    const int frame = blockIdx.y;
    const T* v = values + (size_t)frame * (size_t)n_per_frame * 2;
    unsigned long long* out = counts + (size_t)frame * (size_t)nbins;
    const double d_last = ed[nd], c_last = ec[nc];

    const long long stride = (long long)gridDim.x * blockDim.x;
    // keep whole warps in the loop so match.any sees a full mask
    const long long n_round = ((n_per_frame + 31) / 32) * 32;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
        int bin = -1;
        if (i < n_per_frame) {
            const double d = (double)v[2 * i], c = (double)v[2 * i + 1];
            int bi = searchsorted_right(ed, nd + 1, d);
            int bj = searchsorted_right(ec, nc + 1, c);
            if (d == d_last) --bi;
            if (c == c_last) --bj;
            if (bi >= 1 && bi <= nd && bj >= 1 && bj <= nc) bin = (bi - 1) * nc + (bj - 1);
        }
        const unsigned peers = __match_any_sync(0xffffffffu, bin);
        if (bin >= 0) {
            const int lane = threadIdx.x & 31;
            if (lane == __ffs(peers) - 1) {
                const unsigned add = (unsigned)__popc(peers);
                if (SMEM_COUNTS) atomicAdd(&cnt[bin], add);
                else atomicAdd(&out[bin], (unsigned long long)add);
            }
        }
    }
    if (SMEM_COUNTS) {
        __syncthreads();
        for (int i = threadIdx.x; i < nbins; i += blockDim.x)
            if (cnt[i]) atomicAdd(&out[i], (unsigned long long)cnt[i]);
    }
}

int launch_hist2d(cpet_ctx* c, int n_frames, int64_t n_per_frame, const void* d_values,
                  bool values_f64, int nd, const double* d_edges_dev, int nc,
                  const double* c_edges_dev, unsigned long long* d_counts) {
    c->last_counters[0] = 0;
    c->last_counters[1] = 0;
    c->last_counters[2] = 0;
    const size_t nbins = (size_t)nd * nc;
    CPET_CUDA_TRY(cudaMemsetAsync(d_counts, 0, sizeof(unsigned long long) * nbins * n_frames, c->stream));
    if (n_frames == 0 || n_per_frame == 0 || nbins == 0) return CPET_OK;
    const size_t edge_bytes = sizeof(double) * (size_t)(nd + nc + 2);
    const size_t smem_counts = edge_bytes + sizeof(unsigned) * nbins;
    const bool in_smem = smem_counts <= (size_t)c->max_smem_optin;
    const size_t smem = in_smem ? smem_counts : edge_bytes;
    CPET_REQUIRE(edge_bytes <= (size_t)c->max_smem_optin, CPET_ERR_INVALID,
                 "histogram edges do not fit in shared memory (nd=%d nc=%d)", nd, nc);
    long long bx = (n_per_frame + 256 * 8 - 1) / (256 * 8);
    const long long cap = (long long)c->sm_count * 8 / (n_frames < 1 ? 1 : n_frames) + 1;
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    dim3 grid((unsigned)bx, (unsigned)n_frames, 1);
    KernelTimer timer(c);
#define CPET_K3_LAUNCH(T, S)                                                                   \
    do {                                                                                       \
        auto kern = k3_hist2d_kernel<T, S>;                                                    \
        CPET_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                           (int)smem));                                        \
        kern<<<grid, 256, smem, c->stream>>>((const T*)d_values, (long long)n_per_frame, nd,   \
                                             nc, d_edges_dev, c_edges_dev, d_counts);          \
    } while (0)
    if (values_f64) { if (in_smem) CPET_K3_LAUNCH(double, true); else CPET_K3_LAUNCH(double, false); }
    else            { if (in_smem) CPET_K3_LAUNCH(float, true);  else CPET_K3_LAUNCH(float, false); }
#undef CPET_K3_LAUNCH
    CPET_CUDA_TRY(cudaGetLastError());
    c->last_counters[0] = 1;
    return CPET_OK;
}

// ---------------------------------------------------------------------------------------------
// Radix-select building block for exact order statistics (global min / max and the 25th / 75th
// percentile neighbours that make_histograms' bin widths need: UC:664-670, scipy.stats.iqr).
// One pass = for every target t, the 256-bin histogram of the next 8 key bits over the values
// whose leading `prefix_bits` bits equal prefixes[t].  Keys are the usual monotone map of float32
// (all NaNs last, as NumPy sorts them).  Four passes pin any rank exactly; the pass loop lives on
// the host (and all-reduces the histograms across ranks for multi-GPU selection).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned monotone_key(float v) {
    if (v != v) return 0xffffffffu;
    const unsigned u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

#define CPET_RADIX_MAX_TARGETS 16

__global__ void __launch_bounds__(256) radix_hist_kernel(const float* __restrict__ values, long long n,
                                                         int stride, int offset, int n_targets,
                                                         const unsigned* __restrict__ prefixes,
                                                         int prefix_bits,
                                                         unsigned long long* __restrict__ hist) {
    __shared__ unsigned cnt[CPET_RADIX_MAX_TARGETS][256];
    __shared__ unsigned pre[CPET_RADIX_MAX_TARGETS];
    for (int i = threadIdx.x; i < n_targets * 256; i += blockDim.x) cnt[i >> 8][i & 255] = 0u;
    if (threadIdx.x < n_targets) pre[threadIdx.x] = prefixes[threadIdx.x];
    __syncthreads();
    const int shift = 24 - prefix_bits;
    const int lane = threadIdx.x & 31;
    // whole warps stay in the loop (match.any needs a full mask); the values of a frame cluster in a few
    // exponent bins, so the lanes of a warp are aggregated per (target, bin) before the shared atomic
    const long long n_round = ((n + 31) / 32) * 32;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_round;
         i += (long long)gridDim.x * blockDim.x) {
        unsigned key = 0u, head = 0u;
        const bool live = i < n;
        if (live) {
            key = monotone_key(values[i * stride + offset]);
            head = prefix_bits ? (key >> (32 - prefix_bits)) : 0u;
        }
        const unsigned bin = (key >> shift) & 255u;
        // targets may share a prefix (neighbouring ranks): every matching target gets the count
        const unsigned long long tag = live ? (((unsigned long long)head << 8) | bin) : (1ull << 40);   // NaN keys are all ones
        const unsigned peers = __match_any_sync(0xffffffffu, tag);
        if (live && lane == __ffs(peers) - 1) {
            const unsigned add = (unsigned)__popc(peers);
            for (int t = 0; t < n_targets; ++t)
                if (prefix_bits == 0 || head == pre[t]) atomicAdd(&cnt[t][bin], add);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_targets * 256; i += blockDim.x)
        if (cnt[i >> 8][i & 255]) atomicAdd(&hist[i], (unsigned long long)cnt[i >> 8][i & 255]);
}

int launch_radix_hist(cpet_ctx* c, long long n, const float* d_values, int stride, int offset,
                      int n_targets, const unsigned* d_prefixes, int prefix_bits,
                      unsigned long long* d_hist) {
    CPET_REQUIRE(n_targets >= 1 && n_targets <= CPET_RADIX_MAX_TARGETS, CPET_ERR_INVALID,
                 "radix select handles 1..%d targets per pass", CPET_RADIX_MAX_TARGETS);
    CPET_REQUIRE(prefix_bits == 0 || prefix_bits == 8 || prefix_bits == 16 || prefix_bits == 24,
                 CPET_ERR_INVALID, "prefix_bits must be 0, 8, 16 or 24");
    CPET_CUDA_TRY(cudaMemsetAsync(d_hist, 0, sizeof(unsigned long long) * 256 * n_targets, c->stream));
    c->last_counters[0] = 0;
    if (n <= 0) return CPET_OK;
    long long blocks = (n + 256 * 8 - 1) / (256 * 8);
    if (blocks > (long long)c->sm_count * 8) blocks = (long long)c->sm_count * 8;
    radix_hist_kernel<<<(unsigned)blocks, 256, 0, c->stream>>>(d_values, n, stride, offset, n_targets,
                                                              d_prefixes, prefix_bits, d_hist);
    CPET_CUDA_TRY(cudaGetLastError());
    c->last_counters[0] = 1;
    return CPET_OK;
}

// ---------------------------------------------------------------------------------------------
// chi^2 distance matrix: one CTA per (i, j>i) pair, FP64.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) chi2_kernel(const double* __restrict__ H, int n,
                                                   long long nbins, double* __restrict__ out) {
    const int i = blockIdx.y, j = blockIdx.x;
    if (j < i) return;
    if (j == i) { if (threadIdx.x == 0) out[(size_t)i * n + i] = 0.0; return; }
    const double* a = H + (size_t)i * nbins;
    const double* b = H + (size_t)j * nbins;
    double s = 0.0;
    for (long long k = threadIdx.x; k < nbins; k += blockDim.x) {
        const double x = a[k], y = b[k];
        const double sum = x + y;
        const double diff = x - y;
        if (sum != 0.0) s += diff * diff / sum;
    }
    __shared__ double red[256];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int w = 128; w >= 1; w >>= 1) {
        if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double d = 0.5 * red[0];
        out[(size_t)i * n + j] = d;
        out[(size_t)j * n + i] = d;
    }
}

// Rows [row0, row0 + n_rows) of the same matrix (every column, both triangles): the unit a rank computes when
// the matrix of a trajectory is split over GPUs.  blockIdx.x = column, so the CTAs in flight share row i.
__global__ void __launch_bounds__(256) chi2_rows_kernel(const double* __restrict__ H, int n, long long nbins,
                                                        int row0, double* __restrict__ out) {
    const int i = row0 + blockIdx.y, j = blockIdx.x;
    double* o = out + (size_t)blockIdx.y * n + j;
    if (j == i) { if (threadIdx.x == 0) *o = 0.0; return; }
    // the same operand order as the (i < j) entry of chi2_kernel, so both triangles carry the same bits
    const double* a = H + (size_t)(i < j ? i : j) * nbins;
    const double* b = H + (size_t)(i < j ? j : i) * nbins;
    double s = 0.0;
    for (long long k = threadIdx.x; k < nbins; k += blockDim.x) {
        const double x = a[k], y = b[k];
        const double sum = x + y;
        const double diff = x - y;
        if (sum != 0.0) s += diff * diff / sum;
    }
    __shared__ double red[256];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int w = 128; w >= 1; w >>= 1) {
        if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) *o = 0.5 * red[0];
}

int launch_chi2_rows(cpet_ctx* c, int n_hists, int64_t n_bins, const double* d_H, int row0, int n_rows,
                     double* d_out) {
    c->last_counters[0] = 0;
    if (n_hists == 0 || n_rows == 0) return CPET_OK;
    CPET_REQUIRE(n_hists <= 65535 && n_rows <= 65535, CPET_ERR_INVALID, "chi2 matrix limited to 65535 histograms");
    CPET_REQUIRE(row0 >= 0 && row0 + n_rows <= n_hists, CPET_ERR_INVALID, "row block outside the matrix");
    KernelTimer timer(c);
    dim3 grid((unsigned)n_hists, (unsigned)n_rows, 1);
    chi2_rows_kernel<<<grid, 256, 0, c->stream>>>(d_H, n_hists, (long long)n_bins, row0, d_out);
    CPET_CUDA_TRY(cudaGetLastError());
    c->last_counters[0] = 1;
    return CPET_OK;
}

int launch_chi2(cpet_ctx* c, int n_hists, int64_t n_bins, const double* d_H, double* d_out) {
    c->last_counters[0] = 0;
    if (n_hists == 0) return CPET_OK;
    CPET_REQUIRE(n_hists <= 65535, CPET_ERR_INVALID, "chi2 matrix limited to 65535 histograms");
    KernelTimer timer(c);
    dim3 grid((unsigned)n_hists, (unsigned)n_hists, 1);
    chi2_kernel<<<grid, 256, 0, c->stream>>>(d_H, n_hists, (long long)n_bins, d_out);
    CPET_CUDA_TRY(cudaGetLastError());
    c->last_counters[0] = 1;
    return CPET_OK;
}

}  // namespace cpet
