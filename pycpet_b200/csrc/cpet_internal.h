// cpet_internal.h -- host-side internals of libcpetb200.so (context, scratch, launchers).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/cpet_b200.h"
#include "common.cuh"

namespace cpet {

void set_error(int status, const char* fmt, ...);

#define CPET_CUDA_TRY(expr)                                                                  \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            cpet::set_error(CPET_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,                   \
                            cudaGetErrorString(_e), __FILE__, __LINE__);                     \
            return CPET_ERR_CUDA;                                                            \
        }                                                                                    \
    } while (0)

#define CPET_REQUIRE(cond, status, ...)                                                      \
    do {                                                                                     \
        if (!(cond)) {                                                                       \
            cpet::set_error(status, __VA_ARGS__);                                            \
            return status;                                                                   \
        }                                                                                    \
    } while (0)

// Grow-only device buffer.
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes);
    void release();
    template <typename T>
    T* as() const { return reinterpret_cast<T*>(p); }
};

struct Tuning {
    int k1_threads = 0, k1_points = 0, k1_lanes = 0, k1_tile_pairs = 0, k1_stages = 0;
    int k1_splits = 0;
    int k1_unroll = 0;     // lattice kernel inner-loop unroll (1, 2 or 4; 0 = default 2)
    int k1_lattice = -1;   // -1 auto (detect z-fastest tensor-product meshes in the host entry point), 0 off, 1 on
    int k1_hybrid = -1;    // general field kernel in the hybrid near/far form: -1 auto (lists >= 2,048 points), 0 off, 1 on
    int k1_lat_nodes = -1; // field lattice kernel with node pairs packed: -1 auto (meshes >= 1e5 nodes), 0 off, 1 on
    int k1_esp_mix = -1;   // ESP lattice kernel: -1 auto (large meshes), 0 off, 1 on: every sixth z-node's rsqrt on the FMA pipe
    int k1_softscan = -1;  // lattice kernel: -1 auto (scan for charges on grid nodes when the mesh is large), 0 off, 1 on
    int k2_threads = 0, k2_tile_pairs = 0, k2_stages = 0;
    int k2_sort = -1;   // -1 heuristic (on), 0 off, 1 on
    int k2_cap = 0;     // streamlines per warp (1, 2, 4; 8 in the points-packed kernel; 0 = heuristic)
    int k2_form = 0;    // 0 = auto (points-packed hybrid kernel, charge-pair-packed hybrid for short queues),
                        // 1 = round-1 direct-form kernel, 2 = charge-pair-packed hybrid, 3 = points-packed hybrid
    int k2_unroll = 0;  // hybrid kernel: far-loop unroll of the 4-point pass (3, 4 or 6; 0 = 3)
    int k2_tail4 = -1, k2_tail2 = -1;   // points-packed kernel, end of the queue (experiments; -1 = default 4 / 0)
    int k2_amax = 0;    // hybrid kernel: largest rounding amplification a far charge may have (0 = 8)
    int frames_pin = 0; // cpet_topo_hist_frames: 1 = page-lock pageable result buffers for the call (measured slower: off)
    int timing = 0;
};

}  // namespace cpet

struct cpet_ctx {
    int device = 0;
    int sm_count = 0;
    int max_smem_optin = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    // current frame
    int n_charges = 0;
    int n_pairs = 0;                 // padded
    bool charges_set = false;        // cpet_set_charges* has been called on this context
    cpet::DevBuf charges;            // ChargePair[n_pairs]
    cpet::DevBuf charge_blocks;      // ChargeBlock[ceil(n_pairs / 32)], zero-charge padded
    cpet::DevBuf raw_x, raw_q;       // staging for host uploads
    cpet::DevBuf xblocks, xchunks;   // hybrid streamline kernel: XBlock[near|far-|far+] of the last launch, per-chunk class counts
    // scratch
    cpet::DevBuf in0, in1, out0, out1, work0, work1, work2, counters, flags, totals;
    cpet::Tuning tune;
    int64_t last_counters[3] = {0, 0, 0};
    double last_kernel_ms = 0.0;
    int last_path = 0;               // diagnostic: 1 = the lattice kernel served the last K1 call
    // ring of event pairs: one pair per timed kernel launch (timing=1), read back without forcing a
    // synchronisation inside the timed region (cpet_kernel_times)
    static constexpr int kTimerRing = 256;
    cudaEvent_t ev0[kTimerRing] = {}, ev1[kTimerRing] = {};
    int timer_count = 0;      // launches recorded since the last cpet_kernel_times()
    cpet_ctx* pipe[2] = {nullptr, nullptr};   // child contexts of cpet_topo_hist_frames (own streams)
};

namespace cpet {

// ---- launchers implemented in the .cu files (all enqueue on ctx->stream, never sync) --------
int launch_pack_charges(cpet_ctx* c, int n_charges, const float* d_x, const float* d_q);

// K1.  out_kind: 0 = (N,3) f32, 1 = (N,6) f32 [x0|E], 2 = (N,) f32 phi, 3 = (N,4) f16 [x0|phi],
//                4 = (N,3) f32 p + step*E/|E|
int launch_field_grid(cpet_ctx* c, int mode, int n_points, const float* d_x0, int out_kind,
                      void* d_out, float step = 0.0f);

// K1 on a tensor-product lattice xs x ys x zs (z fastest); same out_kind codes
int launch_field_lattice(cpet_ctx* c, int mode, int nx, int ny, int nz, const float* d_xs,
                         const float* d_ys, const float* d_zs, int out_kind, void* d_out);

int detect_lattice(cpet_ctx* c, int n_points, const float* d_x0, int* is_lattice, int* nx, int* ny,
                   int* nz, const float** d_axes);

// K2
int launch_topo(cpet_ctx* c, int n_lines, const float* d_seeds, const int32_t* d_n_iter,
                float step, const float dims[3], unsigned flags, float* d_out, int32_t* d_steps);

// K2 internals shared by the streamline kernels (topo.cu, topo8.cu)
struct K2XMeta {
    unsigned ext[3];                 // max |seed coordinate| per axis, float bits
    int n_near, n_far;               // charges per class
    int nb_near, nb_far;             // blocks per class, in array order
    int nb_total;
};
int prepare_queue(cpet_ctx* c, int n_lines, const int32_t* d_n_iter, bool do_sort, unsigned int** queue,
                  unsigned long long** evals, const int32_t** order, int* launches);
int pack_hybrid(cpet_ctx* c, int n_lines, const float* d_seeds, float step, const float dims[3], int layout,
                int max_blocks, K2XMeta** meta_out, int* launches);
bool topo8_wants(cpet_ctx* c, int n_lines);
int launch_topo_points_packed(cpet_ctx* c, int n_lines, const float* d_seeds, const int32_t* d_n_iter, float step,
                              const float dims[3], unsigned flags, float* d_out, int32_t* d_steps);

// K3
int launch_hist2d(cpet_ctx* c, int n_frames, int64_t n_per_frame, const void* d_values,
                  bool values_f64, int nd, const double* d_edges_dev, int nc,
                  const double* c_edges_dev, unsigned long long* d_counts);
int launch_chi2(cpet_ctx* c, int n_hists, int64_t n_bins, const double* d_H, double* d_out);
int launch_chi2_rows(cpet_ctx* c, int n_hists, int64_t n_bins, const double* d_H, int row0, int n_rows,
                     double* d_out);
int launch_radix_hist(cpet_ctx* c, long long n, const float* d_values, int stride, int offset,
                      int n_targets, const unsigned* d_prefixes, int prefix_bits,
                      unsigned long long* d_hist);

// misc
int launch_fp32_probe(cpet_ctx* c, int packed, int iters, double* tflops);

struct KernelTimer {
    cpet_ctx* c;
    explicit KernelTimer(cpet_ctx* ctx);
    ~KernelTimer();
};

}  // namespace cpet
