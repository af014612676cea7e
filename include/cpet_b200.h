/*
 * cpet_b200.h -- C ABI of libcpetb200.so, the B200 (sm_100a) drop-in for PyCPET's C-shared
 * math module on the Coulomb field / ESP / streamline / histogram hot path.
 *
 * The reference loads `CPET/utils/math_module.c` with ctypes.CDLL (CPET/utils/c_ops.py:9-12,
 * CPET/utils/calculator.py:19-29) and calls it with plain `float *` / `int` arguments.  This library is
 * opened the same way.  It exports
 *
 *   (A) the reference's legacy symbols with identical names, argument order and ownership rules
 *       (caller allocates every output; nothing is retained), so `Math_ops(shared_loc=...)`
 *       constructs and runs unchanged on top of it; and
 *   (B) batched `cpet_*` entry points (one call per grid / per frame of streamlines / per batch
 *       of histograms) that the calculator-level functions should call instead of looping over
 *       points or forking a Pool.  Each has a host-pointer form and a `_dev` form taking device
 *       pointers (the latter never synchronises and runs on the context's stream).
 *
 * Plain pointers and sizes only; no torch / numpy types cross this boundary.
 * There is NO CPU fallback: every entry point runs hand-written CUDA kernels, and fails (status
 * code < 0, or for the `void` legacy symbols: NaN-filled outputs + cpet_last_status() < 0 + a
 * message on stderr) when no sm_100-class device is usable.
 *
 * Reference citations use C = CPET/utils/math_module.c, OPS = CPET/utils/c_ops.py,
 * UC = CPET/utils/calculator.py, SC = CPET/source/calculator.py, GPU = CPET/utils/gpu.py.
 */
#ifndef CPET_B200_H
#define CPET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CPET_ABI_VERSION 1

/* ---------------------------------------------------------------- status / errors ---------- */
#define CPET_OK 0
#define CPET_ERR_INVALID (-1)   /* bad argument (null pointer, negative size, unknown flag)      */
#define CPET_ERR_CUDA (-2)      /* CUDA runtime/driver error, see cpet_last_error()               */
#define CPET_ERR_NO_DEVICE (-3) /* no CUDA device, or device is not compute capability 10.x       */
#define CPET_ERR_STATE (-4)     /* call order: e.g. field requested before cpet_set_charges()     */

int cpet_abi_version(void);
/* Thread-local message / status of the most recent failing call on this thread ("" / 0 if none).
 * Sticky until cpet_clear_error(); every legacy `void` symbol clears it on entry, so after such a
 * call the status is that call's own outcome. */
const char *cpet_last_error(void);
int cpet_last_status(void);
void cpet_clear_error(void);
int cpet_device_count(void);

/* ---------------------------------------------------------------- contexts ----------------- */
/* A context = one device + one stream + the resident packed charge set of the current frame +
 * grow-only scratch.  Not thread-safe; use one per host thread / per GPU. */
typedef struct cpet_ctx cpet_ctx;

int cpet_create(int device, cpet_ctx **out);
/* Borrow an existing CUstream/cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream). */
int cpet_create_on_stream(int device, void *cuda_stream, cpet_ctx **out);
int cpet_destroy(cpet_ctx *ctx);
int cpet_sync(cpet_ctx *ctx);
int cpet_device_of(cpet_ctx *ctx);
/* Diagnostic: which kernel served the last call on this context.  Field / ESP: 0 general point-list kernel
 * (direct form), 1 lattice kernel, 3 general point-list kernel in the hybrid near/far form.  Streamlines: 11 direct-form kernel (k2w), 12 hybrid, charge pairs packed (k2x),
 * 13 hybrid, points packed (k2p). */
int cpet_last_path(cpet_ctx *ctx);
/* Tuning knobs for experiments (the defaults are measured heuristics, profiles/round1_sweep.md):
 * "k1_threads","k1_points" (points or z-nodes per thread),"k1_lanes" (lanes per point: 1, 8, 32),
 * "k1_tile_pairs","k1_stages","k1_splits" (charge-range splits),"k1_unroll" (lattice kernel),
 * "k1_lattice" (-1 auto-detect meshes in the host entry point, 0 off, 1 on),
 * "k1_softscan" (-1 auto, 0 off, 1 on: prove on the device that the softening cannot act on a mesh
 * and run the unsoftened kernel, bit-identical),
 * "k1_esp_mix" (-1 auto, 0 off, 1 on: ESP lattice kernel with every sixth z-node's rsqrt on the FMA pipe),
 * "k1_lat_nodes" (-1 auto, 0 off, 1 on: field lattice kernel with two z-nodes per packed register),
 * "k1_hybrid" (-1 auto, 0 off, 1 on: field sums over >= 2,048 listed points in the hybrid near/far form),
 * "k2_form" (0 auto by queue length, 1 direct-form kernel, 2 hybrid near/far kernel with charge pairs packed,
 * 3 hybrid kernel with point pairs packed), "k2_cap" (streamlines per warp: 1, 2, 4, and 8 in the points-packed
 * kernel),"k2_threads","k2_tile_pairs","k2_stages" (a tile size or stage count forces the streamed charge
 * ring),"k2_unroll" (far-loop unroll of the hybrid kernels),"k2_amax" (largest rounding amplification a charge
 * may have to take the expanded far form; default 8),"k2_tail4","k2_tail2" (end-of-queue policy of the
 * points-packed kernel),"k2_sort" (-1 auto, 0, 1),"frames_pin" (cpet_topo_hist_frames: page-lock pageable
 * result buffers for the call; default 0),"timing" (record kernel events).
 * value <= 0 restores the built-in heuristic (except the three-state keys). */
int cpet_set_tuning(cpet_ctx *ctx, const char *key, int value);
/* Counters of the last kernel-launching call: [0]=kernels launched, [1]=pair evaluations
 * (algorithmic: points x charges, or sum over lines of (K+2) x charges), [2]=field evaluations. */
int cpet_last_counters(cpet_ctx *ctx, int64_t out[3]);

/* Upload (and pack) the charge set of one frame: x (M,3) float32 row-major, Q (M,) float32 --
 * the `x`, `Q` arguments every reference entry point takes (OPS:250-375).  Done once per frame;
 * all following calls on this context use it.  The host form enqueues the upload and returns: with
 * pageable memory (NumPy arrays) the data has been staged by then; page-locked x / Q must stay
 * untouched until the next host-pointer call on this context or cpet_sync() returns.  Every other
 * host-pointer entry point drains its stream before returning, on error paths too. */
int cpet_set_charges(cpet_ctx *ctx, int n_charges, const float *x, const float *Q);
int cpet_set_charges_dev(cpet_ctx *ctx, int n_charges, const float *d_x, const float *d_Q);

/* ---------------------------------------------------------------- K1: grids ---------------- */
#define CPET_FIELD_SOFTEN 1u  /* r^2 := max(r^2, 1e-6) as compute_looped_field (C:407,433)         */
#define CPET_OUT_CONCAT 2u    /* field: (N,6) f32 rows [x0|E] (UC:446-447)                          */
                              /* esp:   (N,4) f16 rows [x0|phi] (UC:473-475)                        */

/* E(p_i) = k sum_j q_j (p_i - x_j)/|p_i - x_j|^3, k = 14.3996451 (C:412).
 * Replaces compute_looped_field (C:405-451; with CPET_FIELD_SOFTEN) and per-point loops over
 * calc_field / calc_field_base (C:255-333; without).  x0: (N,3) f32.  out: (N,3) f32, or (N,6)
 * f32 with CPET_OUT_CONCAT. */
int cpet_field_grid(cpet_ctx *ctx, int n_points, const float *x0, unsigned flags, float *out);
int cpet_field_grid_dev(cpet_ctx *ctx, int n_points, const float *d_x0, unsigned flags,
                        float *d_out);

/* phi(p_i) = k sum_j q_j/|p_i - x_j| (no softening).  Replaces the Python loop over
 * calc_esp_base in compute_ESP_on_grid (UC:450-475, C:453-486).  out: (N,) f32, or (N,4) f16 with
 * CPET_OUT_CONCAT (each value rounded f64->f32->f16 as the reference's casts do). */
int cpet_esp_grid(cpet_ctx *ctx, int n_points, const float *x0, unsigned flags, void *out);
int cpet_esp_grid_dev(cpet_ctx *ctx, int n_points, const float *d_x0, unsigned flags, void *d_out);

/* The same two sums on a tensor-product lattice: point (i,j,k) = (xs[i], ys[j], zs[k]), flattened
 * with k fastest -- exactly the mesh initialize_box_points_uniform builds (UC:218-233:
 * linspace per axis, meshgrid(indexing="ij"), reshape(-1,3)) and compute_field_on_grid /
 * compute_ESP_on_grid receive.  dx, dy and dx^2+dy^2 are shared along z, which removes a quarter of
 * the FP32 instructions per pair; with the same charge-range split ("k1_splits") results are
 * bit-identical to the general kernel on the expanded point list, otherwise they differ in the
 * order of the FP64 partial sums only.  The host forms cpet_field_grid / cpet_esp_grid recognise
 * such meshes (>= 4096 points) by themselves and take this kernel; with CPET_FIELD_SOFTEN and
 * >= 2e9 pair-evaluations a device scan first proves that max(r^2, 1e-6) cannot act, and the
 * unsoftened instantiation then serves the call with the same bits ("k1_softscan").
 * Output layouts and flags as above (N = nx*ny*nz rows). */
int cpet_field_lattice(cpet_ctx *ctx, int nx, int ny, int nz, const float *xs, const float *ys,
                       const float *zs, unsigned flags, float *out);
int cpet_field_lattice_dev(cpet_ctx *ctx, int nx, int ny, int nz, const float *d_xs,
                           const float *d_ys, const float *d_zs, unsigned flags, float *d_out);
int cpet_esp_lattice(cpet_ctx *ctx, int nx, int ny, int nz, const float *xs, const float *ys,
                     const float *zs, unsigned flags, void *out);
int cpet_esp_lattice_dev(cpet_ctx *ctx, int nx, int ny, int nz, const float *d_xs, const float *d_ys,
                         const float *d_zs, unsigned flags, void *d_out);

/* One streamline step for N independent points: out_i = p_i + h E(p_i)/|E(p_i)|, no zero guard
 * (propagate_topo, C:489-503; field without softening, C:296-333).  out: (N,3) f32. */
int cpet_propagate(cpet_ctx *ctx, int n_points, const float *x0, float step_size, float *out);
int cpet_propagate_dev(cpet_ctx *ctx, int n_points, const float *d_x0, float step_size,
                       float *d_out);

/* ---------------------------------------------------------------- K2: streamlines ---------- */
#define CPET_TOPO_CURV_SECOND_DIFF 1u /* curvature from FP32 second differences of positions,
                                         literally C:575-580.  Default: the algebraically identical
                                         |e_k x e_k+1| / h from consecutive unit field directions,
                                         which has no cancellation. */

/* All streamlines of one frame.  Per-line semantics of thread_operation (C:523-591): advance the
 * seed by p += h E(p)/|E(p)| until n_iter steps or the point is strictly outside +-dims; take two
 * look-ahead steps at both ends; out[i] = { |seed - end|, (kappa_seed + kappa_end)/2 }.
 * Replaces the Pool over task_complete_thread (SC:675-712) and the torch path (SC:793-978,
 * GPU:25-399).  Rows are returned in SEED order.
 *   seeds (L,3) f32; n_iter (L,) int32; dims (3,) f32 half-widths; out (L,2) f32;
 *   steps (L,) int32 = steps actually taken per line, may be NULL. */
int cpet_topo_batch(cpet_ctx *ctx, int n_lines, const float *seeds, const int32_t *n_iter,
                    float step_size, const float dims[3], unsigned flags, float *out,
                    int32_t *steps);
int cpet_topo_batch_dev(cpet_ctx *ctx, int n_lines, const float *d_seeds, const int32_t *d_n_iter,
                        float step_size, const float dims[3], unsigned flags, float *d_out,
                        int32_t *d_steps);

/* ---------------------------------------------------------------- K3: histogram / chi^2 ---- */
/* Batched np.histogram2d (UC:702-707): for each of n_frames frames, counts[f] (nd,nc) int64 of
 * (dist, curv) pairs with the NumPy rule: edges given explicitly (nd+1 and nc+1 float64, as
 * np.linspace produces them), bin = searchsorted(edges, v, 'right') - 1, right edge inclusive,
 * outliers and NaNs dropped.  values: n_frames x n_per_frame rows of 2.
 *   cpet_hist2d      : host float64 values (what make_histograms parses from .top text)
 *   cpet_hist2d_f32  : host float32 values (the (L,2) array K2 returns; f32->f64 is exact)
 *   cpet_hist2d_dev  : device float32 values (K2 output stays on the GPU) */
int cpet_hist2d(cpet_ctx *ctx, int n_frames, int64_t n_per_frame, const double *values, int nd,
                const double *d_edges, int nc, const double *c_edges, int64_t *counts);
int cpet_hist2d_f32(cpet_ctx *ctx, int n_frames, int64_t n_per_frame, const float *values, int nd,
                    const double *d_edges, int nc, const double *c_edges, int64_t *counts);
int cpet_hist2d_dev(cpet_ctx *ctx, int n_frames, int64_t n_per_frame, const float *d_values,
                    int nd, const double *d_edges_host, int nc, const double *c_edges_host,
                    int64_t *d_counts);

/* One frame end to end: cpet_topo_batch followed by cpet_hist2d on its (L,2) rows WITHOUT the rows
 * leaving the device in between (what run_topo + make_histograms do through a .top text file,
 * TOP:96-127 + UC:596-718).  out_rows (L,2) f32 and steps (L,) i32 may each be NULL when only the
 * histogram is wanted; counts is (nd,nc) int64. */
int cpet_topo_hist(cpet_ctx *ctx, int n_lines, const float *seeds, const int32_t *n_iter,
                   float step_size, const float dims[3], unsigned flags, float *out_rows,
                   int32_t *steps, int nd, const double *d_edges, int nc, const double *c_edges,
                   int64_t *counts);

/* MD-frame batch (BASELINE "1000 frames x topology"; in the reference one CPET.run() iteration per
 * PDB file, TOP:96-127, each ending in compute_topo_complete_c_shared SC:675-712, and later
 * make_histograms over the .top files UC:596-718).  The frames share seeds (n_lines,3), box, step
 * and bin edges; frame f has its own charge set x[f] (n_charges[f],3), Q[f] (n_charges[f],) and,
 * when n_iter_frame_stride != 0, its own n_iter row at n_iter + f*n_iter_frame_stride (0 = one row
 * shared by all frames).  counts is (n_frames,nd,nc) int64; out_rows is (n_frames,n_lines,2) f32 or
 * NULL.  Frames alternate between two internal streams, so the host<->device copies of one frame
 * overlap the kernels of its neighbours.  Pinned (page-locked) host buffers make every copy asynchronous; with
 * plain pageable arrays (the reference's own calling convention) a copy back blocks the enqueueing thread until
 * its frame has finished, which costs 0-1.3 % because the next frame is already queued on the other stream
 * (103,823-line frames 2.003 against 2.004 ms, 1M-line frames 17.53 against 17.30 ms, profiles/round2_frames_pin.txt).
 * Tuning key "frames_pin" = 1 page-locks pageable result buffers with cudaHostRegister for the duration of the
 * call; the registration costs more than it saves at these sizes (2.21 / 18.32 ms per frame), so it is off by default.
 * Every frame's result is identical to a cpet_topo_hist call on that frame alone. */
int cpet_topo_hist_frames(cpet_ctx *ctx, int n_frames, const int *n_charges, const float *const *x,
                          const float *const *Q, int n_lines, const float *seeds,
                          const int32_t *n_iter, int64_t n_iter_frame_stride, float step_size,
                          const float dims[3], unsigned flags, float *out_rows, int nd,
                          const double *d_edges, int nc, const double *c_edges, int64_t *counts);

/* Exact order statistics of float32 values on the device: out[t] = the ranks[t]-th smallest
 * (0-based) of values[i*stride + offset], i < n, NaNs ordered last as NumPy sorts them.  Rank 0 and
 * n-1 give the global min / max, the neighbours of 0.25(n-1) and 0.75(n-1) give scipy.stats.iqr,
 * i.e. everything make_histograms needs for its bin plan (UC:664-685) without a host sort.
 * MSB radix select, four 8-bit passes; up to 16 ranks per call.
 * cpet_radix_hist_dev is the single pass (for every target t: 256-bin histogram of the next 8 key
 * bits among values whose leading prefix_bits bits equal prefixes[t]); a multi-GPU caller
 * all-reduces `hist` between passes (pycpet_b200/sharding.py). */
int cpet_order_stats(cpet_ctx *ctx, int64_t n, const float *values, int stride, int offset,
                     int n_ranks, const int64_t *ranks, float *out);
int cpet_order_stats_dev(cpet_ctx *ctx, int64_t n, const float *d_values, int stride, int offset,
                         int n_ranks, const int64_t *ranks, float *out);
int cpet_radix_hist_dev(cpet_ctx *ctx, int64_t n, const float *d_values, int stride, int offset,
                        int n_targets, const uint32_t *prefixes, int prefix_bits, uint64_t *hist);

/* Pairwise chi^2 distance matrix (UC:975-978, UC:1003-1015):
 * out[i][j] = 1/2 sum_{b: h_i[b]+h_j[b] != 0} (h_i[b]-h_j[b])^2 / (h_i[b]+h_j[b]); diagonal 0.
 * H: (n_hists, n_bins) float64 host; out: (n_hists, n_hists) float64 host. */
int cpet_chi2_matrix(cpet_ctx *ctx, int n_hists, int64_t n_bins, const double *H, double *out);
/* Rows [row0, row0 + n_rows) of that matrix from device-resident histograms (the share of one GPU when
 * construct_distance_matrix, UC:1003-1015, is split over ranks; both triangles carry the same bits).
 * d_H: (n_hists, n_bins) float64 device; d_out: (n_rows, n_hists) float64 device; asynchronous. */
int cpet_chi2_rows_dev(cpet_ctx *ctx, int n_hists, int64_t n_bins, const double *d_H, int row0,
                       int n_rows, double *d_out);

/* ---------------------------------------------------------------- text outputs ------------- */
/* Byte-compatible replacement for the np.savetxt calls that write the path's results: `.top`
 * (TOP:123, default "%.18e") and the body of `_efield.dat` / `_esp.dat` (IO:98-102, "%.3f"; the
 * 7-line header IO:59-85 is passed in `header`).  Host code, all cores; dtype 0 = float32,
 * 1 = float64, 2 = float16; rows are space separated, '\n' terminated. */
int cpet_write_rows(const char *path, const char *header, const void *data, int dtype,
                    int64_t n_rows, int n_cols, const char *fmt, int n_threads);

/* The way back in: the parse make_histograms repeats three times per `.top` file (UC:603-607 count,
 * UC:626-633 read, UC:690-698 read again): lines starting with '#' are skipped (blank lines too),
 * the first n_cols whitespace-separated numbers of every other line are converted with a correctly
 * rounded decimal->double conversion -- bit for bit the value Python's float() returns, "inf" /
 * "-inf" / "nan" as np.savetxt writes them included -- and stored row-major in out (n_rows, n_cols)
 * float64; further columns on a line are ignored like the reference's line[0], line[1].  Host
 * code, all cores (n_threads <= 0).  cpet_count_rows sizes the array; cpet_read_rows fails with
 * CPET_ERR_INVALID if the file holds another number of data lines or a line with fewer numbers. */
int cpet_count_rows(const char *path, int64_t *n_rows, int n_threads);
int cpet_read_rows(const char *path, int n_cols, int64_t n_rows, double *out, int n_threads);

/* ---------------------------------------------------------------- measurement -------------- */
/* Sustained non-tensor FP32 rate of this device from a register-resident FMA loop.
 * packed=0: FFMA, packed=1: FFMA2 (fma.rn.f32x2).  Returns TFLOP/s in *tflops (2 flop/FMA). */
int cpet_fp32_peak_probe(cpet_ctx *ctx, int packed, int iters, double *tflops);
/* Time of the kernels of the last call as measured with CUDA events on the context's stream
 * (milliseconds; 0 if timing was not enabled with cpet_set_tuning(ctx,"timing",1)). */
int cpet_last_kernel_ms(cpet_ctx *ctx, double *ms);
/* Durations (ms) of the dominant-kernel launches recorded since the previous call of this function
 * (up to 256, oldest first), without any synchronisation having happened in between; resets the
 * record.  Needs cpet_set_tuning(ctx,"timing",1). */
int cpet_kernel_times(cpet_ctx *ctx, double *ms, int max_n, int *n_out);

/* ================================================================ (A) legacy symbols ======= */
/* Same names / argument order as the reference's math_module.c so that OPS:8-159 binds them.
 * All use an implicit process-wide context on device $CPET_B200_DEVICE (default 0), created
 * lazily on first call (never at load time: the reference forks Pool workers, SC:690).
 * Hot-path symbols: */
void compute_looped_field(int total_points, int n_charges, float *x_0, float *x, float *Q,
                          float *E); /* C:405  OPS:151-159 */
void compute_batched_field(int total_points, int batch_size, int n_charges, float *x_0, float *x,
                           float *Q, float *E); /* C:374  OPS:140-149 (no softening) */
void calc_field(float *E, float *x_init, int n_charges, float *x, float *Q);      /* C:255 OPS:113 */
void calc_field_base(float *E, float *x_init, int n_charges, float *x, float *Q); /* C:296 OPS:122;
                                                                       ADDS into E like the reference */
void calc_esp_base(float *ESP, float *x_init, int n_charges, float *x, float *Q); /* C:453 OPS:131;
                                                                       ADDS into ESP[0]              */
void thread_operation(int n_charges, int n_iter, float step_size, float *x_0, float *dimensions,
                      float *x, float *Q, float *ret); /* C:523  OPS:85-95 */
/* Helper symbols Math_ops.__init__ sets argtypes on (OPS:26-83); small device kernels. */
void einsum_ij_i(int rows, int cols, float *A, float *ret);                          /* C:237 */
void einsum_ij_i_batch(int batch, int rows, int cols, float *A, float *ret);         /* C:214 */
void einsum_operation(int rows, float *r_mag, float *Q, float *R, float *result);    /* C:181 */
void einsum_operation_batch(int batch, int rows, float *r_mag, float *Q, float *R,
                            float *result);                                          /* C:135 */
void vecaddn(float *ret, float *A, float *B, int lenA);                              /* C:124 */
void dot(double *ret, double *A, double *B, int rows, int cols);                     /* C:49  */
void sparse_dot(double *ret, int *indptr, int indptrlen, int *indA, int lenindA, double *A,
                int lenA, double *B, int size_array);                                /* C:15  */
/* Development-only dipole tracer of the reference (C:593-660, OPS:97-111): exported so that
 * Math_ops.__init__ binds; point dipoles mu (n,3). */
void thread_operation_dipole(int n_dipoles, int n_iter, float step_size, float *x_0,
                             float *dimensions, float *x, float *mu, float *ret);

#ifdef __cplusplus
}
#endif
#endif /* CPET_B200_H */
