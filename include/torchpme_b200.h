/*
 * torchpme_b200 -- C ABI of the B200-native PME / P3M hot path.
 *
 * The reference (lab-cosmo/torch-pme) has no FFI layer: its hot path is the Python
 * call  PMECalculator/P3MCalculator.forward()  built from the ops cited below.  This
 * header is the boundary a maintainer binds with ctypes: plain
 * pointers and sizes, no torch types.  All array pointers are DEVICE pointers unless
 * marked "host"; every call is asynchronous on `stream` (a cudaStream_t passed as
 * void*), allocates nothing visible to the caller and is CUDA-graph capturable
 * (plan creation excepted).  Return value: 0 on success, non-zero on failure
 * (tpme_last_error() gives the message).
 *
 * Ordering: the calls are ordered like any other work on `stream`.  The float64 kernels
 * of the mesh entry points (tile sort / spread / gather, FFT, filter) are launched with
 * programmatic stream serialisation: a kernel may become resident while the kernel before
 * it on the same stream is still running, but it touches nothing but its own shared memory
 * and read-only inputs of the whole step until that kernel has completed
 * (griddepcontrol.wait), so callers need no extra synchronisation.  TPME_PDL=0 in the
 * environment turns the attribute off.
 *
 * Conventions (same as the reference): cell rows are lattice vectors; `r2u` is the
 * 3x3 row-major matrix with u = r @ r2u = ns * (r @ cell^-1)
 * (src/torchpme/lib/mesh_interpolator.py:326); positions are (N,3) row-major,
 * per-point weights / charges are (N,C) row-major, meshes are (C,nx,ny,nz) with z
 * fastest, half-complex meshes are (C,nx,ny,nz/2+1) interleaved (re,im).
 * dtype: 0 = float32, 1 = float64.   method: 0 = "P3M", 1 = "Lagrange".
 */
#ifndef TORCHPME_B200_H
#define TORCHPME_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TPME_ABI_VERSION 1

int tpme_abi_version(void);
const char* tpme_last_error(void);

/* ---- mesh interpolation -------------------------------------------------------------
 * replaces MeshInterpolator.compute_weights + points_to_mesh
 * (src/torchpme/lib/mesh_interpolator.py:303-377, 379-426).
 * mesh[c, m] (+)= sum_i weights[i, c] * wx * wy * wz.  The mesh is zeroed first unless
 * `accumulate` is non-zero. */
int tpme_spread(int dtype, const void* positions, const void* weights, int64_t n_points,
                int n_channels, const double* r2u_host, int nx, int ny, int nz, int nodes,
                int method, void* mesh, int accumulate, void* stream);

/* Optional fused O(N) epilogue of the gathers (NULL = plain gather).  With it
 *   values[i,c]  = values[i,c] + scale * gathered - add_coef[i,c] * self_half - background * dc[c]
 * which folds the 1/(2 V) factor, the self term and the neutralising-background term of
 * calculators/pme.py:117-143 into the gather (dc[c] = sum_i q[i,c], see tpme_kfilter_apply), and
 *   grad_positions[i,:] = vjp_scale * (vjp + sum_c coef2[i,c] * dvalues2[i,c,:])   (if coef2 != NULL)
 * which adds the saved forward derivative to the backward gather. */
typedef struct tpme_point_epilogue {
  const void* add_coef;   /* (N,C) device */
  const void* dc;         /* (C,) device */
  double scale, self_half, background;
  const void* coef2;      /* (N,C) device, may be NULL */
  const void* dvalues2;   /* (N,C,3) device */
  double vjp_scale;
} tpme_point_epilogue;

/* replaces MeshInterpolator.mesh_to_points (mesh_interpolator.py:428-457).
 *   values[i, c]      = sum_m mesh[c, m] * wx wy wz              (may be NULL)
 *   dvalues[i, c, :]  = d values[i, c] / d positions[i, :]        (may be NULL)
 * i.e. what PyTorch's tape yields by differentiating the weight polynomials. */
int tpme_gather(int dtype, const void* mesh, const void* positions, int64_t n_points,
                int n_channels, const double* r2u_host, int nx, int ny, int nz, int nodes,
                int method, void* values, void* dvalues, const tpme_point_epilogue* epilogue,
                void* stream);

/* vector-Jacobian product shared by the backward of spread and gather:
 *   grad_positions[i, :] (+)= sum_c coef[i, c] * d/dr_i sum_m mesh[c, m] wx wy wz
 *   values[i, c]          = sum_m mesh[c, m] wx wy wz     (same pass; may be NULL)
 *   grad_r2u[b, a]       +=  sum_i r_i[b] * (same sum differentiated w.r.t. u_a)   (may be NULL)
 * grad_r2u is a device array of 9 reals that the caller zeroes. */
int tpme_gather_vjp(int dtype, const void* mesh, const void* positions, const void* coef,
                    int64_t n_points, int n_channels, const double* r2u_host, int nx, int ny,
                    int nz, int nodes, int method, void* grad_positions, void* values,
                    int accumulate, void* grad_r2u, const tpme_point_epilogue* epilogue,
                    void* stream);

/* Slab variants of the three calls above for a mesh decomposed into x slabs (one per GPU,
 * SURVEY.md section 8e): `mesh` is the local slab (C, nx_local, ny, nz) holding the planes
 * x0 .. x0 + nx_local - 1 of the global (nx, ny, nz) mesh.  Stencil nodes outside the slab are
 * skipped, so spreading needs no halo reduction and the gathers return PARTIAL sums whose sum
 * over the slabs is the full result (the caller all-reduces them together with the real-space
 * part).  x0 = 0, nx_local = nx is exactly the plain call. */
int tpme_spread_slab(int dtype, const void* positions, const void* weights, int64_t n_points,
                     int n_channels, const double* r2u_host, int nx, int ny, int nz, int x0,
                     int nx_local, const int* point_list, const int* list_count, int nodes,
                     int method, void* mesh, int accumulate, void* stream);
int tpme_gather_slab(int dtype, const void* mesh, const void* positions, int64_t n_points,
                     int n_channels, const double* r2u_host, int nx, int ny, int nz, int x0,
                     int nx_local, const int* point_list, const int* list_count, int nodes,
                     int method, void* values, void* dvalues, const tpme_point_epilogue* epilogue,
                     void* stream);
int tpme_gather_vjp_slab(int dtype, const void* mesh, const void* positions, const void* coef,
                         int64_t n_points, int n_channels, const double* r2u_host, int nx, int ny,
                         int nz, int x0, int nx_local, const int* point_list,
                         const int* list_count, int nodes, int method, void* grad_positions,
                         void* values, int accumulate, void* grad_r2u,
                         const tpme_point_epilogue* epilogue, void* stream);
/* `point_list` / `list_count` (device, may both be NULL = all points): the indices of the points
 * whose stencil reaches into the slab and their number, as written by
 * tpme_slab_select_points (capacity n_points ints + 1 int).  With a list the kernels stride over
 * it -- a rank of a W-GPU run touches ~1/W of the atoms -- and leave the outputs of all other
 * points untouched (gathers into zero-initialised / accumulated outputs). */
int tpme_slab_select_points(int dtype, const void* positions, int64_t n_points,
                            const double* r2u_host, int nx, int ny, int nz, int x0, int nx_local,
                            int nodes, int* point_list, int* list_count, void* stream);

/* ---- tiled mesh interpolation: cell-sorted atoms + shared-memory mesh tiles moved by TMA bulk copies ----
 * Same operations as tpme_spread / tpme_gather / tpme_gather_vjp (mesh_interpolator.py:303-457) for
 * power-of-two meshes and 4 interpolation nodes.  tpme_tile_plan_make chooses the tiling (returns 3 when
 * the mesh / stencil is not covered: use the direct calls); tpme_tile_sort bins the points once per set
 * of positions; the spread accumulates a pencil of the mesh in shared memory and flushes it with
 * cp.reduce.async.bulk, the gather stages the pencil with cp.async.bulk behind an mbarrier.
 * Workspace (device, caller-allocated): bin_count (tpme_tile_bin_count_ints(plan) int32, 8-byte aligned:
 * the counters followed by the scratch of the single-pass scan), bin_start (n_bins + 1) int32,
 * key_rank (2 N) int32 8-byte aligned, sorted_rec (4 N) reals 16-byte aligned, sorted_idx (N) int32. */
typedef struct tpme_tile_plan {
  int nx, ny, nz, nodes;
  int tx, ty, zw;             /* pencil footprint (first stencil nodes) and z chunk width */
  int npx, npy, nzc;          /* pencils per axis, z chunks per pencil */
  int nzt, gather_nzt;        /* z tiles per pencil of the spread / of the gather: one CTA per (pencil, z tile) */
  int n_bins;                 /* npx * npy * nzc */
  int row_stride, plane_stride; /* shared-memory row stride of the spread tile in elements (informational) */
  int smem_bytes;
  int spread_threads, gather_threads;
  int spread_batch;            /* atoms a warp of the spread fetches and stages at a time (16 or 32) */
  int spread_tiled;            /* 1: tpme_tile_spread is the faster spread for this case; 0: use tpme_spread
                                  (the plan then has one bin per pencil and serves the gathers only) */
} tpme_tile_plan;
int tpme_tile_plan_make(int dtype, int nx, int ny, int nz, int nodes, int method, int64_t n_points,
                        tpme_tile_plan* plan_host);
int64_t tpme_tile_bin_count_ints(const tpme_tile_plan* plan_host);
int tpme_tile_sort(int dtype, const tpme_tile_plan* plan_host, const void* positions, int64_t n_points,
                   const double* r2u_host, int* bin_count, int* bin_start, int* key_rank,
                   void* sorted_rec, int* sorted_idx, void* stream);
int tpme_tile_spread(int dtype, const tpme_tile_plan* plan_host, const void* sorted_rec,
                     const int* sorted_idx, const int* bin_start, const void* weights, int64_t n_points,
                     int n_channels, int method, void* mesh, int accumulate, void* stream);
/* values / dvalues / grad_positions (+ grad_r2u) select the outputs as in tpme_gather / tpme_gather_vjp
 * (`coef`, `positions` only needed for grad_positions / grad_r2u) */
int tpme_tile_gather(int dtype, const tpme_tile_plan* plan_host, const void* mesh, const void* sorted_rec,
                     const int* sorted_idx, const int* bin_start, const void* positions, const void* coef,
                     int64_t n_points, int n_channels, const double* r2u_host, int method, void* values,
                     void* dvalues, void* grad_positions, int accumulate, void* grad_r2u,
                     const tpme_point_epilogue* epilogue, void* stream);

/* ---- reciprocal space ---------------------------------------------------------------
 * replaces KSpaceFilter.update + forward and P3MKSpaceFilter
 * (src/torchpme/lib/kspace_filter.py:97-120, 122-197, 293-329),
 * generate_kvectors_for_mesh (src/torchpme/lib/kvectors.py:24-102) and
 * Potential.lr_from_k_sq (potentials/coulomb.py:122-142, inversepowerlaw.py:108-141). */
typedef struct tpme_fft_plan_s* tpme_fft_plan;

int tpme_fft_plan_create(tpme_fft_plan* plan, int dtype, int nx, int ny, int nz, int batch);
int tpme_fft_plan_destroy(tpme_fft_plan plan);
/* number of hand-written kernels tpme_kfilter_apply launches for this plan (all mesh dimensions
 * powers of two in 8..512): 3 (fused (y,z)-plane passes + x pass with G), 5 (separate z / y / x
 * passes); 0 if it goes through cuFFT + a multiply kernel */
int tpme_fft_plan_uses_own_fft(tpme_fft_plan plan);
/* unnormalised forward / inverse real 3-D transforms over the last three axes */
int tpme_rfft3(tpme_fft_plan plan, const void* mesh, void* mesh_hat, void* stream);
int tpme_irfft3(tpme_fft_plan plan, void* mesh_hat /* destroyed */, void* mesh, void* stream);

/* Green's function description.  kind: 0 = table (device pointer, (nx,ny,nz/2+1) reals),
 * 1 = Coulomb, 2 = inverse power law with `exponent` in 1..6, 3 = cubic spline in k^2 (SplinePotential,
 * potentials/spline.py:108-121,151-157; lib/splines.py CubicSpline): `table` = device doubles
 * [x(n), y(n), y''(n)], `exponent` = n; 4 = the same on a reciprocal axis (CubicSplineReciprocal): `table` =
 * [1/x knots incl. 0 (n), y (n), y'' (n)] followed by the three-knot head spline [x(3), y(3), y''(3)] used below
 * the first grid point.  G = prefactor * spline(k^2).  `p3m_nodes` > 0 multiplies
 * the mode-0 P3M influence function 1/U^2 for that interpolation order.  `scale`
 * multiplies the result (FFT normalisation, prefactors). */
typedef struct tpme_green {
  int kind;
  int exponent;
  int p3m_nodes;
  int p3m_mode;        /* low byte: influence-function mode 0..3 (lib/kspace_filter.py:307-329); for modes 1..3 the
                          next byte is the order 1..6 of the differential operator (:331-347) */
  double smearing;
  double prefactor;
  double scale;
  double recip[9];     /* host: 2 pi (cell^-1)^T, row-major: recip[3*a+b] = B[a][b] */
  double spacing[3];   /* host: |cell_a| / n_a (only used when p3m_nodes > 0) */
  const void* table;   /* device, kind == 0 */
} tpme_green;

/* mesh_hat[c, k] *= scale * G(k) (in place).  If `dc_out` (device, C reals) is non-NULL it
 * receives Re mesh_hat[c, k=0] before the multiply, i.e. the sum of the real-space mesh. */
int tpme_green_multiply(int dtype, void* mesh_hat, int n_channels, int nx, int ny, int nz,
                        const tpme_green* green_host, void* dc_out, void* stream);
/* out = irfft3(G * rfft3(in)); `work_hat` is caller-provided scratch of the half-complex
 * shape; if `keep_hat` is non-NULL the un-multiplied spectrum rfft3(in) is copied there;
 * `dc_out` as above. */
int tpme_kfilter_apply(tpme_fft_plan plan, const void* mesh_in, void* mesh_out, void* work_hat,
                       void* keep_hat, const tpme_green* green_host, void* dc_out, void* stream);
/* writes the filter itself, (nx,ny,nz/2+1) reals: scale * G(k) */
int tpme_green_table(int dtype, void* table_out, int nx, int ny, int nz,
                     const tpme_green* green_host, void* stream);
/* gradient of L w.r.t. a table filter: grad_table[k] = scale * mult_k *
 * sum_c Re(x_hat[c,k] * conj(y_hat[c,k])), mult_k = 1 on the kz = 0 / Nyquist planes, 2 else */
int tpme_green_table_vjp(int dtype, const void* x_hat, const void* y_hat, int n_channels,
                         int nx, int ny, int nz, double scale, void* grad_table, void* stream);

/* ---- slab-decomposed reciprocal space (multi-GPU) -------------------------------------
 * The 3-D transform of lib/kspace_filter.py:169-187 split at the exchange points of a slab
 * decomposition (power-of-two mesh dimensions in 8..512):
 *   tpme_slab_fft_yz       (y,z) passes of `n_planes` = C * nx_local planes; forward:
 *                          real (n_planes, ny, nz) -> half-complex (n_planes, ny, nz/2+1);
 *                          otherwise the reverse (the half-complex input is destroyed)
 *   tpme_slab_fft_x_green  in place on (C, nx, ny_local, nz/2+1), the rows y0 .. y0+ny_local-1:
 *                          forward x transform, multiply by scale * G(k), inverse x transform
 *   tpme_slab_exchange_copy  the transposing block copy on either side of the all-to-all:
 *        dst[p][c * dst_c + a * dst_a + i] = src[c * src_c + p * src_p + a * src_a + i],
 *        c < n_c, p < n_p, a < n_a, i < run  (strides and run in elements of `elem_bytes`).
 *        `dst_host` is a host array of n_p device pointers: blocks of a local send buffer, or
 *        the peers' receive buffers mapped with tpme_peer_buffer_open (NVLink peer stores). */
#define TPME_MAX_RANKS 16
#define TPME_IPC_HANDLE_BYTES 64
int tpme_slab_fft_yz(int dtype, int forward, void* real_mesh, void* mesh_hat, int n_planes, int ny,
                     int nz, void* stream);
int tpme_slab_fft_x_green(int dtype, void* mesh_hat_t, int n_channels, int nx, int ny, int nz,
                          int y0, int ny_local, const tpme_green* green_host, void* stream);
int tpme_slab_exchange_copy(int elem_bytes, const void* src, void* const* dst_host, int n_c, int n_p,
                            int n_a, int64_t run, int64_t src_c, int64_t src_p, int64_t src_a,
                            int64_t dst_c, int64_t dst_a, void* stream);
/* Fused compute + exchange over peer memory.  `peers` lists, for every rank of the node, the
 * x-slab array `hat[p]` (C, nx/W, ny, nz/2+1) and the y-slab array `hat_t[p]` (C, nx, ny/W, nz/2+1)
 * as mapped into this process (tpme_peer_buffer_open).
 *   tpme_slab_fft_yz_push       forward (y,z) passes of the local planes (scratch: hat[rank]); the
 *                               y pass stores its results straight into hat_t of ALL ranks
 *   tpme_slab_fft_x_green_push  x transform . G . inverse x transform of hat_t[rank]; the results
 *                               are stored straight into hat of ALL ranks
 * i.e. the all-to-all transposes travel as NVLink stores issued by the FFT kernels themselves
 * while other thread blocks still compute; the caller separates the phases with
 * tpme_peer_barrier. */
typedef struct tpme_slab_peers {
  int n_ranks, rank;
  void* hat[TPME_MAX_RANKS];
  void* hat_t[TPME_MAX_RANKS];
} tpme_slab_peers;
int tpme_slab_fft_yz_push(int dtype, void* real_mesh, int n_channels, int nx, int ny, int nz,
                          const tpme_slab_peers* peers, void* stream);
int tpme_slab_fft_x_green_push(int dtype, int n_channels, int nx, int ny, int nz,
                               const tpme_green* green_host, const tpme_slab_peers* peers,
                               void* stream);
/* Peer-memory exchange buffers: cudaMalloc'ed, zero-filled, exported as a CUDA IPC handle that
 * the other ranks of the node open; and a device-side barrier over uint32 flags[n_ranks] that
 * live in such a buffer (`flags_host[p]` = rank p's flag array as mapped in this process,
 * `epoch` = one local uint32, `error_flag` = one local int set when a peer does not arrive
 * within `timeout_seconds`).  The barrier is CUDA-graph replayable. */
int tpme_peer_buffer_create(int64_t bytes, void** dev_ptr, unsigned char* handle_out);
int tpme_peer_buffer_open(const unsigned char* handle, void** dev_ptr);
int tpme_peer_buffer_close(void* dev_ptr);
int tpme_peer_buffer_destroy(void* dev_ptr);
int tpme_peer_barrier(void* const* flags_host, int n_ranks, int rank, void* epoch,
                      double timeout_seconds, void* error_flag, void* stream);
/* Sum all-reduce of `n` reals over peer memory (combines the partial potentials / gradients of
 * the slabs and pair-list chunks): rank r pulls slice r of every rank's input region, sums in
 * rank order and stores the result into slice r of every rank's output region, so all ranks
 * hold bitwise identical sums.  `in_host[p]` / `out_host[p]`: rank p's regions as mapped in this
 * process, 16-byte aligned and padded to a multiple of 16 bytes.  Bracket with tpme_peer_barrier. */
int tpme_peer_allreduce(int dtype, void* const* in_host, void* const* out_host, int n_ranks, int rank,
                        int64_t n, void* stream);

/* Same all-reduce through the NVSwitch (NVLS): `multicast_ptr` is the multicast address of a buffer that
 * all `n_ranks` GPUs of the node have bound (e.g. torch.distributed._symmetric_memory: multicast_ptr of a
 * rendezvoused allocation), holding every rank's partial sums at the same offset.  Rank r reduces slice
 * r with multimem.ld_reduce (the switch returns the sum over all copies) and broadcasts it in place
 * with multimem.st.  16-byte aligned, padded to a multiple of 16 bytes.  Bracket with tpme_peer_barrier. */
int tpme_multimem_allreduce(int dtype, void* multicast_ptr, int n_ranks, int rank, int64_t n, void* stream);

/* ---- real space -----------------------------------------------------------------------
 * replaces Calculator._compute_rspace (src/torchpme/calculators/calculator.py:43-87) and
 * Potential.sr_from_dist (potentials/potential.py:106-138).
 * kind: 0 = per-pair values given in `pair_values`; 1 = Coulomb; 2 = inverse power law.
 *   out[i, c] += 1/2 sum_{p: i_p = i} q[j_p, c] v(d_p)  (+ the mirrored term for half lists)
 * `out` must be zeroed by the caller.  `pair_mask` (uint8, may be NULL) zeroes pairs.
 * `exclusion_radius` <= 0 means "not set".  `n_pairs_dev` (device, may be NULL): the number of valid
 * pairs when the list sits in a buffer of capacity `n_pairs` (a neighbor list built on the device without
 * a host synchronisation, tpme_nl_fill): pairs p >= *n_pairs_dev are skipped and their grad_pairs entries
 * are left untouched. */
typedef struct tpme_pair_potential {
  int kind;
  int exponent;
  int exclusion_degree;
  int reserved;
  double smearing;
  double prefactor;
  double exclusion_radius;
} tpme_pair_potential;

int tpme_pair_forward(int dtype, const void* charges, const void* neighbor_indices,
                      int index_is_int64, const void* distances, const void* pair_values,
                      const uint8_t* pair_mask, int64_t n_pairs, const int64_t* n_pairs_dev,
                      int64_t n_atoms, int n_channels, int full_neighbor_list,
                      const tpme_pair_potential* potential_host, void* out, void* stream);
/* backward of the above for L with dL/dout = grad_out:
 *   grad_charges (N,C), accumulated, may be NULL
 *   grad_pairs   (P,): dL/d distances (kind 1,2) or dL/d pair_values (kind 0), may be NULL */
int tpme_pair_backward(int dtype, const void* charges, const void* neighbor_indices,
                       int index_is_int64, const void* distances, const void* pair_values,
                       const uint8_t* pair_mask, const void* grad_out, int64_t n_pairs,
                       const int64_t* n_pairs_dev, int64_t n_atoms, int n_channels, int full_neighbor_list,
                       const tpme_pair_potential* potential_host, void* grad_charges,
                       void* grad_pairs, void* stream);

/* ---- neighbor list on the GPU -----------------------------------
 * The reference takes the neighbor list from the external `vesin` package (tests/helpers.py:240-275,
 * examples/basic-usage.py:166-169).  Cell list: atoms are wrapped into the cell along the periodic
 * directions, binned into slabs between lattice planes (`n_bins[a]` per direction; 1 for non-periodic
 * directions) and sorted by bin; one thread per sorted atom then walks `reach[a]` bins on either side.
 * A pair (i, j, S) is the image r_j + S . cell seen from r_i; half lists keep i < j (any S) and self
 * images with S lexicographically positive; full lists hold both directions.
 *   tpme_nl_sort   wrap + bin + counting sort of the atoms into sorted_rec / sorted_shift / bin_start
 *   tpme_nl_pairs  the search: *n_pairs_dev = number of pairs (stays on the device); with `indices` != NULL
 *                  also indices (capacity,2) int32 / int64, distances (capacity) reals, shifts (capacity,3)
 *                  int32 (w.r.t. the positions as given, not the wrapped ones).  Pairs beyond `capacity`
 *                  are dropped -- compare *n_pairs_dev with the capacity.  The pairs of one atom are
 *                  contiguous; the order of the atoms' runs depends on the scheduling of the CTAs.
 * Workspace (device, caller-allocated, 16-byte aligned): scratch (tpme_nl_scratch_ints int32), bin_start
 * (prod(n_bins) + 1) int32, sorted_rec (N records of 16 bytes in fp32 / 32 bytes in fp64),
 * sorted_shift (N,4) int32.
 * The per-atom code (csrc/neighbors_core.h) also runs on the CPU in tests/test_neighbors.py. */
typedef struct tpme_neighbor_search {
  double cell[9];    /* host, row-major, rows = lattice vectors */
  int n_bins[3];
  int reach[3];      /* bins visited on each side: ceil(cutoff / slab thickness) */
  int periodic[3];
  int full_list;
  double cutoff;
} tpme_neighbor_search;
int64_t tpme_nl_scratch_ints(int64_t n_atoms, const tpme_neighbor_search* search_host);
int tpme_nl_sort(int dtype, const void* positions, int64_t n_atoms, const tpme_neighbor_search* search_host,
                 int* scratch, int* bin_start, void* sorted_rec, int* sorted_shift, void* stream);
int tpme_nl_pairs(int dtype, const void* sorted_rec, const int* sorted_shift, const int* bin_start,
                  int64_t n_atoms, const tpme_neighbor_search* search_host, int64_t capacity,
                  int index_is_int64, void* indices, void* distances, int* shifts, int64_t* n_pairs_dev,
                  void* stream);
/* d_p = |r_j + S_p . cell - r_i| for a pair list with image shifts, and its vector-Jacobian product:
 * grad_positions (N,4) -- padded to 4 reals per atom so that the three components travel as one 16-byte
 * vector reduction, 16-byte aligned -- accumulated, grad_cell (9 reals, accumulated, may be NULL).  What the reference's
 * users get from vesin's torch front end (examples/basic-usage.py:161-169) so that forces
 * reach the positions through neighbor_distances.  `n_pairs_dev` as for tpme_pair_forward. */
int tpme_pair_distances(int dtype, const void* positions, const double* cell_host,
                        const void* neighbor_indices, int index_is_int64, const int* shifts, int64_t n_pairs,
                        const int64_t* n_pairs_dev, void* distances, void* stream);
int tpme_pair_distances_backward(int dtype, const void* positions, const double* cell_host,
                                 const void* neighbor_indices, int index_is_int64, const int* shifts,
                                 const void* grad_distances, int64_t n_pairs, const int64_t* n_pairs_dev,
                                 void* grad_positions, void* grad_cell, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TORCHPME_B200_H */
