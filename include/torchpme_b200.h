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
 * 1 = Coulomb, 2 = inverse power law with `exponent` in 1..6.  `p3m_nodes` > 0 multiplies
 * the mode-0 P3M influence function 1/U^2 for that interpolation order.  `scale`
 * multiplies the result (FFT normalisation, prefactors). */
typedef struct tpme_green {
  int kind;
  int exponent;
  int p3m_nodes;
  int reserved;
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

/* ---- real space -----------------------------------------------------------------------
 * replaces Calculator._compute_rspace (src/torchpme/calculators/calculator.py:43-87) and
 * Potential.sr_from_dist (potentials/potential.py:106-138).
 * kind: 0 = per-pair values given in `pair_values`; 1 = Coulomb; 2 = inverse power law.
 *   out[i, c] += 1/2 sum_{p: i_p = i} q[j_p, c] v(d_p)  (+ the mirrored term for half lists)
 * `out` must be zeroed by the caller.  `pair_mask` (uint8, may be NULL) zeroes pairs.
 * `exclusion_radius` <= 0 means "not set". */
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
                      const uint8_t* pair_mask, int64_t n_pairs, int64_t n_atoms,
                      int n_channels, int full_neighbor_list,
                      const tpme_pair_potential* potential_host, void* out, void* stream);
/* backward of the above for L with dL/dout = grad_out:
 *   grad_charges (N,C), accumulated, may be NULL
 *   grad_pairs   (P,): dL/d distances (kind 1,2) or dL/d pair_values (kind 0), may be NULL */
int tpme_pair_backward(int dtype, const void* charges, const void* neighbor_indices,
                       int index_is_int64, const void* distances, const void* pair_values,
                       const uint8_t* pair_mask, const void* grad_out, int64_t n_pairs,
                       int64_t n_atoms, int n_channels, int full_neighbor_list,
                       const tpme_pair_potential* potential_host, void* grad_charges,
                       void* grad_pairs, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TORCHPME_B200_H */
