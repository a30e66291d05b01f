/*
 * ho_b200.h -- C ABI of the B200-native Berreman reflection engine (libho_b200.so).
 *
 * The reference (hyperbolic-optics 0.3.0) is pure Python/NumPy and has no FFI of its own;
 * the seam this library sits behind is the Python object protocol of
 *   Structure.execute(payload, backend)                    reference structure.py:370-419
 * Each entry point below names the reference call it replaces.  Host code (Python, via
 * ctypes -- see INTEGRATION.md) builds the small per-axis tables exactly like the
 * reference's layer constructors do and hands the O(A*B*F*T) work to one fused kernel.
 *
 * Conventions
 *   - complex numbers are interleaved (re, im) doubles (ho_c128) or floats;
 *   - batch point index n = ((f*A + a)*B + b)*T + t, i.e. the reference's presentation
 *     order (F, A, B, T) of axes.py:83-95, so every output is already "presented";
 *   - all pointers in ho_problem_t / ho_outputs_t are DEVICE pointers owned by the caller
 *     (PyTorch tensors used as plain buffers); the library never allocates outputs;
 *   - functions return 0 on success or a negative ho_status; ho_last_error() gives text.
 *     Numerical failure is not an error: inf/NaN flow into the outputs like the reference
 *     (tests/test_scattering.py:86-91 of the reference require exactly that).
 */
#ifndef HO_B200_H
#define HO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HO_ABI_VERSION 15
#define HO_MAX_LAYERS 24
#define HO_MAX_PEERS 15

typedef struct { double re, im; } ho_c128;

enum ho_status {
  HO_OK = 0,
  HO_ERR_BAD_ARGUMENT = -1, /* malformed descriptor (layer order, sizes, null tables) */
  HO_ERR_CUDA = -2,         /* a CUDA runtime call failed */
  HO_ERR_NO_DEVICE = -3
};

/* layer kinds below the prism, top to bottom (reference layers.py:669-683 factory keys) */
enum ho_layer_kind {
  HO_ISO_FILM = 1,   /* "Isotropic Middle-Stack Layer"      layers.py:463-540 */
  HO_ANISO_FILM = 2, /* "Crystal Layer"                     layers.py:543-579 */
  HO_ANISO_EXIT = 3, /* "Semi Infinite Anisotropic Layer"   layers.py:582-617 */
  HO_ISO_EXIT = 4    /* "Semi Infinite Isotropic Layer"     layers.py:620-666 */
};

/* backends of Structure.execute(payload, backend=...)        structure.py:395-410 */
enum ho_backend {
  HO_BACKEND_TRANSFER = 0,  /* ordered 4x4 product semantics (overflows like the reference) */
  HO_BACKEND_SCATTERING = 1 /* stable recursion, only decaying exponentials (scattering.py) */
};

typedef struct {
  int32_t kind;        /* ho_layer_kind */
  int32_t n_eps;       /* rows of eps table: 1 (constant) or F */
  int32_t n_mu;        /* rows of mu table: 1 or F */
  int32_t mu_scalar;   /* 1: mu = mu[.,0] * I (rotation invariant)  0: full symmetric tensor */
  int32_t n_beta;      /* entries of cos_beta/sin_beta: 1 or B */
  int32_t n_thickness; /* entries of thickness: 1 or T (0 for semi-infinite layers) */
  int32_t eps_scalar;  /* 1: the crystal is isotropic, eps = eps[.,0] * I at every frequency (e.g. SiC):
                          with mu_scalar the layer has closed-form modes kz^2 = eps mu - kx^2 and is
                          evaluated like the isotropic film / exit, no quartic */
  /* packed symmetric tensors (xx,xy,xz,yy,yz,zz) per frequency, already rotated by
     Ry(rotY+1e-8)*Rx(rotX) on the host (layers.py:26-66,283); [n_eps][6], [n_mu][6] */
  const ho_c128* eps;
  const ho_c128* mu;
  /* z rotation beta = azimuth[b] + rotZ + 1e-9 (layers.py:284,318-351) as cos/sin tables */
  const double* cos_beta;
  const double* sin_beta;
  const double* thickness; /* cm (layers.py:301-308) */
  ho_c128 iso_eps, iso_mu; /* HO_ISO_FILM scalars (layers.py:484-523) */
  const ho_c128* exit_cos; /* HO_ISO_EXIT: cos(theta_f)[A], complex principal root (layers.py:190-238) */
  double exit_n;           /* HO_ISO_EXIT: sqrt(eps_exit) */
  /* Uniaxial crystal (two equal principal values eps_o, any orientation; scalar mu): the ordinary
     principal value per frequency, [n_eps], else NULL.  A performance hint only: the characteristic
     quartic of such a medium factorises exactly into the ordinary and the extraordinary quadratic,
       (x^2 + kx^2 - mu eps_o) (x^2 + c3 x + q_e),
     so the start values of the spectral split are two square roots instead of the Ferrari / Cardano
     closed form (np.linalg.eig in the reference, waves.py:256-296).  The polish and its convergence
     check are the same; a failed check falls back to the general start values. */
  const ho_c128* eps_ordinary;
} ho_layer_t;

enum ho_kernel_hint { HO_HINT_NONE = 0, HO_HINT_UNTILTED = 1, HO_HINT_TILTED_THIN = 2 };

/* How ho_reflect launches the point-per-thread kernels (a launch detail: results are the same) */
enum ho_protocol {
  HO_PROTOCOL_AUTO = 0,      /* fast-path kernel + fix-up launch when hint_fast_path is set, a work list is
                                supplied and the range has >= 2^22 points; the general kernel alone otherwise */
  HO_PROTOCOL_SINGLE = 1,    /* the general kernel alone */
  HO_PROTOCOL_FAST_FIXUP = 2 /* always the two-launch protocol (needs r_pp; any size, any stack) */
};

typedef struct {
  int32_t A, B, F, T;            /* canonical batch sizes (axes.py:21) */
  int32_t n_layers;              /* layers below the prism, <= HO_MAX_LAYERS */
  int32_t backend;               /* ho_backend */
  int32_t hint_fast_path;        /* ho_kernel_hint.  HO_HINT_UNTILTED: every crystal layer is un-tilted (z is a principal
                                    axis of its rotated tensors up to the reference's 1e-8 offsets): the fast-path
                                    kernel applies to (almost) every point.  HO_HINT_TILTED_THIN: tilted crystals whose
                                    films are optically thin: the lean kernel (general start values, films as one
                                    cubic in Delta) applies.  A performance hint only -- points the chosen kernel
                                    cannot evaluate are finished by the general kernel either way */
  int32_t protocol;              /* ho_protocol */
  int32_t tsweep_groups;         /* thickness axis (T > 1): 0 = warp-cooperative kernel, groups per warp chosen by
                                    the library; 1..32 = that many (f,a,b) groups per warp; -1 = one thread per
                                    point, no reuse of the t-independent sub-stack */
  int32_t max_ctas_per_sm;       /* 0: the library's launch shape.  k > 0: the point-per-thread kernels run as ONE wave of
                                    k CTAs per SM (grid-stride), leaving the rest of every SM to kernels on other
                                    streams -- used by the fused all-gather once the step is NVLink-bound, so that the
                                    receivers' Mueller rebuild overlaps it */
  /* Work list of the fast-path protocol: caller-owned device scratch of work_capacity + 1 uint32
     (entry 0 = number of deferred points of the last launch, then their indices relative to
     n_begin).  Launches that share a list must be ordered (same stream).  NULL: the general
     kernel alone.  A list shorter than the deferred set is legal: the fix-up launch then finds the
     points by a sentinel in r_pp instead. */
  uint32_t* work;
  int64_t work_capacity;
  const double* kx;              /* [A] sqrt(eps_prism) sin(theta)       structure.py:183 */
  const double* k0;              /* [F] 2 pi f                           structure.py:189 */
  const double* prism_inv_cos;   /* [A] 1/cos(theta)                     layers.py:109-152 */
  const double* prism_inv_ncos;  /* [A] 1/(n cos(theta)) */
  double prism_inv_n;            /* 1/n */
  ho_layer_t layers[HO_MAX_LAYERS];
} ho_problem_t;

typedef struct {
  /* each [n_end - n_begin] complex128, presentation order; replaces structure.r_pp ...
     (structure.py:272-314 / scattering.py:155-164).  Any pointer may be NULL to skip. */
  ho_c128* r_pp;
  ho_c128* r_ps;
  ho_c128* r_sp;
  ho_c128* r_ss;
  /* [n][16] float64 row-major Mueller matrix M = Re(A F A^-1)   mueller.py:256-325 */
  double* mueller;
  /* [n][4] Stokes vector M * incident_stokes (mueller.py:479-500); needs mueller math only */
  double* stokes;
  double incident_stokes[4];
  /* Power bookkeeping, replaces FieldProfile.reflectance / transmittance / layer_absorption
     (fields.py:120-282) for pure p and s incidence, valid for BOTH backends (the reference's
     FieldProfile needs the transfer matrix, fields.py:87-88).  2*(n_layers+1) planes of
     power_stride doubles each, plane index pol*(n_layers+1) + k with pol 0 = p, 1 = s and
     k = 0: R = 1 - S_z(G_1)/S_inc, k = 1: T = S_z(G_exit)/S_inc, k = 2+i: absorptance of the
     i-th finite layer below the prism (i = 0 .. n_layers-2).  NULL to skip. */
  double* power;
  int64_t power_stride; /* >= n_end - n_begin */
  /* Transmission amplitude coefficients, each [n_end - n_begin] complex128; all four or none.
     Replaces Structure.calculate_transmissivity (structure.py:326-347: inverse of Gamma's forward
     2x2 block) for the transfer backend and the t picks of scattering_coefficients
     (scattering.py:155-164) for the scattering backend.  For a crystal exit they are amplitudes of
     the reference's exit eigenmodes: numpy/LAPACK normalisation (unit 2-norm, largest component
     real positive, waves.py:270) ordered by sort_poynting_indices (waves.py:474-510); when the
     two forward modes are degenerate (isotropic crystal) that basis is not defined and the unit-E
     basis of the forward plane is used. */
  ho_c128* t_pp;
  ho_c128* t_ps;
  ho_c128* t_sp;
  ho_c128* t_ss;
  /* Mode-resolved transmittance (needs the t_* outputs): 4 planes of t_mode_stride doubles, plane
     2*pol + k with pol 0 = p, 1 = s incidence and k = 0: the s-like, k = 1: the p-like forward exit
     mode; value = S_z(amplitude_k * mode_k) / S_inc.  Replaces the per-mode flux of
     FieldProfile.polarization_resolved (fields.py:286-356; its T_co / T_cross).  NULL to skip. */
  double* t_mode;
  int64_t t_mode_stride; /* >= n_end - n_begin */
  /* Number of incident states of the power block: 0 or 2 = p and s (above); 4 = additionally the
     states (a_p, a_s) = (1, 1)/sqrt2 and (1, i)/sqrt2 in planes 2*(n_layers+1) .. 4*(n_layers+1)-1.
     Fluxes are Hermitian forms in (a_p, a_s), so these four determine R, T, A_i of ANY Jones
     state (fields.py:102-118 accepts a complex (a_s, a_p) pair). */
  int32_t power_states;
  /* Fused all-gather over NVLink (multi-GPU): besides the local r_pp .. r_ss planes the
     kernel stores every reflection coefficient into n_peers further destinations -- pointers into
     the result buffers of the other GPUs of the box, mapped into this process (peer memory over
     NVSwitch: CUDA IPC / symmetric memory), each addressing element 0 of the launched range like
     r_pp does.  The transfer then overlaps the arithmetic point by point inside ONE kernel; the
     stores are visible to the owner after this stream's work has completed and the ranks have
     synchronised.  0: none.  (Mueller planes are not mirrored: they are a pure function of r_ij and
     HBM is 7x faster than an NVLink port -- the receiver rebuilds them with ho_polarimetry.) */
  int32_t n_peers;
  ho_c128* peer_r[HO_MAX_PEERS][4]; /* [peer][pp, ps, sp, ss] */
  /* The receiving half of that all-gather, fused into the same kernel: n_rebuild further ranges of
     reflection coefficients that ALREADY lie in this GPU's planes (a chunk every peer pushed during
     the previous launch) are turned into their Mueller matrices while this launch computes its own
     points -- thread i of the launch also handles element i of every range, so the HBM-bound
     rebuild (64 B read, 128 B written, ~60 flops) rides in the load/store slots of the FP64-bound
     arithmetic instead of competing with it as a second kernel.  rebuild_n[k] <= n_end - n_begin;
     rebuild_r[k][j] = element 0 of plane j of range k, rebuild_mueller[k] = its [n][16] output. */
  int32_t n_rebuild;
  int64_t rebuild_n[HO_MAX_PEERS];
  const ho_c128* rebuild_r[HO_MAX_PEERS][4];
  double* rebuild_mueller[HO_MAX_PEERS];
} ho_outputs_t;

/* Replaces Structure.execute()'s numerical body (get_layers .. calculate_reflectivity /
   calculate_scattering) for batch points [n_begin, n_end) of the presentation-ordered grid.
   Output pointers address element 0 of that range.  `cuda_stream` is a cudaStream_t (or NULL
   for the default stream).  Asynchronous: returns after the launch. */
int ho_reflect(const ho_problem_t* problem, const ho_outputs_t* out,
               int64_t n_begin, int64_t n_end, void* cuda_stream);

/* Same computation in complex64 arithmetic (opt-in; tolerance 1e-4).  Tables stay float64 /
   complex128; outputs are complex64 (interleaved floats) and float32. */
typedef struct {
  float* r_pp; float* r_ps; float* r_sp; float* r_ss; /* interleaved (re, im) */
  float* mueller;                                    /* [n][16] */
} ho_outputs_f32_t;
int ho_reflect_f32(const ho_problem_t* problem, const ho_outputs_f32_t* out,
                   int64_t n_begin, int64_t n_end, void* cuda_stream);

/* Polarimetry epilogue over Jones planes that are resident in HBM (the r_ij or t_ij planes of
   ho_outputs_t, or any other [n] complex128 planes): one pass, 64 B read per point, only the
   requested planes written.  Replaces the per-point NumPy post-processing of the reference's
   consumer classes; host code folds the ideal components (polarisers, wave plates, rotators:
   constant matrices) before / after the sample into s_pre, m_post, e_pre, j_post.
   All pointers are device pointers; NULL output = not wanted. */
typedef struct {
  const ho_c128* j_pp; const ho_c128* j_ps; const ho_c128* j_sp; const ho_c128* j_ss; /* [n] each */
  /* Mueller chain  s_out = m_post * M(J) * s_pre               mueller.py:446-500 */
  double s_pre[4];
  double m_post[16];      /* row-major */
  /* Jones chain    e_out = j_post * J * e_pre                  jones.py:153-192 */
  ho_c128 e_pre[2];
  ho_c128 j_post[4];      /* row-major 2x2 */
  double* stokes;         /* [4][n] S0..S3 of the Mueller chain                      mueller.py:479-500 */
  double* polarisation;   /* [3][n] DOP, ellipticity, azimuth of that Stokes vector  mueller.py:516-583 */
  ho_c128* jones_vector;  /* [2][n] E_p, E_s of the Jones chain                      jones.py:182-192 */
  double* jones_stokes;   /* [6][n] S0..S3, azimuth, ellipticity of that vector      jones.py:199-226 */
  double* ellipsometry;   /* [6][n] Psi, Delta, Psi_ps, Delta_ps, Psi_sp, Delta_sp (degrees)  jones.py:290-320 */
  ho_c128* eigenvalues;   /* [2][n] tr/2 +- sqrt(discriminant) (principal root)      jones.py:228-264 */
  ho_c128* eigenvectors;  /* [4][n] planes (row, col) = (0,0) (0,1) (1,0) (1,1); column k is the unit
                             eigenvector of eigenvalue k, largest component real positive (LAPACK) */
  ho_c128* discriminant;  /* [n]    ((J_pp - J_ss)/2)^2 + J_ps J_sp */
  double* overlap;        /* [n]    |<v0|v1>| */
  double* decomposition;  /* [5][n] diattenuation, polarizance, retardance (rad), depolarization,
                             transmittance of the Lu-Chipman decomposition          mueller.py:368-444 */
  double* decomposition_matrices; /* [3][n][16] diattenuator, retarder, depolarizer (normalised) */
  double* mueller;        /* [n][16] Mueller matrix of these Jones planes (e.g. the transmission
                             Mueller matrix, mueller.py:327-356) */
} ho_polarimetry_t;
int ho_polarimetry(const ho_polarimetry_t* request, int64_t n, void* cuda_stream);

/* Device-side material tables: eps(f) of a built-in dispersion model evaluated and rotated on the
   GPU straight into the table a layer descriptor points at, instead of NumPy on the host + H2D
   (reference materials.py:176-213 TO/LO product per principal axis, :614-651 monoclinic beta-Ga2O3,
   layers.py:402-417 R eps R^T with the batch-independent Ry*Rx part).  One thread per frequency. */
enum ho_material_kind { HO_MAT_TOLO = 1, HO_MAT_MONOCLINIC = 2 };
#define HO_MAX_OSC 8
typedef struct {
  int32_t kind;
  int32_t n_osc[3];            /* TOLO: oscillators of the x, y, z axes; MONOCLINIC: {n_bu, n_au, 0} */
  double eps_inf[4];           /* TOLO: x, y, z; MONOCLINIC: Bu xx, yy, xy and Au zz */
  double w_to[3][HO_MAX_OSC];  /* TOLO: per axis; MONOCLINIC: [0] = Bu modes, [1] = Au modes */
  double g_to[3][HO_MAX_OSC];
  double w_lo[3][HO_MAX_OSC];  /* TOLO only */
  double g_lo[3][HO_MAX_OSC];
  double amp[2][HO_MAX_OSC];   /* MONOCLINIC: [0] = Bu, [1] = Au amplitudes */
  double alpha[HO_MAX_OSC];    /* MONOCLINIC: Bu mode angles, radians */
  double rot[9];               /* row-major Q = Ry*Rx; the table holds Q eps Q^T (identity: unrotated) */
} ho_material_t;
/* freq: device [F] float64 (cm^-1); eps_out: device [F][6] complex128 (xx, xy, xz, yy, yz, zz) */
int ho_material_table(const ho_material_t* material, const double* freq, int64_t F, ho_c128* eps_out,
                      void* cuda_stream);

/* Opt-in materialised objects.  The reflection kernels never form
   eigenvectors or 4x4 matrices; the reference, however, exposes them per layer and its own
   consumers read them: FieldProfile (fields.py:78-97, 156-169, 358-390, 470-479) and
   scattering_coefficients (scattering.py:45-63).  These two entry points write exactly those
   arrays in the reference's CANONICAL [A, B, F, T, ...] layout (axes.py:21-22), HBM-bound at
   256 B per point and layer.

   ho_layer_modes: one Wave-built layer (HO_ISO_FILM, HO_ANISO_FILM, HO_ANISO_EXIT) ->
     matrix   layer.matrix: exp(-i Delta k0 d) of a film (waves.py:298-335), [t0, 0, t1, 0] of a
              semi-infinite crystal (waves.py:512-536); batch shape [mA][mB][mF][mT], an axis the
              layer does not depend on has size 1 like in the reference's broadcast
     modes    WaveProfile.tangential_modes()[0] (waves.py:72-96): rows Ex, Ey, Hx, Hy, columns
              t0, t1, r0, r1 = numpy.linalg.eig's unit-norm eigenvectors (largest component real
              positive, waves.py:270), split and ordered by wave_sorting (waves.py:256-296) and
              sort_poynting_indices (waves.py:474-510); batch shape [pA][pB][pF]
     kz       the same columns' eigenvalues
     fields   WaveProfile.full_modes()[0]: rows Ex, Ey, Ez, Hx, Hy, Hz (waves.py:382-408)
     poynting time-averaged Poynting vector, rows Px, Py, Pz (waves.py:410-420)
   Any output pointer may be NULL.  In a degenerate eigenspace (isotropic medium: two equal kz) the
   reference's basis is whatever LAPACK returns; here it is the (p, s) pair of SURVEY A.13. */
typedef struct {
  ho_c128* matrix;   /* [mA][mB][mF][mT][4][4] row-major */
  ho_c128* modes;    /* [pA][pB][pF][4][4] */
  ho_c128* kz;       /* [pA][pB][pF][4] */
  ho_c128* fields;   /* [pA][pB][pF][6][4] */
  double* poynting;  /* [pA][pB][pF][3][4] */
  int32_t mA, mB, mF, mT; /* each 1 or the problem's A, B, F, T */
  int32_t pA, pB, pF;     /* pA = mA, pB = mB, pF = 1 or mF */
  int32_t reserved;
} ho_layer_modes_t;
int ho_layer_modes(const ho_problem_t* problem, int32_t layer, const ho_layer_modes_t* out, void* cuda_stream);

/* structure.transfer_matrix (structure.py:259-270): Gamma = M_prism M_1 ... M_exit for canonical
   points [n_begin, n_end), n = ((a*B + b)*F + f)*T + t; gamma: device [n_end - n_begin][4][4]
   complex128 row-major.  Transfer-matrix semantics (overflows like the reference).  The exit
   columns are in the reference's mode basis (see ho_layer_modes), so FieldProfile and
   calculate_transmissivity run on it unchanged. */
int ho_transfer_matrix(const ho_problem_t* problem, ho_c128* gamma, int64_t n_begin, int64_t n_end,
                       void* cuda_stream);

/* FP64 FMA-chain microbenchmark used as the roofline denominator: runs `iters` dependent
   FMA rounds on every SM and returns achieved FLOP/s (2 flops per FMA) in *flops_per_s. */
int ho_measure_fp64_peak(int iters, double* flops_per_s, void* cuda_stream);

/* Number of kernel launches issued by this library since load (bench.py's gpu_launches). */
int64_t ho_launch_count(void);

int ho_abi_version(void);
const char* ho_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* HO_B200_H */
