/* tsff.h -- C ABI of libtsff: B200-native (sm_100a) Thomson-scattering form-factor kernels.
 *
 * The reference (ergodicio/tsadar) is pure Python/JAX and has NO plugin / FFI interface; the only contracts on
 * its hot path are Python signatures.  Each entry point below names the reference call it replaces; the
 * jax.ffi / ctypes bindings a maintainer would add are shown in INTEGRATION.md.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless marked "host"; the caller (XLA / torch) owns all buffers.
 *   - no allocation, no host synchronisation, no exceptions inside a call; work is enqueued on `stream`
 *     (a cudaStream_t passed as void*).  Return 0 or a negative TSFF_E* code; tsff_last_error() gives the text.
 *   - a tsff_ctx is immutable after creation (device-resident static tables): re-entrant per (ctx, stream).
 *   - arrays are C-contiguous; B = lineouts in the call, W = wavelength points, A = scattering angles,
 *     G = gradient points, I = ion species, V = f-table length, P = G*W*A poles per lineout.
 *   - NaN policy: propagate (as the reference does).
 *
 * Parameter block  params[B][NP], NP = TSFF_P_ION0 + 4*I, float64, PHYSICAL units as produced by the reference's
 * ThomsonParams.__call__ (ts_params.py:583-603): Te [keV], ne [1e20 cm^-3], lam [nm], Va, ud [1e6 cm/s],
 * gradients [%], amp1..3, then per ion (A, Z, Ti [keV], fract).
 */
#ifndef TSFF_H_
#define TSFF_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TSFF_ABI_VERSION 3
#define TSFF_MAX_IONS 4
#define TSFF_MAX_LEAVES 24 /* 10 + 3 * TSFF_MAX_IONS + 1, rounded up */

enum {
  TSFF_OK = 0,
  TSFF_E_INVALID = -1, /* bad argument / unsupported shape */
  TSFF_E_CUDA = -2,    /* a CUDA runtime call or launch failed */
  TSFF_E_NOMEM = -3,
  TSFF_E_NODEVICE = -4 /* no sm_100 device: there is no CPU fallback */
};

enum { TSFF_P_TE = 0, TSFF_P_NE, TSFF_P_LAM, TSFF_P_VA, TSFF_P_UD, TSFF_P_NE_GRAD, TSFF_P_TE_GRAD,
       TSFF_P_AMP1, TSFF_P_AMP2, TSFF_P_AMP3, TSFF_P_ION0 };
enum { TSFF_ION_A = 0, TSFF_ION_Z, TSFF_ION_TI, TSFF_ION_FRACT, TSFF_ION_STRIDE };

/* electron-susceptibility mode */
enum {
  TSFF_MODE_TABLE = 0, /* FormFactor.__call__ (form_factor.py:163-298): f resampled on xi1 (1024), PV table on the
                          fixed grid xi2 (1640) via ratintn, lerped to xi_e; Im chi_e from the forward difference
                          of exp(cubic(log f)) along omega */
  TSFF_MODE_DIRECT = 1, /* calc_chi_vals semantics (form_factor.py:369-388) on a given 1-D table: pole = xi_e,
                          nodes = the f grid, gradient/lerp of f;  the synthetic-sweep workload of SURVEY.md 8(d) */
  TSFF_MODE_2V = 2     /* FormFactor.calc_in_2D (form_factor.py:449-587): fe is a 2-D table [V][V] (float64) on vx x vx,
                          rotated / projected per pole (rotate :300-324, calc_chi_vals :349-388).  tsff_ff_fwd produces
                          ff_out only (no fused modl); tsff_ff_bwd takes ff_bar and returns params_bar, fe_bar [B][V][V] */
};
enum { TSFF_F32 = 0, TSFF_F64 = 1 };
/* precision of the PV inner loop */
enum { TSFF_PV_FP32 = 0 /* MUFU.LG2 fast path */, TSFF_PV_FP64 = 1 /* validation path, ~10x slower */ };

typedef struct tsff_ctx tsff_ctx;

typedef struct tsff_static_cfg {
  int32_t abi_version;  /* TSFF_ABI_VERSION */
  int32_t mode;         /* TSFF_MODE_* */
  int32_t W, A, G, I, V;
  int32_t pv_precision; /* TSFF_PV_* */
  double lam_min, lam_max; /* FormFactor(lambda_range=...) : lamAxis = linspace(lam_min, lam_max, W)  (form_factor.py:132) */
  double lam_shift;        /* FormFactor(lam_shift=...)    (form_factor.py:196) */
  double v0, dv;           /* f-table grid v_i = v0 + i dv  (DistributionFunction1V, base.py:149-151) */
  const double* sa_deg;    /* host [A]  scattering_angles["sa"] */
  const double* weights;   /* host [A]  angular weights applied in the angle sum (generate_spectra.py:165,197) */
  const double* jmul;      /* host [W] or NULL: static per-wavelength multiplier (IAW filter, generate_spectra.py:210-216) */
  const double* zp_x;      /* host [zp_n]  Z' table abscissa (rdWT.txt / idWT.txt, form_factor.py:33-34) */
  const double* zp_re;     /* host [zp_n] */
  const double* zp_im;     /* host [zp_n] */
  int32_t zp_n;
  int32_t reserved;
  double ud_angle_deg, va_angle_deg; /* FormFactor(ud_ang=..., va_ang=...): directions of drift and flow, 2V mode (form_factor.py:501-504) */
  /* Wavelength shard (ARTS multi-GPU, the W-axis sharding of form_factor.py:431-447): this context covers the W points
   * [w_offset, w_offset + W) of lamAxis = linspace(lam_min, lam_max, W_total), computed exactly as the full axis would be.
   * W_total == 0 means the whole axis (W_total = W, w_offset = 0).  jmul is the LOCAL slice [W].  In table mode the
   * forward difference along omega (form_factor.py:258-261) of the last local point is 0, as at the end of the full
   * axis: a shard that is not the last one includes one halo point and drops its output. */
  int32_t W_total, w_offset;
} tsff_static_cfg;

/* ---- context ------------------------------------------------------------------------------------------ */
/* replaces FormFactor.__init__ (form_factor.py:120-161): builds omgs, xi1, xi2, Zpi on `device`. */
int tsff_ctx_create(int device, const tsff_static_cfg* cfg, tsff_ctx** out);
void tsff_ctx_destroy(tsff_ctx* ctx);
const char* tsff_last_error(void); /* thread-local, host pointer */
int tsff_abi_version(void);

/* Optional instrumentation: when both events (cudaEvent_t, created by the caller with timing enabled) are non-NULL,
 * tsff_ff_fwd records them on the launch stream immediately before / after its dominant kernel (the pole sweep in
 * direct mode, the per-(omega,angle) assembly in table mode) and tsff_ff_bwd around the adjoint node sweep
 * (ev_bwd_*).  Used by bench.py for the live per-kernel roofline number; pass NULLs to switch off. */
int tsff_ctx_set_profile_events(tsff_ctx* ctx, void* ev_fwd_start, void* ev_fwd_stop, void* ev_bwd_start, void* ev_bwd_stop);

/* Second-order path (TSFF_MODE_TABLE).  The reference takes the Hessian with jax.hessian (loss_function.py:110, 170-188), which
 * differentiates jnp.interp with the cell index held constant.  mode 1: tsff_ff_fwd records, per (lineout, gradient point,
 * wavelength, angle), the cells of its linear interpolations (PV table at xi_e, Z' table at every xi_i) into
 * cells [B][G][W][A][1 + TSFF_MAX_IONS] (int32, device, tsff_ff_cells_bytes); mode 2: tsff_ff_fwd and tsff_ff_bwd extend those
 * recorded cells linearly instead of looking the cell up, so that finite differences of the adjoint gradient around the
 * recording point converge to the reference's Hessian; mode 0: off.  Not re-entrant while switched on. */
size_t tsff_ff_cells_bytes(const tsff_ctx* ctx, int64_t B);
int tsff_ctx_set_frozen_cells(tsff_ctx* ctx, int mode, int32_t* cells, int64_t B);

/* bytes the caller must provide: `saved` lives from *_fwd to the matching *_bwd, `ws` is scratch per call */
size_t tsff_ff_saved_bytes(const tsff_ctx* ctx, int64_t B);
size_t tsff_ff_workspace_bytes(const tsff_ctx* ctx, int64_t B);

/* ---- B3: form factor ------------------------------------------------------------------------------------ */
/* replaces FormFactor.__call__(params) -> formfactor[G,W,A] (form_factor.py:163-298) under the reference's vmap
 * over lineouts, fused with FitModel's mean over G and weighted angle sum (generate_spectra.py:164-165,193,197):
 *   ff_out   [B][G][W][A] float64 or NULL   (= formfactor)
 *   modl_out [B][W]       float64 or NULL   (= jmul[j] * sum_a weights[a] * mean_g formfactor)
 *   fe       [B][V] float32/float64 (fe_dtype), the electron distribution table per lineout
 *   saved, ws: see *_bytes above; 256-byte aligned. */
int tsff_ff_fwd(tsff_ctx* ctx, int64_t B, const double* params, const void* fe, int fe_dtype, double* modl_out,
                double* ff_out, void* saved, void* ws, void* stream);

/* VJP of tsff_ff_fwd (replaces XLA's reverse-mode of the same graph, loss_function.py:107-108):
 *   modl_bar [B][W] or NULL, ff_bar [B][G][W][A] or NULL  (cotangents; at least one non-NULL)
 *   params_bar [B][NP] float64 (overwritten; amp1..3 and A entries are 0),  fe_bar [B][V] (fe_dtype, overwritten)
 *   TSFF_MODE_2V only: params_bar may be NULL when no kinematic parameter is trainable (the reference's arts-2d deck fits
 *   the table alone); the d/dbeta gather and the kinematics reverse are then skipped. */
int tsff_ff_bwd(tsff_ctx* ctx, int64_t B, const double* params, const void* fe, int fe_dtype, const void* saved,
                const double* modl_bar, const double* ff_bar, double* params_bar, void* fe_bar, void* ws,
                void* stream);

/* ---- B3': two spectral windows of one plasma ----------------------------------------------------------------------- */
/* The reference evaluates the electron-feature and the ion-feature spectra with two FormFactor instances on the same
 * parameters and the same f (generate_spectra.py:136-165; FitModel.__call__ :332-336); each re-derives the f-dependent tables --
 * log f, the cubic's node slopes, ratmod on xi1, its gradient and the principal-value table on xi2 (form_factor.py:256-270) --
 * that depend on neither the wavelength window nor the kinematic parameters.  These two calls build them once: ctx_a's tables
 * serve both windows, and in the adjoint both windows' table cotangents are accumulated before ONE principal-value adjoint
 * sweep.  Both contexts TSFF_MODE_TABLE on the same device with the same V / velocity grid, G, I and PV precision (W, A, the
 * wavelength range, lam_shift, weights and jmul may differ); frozen cells off.
 *   modl_a [B][W_a], modl_b [B][W_b]; saved_a / saved_b / ws_b sized by tsff_ff_saved_bytes / tsff_ff_workspace_bytes of the
 *   respective context, ws_a by the LARGER of the two contexts' tsff_ff_workspace_bytes (both windows' forward passes use it); params_bar [B][NP] and fe_bar [B][V] are the SUMS over the two windows (overwritten). */
int tsff_ff_pair_fwd(tsff_ctx* ctx_a, tsff_ctx* ctx_b, int64_t B, const double* params, const void* fe, int fe_dtype,
                     double* modl_a, double* modl_b, void* saved_a, void* saved_b, void* ws_a, void* stream);
int tsff_ff_pair_bwd(tsff_ctx* ctx_a, tsff_ctx* ctx_b, int64_t B, const double* params, const void* fe, int fe_dtype,
                     const void* saved_a, const void* saved_b, const double* modl_bar_a, const double* modl_bar_b,
                     double* params_bar, void* fe_bar, void* ws_a, void* ws_b, void* stream);

/* ---- B2: electron susceptibility pieces of the 2V path ---------------------------------------------------------- */
/* replaces FormFactor.calc_all_chi_vals(vx, DF, beta, xie_mag, klde_mag) -> (fe_vphi, chiEI, chiERrat)
 * (form_factor.py:390-447; per pole calc_chi_vals :349-388 with rotate :300-324) for a TSFF_MODE_2V context:
 *   fe [V][V] float64 (the table DF on vx x vx),  beta / xie_mag / klde_mag [P] float64 (any [G,W,A] shape, flattened)
 *   chi_out [3][P] float64: fe_vphi | chiEI | chiERrat.  Forward only: the reference differentiates this stage through
 *   calc_in_2D, whose adjoint is tsff_ff_bwd. */
int tsff_chi2v_fwd(tsff_ctx* ctx, const double* fe, const double* beta, const double* xie_mag, const double* klde_mag,
                   int64_t P, double* chi_out, void* stream);

/* ---- B1: principal-value integral -------------------------------------------------------------------- */
/* replaces vmap(ratintn.ratintn)(f, z[None,:] - pole[:,None], z) (ratintn.py:4-23; form_factor.py:266-268,385-386)
 * for uniform nodes z_i = z0 + i h, i < N:
 *   f [B][N] float64 node values, pole [B][P] float64  ->  out [B][P], dout_dpole [B][P] (NULL to skip), float64
 *   ws: tsff_pv_workspace_bytes(B, N). */
size_t tsff_pv_workspace_bytes(int64_t B, int64_t N, int64_t P);
int tsff_pv_fwd(int64_t B, int64_t N, int64_t P, const double* f, double z0, double h, const double* pole,
                double* out, double* dout_dpole, int pv_precision, void* ws, void* stream);
/* VJP: out_bar [B][P] -> f_bar [B][N], pole_bar [B][P] (NULL to skip) */
int tsff_pv_bwd(int64_t B, int64_t N, int64_t P, const double* f, double z0, double h, const double* pole,
                const double* out_bar, double* f_bar, double* pole_bar, void* ws, void* stream);

/* ---- B4: instrument response, pixel binning, amplitude scaling ------------------------------------------------ */
/* replaces irf.add_electron_IRF (kind 0, irf.py:90-132) / irf.add_ion_IRF (kind 1, irf.py:50-87) under the reference's
 * vmap over lineouts (thomson_diagnostic.py:35-36, 42-76), plus the noise add of thomson_diagnostic.py:139-140.
 * PhysParams.norm == 0 (all reference decks): amps / max scaling and the amp1 | amp2 split after the binning (irf.py:125-130);
 * norm > 0 (irf.py:117-124; un-jittable in the reference: boolean-mask indexing with a traced mask): the blue and the red side of
 * the probe wavelength are normalised to their own maxima at full resolution, then binned; the ion spectrum is binned only.
 * The Gaussian taps are truncated at cut_sigma (<=0: 12). */
typedef struct tsff_irf_cfg {
  int32_t W, nbins;       /* model samples, CCD pixels (1024 in the reference, irf.py:74,124); W % nbins == 0 */
  int32_t norm, kind;     /* PhysParams.norm (>= 0);  0 electron / 1 ion */
  double lam_min, lam_max; /* wavelength axis of the model spectrum in nm: linspace(lam_min, lam_max, W) */
  double stddev;          /* widIRF.spect_stddev_ele / spect_stddev_ion [nm] */
  double cut_sigma;
} tsff_irf_cfg;
size_t tsff_irf_workspace_bytes(const tsff_irf_cfg* cfg, int64_t B);
size_t tsff_irf_saved_bytes(const tsff_irf_cfg* cfg, int64_t B);
/* modl [B][W], params [B][NP] (uses lam, amp1, amp2, amp3), amps [B] (batch["e_amps"] / ["i_amps"]),
 * noise [B][nbins] or NULL  ->  thry [B][nbins] */
int tsff_irf_fwd(const tsff_irf_cfg* cfg, int64_t B, const double* modl, const double* params, int32_t NP,
                 const double* amps, const double* noise, double* thry, void* saved, void* ws, void* stream);
/* VJP: thry_bar [B][nbins] -> modl_bar [B][W], amp_bar [B][3] (cotangents of amp1, amp2, amp3) */
int tsff_irf_bwd(const tsff_irf_cfg* cfg, int64_t B, const double* params, int32_t NP, const double* amps,
                 const void* saved, const double* thry_bar, double* modl_bar, double* amp_bar, void* ws, void* stream);

/* ---- B4 (ARTS): angular instrument response + reduction to resolution units ------------------------------------ */
/* replaces irf.add_ATS_IRF (irf.py:5-47, norm == 0) followed by ThomsonScatteringDiagnostic.reduce_ATS_to_resunit and
 * the noise add (thomson_diagnostic.py:78-107, 139) for spectype "angular_full":
 *   modl [NA][W] (= weights @ formfactor^T, generate_spectra.py:194-195)  ->  thry [row_end-row_start][ceil(W/lam_step)]
 * taps_ang [NA], taps_lam [W]: DEVICE arrays with the reference's full-length Gaussian tap vectors (irf.py:26-33);
 * only indices [t0, t1] are visited (the caller truncates where the taps are negligible). */
typedef struct tsff_ats_cfg {
  int32_t NA, W;             /* angle rows (1024), wavelength samples (npts) */
  int32_t lam_step, ang_step; /* round(W / e_data.shape[1]), round(NA / CCDsize[0])   (thomson_diagnostic.py:93-94) */
  int32_t row_start, row_end; /* data.lineouts.start / end, in angle units          (thomson_diagnostic.py:101) */
  int32_t norm, reserved;    /* PhysParams.norm (must be 0) */
  int32_t ang_t0, ang_t1, lam_t0, lam_t1; /* tap supports, inclusive */
  double lam_min, lam_max;   /* wavelength axis linspace(lam_min, lam_max, W) in nm */
  const double* taps_ang;    /* device [NA] */
  const double* taps_lam;    /* device [W] */
} tsff_ats_cfg;
size_t tsff_ats_saved_bytes(const tsff_ats_cfg* cfg);
size_t tsff_ats_workspace_bytes(const tsff_ats_cfg* cfg);
/* params: one row [NP] (uses lam, amp1, amp2), e_amps [rows], noise [rows][units] or NULL */
int tsff_ats_fwd(const tsff_ats_cfg* cfg, const double* modl, const double* params, const double* e_amps,
                 const double* noise, double* thry, void* saved, void* ws, void* stream);
/* VJP: thry_bar -> modl_bar [NA][W], amp_bar [2] (cotangents of amp1, amp2) */
int tsff_ats_bwd(const tsff_ats_cfg* cfg, const double* params, const double* e_amps, const void* saved,
                 const double* thry_bar, double* modl_bar, double* amp_bar, void* ws, void* stream);

/* ---- a7 (ARTS): angular weighting of FitModel.electron_spectrum, spectype "angular_full" -------------------------------- */
/* replaces  modlE = jnp.matmul(weights, mean(formfactor, axis=0).T) * iaw_filter   (generate_spectra.py:193-197, 210-216):
 *   ff [G][W][A] (one parameter set, as tsff_ff_fwd's ff_out with B = 1), weights [NA][A] (DEVICE: the [1024, 241] matrix
 *   angleWghtsFredfine), jmul [W] (DEVICE, the static IAW-filter multiplier, or NULL)  ->  modl [NA][W].
 * Hand-written FP64 tiled contraction (no cuBLAS); deterministic. */
int tsff_arts_weights_fwd(const double* ff, int32_t G, int32_t W, int32_t A, const double* weights, int32_t NA,
                          const double* jmul, double* modl, void* stream);
/* VJP: modl_bar [NA][W] -> ff_bar [G][W][A] (overwritten) */
int tsff_arts_weights_bwd(const double* modl_bar, int32_t G, int32_t W, int32_t A, const double* weights, int32_t NA,
                          const double* jmul, double* ff_bar, void* stream);

/* ---- B5: loss -------------------------------------------------------------------------------------------------- */
/* replaces LossFunction.calc_ei_error + loss_functionals (loss_function.py:190-267, 386-418) and the seed of the
 * reverse pass: loss += scale * sum_{b,q} weight[q] * err(data, theory), theory_bar = d loss / d theory.
 * The fit-window masks and nanmean denominators are static, so they arrive folded into weight[n] (device) and scale.
 * method: 0 l2, 1 l1, 2 log-cosh, 3 poisson.  *loss_out (device scalar) is ACCUMULATED into: zero it first. */
int tsff_loss_fwd_bwd(int64_t B, int32_t n, const double* theory, const double* data, const double* weight,
                      double uncert, double scale, int method, double* loss_out, double* theory_bar, void* stream);

/* ---- N1 / N3: parameter transforms, DLM1V producer, optimiser update (the stages either side of the path on a fit step) -- */
/* replaces ThomsonParams.__call__ (ts_params.py:583-603; ElectronParams :202-218, IonParams :308-326, GeneralParams :459-495,
 * ion-fraction renormalisation and tied Ti :543-563) fused with DLM1V.__call__ (distribution_functions/base.py:277-294).
 * Leaf columns, NL = 10 + 3 I + 1:  Te ne | lam Va ud ne_gradient Te_gradient amp1 amp2 amp3 | per ion: Z Ti fract | m.
 * active_slot[k] >= 0: leaf k is trainable, its normalised value is x_active[b][slot] and physical = sigmoid(x) scale + shift
 * (ts_params.py:93-104, 329-350); otherwise its value is x_static[b][k] and physical = x scale + shift.
 * f table: f_vx_m [nm][V] (DEVICE, row i = the projected super-Gaussian of order m0 + i dm on the lineout's velocity grid);
 * fe = lerp in m (+ m_offset), normalised to sum fe dv = 1.  nm == 1: a fixed table (Maxwellian), no m leaf; V == 0: no fe. */
typedef struct tsff_params_cfg {
  int32_t I, V, nm, fe_dtype, NLA, reserved;
  double dv, m_offset, m0, dm;
  int32_t active_slot[TSFF_MAX_LEAVES];
  double scale[TSFF_MAX_LEAVES], shift[TSFF_MAX_LEAVES];
  double ionA[TSFF_MAX_IONS];       /* atomic masses (static in the reference, ts_params.py:290-298) */
  int32_t ti_same[TSFF_MAX_IONS];   /* 1: Ti tied to ion-1's (ts_params.py:555-557) */
  const double* f_vx_m;             /* DEVICE [nm][V] */
} tsff_params_cfg;
/* x_active [B][NLA], x_static [B][NL]  ->  params [B][NP] (physical block of tsff_ff_fwd), fe [B][V] (fe_dtype) or NULL */
int tsff_params_fwd(const tsff_params_cfg* cfg, int64_t B, const double* x_active, const double* x_static, double* params,
                    void* fe, void* stream);
/* VJP: params_bar [B][NP], fe_bar [B][V] or NULL  ->  x_active_bar [B][NLA] (overwritten) */
int tsff_params_bwd(const tsff_params_cfg* cfg, int64_t B, const double* x_active, const double* x_static,
                    const double* params_bar, const void* fe_bar, double* x_active_bar, void* stream);
/* replaces optax.adam(learning_rate).update + apply_updates (inverse/loops.py:87-89, 239-241) for all lineouts in one launch:
 * x, grad, mu, nu [B][n_active]; count [B] = steps taken so far per lineout (device; advanced by the call). */
int tsff_adam_step(int64_t B, int32_t n_active, double* x, const double* grad, double* mu, double* nu, double* count,
                   double lr, double b1, double b2, double eps, void* stream);

/* ---- N4: data-side stage in front of the fit --------------------------------------------------------------------------------- */
/* replaces the lineout extraction of get_lineouts (tsadar/utils/process/lineouts.py:85-165) for one spectrometer image:
 *   image [NY][NX] float64 (wavelength x time/space, background already subtracted), pixels [L] int32 lineout centres (DEVICE,
 *   and the same list on the HOST for the bounds check), dpixel, gain (other.gain), window [NY] uint8 fit-window mask (DEVICE,
 *   or NULL = all)  ->  data [L][NY] (column sum over [a - dpixel, a + dpixel), boxcar of 2 dpixel + 1 rows, / gain),
 *   amps [L] (max over the window; NULL to skip). */
int tsff_lineouts_fwd(const double* image, int32_t NY, int32_t NX, const int32_t* pixels, const int32_t* pixels_host, int32_t L,
                      int32_t dpixel, double gain, const unsigned char* window, double* data, double* amps, void* stream);

/* ---- microbenchmarks used for the roofline denominators --------------------------------- */
/* runs `iters` dependent-chain FFMA (kind 0) or MUFU.LG2 (kind 1) per thread on the whole device; returns the
 * number of operations issued (FFMA counts 1 op = 2 flop) in *ops; time it with events on `stream`. */
int tsff_microbench(int kind, int64_t iters, double* ops, float* sink, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TSFF_H_ */
