/*
 * picaso_b200.h - C ABI of the B200-native PICASO radiative-transfer hot path.
 *
 * The reference (natashabatalha/picaso, commit 0369089) has no FFI: its seam is
 * Python name binding in picaso/justdoit.py:2,8,9 (`from .fluxes import
 * get_reflected_1d, get_thermal_1d, get_transit_1d, ...`, `from .disco import
 * compress_disco, compress_thermal`).  Each entry point below replaces one of those
 * numba functions; picaso_b200/fluxes.py and picaso_b200/disco.py bind them with
 * ctypes behind the reference's exact Python signatures.
 *
 * Conventions
 *  - every array is float64, wavelength on the fastest axis.  A "layer array" is
 *    [nlayer][nwno], a "level array" [nlevel][nwno] (nlevel = nlayer + 1); `ld` is
 *    the element distance between consecutive layers/levels (>= nwno), the same for
 *    all layer and level inputs of one call.
 *  - batched calls (nbatch > 1) take [nbatch] stacked atmospheres: layer arrays are
 *    [nbatch][nlayer][ld], level arrays [nbatch][nlevel][ld], per-wave vectors
 *    [nbatch][nwno], outputs [nbatch][...].  Geometry is shared by the batch.
 *  - `memspace` says where ALL array pointers of the call live: PB_HOST (the library
 *    stages them through the context's device buffers: H2D, kernels, D2H, and returns
 *    when the outputs are valid) or PB_DEVICE (pointers from pb_dev_alloc; kernels are
 *    enqueued on the context stream and the call returns without synchronising).
 *  - small geometry vectors (ubar0, ubar1, gweight, tweight, tlevel, plevel, z, ...)
 *    are ALWAYS host pointers.
 *  - every function returns PB_OK (0) or an error code; pb_last_error() has the text.
 *    Numerical NaN/inf propagate exactly as in the reference (no extra clamping).
 *  - there is no CPU fallback: without a CUDA device pb_create fails.
 */
#ifndef PICASO_B200_H
#define PICASO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pb_ctx pb_ctx;

enum pb_status {
    PB_OK = 0,
    PB_ERR_CUDA = 1,    /* CUDA runtime failure (text in pb_last_error) */
    PB_ERR_ARG = 2,     /* invalid argument */
    PB_ERR_NOMEM = 3,   /* device / pinned allocation failed */
    PB_ERR_UNSUPPORTED = 4
};

enum pb_memspace { PB_HOST = 0, PB_DEVICE = 1 };

/* ---- context, memory, timing ------------------------------------------------------ */
int pb_version(void);
int pb_device_count(int *count);
int pb_create(int device, pb_ctx **out);
void pb_destroy(pb_ctx *ctx);
const char *pb_last_error(const pb_ctx *ctx); /* ctx may be NULL: last pb_create error */
int pb_device_name(pb_ctx *ctx, char *buf, size_t buflen);
int pb_sm_count(pb_ctx *ctx, int *count);

int pb_dev_alloc(pb_ctx *ctx, size_t bytes, void **out);
int pb_dev_free(pb_ctx *ctx, void *ptr);
int pb_host_alloc(pb_ctx *ctx, size_t bytes, void **out); /* page-locked host memory */
int pb_host_free(pb_ctx *ctx, void *ptr);
int pb_memcpy_h2d(pb_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes); /* async */
int pb_memcpy_d2h(pb_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes); /* async */
int pb_memset(pb_ctx *ctx, void *dst_dev, int value, size_t bytes);                /* async */
int pb_sync(pb_ctx *ctx);
/* run every later kernel/copy of this context on a caller-owned cudaStream_t (e.g. the
 * stream a collective library uses) so that ordering needs no host synchronisation;
 * NULL restores the context's own stream */
int pb_set_stream(pb_ctx *ctx, void *cuda_stream);

/* CUDA-event stopwatch on the context stream (the stream every kernel is launched on) */
int pb_timer_start(pb_ctx *ctx);
int pb_timer_stop(pb_ctx *ctx, float *elapsed_ms); /* records, synchronises, returns ms */
/* kernels launched by this context since creation */
uint64_t pb_launch_count(const pb_ctx *ctx);

/* ---- reflected light, Toon89 two-stream ------------------------------------------- */
/* replaces get_reflected_1d, picaso/fluxes.py:1010-1413 (incl. setup_tri_diag :89-183,
 * tri_diag_solve :289-323) and, when `albedo` is given, compress_disco, disco.py:118-149 */
/* All-gather of the per-rank result slab fused into the producing kernel (the
 * wavelength grid shards over GPUs, one process per GPU; the only exchange is the final [nwno] vector).
 * Every rank owns a gathered buffer [nranks][nwno] and an arrival-flag array [nranks], both cudaMalloc'ed
 * and mapped into the peers with pb_ipc_export / pb_ipc_open.  The kernel's epilogue stores its albedo
 * slab into row `rank` of EVERY rank's buffer (P2P stores), and the last CTA to finish publishes `step`
 * in flags[r][rank] on every rank r (fence.sys + release store).  A consumer waits with pb_gather_wait
 * (stream-ordered) until all of its flags reached `step`.  `wait_step` makes the kernel itself hold its
 * peer stores until all local flags reached that value: with 3 rotating buffers and wait_step = step - 2
 * a rank never overwrites a row a peer may still be reading, and ranks may run one step apart.
 * Reader contract: the slab of step t is complete on this rank once pb_gather_wait(t) has passed.  A peer overwrites
 * the slot of step t (with step t + nbuf) only after it has seen every rank publish wait_step = t + nbuf - (nbuf - 1)
 * = t + 1, so a rank must enqueue its reads of step t (stream-ordered after that wait) BEFORE the launch that publishes
 * its step t + 1: the launch of step t + 1 itself in the fused / push modes, the launch of step t + 2 in the lazy /
 * deferred modes (there the flags of a step are published by the following launch). */
typedef struct pb_peer_gather {
    int nranks, rank;                       /* nranks <= 8 */
    double *const *albedo;                  /* host array [nranks]: rank r's gathered buffer, as mapped here */
    unsigned long long *const *flags;       /* host array [nranks]: rank r's flag array, as mapped here */
    unsigned long long step, wait_step;
    unsigned int *done_counter;             /* device, this rank, 8 words, zero-initialised: CTA counter | timeout flag |
                                             * (pb_gather_wait's timeout flag) | - | go word of the current launch |
                                             * duration of the last push = 3 courier in ns (diagnostic) | courier counter */
    /* push = 0: the stores and the flag publication happen in the solver kernel's epilogue (one launch; costs
     *   4-10 us at the kernel tail: system-scope fence + NVLink round trip before the grid can retire).
     * push = 1: the solver kernel writes its slab into row `rank` of the LOCAL gathered buffer only; a small
     *   copy kernel on the context's side stream (event-ordered after it) pushes that row to the peers and
     *   publishes the flags, overlapping the next launch.  `slot` (0..7, the rotating-buffer index) names the
     *   event that keeps a later launch from overwriting a row whose push is still in flight.
     * push = 2 ("lazy"): as push = 0, but the epilogue only stores; the flags of step s are published by the first
     *   CTA of the launch of step s + 1 (a finished grid's stores are performed system-wide), the last step's by
     *   pb_peer_signal.  No fence, counter or NVLink round trip on any launch's critical path.
     * push = 3 ("deferred"): the solver CTAs write the slab into row `rank` of the LOCAL gathered buffer only.  The
     *   launch of step s + 1 carries four extra CTAs (the tile width of a single-wave launch is chosen so that the
     *   3 x 148 resident slots leave room for them) that push the slab of step s (buffers `albedo_prev`) to the
     *   peers, fence, and - the last one to finish - publish the flags while the solver CTAs compute: no peer store, fence or flag poll sits in any solver CTA, and the
     *   grid does not wait for an NVLink round trip at its tail.  pb_peer_flush delivers the last step.
     *   refl_toa_kernel5 only. */
    int push, slot;
    double *const *albedo_prev;             /* push = 3: host array [nranks], the buffers of step - 1 (NULL for step 1) */
} pb_peer_gather;

typedef struct pb_reflected_args {
    int nlayer, nwno, numg, numt, nbatch;
    int64_t ld;
    /* layer arrays */
    const double *dtau, *w0, *cosb, *gcos2, *ftau_cld, *ftau_ray, *dtau_og, *w0_og, *cosb_og;
    /* level arrays */
    const double *tau, *tau_og;
    /* per-wave vectors [nwno]; NULL = 0 (surf_reflect, b_top) or 1 (F0PI) */
    const double *surf_reflect, *F0PI, *b_top;
    /* geometry, host: ubar0/ubar1 [numg*numt], gweight [numg], tweight [numt] */
    const double *ubar0, *ubar1, *gweight, *tweight;
    double cos_theta;
    int single_phase;      /* 0 cahoy, 1 OTHG, 2 TTHG, 3 TTHG_ray  (justdoit.py:5512-5658) */
    int multi_phase;       /* 0 N=2, 1 N=1 */
    int toon_coefficients; /* 0 quadrature, 1 eddington */
    double frac_a, frac_b, frac_c, constant_back, constant_forward;
    int get_toa_intensity, get_lvl_flux;
    /* outputs (NULL = not wanted) */
    double *xint_at_top;                                        /* [numg*numt][nwno] */
    double *albedo;                                             /* [nwno], fused compress_disco */
    double *flux_minus, *flux_plus, *flux_minus_mdpt, *flux_plus_mdpt; /* [numg*numt][nlevel][nwno] */
    /* 0: get_reflected_1d.  1: the per-facet formulas of get_reflected_3d (fluxes.py:355-660): the
     * nbatch entries are the ng*nt facets (each with its own opacities), numg = numt = 1, ubar0/ubar1
     * hold one value per facet, |ubar| is used, exponents clip at 40, quadrature coefficients, 3-D
     * 'cahoy' phase function; TOA intensity only. */
    int variant;
    /* NULL, or: also deliver the fused albedo slab of this rank to every rank's gathered buffer over
     * peer memory (NVLink) from inside the kernel - see pb_peer_gather below.  Needs `albedo`, nbatch 1,
     * numg*numt <= 8, PB_DEVICE. */
    const struct pb_peer_gather *gather;
} pb_reflected_args;

int pb_reflected_toon_1d(pb_ctx *ctx, const pb_reflected_args *args, int memspace);

/* ---- reflected light, spherical harmonics (P1 "SH2" / P3 "SH4") ---------------------- */
/* replaces get_reflected_SH, picaso/fluxes.py:2675-2976 (setup_2_stream_fluxes :3189,
 * setup_4_stream_fluxes :3336, solve_4_stream_banded :3610 = scipy/LAPACK dgbsv, legP :3639)
 * and, when `albedo` is given, compress_disco.  flx = 1 (calculate_fluxes, off by default) also returns the
 * layer fluxes F.X + G in `flux`.  The reference scales its f_deltaM argument in place, once
 * per angle, when a TTHG form is active; results reproduce that, the
 * input array is NOT modified, and the final scaled array is written to `f_deltaM_out` if
 * that is non-NULL. */
typedef struct pb_sh_args {
    int nlayer, nwno, numg, numt, nbatch;
    int64_t ld;
    const double *dtau, *w0, *ftau_cld, *ftau_ray, *f_deltaM, *dtau_og, *w0_og, *cosb_og; /* layer */
    const double *tau, *tau_og;                                                          /* level */
    const double *surf_reflect, *F0PI, *b_top;                                           /* [nwno] or NULL */
    const double *ubar0, *ubar1, *gweight, *tweight;                                     /* host */
    double cos_theta;
    int w_single_form, w_multi_form, psingle_form;             /* 0 TTHG, 1 OTHG */
    int w_single_rayleigh, w_multi_rayleigh, psingle_rayleigh; /* 0 off, 1 on */
    double frac_a, frac_b, frac_c, constant_back, constant_forward;
    int stream;      /* 2 or 4 */
    int flx;         /* 0 | 1: also return the layer fluxes (calculate_fluxes, justdoit.py:4638; fluxes.py:2889-2890) */
    int single_form; /* 0 explicit, 1 legendre */
    double *xint_at_top;  /* [numg*numt][nwno] or NULL */
    double *albedo;       /* [nwno] or NULL */
    double *f_deltaM_out; /* [nlayer][nwno] or NULL */
    double *flux;         /* flx = 1: [nbatch][numg*numt][stream*nlevel][nwno] = F.X + G; follows memspace */
} pb_sh_args;

int pb_reflected_sh(pb_ctx *ctx, const pb_sh_args *args, int memspace);

/* replaces get_thermal_SH, fluxes.py:2979-3186 (flx = 0) and, optionally, compress_thermal.
 * The reference accepts but never reads tau, dtau_og, tau_og, w0_og and w0_no_raman; they are
 * not part of this struct.  `cosb` is only compared with `cosb_og` (np.array_equal decides
 * whether the delta-M fraction is 0, fluxes.py:3044-3047). */
typedef struct pb_thermal_sh_args {
    int nlayer, nwno, numg, numt, nbatch;
    int64_t ld;
    const double *dtau, *w0, *cosb, *cosb_og;  /* layer arrays */
    const double *wno;                          /* [nwno] */
    const double *surf_reflect;                 /* [nwno] or NULL */
    const double *tlevel, *plevel;              /* host, [nbatch][nlevel] */
    const double *ubar1, *gweight, *tweight;    /* host */
    int stream, hard_surface, flx;
    double *xint_at_top;                        /* [numg*numt][nwno] or NULL */
    double *thermal;                            /* [nwno] fused compress_thermal, or NULL */
} pb_thermal_sh_args;

int pb_thermal_sh(pb_ctx *ctx, const pb_thermal_sh_args *args, int memspace);

/* ---- thermal emission, Toon89 two-stream + source function ------------------------- */
/* replaces get_thermal_1d, fluxes.py:1683-1912 (blackbody :1661, blackbody_integrated
 * :1609) and, when `thermal` is given, compress_thermal, disco.py:152-180 */
typedef struct pb_thermal_args {
    int nlayer, nwno, numg, numt, nbatch;
    int64_t ld;
    const double *dtau, *w0, *cosb;        /* layer arrays */
    const double *wno, *dwno;              /* [nwno] (shared by the batch); dwno may be NULL if calc_type==0 */
    const double *surf_reflect;            /* [nwno] or NULL */
    const double *tlevel, *plevel;         /* host, [nbatch][nlevel] */
    const double *ubar1, *gweight, *tweight; /* host */
    int hard_surface, calc_type;
    double *flux_at_top;                   /* [numg*numt][nwno] or NULL */
    double *thermal;                       /* [nwno] fused compress_thermal, or NULL */
    double *flux_minus, *flux_plus, *flux_minus_mdpt, *flux_plus_mdpt; /* [numg*numt][nlevel][nwno] or NULL */
    /* 0: get_thermal_1d.  1: get_thermal_3d (fluxes.py:2148-2352): nbatch = ng*nt facets with their own
     * tlevel/plevel/opacities, numg = numt = 1, ubar1 one value per facet, pi-based boundary terms,
     * calc_type 0, TOA only. */
    int variant;
    /* > 0: the nbatch entries share `opacity_period` blocks of dtau / w0 / cosb / surf_reflect - entry b reads block
     * b % opacity_period - while tlevel / plevel / outputs stay per entry (Jacobian batching of the climate solver:
     * many temperature profiles over one set of opacities) */
    int opacity_period;
} pb_thermal_args;

int pb_thermal_toon_1d(pb_ctx *ctx, const pb_thermal_args *args, int memspace);

/* ---- transmission ------------------------------------------------------------------ */
/* replaces get_transit_1d, fluxes.py:2582-2663 */
typedef struct pb_transit_args {
    int nlevel, nwno, nbatch;
    int64_t ld;
    const double *DTAU;                          /* layer array */
    const double *z, *dz, *player, *tlayer;      /* host [nbatch][nlevel] */
    const double *mmw, *colden;                  /* host [nbatch][nlevel-1] */
    double rstar, k_b, amu;
    double *F;                                   /* [nbatch][nwno] */
} pb_transit_args;

int pb_transit_1d(pb_ctx *ctx, const pb_transit_args *args, int memspace);

/* ---- disk integration ---------------------------------------------------------------- */
/* replaces compress_disco, disco.py:118-149: xint [ng*nt][nwno] -> albedo [nwno] */
int pb_compress_disco(pb_ctx *ctx, int nwno, double cos_theta, const double *xint_at_top,
                      const double *gweight, int ng, const double *tweight, int nt,
                      const double *F0PI, double *albedo, int memspace);
/* replaces compress_thermal, disco.py:152-180: flux [ng*nt][n] -> out [n] */
int pb_compress_thermal(pb_ctx *ctx, int64_t n, const double *flux_at_top, const double *gweight,
                        int ng, const double *tweight, int nt, double *out, int memspace);

/* ---- opacity state + per-layer optical properties -------------------------------------- */
/* Device-resident replacement of the opacity data the reference re-reads from sqlite on every
 * call (RetrieveOpacities, picaso/optics.py:1877-2368).  Tables are uploaded once; all are
 * [rows][nwno] float64 on the wavenumber grid of the connection.
 *   molecular  : rows = (T,P) grid points in ptid order (T-major), raw cross-sections as stored
 *                in the DB.  store bit 1 keeps the raw rows (nearest-neighbour queries, the
 *                reference default), bit 2 keeps log10(k != 0 ? k : 1e-50) rows (bilinear
 *                queries, optics.py:2281-2293); 3 keeps both.
 *   continuum  : rows = unique CIA temperatures, ascending (optics.py:2298)
 *   rayleigh   : one row of cross-sections per scatterer (optics.py:2041-2046)
 *   raman      : Oklopcic+2016 transitions (c, ji, deltanu) + stellar_shifts[nwno][ntrans]
 *                (optics.py:467-494, :2370-2402) */
typedef struct pb_optab pb_optab;
int pb_optab_create(pb_ctx *ctx, int nwno, int nmol, int ncont, int nray, pb_optab **out);
int pb_optab_destroy(pb_ctx *ctx, pb_optab *tab);
int pb_optab_set_molecular(pb_ctx *ctx, pb_optab *tab, int imol, const double *table, int npt, int store);
int pb_optab_set_continuum(pb_ctx *ctx, pb_optab *tab, int icont, const double *table, int ntemp);
int pb_optab_set_rayleigh(pb_ctx *ctx, pb_optab *tab, int iray, const double *sigma);
int pb_optab_set_raman(pb_ctx *ctx, pb_optab *tab, const double *wno, int ntrans, const double *c,
                       const int *ji, const double *deltanu, const double *stellar_shifts);
/* pre-mixed correlated-k table of RetrieveCKs (optics.py:737-753): ln(kappa) [npress][ntemp][nwno][ngauss] */
int pb_optab_set_ck(pb_ctx *ctx, pb_optab *tab, const double *lnkappa, int npress, int ntemp, int ngauss);
int pb_optab_bytes(const pb_optab *tab, size_t *bytes);

/* replaces get_opacities / get_opacities_nearest (optics.py:2241-2368) + compute_opacity
 * (optics.py:147-431, ngauss = 1, test_mode = None) + compute_raman (optics.py:435-494) in one
 * kernel.  All per-layer vectors are HOST pointers (O(nlayer) data); only the optional cloud
 * arrays, raman_pollack and the 13 outputs follow `memspace`.  Layer multipliers are formed by
 * the caller exactly as the reference parenthesises them, e.g. mol_scale = colden*x_mol/mmw,
 * cont_scale = COEF1*x_a*x_b (or the H-, H2- expressions of optics.py:175-213). */
typedef struct pb_opacity_args {
    int nlayer;
    int query;                 /* 0 nearest (pt_index[l][0]), 1 bilinear */
    const int *pt_index;       /* [nlayer][4] 0-based table rows in the reference's term order:
                                  (t_low,p_low), (t_hi,p_low), (t_hi,p_hi), (t_low,p_hi) */
    const double *weights;     /* [nlayer][4] (1-t)(1-p), t(1-p), t p, (1-t) p   (query = 1) */
    const double *mol_scale;   /* [nmol][nlayer] */
    const int *cont_index;     /* [nlayer] row of the nearest CIA temperature */
    const double *cont_scale;  /* [ncont][nlayer] */
    const double *ray_scale;   /* [nray][nlayer] */
    int raman;                 /* 0 oklopcic, 1 pollack, 2 none  (justdoit.py:5512-5658) */
    const double *jfrac;       /* [10][nlayer] j_fraction(J, T_layer), optics.py:570 (raman = 0) */
    const double *raman_pollack; /* [nwno] (raman = 1), follows memspace */
    const double *cloud_opd, *cloud_w0, *cloud_g0; /* [nlayer][cloud_ld] or NULL, follow memspace */
    int64_t cloud_ld;
    double fthin_cld;
    int do_holes;
    int stream;                /* 2 or 4: exponent of the delta-Eddington f = COSB**stream */
    int delta_eddington;
    /* outputs in the reference's return order (optics.py:423-431); NULL = not wanted.
     * TAU, TAU_OG are [nlayer+1][nwno], the others [nlayer][nwno] */
    double *DTAU, *TAU, *W0, *COSB, *ftau_cld, *ftau_ray, *GCOS2, *DTAU_OG, *TAU_OG, *W0_OG, *COSB_OG,
        *W0_no_raman, *f_deltaM;
    /* correlated-k (ngauss > 1; RetrieveCKs.get_pre_mix_ck optics.py:1081-1161, compute_opacity :257-262):
     * outputs become [nlayer|nlevel][nwno][ngauss] (gauss point fastest, the reference's layout) */
    int ngauss;                /* 0/1 monochromatic, else must equal the ngauss of pb_optab_set_ck */
    const int *ck_index;       /* [nlayer][4] flat rows p*ntemp + t in the reference's term order:
                                  (p_low,t_low), (p_low,t_hi), (p_hi,t_hi), (p_hi,t_low) */
    const double *ck_weights;  /* [nlayer][4] (1-t)(1-p), t(1-p), t p, (1-t) p */
    const double *ck_scale;    /* [nlayer] colden/mmw */
    /* continuum lookup: 0 nearest temperature (RetrieveOpacities), 1 log-linear in 1/T between rows
     * cont_index and cont_index_hi with weight cont_t (RetrieveCKs.get_continuum, optics.py:1471-1497) */
    int cont_mode;
    const int *cont_index_hi;  /* [nlayer] */
    const double *cont_t;      /* [nlayer] */
    /* resort-rebin path: molecular_opa [nlayer][nwno][ngauss] written by pb_ck_mix with PB_DEVICE
     * (ALWAYS a device pointer), used instead of the pre-mixed table; needs ck_scale */
    const double *ck_direct;
    /* full_output (optics.py:322-325, atmosphere.taugas / tauray / taucld): the three per-layer optical depths the
     * totals are built from, [nlayer][nwno(*ngauss)]; NULL = not wanted */
    double *TAUGAS, *TAURAY, *TAUCLD;
    /* test modes of compute_opacity (optics.py:372-399, the Dlugach & Yanovitskij checks): 0 = off (test_mode None),
     * 1 = 'rayleigh' (DTAU = TAURAY, GCOS2 = 0.5, ftau_ray = 1, ftau_cld = 0), 2 = any other string (DTAU = cloud opd,
     * GCOS2 = 0, ftau_ray = 0, ftau_cld = 1); in both W0 = W0_no_raman = cloud w0 (values <= 0 -> 1e-10), COSB = cloud g0,
     * DTAU <= 0 -> 1e-10.  Needs the cloud arrays. */
    int test_mode;
} pb_opacity_args;

/* Host-side planning of a bilinear opacity query, no device involved: replaces the per-layer loop of
 * RetrieveOpacities.find_needed_pts + the index / weight bookkeeping of get_opacities (optics.py:2048-2123, :2265,
 * :2269-2298).  t_inv = 1/T_layer and p_log = log10(P_layer [bar]) are formed by the caller; the grid is described as
 * get_available_data does (optics.py:2019-2025): 1/T of the unique temperatures, log10 of the unique pressures, nc_p
 * pressures per temperature, row_offset = [0, cumsum(nc_p)].  t_mono / p_mono: the 1/T grid is strictly descending /
 * the pressure grid strictly ascending (binary search; otherwise the reference's "last True" scan).  Outputs are the
 * pt_index / weights / cont_index arrays of pb_opacity_args and the sorted unique 1-based rows (layer['pt_opa_index']). */
int pb_host_plan_bilinear(int nlayer, const double *t_inv, const double *p_log, const double *tlayer,
                          int nT, const double *t_inv_grid, int nPg, const double *p_log_grid,
                          const int64_t *nc_p, const int64_t *row_offset, int t_mono, int p_mono,
                          int ncia, const double *cia_unique, int32_t *idx, double *wts, int32_t *cia,
                          int64_t *rows_used, int *nrows_used);

int pb_compute_opacity(pb_ctx *ctx, pb_optab *tab, const pb_opacity_args *args, int memspace);

/* One call per spectrum: what picaso() does between `get_opacities(atm)` and `returns['albedo']` for a reflected-light
 * Toon run (justdoit.py:243-310, :530) - compute_opacity, get_reflected_1d, compress_disco - with nothing of size
 * O(nlayer x nwno) crossing the C boundary: the opacity kernel writes its 11 arrays into a context-owned workspace
 * in HBM, the flux kernel (fused disk integration) reads them there, and ONE device-to-host copy returns the albedo
 * (and the TOA intensities if wanted).  `opacity` is filled as for pb_compute_opacity with PB_DEVICE (its output
 * pointers are ignored; ngauss must be 0 or 1); the remaining fields are those of pb_reflected_args. */
typedef struct pb_spectrum_args {
    pb_opacity_args opacity;
    int nwno, numg, numt;
    const double *ubar0, *ubar1, *gweight, *tweight;  /* host: [numg*numt], [numg*numt], [numg], [numt] */
    double cos_theta;
    const double *surf_reflect, *F0PI, *b_top;        /* DEVICE [nwno] vectors or NULL (= 0, 1, 0) */
    int single_phase, multi_phase, toon_coefficients;
    double frac_a, frac_b, frac_c, constant_back, constant_forward;
    double *albedo;       /* host [nwno] */
    double *xint_at_top;  /* host [numg*numt][nwno] or NULL */
} pb_spectrum_args;

int pb_spectrum_reflected(pb_ctx *ctx, pb_optab *tab, const pb_spectrum_args *args);

/* The same for thermal emission (justdoit.py:243, :337-342, :567: compute_opacity -> get_thermal_1d on DTAU_OG,
 * W0_no_raman, COSB_OG -> compress_thermal) and for transmission (:243, :388-396: compute_opacity -> get_transit_1d
 * on DTAU_OG).  `opacity` as for pb_spectrum_reflected; the remaining fields are those of pb_thermal_args /
 * pb_transit_args with nbatch = 1. */
typedef struct pb_spectrum_thermal_args {
    pb_opacity_args opacity;
    int nwno, numg, numt;
    const double *tlevel, *plevel;            /* host [nlevel] */
    const double *ubar1, *gweight, *tweight;  /* host */
    const double *wno;                        /* DEVICE [nwno] */
    const double *surf_reflect;               /* DEVICE [nwno] or NULL (= 0) */
    int hard_surface;
    double *thermal;      /* host [nwno] */
    double *flux_at_top;  /* host [numg*numt][nwno] or NULL */
} pb_spectrum_thermal_args;

int pb_spectrum_thermal(pb_ctx *ctx, pb_optab *tab, const pb_spectrum_thermal_args *args);

typedef struct pb_spectrum_transit_args {
    pb_opacity_args opacity;
    int nwno;
    const double *z, *dz, *player, *tlayer;   /* host [nlevel] (level pressures / temperatures, justdoit.py:394-395) */
    const double *mmw, *colden;               /* host [nlayer] */
    double rstar, k_b, amu;
    double *F;            /* host [nwno]: (rp/rs)^2 */
} pb_spectrum_transit_args;

int pb_spectrum_transit(pb_ctx *ctx, pb_optab *tab, const pb_spectrum_transit_args *args);

/* ---- resort-rebin mixing of per-gas correlated-k tables ----------------------------------
 * Replaces deq_chem.mix_all_gases_gasesfly / do_mixing_mono_gasesfly / mix_2_gases
 * (deq_chem.py:334-386, :388-432, :538-597) and the interpolation + exp * N_A of
 * optics.RetrieveCKs.mix_my_opacities_gasesfly (optics.py:1164-1197).  The per-gas tables stay
 * resident in HBM (pb_dev_alloc + pb_memcpy_h2d); the small per-layer vectors are host memory;
 * the outputs follow `memspace`. */
typedef struct pb_ck_mix_args {
    int nlayer, nwno, ngauss, ngas; /* ngauss <= 8, ngas <= 32 */
    int np, nt;                     /* table grid: kappas[g] is ln(kappa) [np][nt][nwno][ngauss] */
    const double *const *kappas;    /* HOST array of ngas DEVICE pointers, in gases_fly order */
    const double *mixes;            /* HOST [ngas][nlayer] volume mixing ratios */
    const int *indices;             /* HOST [4][nlayer]: p_low, p_hi, t_low, t_hi (get_mixing_indices, optics.py:1199-1277) */
    const double *t_interp, *p_interp; /* HOST [nlayer] */
    const double *gauss_pts, *gauss_wts; /* HOST [ngauss] */
    double *molecular_opa;          /* out [nlayer][nwno][ngauss] = exp(bilinear ln kappa_mixed) * N_A */
    double *ln_mixed;               /* out, optional [nlayer][nwno][ngauss][4] = ln kappa_mixed (deq_chem.py:386) */
} pb_ck_mix_args;

int pb_ck_mix(pb_ctx *ctx, const pb_ck_mix_args *args, int memspace);

/* ---- climate solver: level fluxes of all correlated-k gauss points in one call ------- */
/* replaces get_fluxes, picaso/climate.py:1686-1952 (the cloudy or the clear column of it; the
 * do_holes mix (1-fhole)*cloudy + fhole*clear is linear in the outputs and done by the caller).
 * Opacity arrays are [nlayer | nlevel][nwno][ngauss], gauss point fastest - the reference's layout
 * (optics.py:423-431) - in `memspace`; every other pointer (gauss_wts, wno, dwno, surf_reflect, F0PI,
 * tlevel, plevel, geometry) and ALL outputs are host pointers in both memory spaces: the solver
 * consumes the fluxes on the host.  reflected: one mu0 = mu1 = 0.5 stream, quadrature coefficients,
 * b_top = 0 (climate.py:1797-1816).  thermal: DTAU_OG / W0_no_raman / COSB_OG, hard_surface = 0,
 * calc_type = 1 over the numg x numt disk angles, then compress_thermal (climate.py:1887-1932). */
typedef struct pb_climate_args {
    int nlayer, nwno, ngauss, numg, numt;
    int reflected, thermal;
    const double *DTAU, *TAU, *W0, *COSB, *ftau_cld, *ftau_ray, *GCOS2, *W0_no_raman; /* OpacityWEd */
    const double *DTAU_OG, *TAU_OG, *W0_OG, *COSB_OG;                                  /* OpacityNoEd */
    const double *gauss_wts;            /* [ngauss] */
    const double *wno, *dwno;           /* [nwno]; dwno = Opagrid.delta_wno */
    const double *surf_reflect, *F0PI;  /* [nwno] or NULL (0 / 1) */
    const double *tlevel, *plevel;      /* [nlevel] */
    const double *ubar1, *gweight, *tweight; /* [numg*numt], [numg], [numt] (thermal disk angles) */
    double cos_theta;
    int single_phase, multi_phase;
    double frac_a, frac_b, frac_c, constant_back, constant_forward;
    /* outputs; the reference broadcasts the visible ones over (ng, nt): here one copy */
    double *flux_net_v_layer, *flux_net_v;   /* [nlevel] */
    double *flux_plus_v, *flux_minus_v;      /* [nlevel][nwno], or NULL: not copied back */
    double *flux_net_ir_layer, *flux_net_ir; /* [nlevel] */
    double *flux_plus_ir, *flux_minus_ir;    /* [nlevel][nwno], or NULL: not copied back */
    /* Alternative to the eight pointers above (which are then ignored): ONE host block, one device-to-host copy.
     * Layout: [4][nlevel] = flux_net_v_layer, flux_net_v, flux_net_ir_layer, flux_net_ir; if packed_full != 0
     * followed by [4][nlevel][nwno] = flux_plus_v, flux_minus_v, flux_plus_ir, flux_minus_ir.  The half that
     * is not computed (reflected = 0 or thermal = 0) reads as zeros. */
    double *packed;
    int packed_full;
    /* Jacobian batching (climate.py:1108-1180: t_start perturbs one level temperature per get_fluxes call, thermal
     * only - the opacities stay fixed, only the Planck terms change).  nprofiles > 0: the thermal half is run for
     * `nprofiles` temperature profiles tlevels[nprofiles][nlevel] (host) in ONE launch sequence over the shared
     * opacity arrays (batch = nprofiles x ngauss), `reflected` is ignored, and jac_out (host) receives
     * [nprofiles][2][nlevel] = flux_net_ir_layer, flux_net_ir of every profile with one device-to-host copy.
     * tlevel / the eight pointers / packed above are not used in this mode. */
    int nprofiles;
    const double *tlevels;
    double *jac_out;
} pb_climate_args;

int pb_climate_get_fluxes(pb_ctx *ctx, const pb_climate_args *args, int memspace);

/* A climate call bound once and re-run with two integers - callable from numba nopython code through ctypes
 * (the reference's solver loop t_start and get_fluxes are jitted, climate.py:804, :1686, so Python rebinding cannot
 * reach them).  pb_climate_bind copies the argument struct (the arrays it points to stay owned by the caller and
 * must outlive the binding; the solver writes new temperatures into the bound tlevel / tlevels array in place);
 * pb_climate_run_bound(ctx address, handle) runs it and returns the status code. */
int pb_climate_bind(pb_ctx *ctx, const pb_climate_args *args, int memspace, int *handle_out);
int pb_climate_run_bound(unsigned long long ctx_address, int handle);
int pb_climate_unbind(pb_ctx *ctx, int handle);

/* ---- rebinning onto a data grid ---------------------------------------------------------- */
/* replaces mean_regrid, picaso/justplotit.py:31-63 = scipy.stats.binned_statistic(x, y, 'mean', bins):
 * out[b][i] = mean of scale*y[b][start[i] .. start[i]+count[i]) (NaN for an empty bin).  The (start,
 * count) ranges are planned once per (model grid, data grid) pair on the host (picaso_b200/regrid.py:
 * np.digitize + scipy's right-edge rule; x must be monotonic) and kept in HBM by the plan.  `scale`
 * folds the 1e-8 (R/d)^2 factor the retrieval driver applies (picaso/driver.py:226). */
typedef struct pb_regrid_plan pb_regrid_plan;
int pb_regrid_plan_create(pb_ctx *ctx, int nbins, const int *start, const int *count, pb_regrid_plan **out);
int pb_regrid_plan_destroy(pb_ctx *ctx, pb_regrid_plan *plan);
int pb_mean_regrid(pb_ctx *ctx, const pb_regrid_plan *plan, int nbatch, int nwno, int64_t ld, const double *y,
                   double scale, double *out, int memspace);

/* ---- peer memory for the fused all-gather ------------------------------------------------ */
/* cudaIpcGetMemHandle / cudaIpcOpenMemHandle on pb_dev_alloc'ed blocks (handle = 64 bytes) */
int pb_ipc_export(pb_ctx *ctx, void *dev_ptr, void *handle64);
int pb_ipc_open(pb_ctx *ctx, const void *handle64, void **dev_ptr);
int pb_ipc_close(pb_ctx *ctx, void *dev_ptr);
/* stream-ordered: returns (on the stream) once flags[0..nranks) >= step; gives up after ~2 s of spinning
 * and sets *timed_out_dev (device int, may be NULL) */
int pb_gather_wait(pb_ctx *ctx, const unsigned long long *flags, int nranks, unsigned long long step,
                   int *timed_out_dev);
/* stream-ordered: after everything queued before it, store `value` into word `offset + rank` of every rank's flag
 * array (flags: host array [nranks] of device pointers as mapped here; release at system scope).  Publishes the last
 * step of a push = 2 ("lazy") gather, and - with pb_gather_wait on flags + offset - makes a device-side barrier. */
/* push = 3 ("deferred"): deliver the slab of step `g->step` (row `rank` of the local buffer g->albedo[rank]) to every
 * peer and publish its flags, stream-ordered - what the next solver launch would have done for it. */
int pb_peer_flush(pb_ctx *ctx, const struct pb_peer_gather *g, int nwno);
int pb_peer_signal(pb_ctx *ctx, unsigned long long *const *flags, int nranks, int rank, int offset,
                   unsigned long long value);

/* ---- self test ------------------------------------------------------------------------- */
/* evaluates the kernels' branch-free exp() and 1/x on x[n] (host pointers); test hook */
int pb_selftest_math(pb_ctx *ctx, const double *x, int n, double *exp_out, double *rcp_out);
/* evaluates the flux kernels' shared-memory table exp() on x[n] (host pointers, x <= 709); test hook */
int pb_selftest_exp_tab(pb_ctx *ctx, const double *x, int n, double *exp_out);
/* machine numbers of the fp64 pipe (the second roof in bench.py's roofline block; no reference counterpart).
 * which: 0 DFMA throughput, 1 DFMA latency, 2 DFMA throughput with half-active warps, 3 LDS.64 throughput,
 * 4 reciprocal latency, 5 DFMA throughput at 3 warps per SM sub-partition.
 * out[4] = {elapsed ms, operations per thread, threads, SM cycles per iteration (thread 0)} */
int pb_microbench(pb_ctx *ctx, int which, int iters, double *out);

#ifdef __cplusplus
}
#endif
#endif /* PICASO_B200_H */
