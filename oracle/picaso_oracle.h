/*
 * picaso_oracle.h - CPU restatement of PICASO's per-wavelength radiative-transfer
 * hot path.  TEST INFRASTRUCTURE ONLY: nothing in the shipped product path
 * (picaso_b200/) may link, load or call this library.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
 *
 * Parity status: PINNED.  Every function below is checked against outputs of the
 * unmodified reference (numba functions of /root/reference/picaso/fluxes.py and
 * disco.py, commit 0369089, imported by file path) on seeded inputs; the generated
 * vectors live in tests/golden/ together with the script that made them
 * (tests/golden/make_golden.py).  The reference's own tests hold no golden vectors
 * at this boundary.
 *
 * All arrays are float64, C-order, wavelength on the fastest axis: a "layer array"
 * is [nlayer][nwno], a "level array" is [nlevel][nwno], nlevel = nlayer + 1.
 * nthreads > 1 splits the wavelength axis with OpenMP (the reference itself is
 * single-threaded; wavelengths are independent).
 */
#ifndef PICASO_ORACLE_H
#define PICASO_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* follows fluxes.py:1010-1413 (get_reflected_1d) with setup_tri_diag :139-183 and
 * tri_diag_solve :311-323.  Outputs: xint_at_top [ng*nt][nwno]; the four level
 * arrays [ng*nt][nlevel][nwno] may be NULL when get_lvl_flux == 0. */
void orc_get_reflected_1d(
    int nlevel, int nwno, int numg, int numt,
    const double *dtau, const double *tau, const double *w0, const double *cosb,
    const double *gcos2, const double *ftau_cld, const double *ftau_ray,
    const double *dtau_og, const double *tau_og, const double *w0_og, const double *cosb_og,
    const double *surf_reflect /*[nwno]*/, const double *ubar0, const double *ubar1,
    double cos_theta, const double *F0PI /*[nwno]*/,
    int single_phase, int multi_phase,
    double frac_a, double frac_b, double frac_c, double constant_back, double constant_forward,
    int get_toa_intensity, int get_lvl_flux, int toon_coefficients,
    const double *b_top /*[nwno]*/,
    double *xint_at_top, double *flux_minus, double *flux_plus,
    double *flux_minus_mdpt, double *flux_plus_mdpt, int nthreads);

/* follows fluxes.py:1683-1912 (get_thermal_1d), blackbody :1661-1680,
 * blackbody_integrated :1609-1658.  Level arrays may be NULL. */
void orc_get_thermal_1d(
    int nlevel, const double *wno, int nwno, int numg, int numt,
    const double *tlevel, const double *dtau, const double *w0, const double *cosb,
    const double *plevel, const double *ubar1, const double *surf_reflect /*[nwno]*/,
    int hard_surface, const double *dwno, int calc_type,
    double *flux_at_top, double *flux_minus, double *flux_plus,
    double *flux_minus_mdpt, double *flux_plus_mdpt, int nthreads);

/* follows fluxes.py:2582-2663 (get_transit_1d).  player/tlayer have nlevel entries
 * as passed by picaso() (justdoit.py:392-396). */
void orc_get_transit_1d(
    const double *z, const double *dz, int nlevel, int nwno, double rstar,
    const double *mmw, double k_b, double amu, const double *player, const double *tlayer,
    const double *colden, const double *DTAU, double *F, int nthreads);

/* follows disco.py:118-149 (compress_disco) */
void orc_compress_disco(int nwno, double cos_theta, const double *xint_at_top,
                        const double *gweight, int ng, const double *tweight, int nt,
                        const double *F0PI, double *albedo);

/* follows disco.py:152-180 (compress_thermal); n = trailing size (nwno or nlevel*nwno) */
void orc_compress_thermal(int n, const double *flux_at_top, const double *gweight, int ng,
                          const double *tweight, int nt, double *flux);

/* follows fluxes.py:2675-2976 (get_reflected_SH; setup_2_stream_fluxes :3189,
 * setup_4_stream_fluxes :3336, solve_4_stream_banded :3610 -> LAPACK dgbsv restated).
 * stream in {2,4}; flx=0 only.  f_deltaM [nlayer][nwno] is MODIFIED exactly like the
 * reference modifies its argument. */
void orc_get_reflected_SH(
    int nlevel, int nwno, int numg, int numt,
    const double *dtau, const double *tau, const double *w0, const double *cosb,
    const double *ftau_cld, const double *ftau_ray, double *f_deltaM,
    const double *dtau_og, const double *tau_og, const double *w0_og, const double *cosb_og,
    const double *surf_reflect, const double *ubar0, const double *ubar1, double cos_theta,
    const double *F0PI, int w_single_form, int w_multi_form, int psingle_form,
    int w_single_rayleigh, int w_multi_rayleigh, int psingle_rayleigh,
    double frac_a, double frac_b, double frac_c, double constant_back, double constant_forward,
    int stream, const double *b_top, int single_form, double *xint_at_top,
    double *flux /* flx = 1: [numg*numt][stream*nlevel][nwno]; NULL: flx = 0 */, int nthreads);

/* follows fluxes.py:2979-3186 (get_thermal_SH, flx = 0).  Arguments the reference accepts
 * but never reads (tau, dtau_og, tau_og, w0_og, w0_no_raman) are not part of this signature. */
void orc_get_thermal_SH(
    int nlevel, const double *wno, int nwno, int numg, int numt, const double *tlevel,
    const double *dtau, const double *w0, const double *cosb, const double *cosb_og,
    const double *plevel, const double *ubar1, const double *surf_reflect, int stream,
    int hard_surface, double *xint_at_top, int nthreads);

/* follow fluxes.py:355-660 (get_reflected_3d) and :2148-2352 (get_thermal_3d); facet-major
 * inputs [numg*numt][nlayer|nlevel][nwno] (tlevel/plevel [numg*numt][nlevel]). */
void orc_get_reflected_3d(
    int nlevel, int nwno, int numg, int numt,
    const double *dtau, const double *tau, const double *w0, const double *cosb, const double *gcos2,
    const double *ftau_cld, const double *ftau_ray, const double *dtau_og, const double *tau_og,
    const double *w0_og, const double *cosb_og, const double *surf_reflect, const double *ubar0,
    const double *ubar1, double cos_theta, const double *F0PI, int single_phase, int multi_phase,
    double frac_a, double frac_b, double frac_c, double constant_back, double constant_forward,
    double *xint_at_top, int nthreads);
void orc_get_thermal_3d(
    int nlevel, const double *wno, int nwno, int numg, int numt, const double *tlevel,
    const double *dtau, const double *w0, const double *cosb, const double *plevel,
    const double *ubar1, const double *surf_reflect, int hard_surface, double *int_at_top,
    int nthreads);

#ifdef __cplusplus
}
#endif
#endif
