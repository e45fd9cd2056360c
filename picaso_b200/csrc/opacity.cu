// opacity.cu - device-resident opacity state + fused interpolation / mixing / Raman /
// delta-Eddington kernel for sm_100a.
//
// Replaces, for monochromatic (ngauss = 1) opacities:
//   RetrieveOpacities.get_opacities          picaso/optics.py:2241-2308  (bilinear in 1/T, log10 P)
//   RetrieveOpacities.get_opacities_nearest  optics.py:2310-2368         (the reference default)
//   compute_opacity                          optics.py:147-431
//   compute_raman                            optics.py:435-494
//
// B200 design.  The reference fetches sqlite blobs per call, builds one [nlayer, nwno] array
// per molecule with Python loops, then sums NumPy temporaries; a drop-in that takes those
// arrays from the host would move (N_mol + N_cia + 13) * L * W * 8 bytes over PCIe per
// spectrum.  Here the cross-section tables are uploaded ONCE into HBM (180 GB holds a full
// R ~ 1e5 database) - both raw (nearest lookups) and as log10(max-guarded) rows so that the
// bilinear interpolation needs no log per call - and a call ships only O(nlayer) scalars:
// table-row indices, interpolation weights and the per-layer multipliers
// colden * x_mol / mmw etc.  One thread per (layer, wavelength) - 4 M threads at 80 x 50 000 -
// gathers 4 table rows per molecule (coalesced 256-B segments; layers sharing (T, P) neighbours
// share rows through L2), applies 10**bilinear, mixes continuum + molecular + Rayleigh + cloud,
// the Raman factor and delta-Eddington, and writes only the outputs the caller asked for
// (transit needs 1 array, thermal 3, reflected 11).  A second, tiny kernel accumulates the
// running optical depths in the reference's summation order.  Per-wavelength Raman sums are
// built once per star (pb_optab_set_raman), not per call.
#include <string>
#include <vector>

#include "pb_common.cuh"
#include "pb_math.cuh"

struct pb_optab {
    int W = 0, nmol = 0, ncont = 0, nray = 0, ntrans = 0;
    std::vector<double *> mol_raw, mol_log;  // [npt][W] device
    std::vector<int> mol_npt;
    std::vector<double *> cont;              // [ntemp][W]
    std::vector<double *> cont_ln;           // ln of the same rows (log-linear interpolation in 1/T)
    double *ck = nullptr;                    // pre-mixed correlated-k ln(kappa) [nP][nT][W][K]
    int ck_np = 0, ck_nt = 0, K = 1;
    std::vector<int> cont_nt;
    std::vector<double *> ray;               // [W]
    double *wno = nullptr;                   // [W]
    double *shifts = nullptr;                // [W][ntrans]
    double *raman_c = nullptr, *raman_dnu = nullptr;
    int *raman_ji = nullptr;
    double *RA = nullptr, *RB = nullptr;       // [10][W] per-level Raman sums
    // device-side pointer tables rebuilt when a table changes
    const double **d_mol_raw = nullptr, **d_mol_log = nullptr, **d_cont = nullptr, **d_ray = nullptr,
                 **d_cont_ln = nullptr;
    bool dirty = true;
    size_t bytes = 0;
};

namespace {

constexpr int kMaxJ = 10;

struct OpaParams {
    int L, W, nmol, ncont, nray, ntrans;
    int query;  // 0 nearest, 1 bilinear
    const double *const *mol_raw, *const *mol_log, *const *cont, *const *ray, *const *cont_ln;
    int K;                     // gauss points per wavelength (1 = monochromatic)
    const double *ck;          // ln kappa [nP*nT][W*K] or null
    const int *ck_index;       // [L][4]
    const double *ck_wts;      // [L][4]
    const double *ck_scale;    // [L] colden/mmw
    const double *ck_direct;   // molecular_opa [L][W*K] from pb_ck_mix, or null
    int cont_mode;             // 0 nearest row, 1 log-linear between cont_index and cont_index_hi
    const int *cont_index_hi;  // [L]
    const double *cont_t;      // [L]
    const int *pt_index;       // [L][4]
    const double *wts;         // [L][4] bilinear weights in the reference's term order
    const double *mol_scale;   // [nmol][L]
    const int *cont_index;     // [L]
    const double *cont_scale;  // [ncont][L]
    const double *ray_scale;   // [nray][L]
    int raman;                 // 0 oklopcic, 1 pollack, 2 none
    const double *jfrac;       // [10][L]
    const double *RA, *RB, *pollack;
    const double *cld_opd, *cld_w0, *cld_g0;  // [L][ld] or null
    int64_t ld;
    double fthin;
    int do_holes, stream, dedd;
    double *o[13];
    double *x[3];              // full_output: TAUGAS, TAURAY, TAUCLD
    int test_mode;             // 0 off, 1 'rayleigh', 2 other (optics.py:372-399)
    int64_t bs_out;
};

__global__ void log_table_kernel(int64_t n, const double *raw, double *lg)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double a = raw[i];
    lg[i] = log10(a != 0 ? a : 1e-50);  // optics.py:2282
}

__global__ void ln_table_kernel(int64_t n, const double *raw, double *ln)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ln[i] = log(raw[i]);  // optics.py:1483-1484
}

// per-wavelength Raman sums grouped by initial rotational level (optics.py:478-491); they depend
// only on the wavenumber grid and the stellar shifts, so they are built once per star:
// RA[j][w] = sum_{i: ji=j} Q_i * (deltanu_i == 0 ? 1 : shift_i),  RB[j][w] = sum_{i: ji=j} Q_i
__global__ void raman_sums_kernel(int W, int ntrans, const double *wno, const double *shifts, const double *c,
                                  const double *dnu_, const int *ji_, double *RA, double *RB)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= W) return;
    double ra[kMaxJ], rb[kMaxJ];
#pragma unroll
    for (int j = 0; j < kMaxJ; ++j) ra[j] = rb[j] = 0.0;
    const double wn = wno[w];
    const double w3 = wn * wn * wn;
    const double *sh = shifts + (int64_t)w * ntrans;
    for (int i = 0; i < ntrans; ++i) {
        const double dnu = dnu_[i];
        const double Q = c[i] / w3 / (wn + dnu);
        const int ji = ji_[i];
        const double qa = (dnu == 0.0) ? Q : Q * sh[i];
#pragma unroll
        for (int j = 0; j < kMaxJ; ++j)
            if (j == ji) { ra[j] += qa; rb[j] += Q; }
    }
#pragma unroll
    for (int j = 0; j < kMaxJ; ++j) { RA[(int64_t)j * W + w] = ra[j]; RB[(int64_t)j * W + w] = rb[j]; }
}

// One thread per (layer, VEC consecutive wavelengths): blockIdx.y = layer, so everything that
// depends on the layer only - table-row base pointers, interpolation weights, multipliers,
// j-fractions - is resolved ONCE per CTA into shared memory; a thread then issues one 16-byte
// (VEC = 2) load per table row.  VEC = 2 needs an even nwno (rows are then 16-byte aligned).
template <int VEC>
struct Vec;
template <>
struct Vec<1> {
    double v[1];
    __device__ __forceinline__ void load(const double *p) { v[0] = __ldg(p); }
    __device__ __forceinline__ void store(double *p) const { p[0] = v[0]; }
};
template <>
struct Vec<2> {
    double v[2];
    __device__ __forceinline__ void load(const double *p)
    {
        const double2 t = __ldg(reinterpret_cast<const double2 *>(p));
        v[0] = t.x; v[1] = t.y;
    }
    __device__ __forceinline__ void store(double *p) const { *reinterpret_cast<double2 *>(p) = make_double2(v[0], v[1]); }
};

#ifndef PB_OPA_PF
#define PB_OPA_PF 1
#endif
template <int VEC>
__global__ void __launch_bounds__(128) opacity_layer_kernel(OpaParams p)
{
    extern __shared__ unsigned char s_raw[];
    const int L = p.L, W = p.W, l = blockIdx.y;
    const int nrow = (p.query == 1) ? 4 : 1;
    // shared layout: row pointers [nmol*nrow + ncont + nray], then doubles [4 + nmol + ncont + nray + 10]
    const double **s_ptr = reinterpret_cast<const double **>(s_raw);
    const int K = p.K, C = W * K;
    const int ncp = p.ncont * (p.cont_mode ? 2 : 1);  // continuum rows per pair: nearest | (low, high)
    const int nptr = p.nmol * nrow + ncp + p.nray + 4;
    double *s_d = reinterpret_cast<double *>(s_raw + sizeof(double *) * (size_t)nptr);
    double *s_w = s_d, *s_ms = s_d + 4, *s_cs = s_ms + p.nmol, *s_rs = s_cs + p.ncont, *s_jf = s_rs + p.nray,
           *s_ck = s_jf + kMaxJ;  // 4 ck weights, ck scale, continuum t
    for (int i = threadIdx.x; i < nptr; i += blockDim.x) {
        if (i < p.nmol * nrow) {
            const int m = i / nrow, k = i - m * nrow;
            const double *base = (p.query == 1) ? p.mol_log[m] : p.mol_raw[m];
            s_ptr[i] = base + (int64_t)p.pt_index[4 * l + k] * W;
        } else if (i < p.nmol * nrow + ncp) {
            const int c = i - p.nmol * nrow;
            if (p.cont_mode == 0) s_ptr[i] = p.cont[c] + (int64_t)p.cont_index[l] * W;
            else s_ptr[i] = p.cont_ln[c >> 1] + (int64_t)((c & 1) ? p.cont_index_hi[l] : p.cont_index[l]) * W;
        } else if (i < p.nmol * nrow + ncp + p.nray) {
            s_ptr[i] = p.ray[i - p.nmol * nrow - ncp];
        } else {
            const int k = i - (p.nmol * nrow + ncp + p.nray);
            s_ptr[i] = p.ck ? p.ck + (int64_t)p.ck_index[4 * l + k] * C : nullptr;
        }
    }
    for (int i = threadIdx.x; i < 4 + p.nmol + p.ncont + p.nray + kMaxJ + 6; i += blockDim.x) {
        double v;
        const int base_ck = 4 + p.nmol + p.ncont + p.nray + kMaxJ;
        if (i < 4) v = (p.query == 1 && p.nmol) ? p.wts[4 * l + i] : 0.0;
        else if (i < 4 + p.nmol) v = p.mol_scale[(i - 4) * L + l];
        else if (i < 4 + p.nmol + p.ncont) v = p.cont_scale[(i - 4 - p.nmol) * L + l];
        else if (i < 4 + p.nmol + p.ncont + p.nray) v = p.ray_scale[(i - 4 - p.nmol - p.ncont) * L + l];
        else if (i < base_ck) v = (p.raman == 0) ? p.jfrac[(i - 4 - p.nmol - p.ncont - p.nray) * L + l] : 0.0;
        else if (i < base_ck + 4) v = p.ck ? p.ck_wts[4 * l + (i - base_ck)] : 0.0;
        else if (i == base_ck + 4) v = (p.ck || p.ck_direct) ? p.ck_scale[l] : 0.0;
        else v = p.cont_mode ? p.cont_t[l] : 0.0;
        s_d[i] = v;
    }
    __syncthreads();
    // column j runs over (wavelength, gauss point) with the gauss point fastest, the reference's
    // [nlayer, nwno, ngauss] layout; per-wavelength tables are read at w = j / K
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    if (j >= C) return;
    const int w = (K == 1) ? j : j / K;
    const double N_A = 6.02214086e+23;
    double taugas[VEC], tauray[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) taugas[v] = tauray[v] = 0.0;
    // continuum (optics.py:172-233): table row of the nearest CIA temperature x layer factor
    const double **cp = s_ptr + p.nmol * nrow;
    if (p.cont_mode == 0) {
        for (int c = 0; c < p.ncont; ++c) {
            Vec<VEC> k;
            k.load(cp[c] + w);
            const double sc = s_cs[c];
#pragma unroll
            for (int v = 0; v < VEC; ++v) taugas[v] += k.v[v] * sc;
        }
    } else {
        // RetrieveCKs.get_continuum (optics.py:1471-1497): log-linear in 1/T between the bracketing rows
        const double t = s_ck[5];
        for (int c = 0; c < p.ncont; ++c) {
            Vec<VEC> lo, hi;
            lo.load(cp[2 * c] + w);
            hi.load(cp[2 * c + 1] + w);
            const double sc = s_cs[c];
#pragma unroll
            for (int v = 0; v < VEC; ++v) taugas[v] += exp(((1 - t) * lo.v[v]) + ((t)*hi.v[v])) * sc;
        }
    }
    // pre-mixed correlated-k (RetrieveCKs.get_pre_mix_ck, optics.py:1151-1161; compute_opacity :257-262)
    if (p.ck) {
        const double **kp = s_ptr + p.nmol * nrow + ncp + p.nray;
        Vec<VEC> a1, a2, a3, a4;
        a1.load(kp[0] + j); a2.load(kp[1] + j); a3.load(kp[2] + j); a4.load(kp[3] + j);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const double e = ((s_ck[0] * a1.v[v]) + (s_ck[1] * a2.v[v]) + (s_ck[2] * a3.v[v]) + (s_ck[3] * a4.v[v]));
            taugas[v] += (exp(e) * N_A) * s_ck[4];
        }
    }
    // resort-rebin mixed k-coefficients (optics.py:1197 molecular_opa, compute_opacity :257-262)
    if (p.ck_direct) {
        Vec<VEC> k;
        k.load(p.ck_direct + (int64_t)l * C + j);
#pragma unroll
        for (int v = 0; v < VEC; ++v) taugas[v] += k.v[v] * s_ck[4];
    }
    // molecular (optics.py:243-250)
    if (p.query == 1) {
        const double w1 = s_w[0], w2 = s_w[1], w3 = s_w[2], w4 = s_w[3];
#if PB_OPA_PF
        // software pipeline over the molecules: the four table rows of molecule m + 1 are requested before molecule m's
        // exponentials are evaluated (ncu r1: long_scoreboard 8.0 of the stall cycles per issue - each trip of the plain
        // loop waited for its own four gathers)
        Vec<VEC> n1, n2, n3, n4;
        if (p.nmol > 0) {
            n1.load(s_ptr[0] + w); n2.load(s_ptr[1] + w); n3.load(s_ptr[2] + w); n4.load(s_ptr[3] + w);
        }
        for (int m = 0; m < p.nmol; ++m) {
            const Vec<VEC> a1 = n1, a2 = n2, a3 = n3, a4 = n4;
            if (m + 1 < p.nmol) {
                n1.load(s_ptr[4 * m + 4] + w);
                n2.load(s_ptr[4 * m + 5] + w);
                n3.load(s_ptr[4 * m + 6] + w);
                n4.load(s_ptr[4 * m + 7] + w);
            }
#else
        for (int m = 0; m < p.nmol; ++m) {
            Vec<VEC> a1, a2, a3, a4;
            a1.load(s_ptr[4 * m] + w);
            a2.load(s_ptr[4 * m + 1] + w);
            a3.load(s_ptr[4 * m + 2] + w);
            a4.load(s_ptr[4 * m + 3] + w);
#endif
            const double sc = s_ms[m];
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                // 10**((1-t)(1-p) l1 + t(1-p) l2 + t p l3 + (1-t) p l4), optics.py:2290-2293
                const double e = ((w1 * a1.v[v]) + (w2 * a2.v[v]) + (w3 * a3.v[v]) + (w4 * a4.v[v]));
                // 10**e as exp(e ln 10): |e ln 10| < 120 adds < 2e-14 relative error, far inside 1e-6
                taugas[v] += (exp(e * 2.302585092994045684) * N_A) * sc;
            }
        }
    } else {
#pragma unroll 4
        for (int m = 0; m < p.nmol; ++m) {
            Vec<VEC> k;
            k.load(s_ptr[m] + w);
            const double sc = s_ms[m];
#pragma unroll
            for (int v = 0; v < VEC; ++v) taugas[v] += (k.v[v] * N_A) * sc;
        }
    }
    // Rayleigh (optics.py:265-271)
    const double **rp = cp + ncp;
    for (int m = 0; m < p.nray; ++m) {
        Vec<VEC> k;
        k.load(rp[m] + w);
        const double sc = s_rs[m];
#pragma unroll
        for (int v = 0; v < VEC; ++v) tauray[v] += k.v[v] * sc;
    }
    // Raman factor (optics.py:287-306), capped at 0.99999
    double rf[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) rf[v] = 0.99999;
    if (p.raman == 0) {
        double num[VEC], den[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) num[v] = den[v] = 0.0;
#pragma unroll
        for (int j = 0; j < kMaxJ; ++j) {
            Vec<VEC> ra, rb;
            ra.load(p.RA + (int64_t)j * W + w);
            rb.load(p.RB + (int64_t)j * W + w);
            const double f = s_jf[j];
#pragma unroll
            for (int v = 0; v < VEC; ++v) { num[v] = fma(f, ra.v[v], num[v]); den[v] = fma(f, rb.v[v], den[v]); }
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v) rf[v] = fmin(num[v] / den[v], 0.99999);
    } else if (p.raman == 1) {
        Vec<VEC> pl;
        pl.load(p.pollack + w);
#pragma unroll
        for (int v = 0; v < VEC; ++v) rf[v] = fmin(pl.v[v], 0.99999);
    }
    // cloud (optics.py:309-315)
    Vec<VEC> opd, cw0, cg0, opd_raw;
#pragma unroll
    for (int v = 0; v < VEC; ++v) opd.v[v] = cw0.v[v] = cg0.v[v] = opd_raw.v[v] = 0.0;
    if (p.cld_opd) {
        const int64_t ic = (int64_t)l * p.ld + w;
        if (VEC == 1 || (p.ld & 1) == 0) {
            opd.load(p.cld_opd + ic); cw0.load(p.cld_w0 + ic); cg0.load(p.cld_g0 + ic);
        } else {
#pragma unroll
            for (int v = 0; v < VEC; ++v) { opd.v[v] = __ldg(p.cld_opd + ic + v); cw0.v[v] = __ldg(p.cld_w0 + ic + v); cg0.v[v] = __ldg(p.cld_g0 + ic + v); }
        }
        opd_raw = opd;   // atm.layer['cloud']['opd'] itself: the test modes ignore fthin_cld (optics.py:386)
        if (p.do_holes) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) opd.v[v] = p.fthin * opd.v[v];
        }
    }
    // totals (optics.py:329-350) and delta-Eddington (optics.py:412-420)
    Vec<VEC> o0, o2, o3, o4, o5, o6, o7, o9, o10, o11, o12;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        const double taucld = opd.v[v], w0c = cw0.v[v], g0 = cg0.v[v];
        double dtau = taugas[v] + tauray[v] + taucld;
        const double sc = w0c * taucld;
        double w0 = (tauray[v] * rf[v] + taucld * w0c) / dtau;
        const double fray = tauray[v] / (tauray[v] + sc);
        o4.v[v] = sc / (sc + tauray[v]);
        o5.v[v] = fray;
        o6.v[v] = 0.5 * fray;
        o11.v[v] = (tauray[v] * 0.99999 + taucld * w0c) / dtau;
        if (p.test_mode) {
            // optics.py:372-399: Rayleigh-only / cloud-only optical depth, the cloud's w0 and g0 everywhere
            const bool ray = p.test_mode == 1;
            dtau = ray ? tauray[v] : opd_raw.v[v];
            if (dtau <= 0) dtau = 1e-10;
            o4.v[v] = ray ? 0.0 : 1.0;
            o5.v[v] = ray ? 1.0 : 0.0;
            o6.v[v] = ray ? 0.5 : 0.0;
            w0 = w0c <= 0 ? 1e-10 : w0c;
            o11.v[v] = w0;
        }
        o7.v[v] = dtau;
        o9.v[v] = w0;
        o10.v[v] = g0;
        if (p.dedd) {
            double f = 1.0;
            for (int s = 0; s < p.stream; ++s) f *= g0;
            o0.v[v] = dtau * (1. - w0 * f);
            o2.v[v] = w0 * (1. - f) / (1.0 - w0 * f);
            o3.v[v] = (g0 - f) / (1. - f);
            o12.v[v] = f;
        } else {
            o0.v[v] = dtau; o2.v[v] = w0; o3.v[v] = g0; o12.v[v] = 0 * g0;
        }
    }
    const int64_t io = (int64_t)l * C + j;
    if (p.x[0] || p.x[1] || p.x[2]) {
        Vec<VEC> xg, xr;
#pragma unroll
        for (int v = 0; v < VEC; ++v) { xg.v[v] = taugas[v]; xr.v[v] = tauray[v]; }
        if (p.x[0]) xg.store(p.x[0] + io);
        if (p.x[1]) xr.store(p.x[1] + io);
        if (p.x[2]) opd.store(p.x[2] + io);
    }
    if (p.o[0]) o0.store(p.o[0] + io);
    if (p.o[2]) o2.store(p.o[2] + io);
    if (p.o[3]) o3.store(p.o[3] + io);
    if (p.o[4]) o4.store(p.o[4] + io);
    if (p.o[5]) o5.store(p.o[5] + io);
    if (p.o[6]) o6.store(p.o[6] + io);
    if (p.o[7]) o7.store(p.o[7] + io);
    if (p.o[9]) o9.store(p.o[9] + io);
    if (p.o[10]) o10.store(p.o[10] + io);
    if (p.o[11]) o11.store(p.o[11] + io);
    if (p.o[12]) o12.store(p.o[12] + io);
}

// TAU[0] = 0, TAU[l+1] = TAU[l] + DTAU[l]  (numba_cumsum, optics.py:353-354, :419-420): one
// wavelength per thread, sequential in l so that the summation order is the reference's
__global__ void opacity_cumsum_kernel(int L, int W, const double *dtau, double *tau)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= W) return;
    double acc = 0.0;
    tau[w] = 0.0;
    for (int l = 0; l < L; ++l) {
        acc += __ldg(dtau + (int64_t)l * W + w);
        tau[(int64_t)(l + 1) * W + w] = acc;
    }
}

int upload_table(pb_ctx *ctx, const double *host, size_t n, double **dev, size_t *bytes)
{
    if (*dev) {
        PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        PB_CUDA(ctx, cudaFree(*dev));
        *dev = nullptr;
    }
    cudaError_t e = cudaMalloc((void **)dev, n * sizeof(double));
    if (e != cudaSuccess) return pb_fail(ctx, PB_ERR_NOMEM, "opacity table cudaMalloc(%zu) -> %s", n * 8, cudaGetErrorString(e));
    PB_CUDA(ctx, cudaMemcpyAsync(*dev, host, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *bytes += n * sizeof(double);
    return PB_OK;
}

int sync_pointer_tables(pb_ctx *ctx, pb_optab *t)
{
    if (!t->dirty) return PB_OK;
    auto push = [&](const std::vector<double *> &v, const double ***d) -> int {
        if (*d) { PB_CUDA(ctx, cudaFree((void *)*d)); *d = nullptr; }
        if (v.empty()) return PB_OK;
        PB_CUDA(ctx, cudaMalloc((void **)d, v.size() * sizeof(double *)));
        PB_CUDA(ctx, cudaMemcpy((void *)*d, v.data(), v.size() * sizeof(double *), cudaMemcpyHostToDevice));
        return PB_OK;
    };
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    PB_TRY(push(t->mol_raw, &t->d_mol_raw));
    PB_TRY(push(t->mol_log, &t->d_mol_log));
    PB_TRY(push(t->cont, &t->d_cont));
    PB_TRY(push(t->cont_ln, &t->d_cont_ln));
    PB_TRY(push(t->ray, &t->d_ray));
    t->dirty = false;
    return PB_OK;
}

} // namespace

extern "C" int pb_optab_create(pb_ctx *ctx, int nwno, int nmol, int ncont, int nray, pb_optab **out)
{
    if (!ctx || !out || nwno < 1 || nmol < 0 || ncont < 0 || nray < 0) return pb_fail(ctx, PB_ERR_ARG, "optab_create: bad arguments");
    pb_optab *t = new pb_optab();
    t->W = nwno; t->nmol = nmol; t->ncont = ncont; t->nray = nray;
    t->mol_raw.assign(nmol, nullptr); t->mol_log.assign(nmol, nullptr); t->mol_npt.assign(nmol, 0);
    t->cont.assign(ncont, nullptr); t->cont_nt.assign(ncont, 0); t->cont_ln.assign(ncont, nullptr);
    t->ray.assign(nray, nullptr);
    *out = t;
    return PB_OK;
}

extern "C" int pb_optab_destroy(pb_ctx *ctx, pb_optab *t)
{
    if (!ctx || !t) return PB_ERR_ARG;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto p : t->mol_raw) if (p) cudaFree(p);
    for (auto p : t->mol_log) if (p) cudaFree(p);
    for (auto p : t->cont) if (p) cudaFree(p);
    for (auto p : t->cont_ln) if (p) cudaFree(p);
    if (t->ck) cudaFree(t->ck);
    if (t->d_cont_ln) cudaFree((void *)t->d_cont_ln);
    for (auto p : t->ray) if (p) cudaFree(p);
    if (t->wno) cudaFree(t->wno);
    if (t->shifts) cudaFree(t->shifts);
    if (t->raman_c) cudaFree(t->raman_c);
    if (t->raman_dnu) cudaFree(t->raman_dnu);
    if (t->raman_ji) cudaFree(t->raman_ji);
    if (t->RA) cudaFree(t->RA);
    if (t->RB) cudaFree(t->RB);
    if (t->d_mol_raw) cudaFree((void *)t->d_mol_raw);
    if (t->d_mol_log) cudaFree((void *)t->d_mol_log);
    if (t->d_cont) cudaFree((void *)t->d_cont);
    if (t->d_ray) cudaFree((void *)t->d_ray);
    delete t;
    return PB_OK;
}

extern "C" int pb_optab_set_molecular(pb_ctx *ctx, pb_optab *t, int imol, const double *table, int npt, int store)
{
    if (!ctx || !t || !table || imol < 0 || imol >= t->nmol || npt < 1 || store < 1 || store > 3)
        return pb_fail(ctx, PB_ERR_ARG, "optab_set_molecular: bad arguments");
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t n = (size_t)npt * t->W;
    PB_TRY(upload_table(ctx, table, n, &t->mol_raw[imol], &t->bytes));
    t->mol_npt[imol] = npt;
    if (store & 2) {
        if (t->mol_log[imol]) { PB_CUDA(ctx, cudaFree(t->mol_log[imol])); t->mol_log[imol] = nullptr; }
        cudaError_t e = cudaMalloc((void **)&t->mol_log[imol], n * sizeof(double));
        if (e != cudaSuccess) return pb_fail(ctx, PB_ERR_NOMEM, "opacity log-table cudaMalloc -> %s", cudaGetErrorString(e));
        log_table_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((int64_t)n, t->mol_raw[imol], t->mol_log[imol]);
        PB_CHECK_LAUNCH(ctx);
        PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        t->bytes += n * sizeof(double);
    }
    if (!(store & 1)) {  // keep only the log table
        PB_CUDA(ctx, cudaFree(t->mol_raw[imol]));
        t->mol_raw[imol] = nullptr;
        t->bytes -= n * sizeof(double);
    }
    t->dirty = true;
    return PB_OK;
}

extern "C" int pb_optab_set_continuum(pb_ctx *ctx, pb_optab *t, int icont, const double *table, int ntemp)
{
    if (!ctx || !t || !table || icont < 0 || icont >= t->ncont || ntemp < 1) return pb_fail(ctx, PB_ERR_ARG, "optab_set_continuum: bad arguments");
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t n = (size_t)ntemp * t->W;
    PB_TRY(upload_table(ctx, table, n, &t->cont[icont], &t->bytes));
    t->cont_nt[icont] = ntemp;
    if (t->cont_ln[icont]) { PB_CUDA(ctx, cudaFree(t->cont_ln[icont])); t->cont_ln[icont] = nullptr; }
    PB_CUDA(ctx, cudaMalloc((void **)&t->cont_ln[icont], n * sizeof(double)));
    ln_table_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((int64_t)n, t->cont[icont], t->cont_ln[icont]);
    PB_CHECK_LAUNCH(ctx);
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    t->bytes += n * sizeof(double);
    t->dirty = true;
    return PB_OK;
}

extern "C" int pb_optab_set_rayleigh(pb_ctx *ctx, pb_optab *t, int iray, const double *sigma)
{
    if (!ctx || !t || !sigma || iray < 0 || iray >= t->nray) return pb_fail(ctx, PB_ERR_ARG, "optab_set_rayleigh: bad arguments");
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    PB_TRY(upload_table(ctx, sigma, (size_t)t->W, &t->ray[iray], &t->bytes));
    t->dirty = true;
    return PB_OK;
}

extern "C" int pb_optab_set_raman(pb_ctx *ctx, pb_optab *t, const double *wno, int ntrans, const double *c,
                                  const int *ji, const double *deltanu, const double *stellar_shifts)
{
    if (!ctx || !t || !wno || ntrans < 1 || !c || !ji || !deltanu || !stellar_shifts)
        return pb_fail(ctx, PB_ERR_ARG, "optab_set_raman: bad arguments");
    for (int i = 0; i < ntrans; ++i)
        if (ji[i] < 0 || ji[i] >= kMaxJ) return pb_fail(ctx, PB_ERR_ARG, "optab_set_raman: ji[%d]=%d outside 0..9", i, ji[i]);
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    PB_TRY(upload_table(ctx, wno, (size_t)t->W, &t->wno, &t->bytes));
    PB_TRY(upload_table(ctx, stellar_shifts, (size_t)t->W * ntrans, &t->shifts, &t->bytes));
    PB_TRY(upload_table(ctx, c, (size_t)ntrans, &t->raman_c, &t->bytes));
    PB_TRY(upload_table(ctx, deltanu, (size_t)ntrans, &t->raman_dnu, &t->bytes));
    if (t->raman_ji) { PB_CUDA(ctx, cudaFree(t->raman_ji)); t->raman_ji = nullptr; }
    PB_CUDA(ctx, cudaMalloc((void **)&t->raman_ji, ntrans * sizeof(int)));
    PB_CUDA(ctx, cudaMemcpy(t->raman_ji, ji, ntrans * sizeof(int), cudaMemcpyHostToDevice));
    t->ntrans = ntrans;
    if (!t->RA) {
        PB_CUDA(ctx, cudaMalloc((void **)&t->RA, (size_t)kMaxJ * t->W * sizeof(double)));
        PB_CUDA(ctx, cudaMalloc((void **)&t->RB, (size_t)kMaxJ * t->W * sizeof(double)));
        t->bytes += 2 * (size_t)kMaxJ * t->W * sizeof(double);
    }
    raman_sums_kernel<<<(t->W + 127) / 128, 128, 0, ctx->stream>>>(t->W, ntrans, t->wno, t->shifts, t->raman_c,
                                                                    t->raman_dnu, t->raman_ji, t->RA, t->RB);
    PB_CHECK_LAUNCH(ctx);
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PB_OK;
}

extern "C" int pb_optab_set_ck(pb_ctx *ctx, pb_optab *t, const double *lnkappa, int npress, int ntemp, int ngauss)
{
    if (!ctx || !t || !lnkappa || npress < 2 || ntemp < 2 || ngauss < 1)
        return pb_fail(ctx, PB_ERR_ARG, "optab_set_ck: bad arguments");
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    PB_TRY(upload_table(ctx, lnkappa, (size_t)npress * ntemp * t->W * ngauss, &t->ck, &t->bytes));
    t->ck_np = npress; t->ck_nt = ntemp; t->K = ngauss;
    return PB_OK;
}

extern "C" int pb_optab_bytes(const pb_optab *t, size_t *bytes)
{
    if (!t || !bytes) return PB_ERR_ARG;
    *bytes = t->bytes;
    return PB_OK;
}

extern "C" int pb_compute_opacity(pb_ctx *ctx, pb_optab *t, const pb_opacity_args *a, int memspace)
{
    if (!ctx || !t || !a) return PB_ERR_ARG;
    const int L = a->nlayer, W = t->W;
    const bool direct = a->ck_direct != nullptr;
    const bool ck = !direct && (a->ngauss > 1 || (t->ck && t->nmol == 0 && a->ck_index));
    const int K = direct ? (a->ngauss > 1 ? a->ngauss : 1) : ck ? t->K : 1;
    if (direct && !a->ck_scale) return pb_fail(ctx, PB_ERR_ARG, "compute_opacity: ck_direct needs ck_scale");
    if (!direct && a->ngauss > 1 && (!t->ck || a->ngauss != t->K)) return pb_fail(ctx, PB_ERR_ARG, "compute_opacity: ngauss=%d needs pb_optab_set_ck tables with the same number of gauss points", a->ngauss);
    if (ck && (!a->ck_index || !a->ck_weights || !a->ck_scale)) return pb_fail(ctx, PB_ERR_ARG, "compute_opacity: correlated-k call needs ck_index, ck_weights, ck_scale");
    if (a->cont_mode != 0 && a->cont_mode != 1) return pb_fail(ctx, PB_ERR_ARG, "compute_opacity: cont_mode must be 0 or 1");
    if (a->cont_mode == 1 && (!a->cont_index_hi || !a->cont_t)) return pb_fail(ctx, PB_ERR_ARG, "compute_opacity: cont_mode=1 needs cont_index_hi and cont_t");
    if (L < 1) return pb_fail(ctx, PB_ERR_ARG, "compute_opacity: nlayer < 1");
    if (a->query != 0 && a->query != 1) return pb_fail(ctx, PB_ERR_ARG, "compute_opacity: query must be 0 (nearest) or 1 (bilinear)");
    if (a->raman < 0 || a->raman > 2) return pb_fail(ctx, PB_ERR_ARG, "compute_opacity: raman must be 0, 1 or 2");
    if (a->stream != 2 && a->stream != 4) return pb_fail(ctx, PB_ERR_ARG, "compute_opacity: stream must be 2 or 4");
    if ((t->nmol && !a->pt_index) || !a->cont_index || (t->nmol && !a->mol_scale) || (t->ncont && !a->cont_scale) || (t->nray && !a->ray_scale))
        return pb_fail(ctx, PB_ERR_ARG, "compute_opacity: NULL per-layer vector");
    if (a->query == 1 && t->nmol && !a->weights) return pb_fail(ctx, PB_ERR_ARG, "compute_opacity: bilinear query needs weights");
    if (a->raman == 0 && (!t->shifts || !a->jfrac)) return pb_fail(ctx, PB_ERR_ARG, "compute_opacity: raman=0 needs pb_optab_set_raman and jfrac");
    if (a->raman == 1 && !a->raman_pollack) return pb_fail(ctx, PB_ERR_ARG, "compute_opacity: raman=1 needs raman_pollack[nwno]");
    if (a->test_mode < 0 || a->test_mode > 2) return pb_fail(ctx, PB_ERR_ARG, "compute_opacity: test_mode must be 0, 1 or 2");
    if (a->test_mode && (!a->cloud_opd || !a->cloud_w0 || !a->cloud_g0))
        return pb_fail(ctx, PB_ERR_ARG, "compute_opacity: test modes read the cloud arrays (optics.py:386-395)");
    const bool cloud = a->cloud_opd != nullptr;
    if (cloud && (!a->cloud_w0 || !a->cloud_g0)) return pb_fail(ctx, PB_ERR_ARG, "compute_opacity: cloud needs opd, w0 and g0");
    for (int m = 0; m < t->nmol; ++m) {
        if (a->query == 1 && !t->mol_log[m]) return pb_fail(ctx, PB_ERR_ARG, "compute_opacity: molecule %d has no log table (store & 2)", m);
        if (a->query == 0 && !t->mol_raw[m]) return pb_fail(ctx, PB_ERR_ARG, "compute_opacity: molecule %d has no raw table (store & 1)", m);
        for (int l = 0; l < L; ++l)
            for (int k = 0; k < (a->query ? 4 : 1); ++k) {
                const int r = a->pt_index[4 * l + k];
                if (r < 0 || r >= t->mol_npt[m]) return pb_fail(ctx, PB_ERR_ARG, "compute_opacity: pt_index[%d][%d]=%d outside table of molecule %d", l, k, r, m);
            }
    }
    for (int c = 0; c < t->ncont; ++c) {
        if (!t->cont[c]) return pb_fail(ctx, PB_ERR_ARG, "compute_opacity: continuum table %d not set", c);
        for (int l = 0; l < L; ++l)
            if (a->cont_index[l] < 0 || a->cont_index[l] >= t->cont_nt[c] ||
                (a->cont_mode == 1 && (a->cont_index_hi[l] < 0 || a->cont_index_hi[l] >= t->cont_nt[c])))
                return pb_fail(ctx, PB_ERR_ARG, "compute_opacity: cont_index[%d] out of range", l);
    }
    for (int m = 0; m < t->nray; ++m)
        if (!t->ray[m]) return pb_fail(ctx, PB_ERR_ARG, "compute_opacity: rayleigh table %d not set", m);
    if (ck)
        for (int l = 0; l < L; ++l)
            for (int k = 0; k < 4; ++k)
                if (a->ck_index[4 * l + k] < 0 || a->ck_index[4 * l + k] >= t->ck_np * t->ck_nt)
                    return pb_fail(ctx, PB_ERR_ARG, "compute_opacity: ck_index[%d][%d] out of range", l, k);
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    PB_TRY(sync_pointer_tables(ctx, t));
    const bool host = memspace == PB_HOST;
    const size_t nW = (size_t)W * sizeof(double);
    const size_t nC = nW * K;  // one output row: (wavelength, gauss point) columns
    double *const outs[13] = {a->DTAU, a->TAU, a->W0, a->COSB, a->ftau_cld, a->ftau_ray, a->GCOS2, a->DTAU_OG,
                              a->TAU_OG, a->W0_OG, a->COSB_OG, a->W0_no_raman, a->f_deltaM};
    const bool is_level[13] = {false, true, false, false, false, false, false, false, true, false, false, false, false};
    size_t need = 96 * 256 + 2 * pb_align((size_t)L * nC) + 8 * pb_align((size_t)L * 8) + pb_align(4 * (size_t)L * 12) + pb_align((size_t)L * 8) +
                  (size_t)(t->nmol + t->ncont + t->nray + kMaxJ + 4) * pb_align((size_t)L * 8);
    if (host) {
        if (cloud) need += 3 * pb_align((size_t)L * nW);
        if (a->raman == 1) need += pb_align(nW);
        for (int k = 0; k < 13; ++k)
            if (outs[k]) need += pb_align((size_t)(L + (is_level[k] ? 1 : 0)) * nC);
        need += 3 * pb_align((size_t)L * nC);  // full_output arrays
    }
    pb_arena_reset(ctx);
    PB_TRY(pb_arena_reserve(ctx, need));
    PB_TRY(pb_pinned_reserve(ctx, (size_t)(t->nmol + t->ncont + t->nray + kMaxJ + 32) * ((size_t)L * 8 + 64)));
    OpaParams p;
    memset(&p, 0, sizeof(p));
    p.L = L; p.W = W; p.nmol = t->nmol; p.ncont = t->ncont; p.nray = t->nray; p.ntrans = t->ntrans;
    p.query = a->query;
    p.mol_raw = t->d_mol_raw; p.mol_log = t->d_mol_log; p.cont = t->d_cont; p.ray = t->d_ray;
    p.cont_ln = t->d_cont_ln; p.K = K; p.ck = ck ? t->ck : nullptr; p.cont_mode = a->cont_mode;
    const double *tmp;
    // ints travel through the same pinned bounce as doubles (sizes rounded up to 8 bytes)
    if (t->nmol) {
        PB_TRY(pb_upload_small(ctx, (const double *)a->pt_index, ((size_t)4 * L * sizeof(int) + 7) / 8, &tmp));
        p.pt_index = (const int *)tmp;
    }
    if (ck) {
        PB_TRY(pb_upload_small(ctx, (const double *)a->ck_index, ((size_t)4 * L * sizeof(int) + 7) / 8, &tmp));
        p.ck_index = (const int *)tmp;
        PB_TRY(pb_upload_small(ctx, a->ck_weights, (size_t)4 * L, &p.ck_wts));
        PB_TRY(pb_upload_small(ctx, a->ck_scale, (size_t)L, &p.ck_scale));
    }
    if (direct) {
        p.ck_direct = a->ck_direct;
        PB_TRY(pb_upload_small(ctx, a->ck_scale, (size_t)L, &p.ck_scale));
    }
    if (a->cont_mode == 1) {
        std::vector<int> ci(a->cont_index_hi, a->cont_index_hi + L);
        if (ci.size() & 1) ci.push_back(0);
        PB_TRY(pb_upload_small(ctx, (const double *)ci.data(), ci.size() / 2, &tmp));
        p.cont_index_hi = (const int *)tmp;
        PB_TRY(pb_upload_small(ctx, a->cont_t, (size_t)L, &p.cont_t));
    }
    {
        std::vector<int> ci(a->cont_index, a->cont_index + L);
        if (ci.size() & 1) ci.push_back(0);
        PB_TRY(pb_upload_small(ctx, (const double *)ci.data(), ci.size() / 2, &tmp));
        p.cont_index = (const int *)tmp;
    }
    if (a->query == 1 && t->nmol) PB_TRY(pb_upload_small(ctx, a->weights, (size_t)4 * L, &p.wts));
    if (t->nmol) PB_TRY(pb_upload_small(ctx, a->mol_scale, (size_t)t->nmol * L, &p.mol_scale));
    if (t->ncont) PB_TRY(pb_upload_small(ctx, a->cont_scale, (size_t)t->ncont * L, &p.cont_scale));
    if (t->nray) PB_TRY(pb_upload_small(ctx, a->ray_scale, (size_t)t->nray * L, &p.ray_scale));
    p.raman = a->raman;
    if (a->raman == 0) {
        PB_TRY(pb_upload_small(ctx, a->jfrac, (size_t)kMaxJ * L, &p.jfrac));
        p.RA = t->RA; p.RB = t->RB;
    }
    int64_t ldo;
    if (a->raman == 1) PB_TRY(pb_stage_in(ctx, a->raman_pollack, memspace, 1, W, W, &p.pollack, &ldo));
    p.ld = W;
    if (cloud) {
        const int64_t ldc = a->cloud_ld > 0 ? a->cloud_ld : W;
        PB_TRY(pb_stage_in(ctx, a->cloud_opd, memspace, L, W, ldc, &p.cld_opd, &ldo));
        PB_TRY(pb_stage_in(ctx, a->cloud_w0, memspace, L, W, ldc, &p.cld_w0, &ldo));
        PB_TRY(pb_stage_in(ctx, a->cloud_g0, memspace, L, W, ldc, &p.cld_g0, &ldo));
        p.ld = host ? W : ldc;
    }
    p.fthin = a->fthin_cld; p.do_holes = a->do_holes; p.stream = a->stream; p.dedd = a->delta_eddington;
    p.test_mode = a->test_mode;
    for (int k = 0; k < 13; ++k) {
        p.o[k] = nullptr;
        if (!outs[k]) continue;
        if (host) PB_TRY(pb_arena_alloc(ctx, (size_t)(L + (is_level[k] ? 1 : 0)) * nC, (void **)&p.o[k]));
        else p.o[k] = outs[k];
    }
    double *const xouts[3] = {a->TAUGAS, a->TAURAY, a->TAUCLD};
    for (int k = 0; k < 3; ++k) {
        p.x[k] = nullptr;
        if (!xouts[k]) continue;
        if (host) PB_TRY(pb_arena_alloc(ctx, (size_t)L * nC, (void **)&p.x[k]));
        else p.x[k] = xouts[k];
    }
    PB_TRY(pb_upload_flush(ctx));
    // the running optical depths need the per-layer values even if the caller did not ask for them
    double *dtau_d = p.o[0], *dtau_og = p.o[7];
    if (p.o[1] && !dtau_d) { PB_TRY(pb_arena_alloc(ctx, (size_t)L * nC, (void **)&dtau_d)); p.o[0] = dtau_d; }
    if (p.o[8] && !dtau_og) { PB_TRY(pb_arena_alloc(ctx, (size_t)L * nC, (void **)&dtau_og)); p.o[7] = dtau_og; }
    {
        const int nrow = a->query == 1 ? 4 : 1;
        const size_t smem = sizeof(double *) * (size_t)(t->nmol * nrow + 2 * t->ncont + t->nray + 4) +
                            sizeof(double) * (size_t)(4 + t->nmol + t->ncont + t->nray + kMaxJ + 6);
        if (smem > 48 * 1024) return pb_fail(ctx, PB_ERR_UNSUPPORTED, "compute_opacity: too many species for the per-layer shared table");
        // 16-byte vector path: even nwno (every table row and output row is then 16-byte aligned)
        const int C = W * K;
        bool vec2 = (W % 2 == 0) && K == 1;
        for (int k = 0; k < 13; ++k) if (p.o[k] && ((uintptr_t)p.o[k] & 15)) vec2 = false;
        for (int k = 0; k < 3; ++k) if (p.x[k] && ((uintptr_t)p.x[k] & 15)) vec2 = false;
        if (p.pollack && ((uintptr_t)p.pollack & 15)) vec2 = false;
        if (p.cld_opd && (((uintptr_t)p.cld_opd | (uintptr_t)p.cld_w0 | (uintptr_t)p.cld_g0) & 15)) vec2 = false;
        if (vec2) {
            dim3 grid((W / 2 + 127) / 128, L);
            opacity_layer_kernel<2><<<grid, 128, smem, ctx->stream>>>(p);
        } else {
            dim3 grid((C + 127) / 128, L);
            opacity_layer_kernel<1><<<grid, 128, smem, ctx->stream>>>(p);
        }
        PB_CHECK_LAUNCH(ctx);
    }
    if (p.o[1]) {
        opacity_cumsum_kernel<<<(W * K + 127) / 128, 128, 0, ctx->stream>>>(L, W * K, dtau_d, p.o[1]);
        PB_CHECK_LAUNCH(ctx);
    }
    if (p.o[8]) {
        opacity_cumsum_kernel<<<(W * K + 127) / 128, 128, 0, ctx->stream>>>(L, W * K, dtau_og, p.o[8]);
        PB_CHECK_LAUNCH(ctx);
    }
    if (host) {
        for (int k = 0; k < 13; ++k)
            if (outs[k])
                PB_CUDA(ctx, cudaMemcpyAsync(outs[k], p.o[k], (size_t)(L + (is_level[k] ? 1 : 0)) * nC,
                                             cudaMemcpyDeviceToHost, ctx->stream));
        for (int k = 0; k < 3; ++k)
            if (xouts[k]) PB_CUDA(ctx, cudaMemcpyAsync(xouts[k], p.x[k], (size_t)L * nC, cudaMemcpyDeviceToHost, ctx->stream));
        PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return PB_OK;
}
