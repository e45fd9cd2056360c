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
    std::vector<int> cont_nt;
    std::vector<double *> ray;               // [W]
    double *wno = nullptr;                   // [W]
    double *shifts = nullptr;                // [W][ntrans]
    double *raman_c = nullptr, *raman_dnu = nullptr;
    int *raman_ji = nullptr;
    double *RA = nullptr, *RB = nullptr;       // [10][W] per-level Raman sums
    // device-side pointer tables rebuilt when a table changes
    const double **d_mol_raw = nullptr, **d_mol_log = nullptr, **d_cont = nullptr, **d_ray = nullptr;
    bool dirty = true;
    size_t bytes = 0;
};

namespace {

constexpr int kMaxJ = 10;

struct OpaParams {
    int L, W, nmol, ncont, nray, ntrans;
    int query;  // 0 nearest, 1 bilinear
    const double *const *mol_raw, *const *mol_log, *const *cont, *const *ray;
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
    int64_t bs_out;
};

__global__ void log_table_kernel(int64_t n, const double *raw, double *lg)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double a = raw[i];
    lg[i] = log10(a != 0 ? a : 1e-50);  // optics.py:2282
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

// One thread per (layer, wavelength): blockIdx.y = layer, so every per-layer scalar (row indices,
// weights, multipliers) is CTA-uniform; table rows are read as coalesced 256-B segments.
__global__ void __launch_bounds__(256) opacity_layer_kernel(OpaParams p)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    const int l = blockIdx.y;
    if (w >= p.W) return;
    const int L = p.L, W = p.W;
    const double N_A = 6.02214086e+23;
    double taugas = 0.0;
    // continuum (optics.py:172-233): table row of the nearest CIA temperature x layer factor
    const int64_t crow = (int64_t)p.cont_index[l] * W + w;
    for (int c = 0; c < p.ncont; ++c) taugas += __ldg(p.cont[c] + crow) * p.cont_scale[c * L + l];
    // molecular (optics.py:243-250)
    if (p.query == 1) {
        const int *ix = p.pt_index + 4 * l;
        const double w1 = p.wts[4 * l], w2 = p.wts[4 * l + 1], w3 = p.wts[4 * l + 2], w4 = p.wts[4 * l + 3];
        const int64_t r1 = (int64_t)ix[0] * W + w, r2 = (int64_t)ix[1] * W + w, r3 = (int64_t)ix[2] * W + w,
                      r4 = (int64_t)ix[3] * W + w;
        for (int m = 0; m < p.nmol; ++m) {
            const double *t = p.mol_log[m];
            // 10**((1-t)(1-p) l1 + t(1-p) l2 + t p l3 + (1-t) p l4), optics.py:2290-2293
            const double e = ((w1 * __ldg(t + r1)) + (w2 * __ldg(t + r2)) + (w3 * __ldg(t + r3)) +
                              (w4 * __ldg(t + r4)));
            taugas += (exp10(e) * N_A) * p.mol_scale[m * L + l];
        }
    } else {
        const int64_t r1 = (int64_t)p.pt_index[4 * l] * W + w;
        for (int m = 0; m < p.nmol; ++m) taugas += (__ldg(p.mol_raw[m] + r1) * N_A) * p.mol_scale[m * L + l];
    }
    // Rayleigh (optics.py:265-271)
    double tauray = 0.0;
    for (int m = 0; m < p.nray; ++m) tauray += __ldg(p.ray[m] + w) * p.ray_scale[m * L + l];
    // Raman factor (optics.py:287-306), capped at 0.99999
    double rf = 0.99999;
    if (p.raman == 0) {
        double num = 0.0, den = 0.0;
#pragma unroll
        for (int j = 0; j < kMaxJ; ++j) {
            const double f = p.jfrac[j * L + l];
            num = fma(f, __ldg(p.RA + (int64_t)j * W + w), num);
            den = fma(f, __ldg(p.RB + (int64_t)j * W + w), den);
        }
        rf = fmin(num / den, 0.99999);
    } else if (p.raman == 1) {
        rf = fmin(p.pollack[w], 0.99999);
    }
    // cloud (optics.py:309-315)
    double taucld = 0.0, w0c = 0.0, g0 = 0.0;
    if (p.cld_opd) {
        const int64_t ic = (int64_t)l * p.ld + w;
        taucld = __ldg(p.cld_opd + ic);
        w0c = __ldg(p.cld_w0 + ic);
        g0 = __ldg(p.cld_g0 + ic);
        if (p.do_holes) taucld = p.fthin * taucld;
    }
    // totals (optics.py:329-350)
    const double dtau = taugas + tauray + taucld;
    const double sc = w0c * taucld;
    const double w0 = (tauray * rf + taucld * w0c) / dtau;
    const int64_t io = (int64_t)l * W + w;
    if (p.o[4]) p.o[4][io] = sc / (sc + tauray);
    const double fray = tauray / (tauray + sc);
    if (p.o[5]) p.o[5][io] = fray;
    if (p.o[6]) p.o[6][io] = 0.5 * fray;
    if (p.o[7]) p.o[7][io] = dtau;
    if (p.o[9]) p.o[9][io] = w0;
    if (p.o[10]) p.o[10][io] = g0;
    if (p.o[11]) p.o[11][io] = (tauray * 0.99999 + taucld * w0c) / dtau;
    if (p.dedd) {
        // delta-Eddington (optics.py:412-420)
        double f = 1.0;
        for (int s = 0; s < p.stream; ++s) f *= g0;
        if (p.o[0]) p.o[0][io] = dtau * (1. - w0 * f);
        if (p.o[2]) p.o[2][io] = w0 * (1. - f) / (1.0 - w0 * f);
        if (p.o[3]) p.o[3][io] = (g0 - f) / (1. - f);
        if (p.o[12]) p.o[12][io] = f;
    } else {
        if (p.o[0]) p.o[0][io] = dtau;
        if (p.o[2]) p.o[2][io] = w0;
        if (p.o[3]) p.o[3][io] = g0;
        if (p.o[12]) p.o[12][io] = 0 * g0;
    }
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
    t->cont.assign(ncont, nullptr); t->cont_nt.assign(ncont, 0);
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
    PB_TRY(upload_table(ctx, table, (size_t)ntemp * t->W, &t->cont[icont], &t->bytes));
    t->cont_nt[icont] = ntemp;
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
    if (L < 1) return pb_fail(ctx, PB_ERR_ARG, "compute_opacity: nlayer < 1");
    if (a->query != 0 && a->query != 1) return pb_fail(ctx, PB_ERR_ARG, "compute_opacity: query must be 0 (nearest) or 1 (bilinear)");
    if (a->raman < 0 || a->raman > 2) return pb_fail(ctx, PB_ERR_ARG, "compute_opacity: raman must be 0, 1 or 2");
    if (a->stream != 2 && a->stream != 4) return pb_fail(ctx, PB_ERR_ARG, "compute_opacity: stream must be 2 or 4");
    if (!a->pt_index || !a->cont_index || (t->nmol && !a->mol_scale) || (t->ncont && !a->cont_scale) || (t->nray && !a->ray_scale))
        return pb_fail(ctx, PB_ERR_ARG, "compute_opacity: NULL per-layer vector");
    if (a->query == 1 && !a->weights) return pb_fail(ctx, PB_ERR_ARG, "compute_opacity: bilinear query needs weights");
    if (a->raman == 0 && (!t->shifts || !a->jfrac)) return pb_fail(ctx, PB_ERR_ARG, "compute_opacity: raman=0 needs pb_optab_set_raman and jfrac");
    if (a->raman == 1 && !a->raman_pollack) return pb_fail(ctx, PB_ERR_ARG, "compute_opacity: raman=1 needs raman_pollack[nwno]");
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
            if (a->cont_index[l] < 0 || a->cont_index[l] >= t->cont_nt[c]) return pb_fail(ctx, PB_ERR_ARG, "compute_opacity: cont_index[%d] out of range", l);
    }
    for (int m = 0; m < t->nray; ++m)
        if (!t->ray[m]) return pb_fail(ctx, PB_ERR_ARG, "compute_opacity: rayleigh table %d not set", m);
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    PB_TRY(sync_pointer_tables(ctx, t));
    const bool host = memspace == PB_HOST;
    const size_t nW = (size_t)W * sizeof(double);
    double *const outs[13] = {a->DTAU, a->TAU, a->W0, a->COSB, a->ftau_cld, a->ftau_ray, a->GCOS2, a->DTAU_OG,
                              a->TAU_OG, a->W0_OG, a->COSB_OG, a->W0_no_raman, a->f_deltaM};
    const bool is_level[13] = {false, true, false, false, false, false, false, false, true, false, false, false, false};
    size_t need = 64 * 256 + 2 * pb_align((size_t)L * nW) + pb_align(4 * (size_t)L * 12) + pb_align((size_t)L * 8) +
                  (size_t)(t->nmol + t->ncont + t->nray + kMaxJ + 4) * pb_align((size_t)L * 8);
    if (host) {
        if (cloud) need += 3 * pb_align((size_t)L * nW);
        if (a->raman == 1) need += pb_align(nW);
        for (int k = 0; k < 13; ++k)
            if (outs[k]) need += pb_align((size_t)(L + (is_level[k] ? 1 : 0)) * nW);
    }
    pb_arena_reset(ctx);
    PB_TRY(pb_arena_reserve(ctx, need));
    PB_TRY(pb_pinned_reserve(ctx, (size_t)(t->nmol + t->ncont + t->nray + kMaxJ + 16) * ((size_t)L * 8 + 64)));
    OpaParams p;
    memset(&p, 0, sizeof(p));
    p.L = L; p.W = W; p.nmol = t->nmol; p.ncont = t->ncont; p.nray = t->nray; p.ntrans = t->ntrans;
    p.query = a->query;
    p.mol_raw = t->d_mol_raw; p.mol_log = t->d_mol_log; p.cont = t->d_cont; p.ray = t->d_ray;
    const double *tmp;
    // ints travel through the same pinned bounce as doubles (sizes rounded up to 8 bytes)
    PB_TRY(pb_upload_small(ctx, (const double *)a->pt_index, ((size_t)4 * L * sizeof(int) + 7) / 8, &tmp));
    p.pt_index = (const int *)tmp;
    {
        std::vector<int> ci(a->cont_index, a->cont_index + L);
        if (ci.size() & 1) ci.push_back(0);
        PB_TRY(pb_upload_small(ctx, (const double *)ci.data(), ci.size() / 2, &tmp));
        p.cont_index = (const int *)tmp;
    }
    if (a->query == 1) PB_TRY(pb_upload_small(ctx, a->weights, (size_t)4 * L, &p.wts));
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
    for (int k = 0; k < 13; ++k) {
        p.o[k] = nullptr;
        if (!outs[k]) continue;
        if (host) PB_TRY(pb_arena_alloc(ctx, (size_t)(L + (is_level[k] ? 1 : 0)) * nW, (void **)&p.o[k]));
        else p.o[k] = outs[k];
    }
    PB_TRY(pb_upload_flush(ctx));
    // the running optical depths need the per-layer values even if the caller did not ask for them
    double *dtau_d = p.o[0], *dtau_og = p.o[7];
    if (p.o[1] && !dtau_d) { PB_TRY(pb_arena_alloc(ctx, (size_t)L * nW, (void **)&dtau_d)); p.o[0] = dtau_d; }
    if (p.o[8] && !dtau_og) { PB_TRY(pb_arena_alloc(ctx, (size_t)L * nW, (void **)&dtau_og)); p.o[7] = dtau_og; }
    dim3 grid((W + 255) / 256, L);
    opacity_layer_kernel<<<grid, 256, 0, ctx->stream>>>(p);
    PB_CHECK_LAUNCH(ctx);
    if (p.o[1]) {
        opacity_cumsum_kernel<<<(W + 127) / 128, 128, 0, ctx->stream>>>(L, W, dtau_d, p.o[1]);
        PB_CHECK_LAUNCH(ctx);
    }
    if (p.o[8]) {
        opacity_cumsum_kernel<<<(W + 127) / 128, 128, 0, ctx->stream>>>(L, W, dtau_og, p.o[8]);
        PB_CHECK_LAUNCH(ctx);
    }
    if (host) {
        for (int k = 0; k < 13; ++k)
            if (outs[k])
                PB_CUDA(ctx, cudaMemcpyAsync(outs[k], p.o[k], (size_t)(L + (is_level[k] ? 1 : 0)) * nW,
                                             cudaMemcpyDeviceToHost, ctx->stream));
        PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return PB_OK;
}
