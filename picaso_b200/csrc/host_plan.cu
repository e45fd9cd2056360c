// host_plan.cu - host-side planning of an opacity query (no CUDA in this file: plain C++ compiled into the library).
//
// Per spectrum the host has to find, for every layer, the four bilinear neighbours of (1/T, log10 P) on the table grid
// and their weights (RetrieveOpacities.find_needed_pts + get_opacities, picaso/optics.py:2048-2123, :2269-2294), the
// nearest CIA temperature (:2298) and the list of touched table rows (:2265).  In numpy that is ~40 small-array calls
// (70 us per spectrum against ~150 us of kernels); here it is one call.  Same comparisons and the same IEEE
// operations in the same order as picaso_b200.optics.find_needed_pts_grid, so the plan is bit-identical
// (tests/test_host_plan_cpu.py).  1/T and log10 P are formed by the caller with numpy: numpy's log10 and libm's may
// differ in the last bit, and a layer exactly on a grid pressure must land on the same side in both.
#include <algorithm>
#include <cstdint>

#include "pb_common.cuh"

namespace {
// numpy index semantics for a possibly negative index into an array of n elements
inline int64_t wrap(int64_t i, int64_t n) { return i < 0 ? i + n : i; }
}

extern "C" int pb_host_plan_bilinear(int nlayer, const double *t_inv, const double *p_log, const double *tlayer,
                                     int nT, const double *t_inv_grid, int nPg, const double *p_log_grid,
                                     const int64_t *nc_p /* [nT] */, const int64_t *row_offset /* [nT + 1] */,
                                     int t_mono, int p_mono, int ncia, const double *cia_unique,
                                     int32_t *idx /* [nlayer][4]: ll, hl, hh, lh */, double *wts /* [nlayer][4] */,
                                     int32_t *cia /* [nlayer] */, int64_t *rows_used /* [4 nlayer], sorted unique, 1-based */,
                                     int *nrows_used)
{
    if (nlayer < 0 || nT < 2 || nPg < 2 || !t_inv || !p_log || !t_inv_grid || !p_log_grid || !nc_p || !row_offset ||
        !idx || !wts)
        return PB_ERR_ARG;
    for (int l = 0; l < nlayer; ++l) {
        const double ti = t_inv[l], pi = p_log[l];
        // last grid temperature strictly below T: last j with t_inv_grid[j] > 1/T (0 if none)
        int64_t t_low = 0;
        if (t_mono) {   // strictly descending grid: count the entries > ti
            int64_t lo = 0, hi = nT;
            while (lo < hi) {
                const int64_t mid = (lo + hi) >> 1;
                if (t_inv_grid[mid] > ti) lo = mid + 1; else hi = mid;
            }
            t_low = lo > 0 ? lo - 1 : 0;
        } else {
            for (int64_t j = nT - 1; j >= 0; --j)
                if (t_inv_grid[j] > ti) { t_low = j; break; }
        }
        if (t_low == nT - 1) t_low = nT - 2;
        const int64_t t_hi = t_low + 1;
        // last grid pressure <= P (0 if none)
        int64_t p_low = 0;
        if (p_mono) {   // strictly ascending grid: count the entries <= pi
            int64_t lo = 0, hi = nPg;
            while (lo < hi) {
                const int64_t mid = (lo + hi) >> 1;
                if (p_log_grid[mid] <= pi) lo = mid + 1; else hi = mid;
            }
            p_low = lo > 0 ? lo - 1 : 0;
        } else {
            for (int64_t j = nPg - 1; j >= 0; --j)
                if (p_log_grid[j] <= pi) { p_low = j; break; }
        }
        if (p_low > nc_p[t_hi] - 3) p_low = nc_p[t_hi] - 3;
        const int64_t p_hi = p_low + 1;
        const double tl = t_inv_grid[wrap(t_low, nT)], th = t_inv_grid[wrap(t_hi, nT)];
        const double pl = p_log_grid[wrap(p_low, nPg)], ph = p_log_grid[wrap(p_hi, nPg)];
        const double t = (ti - tl) / (th - tl);
        const double p = (pi - pl) / (ph - pl);
        const int64_t ot_low = row_offset[wrap(t_low, nT + 1)], ot_hi = row_offset[wrap(t_hi, nT + 1)];
        idx[4 * l + 0] = (int32_t)(ot_low + p_low);   // ll
        idx[4 * l + 1] = (int32_t)(ot_hi + p_low);    // hl
        idx[4 * l + 2] = (int32_t)(ot_hi + p_hi);     // hh
        idx[4 * l + 3] = (int32_t)(ot_low + p_hi);    // lh
        const double t1 = 1 - t, p1 = 1 - p;
        wts[4 * l + 0] = t1 * p1;
        wts[4 * l + 1] = t * p1;
        wts[4 * l + 2] = t * p;
        wts[4 * l + 3] = t1 * p;
        if (cia && ncia > 0 && tlayer) {   // nearest CIA temperature, first minimum (np.argmin)
            int best = 0;
            double bd = cia_unique[0] - tlayer[l];
            bd = bd < 0 ? -bd : bd;
            for (int j = 1; j < ncia; ++j) {
                double d = cia_unique[j] - tlayer[l];
                d = d < 0 ? -d : d;
                if (d < bd) { bd = d; best = j; }
            }
            cia[l] = best;
        }
    }
    if (rows_used && nrows_used) {   // 1 + np.unique(idx)
        const int n = 4 * nlayer;
        for (int i = 0; i < n; ++i) rows_used[i] = (int64_t)idx[i] + 1;
        std::sort(rows_used, rows_used + n);
        *nrows_used = (int)(std::unique(rows_used, rows_used + n) - rows_used);
    }
    return PB_OK;
}
