"""CPU (numpy) restatement of the reference's resort-rebin k-coefficient mixing - TEST
INFRASTRUCTURE ONLY.  Follows /root/reference/picaso/deq_chem.py:334-386 (mix_all_gases_gasesfly),
:388-432 (do_mixing_mono_gasesfly), :538-597 (mix_2_gases) and the interpolation that follows in
optics.RetrieveCKs.mix_my_opacities_gasesfly (optics.py:1164-1197).  Pinned against the unmodified
reference by tests/golden/make_golden_mix.py."""
import numpy as np

N_A = 6.02214086e+23


def mix_2_gases(k1, k2, mix1, mix2, gauss_pts, gauss_wts):
    """deq_chem.py:565-597: random-overlap mix of two Nk-point k-distributions."""
    mix_t = mix1 + mix2
    Nk = len(gauss_wts)
    kmix = ((mix1 * np.asarray(k1)[:, None] + mix2 * np.asarray(k2)[None, :]) / mix_t).reshape(Nk * Nk)
    wts = (np.asarray(gauss_wts)[:, None] * np.asarray(gauss_wts)[None, :]).reshape(Nk * Nk)
    order = np.argsort(kmix, kind="mergesort")
    ks, ws = kmix[order], wts[order]
    csum = np.cumsum(ws)
    x = csum / np.max(csum)
    return 10 ** np.interp(gauss_pts, x, np.log10(ks)), mix_t


def do_mixing(ln_kappas, mixes, gauss_pts, gauss_wts):
    """deq_chem.py:424-432: fold the gases in pairwise, in list order."""
    kmix = np.exp(ln_kappas[0])
    mix_t = mixes[0]
    for i in range(1, len(ln_kappas)):
        kmix, mix_t = mix_2_gases(kmix, np.exp(ln_kappas[i]), mix_t, mixes[i], gauss_pts, gauss_wts)
    return kmix


def mix_all_gases(kappas, mixes, gauss_pts, gauss_wts, indices):
    """deq_chem.py:360-386: kappas[g] = ln kappa [nP, nT, W, Nk]; mixes[g][L]; indices = (p_low, p_hi,
    t_low, t_hi) [4][L].  Returns ln kappa_mixed [L, W, Nk, 4] (neighbour order (pl,tl),(pl,th),(ph,tl),(ph,th))."""
    Nk = len(gauss_wts)
    L = len(indices[0])
    W = kappas[0].shape[2]
    out = np.zeros((L, W, Nk, 4))
    for il in range(L):
        ct = 0
        for p_ind in (indices[0][il], indices[1][il]):
            for t_ind in (indices[2][il], indices[3][il]):
                for iw in range(W):
                    out[il, iw, :, ct] = do_mixing([k[p_ind, t_ind, iw, :] for k in kappas],
                                                   [m[il] for m in mixes], gauss_pts, gauss_wts)
                ct += 1
    return np.log(out)


def interpolate_mixed(ln_mixed, t_interp, p_interp):
    """optics.py:1189-1197: bilinear in (1/T, log10 P) of ln kappa, then exp * N_A -> [L, W, Nk]."""
    t = np.asarray(t_interp)[:, None, None]
    p = np.asarray(p_interp)[:, None, None]
    kappa = (((1 - t) * (1 - p) * ln_mixed[..., 0]) + ((t) * (1 - p) * ln_mixed[..., 1]) +
             ((t) * (p) * ln_mixed[..., 3]) + ((1 - t) * (p) * ln_mixed[..., 2]))
    return np.exp(kappa) * N_A
