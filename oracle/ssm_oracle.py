"""CPU oracle: a numpy restatement of the reference's hot path (jacobnzw/SSMToybox v0.1.1a0).

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module, and only as the checker / CPU baseline.  The product
package (ssmtoybox_b200/) never imports it: the product path is CUDA-only and fails loudly when
the extension is missing.

Parity status: PINNED.  The oracle is checked (tests/test_oracle_golden.py) against golden vectors
produced by running the unmodified reference in the build container (oracle/gen_golden.py ->
tests/golden/*.npz): filtered / predictive / smoothed moments for every filter family, BQ weights
and kernel expectations, point sets, simulators and scores.

All citations are file:line relative to /root/reference/ssmtoybox/.

Two linear-algebra back-ends:
  'lapack' : numpy.linalg.cholesky + scipy cho_factor/cho_solve, one trajectory at a time, the
             same calls in the same order as the reference (float64 only);
  'loops'  : explicit Cholesky / triangular solves vectorised over a leading trajectory axis and
             dtype-generic (float64 or longdouble).  The longdouble run is the arbiter for the
             un-centred BQ covariances whose float64 noise floor exceeds 1e-9.
"""
import math

import numpy as np
from numpy.polynomial.hermite_e import hermegauss, hermeval
from scipy.linalg import cho_factor, cho_solve, solve as sp_solve
from scipy.special import factorial


# ==============================================================================================
# point sets and classical weights                                              mtran.py:166-578
# ==============================================================================================
def ut_points(dim, kappa=None, alpha=1.0):
    """Unscented unit points [0, c I, -c I], c = sqrt(dim + lambda).            mtran.py:235-258"""
    kappa = max(3.0 - dim, 0.0) if kappa is None else kappa
    lam = alpha ** 2 * (dim + kappa) - dim
    c = np.sqrt(dim + lam)
    return np.hstack((np.zeros((dim, 1)), c * np.eye(dim), -c * np.eye(dim)))


def ut_weights(dim, kappa=None, alpha=1.0, beta=2.0):
    """Unscented mean / covariance weights.                                      mtran.py:261-293"""
    kappa = max(3.0 - dim, 0.0) if kappa is None else kappa
    lam = alpha ** 2 * (dim + kappa) - dim
    wm = 1.0 / (2.0 * (dim + lam)) * np.ones(2 * dim + 1)
    wc = wm.copy()
    wm[0] = lam / (dim + lam)
    wc[0] = wm[0] + (1 - alpha ** 2 + beta)
    return wm, wc


def sr_points(dim):
    """Spherical-radial unit points [c I, -c I], c = sqrt(dim).                 mtran.py:189-204"""
    c = np.sqrt(dim)
    return np.hstack((c * np.eye(dim), -c * np.eye(dim)))


def sr_weights(dim):
    """mtran.py:172-186"""
    return (1 / (2.0 * dim)) * np.ones(2 * dim)


def _cartesian(arrays):
    """Cartesian product, first array slowest (sklearn.utils.extmath.cartesian, used at mtran.py:337,360)."""
    grids = np.meshgrid(*arrays, indexing='ij')
    return np.stack([g.reshape(-1) for g in grids], axis=1)


def gh_points(dim, degree=3):
    """Gauss-Hermite unit points, cartesian product of the 1-D nodes.           mtran.py:340-360"""
    x, _ = hermegauss(degree)
    return _cartesian([x] * dim).T


def gh_weights(dim, degree=3):
    """GH weights recomputed as n!/(n^2 He_{n-1}(x)^2), NOT hermegauss' weights.  mtran.py:316-337"""
    x, _ = hermegauss(degree)
    w = factorial(degree) / (degree ** 2 * hermeval(x, [0] * (degree - 1) + [1]) ** 2)
    return np.prod(_cartesian([w] * dim), axis=1)


def fs_weights(dim, degree=3, kappa=None, dof=4.0):
    """Fully-symmetric Student rule weights, degrees 3 and 5.                   mtran.py:406-463"""
    if degree not in (3, 5):
        degree = 3
    kappa = max(3.0 - dim, 0.0) if kappa is None else kappa
    dof = max(dof, degree)
    if degree == 3:
        w = 1 / (2 * (dim + kappa)) * np.ones(2 * dim + 1)
        w[0] = kappa / (dim + kappa)
        return w
    i2 = dof / (dof - 2)
    i22 = dof ** 2 / ((dof - 2) * (dof - 4))
    i4 = 3 * i22
    a0 = 1 - dim * (i2 / i4) ** 2 * (i4 - 0.5 * (dim - 1) * i22)
    a1 = 0.5 * (i2 / i4) ** 2 * (i4 - (dim - 1) * i22)
    a11 = 0.25 * (i2 / i4) ** 2 * i22
    return np.hstack((a0, a1 * np.ones(2 * dim), a11 * np.ones(2 * dim * (dim - 1))))


def fs_points(dim, degree=3, kappa=None, dof=4.0):
    """Fully-symmetric Student unit points.                                     mtran.py:466-578
    Degree 5 = centre, +-u e_i, and (+-u e_i +- u e_j), i<j, enumerated in the order of the
    reference's recursive symmetric_set(): for i, for j>i, for s in (+,-): v=(u e_i + s u e_j), -v."""
    if degree not in (3, 5):
        degree = 3
    kappa = max(3.0 - dim, 0.0) if kappa is None else kappa
    dof = max(dof, degree)
    eye = np.eye(dim)
    if degree == 3:
        u = np.sqrt(dof / (dof - 2) * (dim + kappa))
        return u * np.hstack((np.zeros((dim, 1)), eye, -eye))
    i2 = dof / (dof - 2)
    i4 = 3 * dof ** 2 / ((dof - 2) * (dof - 4))
    u = np.sqrt(i4 / i2)
    cols = [np.zeros(dim)]
    for i in range(dim):
        cols += [u * eye[i], -u * eye[i]]
    for i in range(dim):
        for j in range(i + 1, dim):
            for s in (1.0, -1.0):
                v = u * eye[i] + s * u * eye[j]
                cols += [v, -v]
    return np.stack(cols, axis=1)


def classical_rule(points, dim, **kw):
    """(unit points, wm, Wc=diag) of the sigma-point transforms built at ssinf.py:360-402, 770-775."""
    if points == 'ut':
        wm, wc = ut_weights(dim, kw.get('kappa'), kw.get('alpha', 1.0), kw.get('beta', 2.0))
        return ut_points(dim, kw.get('kappa'), kw.get('alpha', 1.0)), wm, np.diag(wc)
    if points == 'sr':
        w = sr_weights(dim)
        return sr_points(dim), w, np.diag(w)
    if points == 'gh':
        w = gh_weights(dim, kw.get('degree', 3))
        return gh_points(dim, kw.get('degree', 3)), w, np.diag(w)
    if points == 'fs':
        a = (dim, kw.get('degree', 3), kw.get('kappa'), kw.get('dof', 4.0))
        w = fs_weights(*a)
        return fs_points(*a), w, np.diag(w)
    raise ValueError(points)


# ==============================================================================================
# RBF kernel and its Gaussian expectations                                    bq/bqkern.py:295-454
# ==============================================================================================
def _maha(x, y, V=None):
    """Pairwise Mahalanobis distances of rows of x and y.                        utils.py:385-409"""
    if V is None:
        V = np.eye(x.shape[1])
    x2 = np.sum(x.dot(V) * x, 1)
    y2 = np.sum(y.dot(V) * y, 1)
    return (x2[:, None] + y2[:, None].T) - 2 * x.dot(V).dot(y.T)


def _unpack(par):
    par = np.asarray(par, dtype=float).squeeze()
    return par[0], np.diag(par[1:] ** -1)  # alpha, Lambda^{-1/2}                bqkern.py:438-454


def rbf_eval(par, x1, x2=None, scaling=True):
    """K_ij = alpha^2 exp(-0.5 |Lambda^-1/2 (x_i - x_j)|^2), evaluated in log-space.  bqkern.py:329-343"""
    x2 = x1.copy() if x2 is None else x2
    alpha, sil = _unpack(par)
    alpha = alpha if scaling else 1.0
    return np.exp(2 * np.log(alpha) - 0.5 * _maha(sil.dot(x1).T, sil.dot(x2).T))


def rbf_inv(par, x, jitter=1e-8, scaling=False):
    """inv(K + jitter I) by Cholesky, then symmetrised.               bqkern.py:38-64, 96-120"""
    n = x.shape[1]
    iA = cho_solve(cho_factor(rbf_eval(par, x, scaling=scaling) + jitter * np.eye(n)), np.eye(n))
    return 0.5 * (iA + iA.T)


def rbf_exp_x_kx(par, x, scaling=False):
    """q_i = E_x[k(x, x_i)], x ~ N(0, I).                                       bqkern.py:345-356"""
    alpha, sil = _unpack(par)
    alpha = alpha if scaling else 1.0
    d = x.shape[0]
    inv_lam = sil ** 2
    lam = np.diag(inv_lam.diagonal() ** -1)
    c = alpha ** 2 * np.linalg.det(inv_lam + np.eye(d)) ** -0.5
    xl = np.linalg.inv(lam + np.eye(d)).dot(x)
    return c * np.exp(-0.5 * np.sum(x * xl, axis=0))


def rbf_exp_x_xkx(par, x):
    """R = E_x[x k(x, x_i)] = q_i (Lambda + I)^-1 x_i.                          bqkern.py:358-364"""
    _, sil = _unpack(par)
    lam = np.diag(sil.diagonal() ** -2)
    mu = np.linalg.inv(lam + np.eye(x.shape[0])).dot(x)
    return rbf_exp_x_kx(par, x)[None, :] * mu


def rbf_exp_x_kxkx(par0, par1, x, scaling=False):
    """Q_ij = E_x[k(x, x_i) k(x, x_j)].                                         bqkern.py:366-415"""
    a0, sil0 = _unpack(par0)
    a1, sil1 = _unpack(par1)
    if not scaling:
        a0 = a1 = 1.0
    il0, il1 = sil0 ** 2, sil1 ** 2
    xi0 = sil0.dot(x)
    xi0 = 2 * np.log(a0) - 0.5 * np.sum(xi0 * xi0, axis=0)
    xi1 = sil1.dot(x)
    xi1 = 2 * np.log(a1) - 0.5 * np.sum(xi1 * xi1, axis=0)
    x0, x1 = il0.dot(x), il1.dot(x)
    r = il0 + il1 + np.eye(x.shape[0])
    n = (xi0[:, None] + xi1[None, :]) + 0.5 * _maha(x0.T, -x1.T, V=np.linalg.inv(r))
    return np.linalg.det(r) ** -0.5 * np.exp(n)


def rbf_exp_xy_kxy(par):
    """kbar = alpha^2 |2 Lambda^-1 + I|^-1/2.                                    bqkern.py:421-424"""
    alpha, sil = _unpack(par)
    return alpha ** 2 * np.linalg.det(2 * sil ** 2 + np.eye(sil.shape[0])) ** -0.5


# ----------------------------------------------------------------------------------------------
# RBF kernel expectations under a standard Student-t density               bq/bqkern.py:457-536
# ----------------------------------------------------------------------------------------------
def _gamma_mix_nodes(dof, n=400):
    """Nodes / weights of the Gamma(dof/2, scale 2/dof) mixing variable u of x = z / sqrt(u), z ~ N(0, I)
    (utils.py:349-382): Gauss-Legendre in the probability variable is inaccurate at the u -> 0 end, so the
    integral over u in (0, inf) is done as Gauss-Legendre in t = log u on a range that leaves < 1e-18 outside."""
    from scipy.special import gammaln
    k, th = 0.5 * dof, 2.0 / dof
    t, w = np.polynomial.legendre.leggauss(n)
    # log-density in t is k (t - e^t) + const for the unit-mean Gamma: 40 below its maximum at both ends
    lo, hi = -(40.0 / k + np.sqrt(80.0 / k)), np.log1p(40.0 / k + np.sqrt(80.0 / k))
    t = 0.5 * (hi - lo) * t + 0.5 * (hi + lo)
    w = 0.5 * (hi - lo) * w
    u = np.exp(t)
    logpdf = (k - 1) * t - u / th - gammaln(k) - k * np.log(th)
    return u, w * np.exp(logpdf + t)


def rbf_student_expectations(par, x, dof, n=400):
    """What RBFStudent estimates by 2 * 10^6-sample Monte Carlo (bq/bqkern.py:475-536), computed to quadrature
    accuracy: a standard Student-t variable is a Gaussian scale mixture, x | u ~ N(0, I / u), and under a Gaussian the
    RBF expectations are closed-form (bqkern.py:345-424 with I replaced by I / u), so each expectation is a 1-D
    integral over u (2-D for the pair expectation).  Unscaled kernel (scaling=False, as bq_weights calls it,
    bqmod.py:508-511).  Returns q (N), R (D, N), Q (N, N), kbar = E[k(x, x')] for independent x, x'."""
    par = np.atleast_2d(np.asarray(par, dtype=float))
    l2 = par[0, 1:] ** 2                                        # squared lengthscales (D,)
    u, w = _gamma_mix_nodes(dof, n)
    s2 = 1.0 / u                                                # conditional variance
    D, N = x.shape
    # q_i(u) = prod_d (1 + s2 / l2_d)^-1/2 exp(-1/2 x_id^2 / (l2_d + s2))
    den = l2[None, :] + s2[:, None]                             # (n, D)
    c1 = np.prod(l2[None, :] / den, axis=1) ** 0.5              # (n,)
    qi = c1[:, None] * np.exp(-0.5 * np.einsum('di,nd->ni', x ** 2, 1.0 / den))
    q = w.dot(qi)
    # R_di(u) = q_i(u) x_id s2 / (l2_d + s2)
    R = np.einsum('n,ni,nd,di->di', w, qi, s2[:, None] / den, x)
    # Q_ij(u) = exp(-|x_i - x_j|^2_l / 4) prod_d (1 + 2 s2 / l2_d)^-1/2 exp(-1/2 c_d^2 / (l2_d / 2 + s2)), c = (x_i + x_j) / 2
    diff = x[:, :, None] - x[:, None, :]
    cen = 0.5 * (x[:, :, None] + x[:, None, :])
    pref = np.exp(-0.25 * np.einsum('dij,d->ij', diff ** 2, 1.0 / l2))
    den2 = 0.5 * l2[None, :] + s2[:, None]
    c2 = np.prod(0.5 * l2[None, :] / den2, axis=1) ** 0.5
    Qu = c2[:, None, None] * np.exp(-0.5 * np.einsum('dij,nd->nij', cen ** 2, 1.0 / den2))
    Q = pref * np.einsum('n,nij->ij', w, Qu)
    # kbar: x - x' | u, u' ~ N(0, (1/u + 1/u') I)
    ss = s2[:, None] + s2[None, :]
    kb = np.prod(l2[None, None, :] / (l2[None, None, :] + ss[:, :, None]), axis=2) ** 0.5
    kbar = w.dot(kb).dot(w)
    return q, R, Q, float(kbar)


def rbf_student_exp_xy_kxy_reference(par, kbar, num_samples=2e6):
    """The value RBFStudent.exp_xy_kxy converges to (bq/bqkern.py:527-535): it sums the FULL 200 x 200 kernel matrix
    of each of its 10^4 batches (diagonal included, scaling on) and divides by num_samples instead of by the number
    of pairs, i.e. it returns (2e6 / num_samples) alpha^2 (199 kbar + 1), not kbar."""
    alpha2 = float(np.atleast_2d(par)[0, 0]) ** 2
    batch = int(2e6 // 10000)
    return 2e6 / float(num_samples) * alpha2 * ((batch - 1) * kbar + 1.0)


def student_bq_weights(par, points, dof=4.0):
    """bq_weights (bqmod.py:495-523) with the RBFStudent expectations in the limit of infinitely many samples."""
    par = np.atleast_2d(np.asarray(par, dtype=float))
    p1 = par.copy()
    p1[0, 0] = 1.0
    iK = rbf_inv(p1, points)
    q, R, Q, kbar = rbf_student_expectations(par, points, dof)
    Wc = iK.dot(Q).dot(iK)
    if not np.array_equal(Wc, Wc.T):
        Wc = 0.5 * (Wc + Wc.T)
    return dict(wm=q.dot(iK), Wc=Wc, Wcc=R.dot(iK), iK=iK, q=q, Q=Q, R=R, model_var=par[0, 0] ** 2 * (1 - np.trace(Q.dot(iK))),
                integral_var=rbf_student_exp_xy_kxy_reference(par, kbar) - q.dot(iK).dot(q))


# ==============================================================================================
# Bayesian-quadrature weights                                      bq/bqmod.py:495-523, 893-992
# ==============================================================================================
def gp_weights(par, points):
    """GPQ (and TPQ) weights: wm = q iK, Wc = sym(iK Q iK), Wcc = R iK.          bqmod.py:495-523"""
    par = np.atleast_2d(np.asarray(par, dtype=float))
    x = points
    iK = rbf_inv(par, x)
    q = rbf_exp_x_kx(par, x)
    Q = rbf_exp_x_kxkx(par, par, x)
    R = rbf_exp_x_xkx(par, x)
    wm = q.dot(iK)
    Wc = iK.dot(Q).dot(iK)
    Wcc = R.dot(iK)
    alpha = _unpack(par)[0]
    model_var = alpha ** 2 * (1 - np.trace(Q.dot(iK)))
    integral_var = rbf_exp_xy_kxy(par) - q.T.dot(iK).dot(q)
    if not np.array_equal(Wc, Wc.T):
        Wc = 0.5 * (Wc + Wc.T)
    return dict(wm=wm, Wc=Wc, Wcc=Wcc, model_var=model_var, integral_var=integral_var, iK=iK, q=q, Q=Q, R=R)


def rbf_der_par(par, x):
    """dK/dalpha = 2 K / alpha and, for each lengthscale, K_ij (x_di - x_dj)^2 / l_d^2, as written in
    RBFGauss.der_par.                                                          bqkern.py:426-436"""
    par = np.asarray(par, dtype=float).squeeze()
    alpha, el = par[0], par[1:]
    K = rbf_eval(par, x)
    d_alpha = 2 * alpha ** -1 * K
    d_el = (x[:, None, :] - x[:, :, None]) ** 2 * (el ** -2)[:, None, None] * K[None, :, :]
    return np.concatenate((d_alpha[..., None], d_el.T), axis=2)


def gp_nlml(log_par, fcn_obs, x_obs, jitter, nu=None):
    """Negative log marginal likelihood and gradient of the GP (nu None; bqmod.py:537-596) or Student-t process
    (bqmod.py:1191-1245) regression model.  fcn_obs (N, E), x_obs (D, N), jitter (N, N)."""
    from scipy.special import gammaln
    par = np.exp(np.asarray(log_par, dtype=float))
    N, E = fcn_obs.shape
    K = rbf_eval(par, x_obs) + jitter
    L = cho_factor(K)
    a = cho_solve(L, fcn_obs)
    yda = np.einsum('ij,ij->j', fcn_obs, a)
    hl = np.sum(np.log(np.diag(L[0])))
    dK = rbf_der_par(par, x_obs)
    iK = cho_solve(L, np.eye(N))
    if nu is None:
        nlml = E * hl + 0.5 * (yda.sum() + E * N * np.log(2 * np.pi))
        aoa = a.dot(a.T)
    else:
        const = (N / 2) * np.log((nu - 2) * np.pi) - gammaln((nu + N) / 2) + gammaln(nu / 2)
        nlml = 0.5 * (nu + N) * np.log(1 + yda / (nu - 2)).sum() + E * (hl + const)
        aoa = (a * ((nu + N) / (nu + yda - 2))[None, :]).dot(a.T)
    grad = 0.5 * np.einsum('ij,jip->p', E * iK - aoa, dK)
    return nlml, grad


def _fact2(n):
    """Double factorial with (-1)!! = 0!! = 1 (the convention bqmod.py:656-661 relies on)."""
    n = int(n)
    return 1 if n <= 0 else math.prod(range(n, 0, -2))


def vandermonde(mulind, x):
    """V[n, b] = prod_d x[d, n] ** mulind[d, b].                                 utils.py:478-502"""
    return np.prod(x[:, :, None] ** mulind[:, None, :], axis=0)


def _exp_px(mi):
    """E[p_q(x)] = prod_d (a_d - 1)!! if all a_d even else 0.                   bqmod.py:635-661"""
    return np.array([np.prod([_fact2(a - 1) for a in mi[:, q]]) if np.all(mi[:, q] % 2 == 0) else 0.0
                     for q in range(mi.shape[1])], dtype=float)


def _exp_xpx(mi):
    """E[x_e p_q(x)].                                                           bqmod.py:663-697"""
    dim, nb = mi.shape
    out = np.zeros((dim, nb))
    for d in range(dim):
        for q in range(nb):
            rest = np.delete(mi[:, q], d)
            if (mi[d, q] + 1) % 2 == 0 and np.all(rest % 2 == 0):
                out[d, q] = mi[d, q] * np.prod([_fact2(a - 1) for a in rest])
    return out


def _exp_pxpx(mi):
    """E[p_r(x) p_q(x)].                                                        bqmod.py:699-731"""
    nb = mi.shape[1]
    out = np.zeros((nb, nb))
    for r in range(nb):
        for q in range(nb):
            s = mi[:, r] + mi[:, q]
            if np.all(s % 2 == 0):
                out[r, q] = np.prod([_fact2(a - 1) for a in s])
    return out


def _exp_kxpx(par, mi, x):
    """E[k(x, x_n) p_q(x)] by the binomial closed form.                         bqmod.py:733-797"""
    dim, nb = mi.shape
    n_pts = x.shape[1]
    # NOTE reference quirk: "ell" is diag(Lambda^-1/2) ** -2, i.e. the SQUARED length-scale, and is
    # squared again below (bqmod.py:770-771, 781-786); reproduced as is.
    ell = (np.asarray(par, dtype=float).squeeze()[1:] ** -1) ** -2
    out = np.zeros((n_pts, nb))
    for n in range(n_pts):
        for q in range(nb):
            t = np.zeros(dim)
            for d in range(dim):
                a = int(mi[d, q])
                e = ell[d] * (1 + ell[d] ** 2) ** (-(1 + a) / 2) * np.exp(-x[d, n] ** 2 / (2 * (1 + ell[d] ** 2)))
                b = 0
                for m in range(a // 2 + 1):
                    p1 = math.factorial(a) / ((2 ** m) * math.factorial(m) * math.factorial(a - 2 * m))
                    p2 = (ell[d] ** (2 * m)) * ((x[d, n] / np.sqrt(1 + ell[d] ** 2)) ** (a - 2 * m))
                    b += p1 * p2
                t[d] = e * b
            out[n, q] = np.prod(t)
    return out


def bs_weights(par, points, mulind):
    """Bayes-Sard weights (polynomial prior mean).                              bqmod.py:893-992"""
    par = np.atleast_2d(np.asarray(par, dtype=float))
    mi = np.asarray(mulind)
    x = points
    nb, n_pts = mi.shape[1], x.shape[1]
    iK = rbf_inv(par, x)
    V = vandermonde(mi, x)
    iViKV = cho_solve(cho_factor(V.T.dot(iK).dot(V) + 1e-8 * np.eye(nb)), np.eye(nb))
    px, xpx, pxpx, kxpx = _exp_px(mi), _exp_xpx(mi), _exp_pxpx(mi), _exp_kxpx(par, mi, x)
    q = rbf_exp_x_kx(par, x)
    kxy = rbf_exp_xy_kxy(par)
    a2 = _unpack(par)[0] ** 2
    if nb == n_pts:  # pi-unisolvent special case: classical rule through the inverse Vandermonde
        iV = sp_solve(V, np.eye(nb))
        wm = iV.T.dot(px)
        Wc = iV.T.dot(pxpx).dot(iV)
        Wcc = xpx.dot(iV)
        model_var = a2 * (1 - np.trace(kxpx.T.dot(iV.T) + kxpx.dot(iV) - pxpx.dot(iViKV)))
        integral_var = kxy - q.T.dot(iV.T).dot(px) - px.T.dot(iV).dot(q) + px.T.dot(iViKV).dot(px)
    elif nb < n_pts:
        Q = rbf_exp_x_kxkx(par, par, x)
        R = rbf_exp_x_xkx(par, x)
        Z = V.T.dot(iK)
        A = V.dot(iViKV)
        b = Z.dot(q) - px
        B = Z.dot(Q).dot(Z.T) + pxpx - Z.dot(kxpx) - kxpx.T.dot(Z.T)
        D = R.dot(Z.T) - xpx
        wm = iK.dot(q - A.dot(b))
        Wc = iK.dot(Q - A.dot(B).dot(A.T)).dot(iK)
        Wcc = (R - D.dot(A.T)).dot(iK)
        model_var = a2 * (1 - np.trace(Q.dot(iK)) + np.trace(B.dot(iViKV)))
        integral_var = kxy - q.T.dot(iK).dot(q) + b.T.dot(iViKV).dot(b)
    else:
        raise ValueError('more basis functions than points')
    if not np.array_equal(Wc, Wc.T):
        Wc = 0.5 * (Wc + Wc.T)
    return dict(wm=wm, Wc=Wc, Wcc=Wcc, model_var=model_var, integral_var=integral_var, iK=iK)


# ==============================================================================================
# state-space model functions, vectorised over trailing axes                 ssmod.py:268-1252
# x has shape (dim, ...) ; time is the integer index the reference passes.
# ==============================================================================================
REENTRY = dict(R0=6374.0, H0=13.406, Gm0=3.9860e5, b0=-0.59783)  # ssmod.py:523-526
PEND_G = 9.81  # ssmod.py:351


def dyn_fcn(name, x, q, time, dt):
    """Discrete-time dynamics x_{k+1} = f(x_k, q_k, k)."""
    xp = x
    if np.isscalar(q):
        q = [q] * 5
    if name == 'UNGMTransition':  # ssmod.py:268-269
        return (0.5 * xp[0] + 25 * (xp[0] / (1 + xp[0] ** 2)) + 8 * np.cos(1.2 * time))[None] + q[0]
    if name == 'UNGMNATransition':  # ssmod.py:299-300: non-additive noise
        return (0.5 * xp[0] + 25 * (xp[0] / (1 + xp[0] ** 2)) + 8 * q[0] * np.cos(1.2 * time))[None]
    if name == 'Pendulum2DTransition':  # ssmod.py:357-358
        return np.stack([xp[0] + xp[1] * dt + q[0], xp[1] - PEND_G * dt * np.sin(xp[0]) + q[1]])
    if name == 'ReentryVehicle2DTransition':  # ssmod.py:530-564 (noise enters components 2..4)
        c = REENTRY
        b = c['b0'] * np.exp(xp[4])
        R = np.sqrt(xp[0] ** 2 + xp[1] ** 2)
        V = np.sqrt(xp[2] ** 2 + xp[3] ** 2)
        D = b * np.exp((c['R0'] - R) / c['H0']) * V
        G = -c['Gm0'] / R ** 3
        return np.stack([xp[0] + dt * xp[2],
                         xp[1] + dt * xp[3],
                         xp[2] + dt * (D * xp[2] + G * xp[0]) + q[0],
                         xp[3] + dt * (D * xp[3] + G * xp[1]) + q[1],
                         xp[4] + q[2]])
    if name == 'ReentryVehicle1DTransition':  # ssmod.py:418-421, Gamma = 1 / 6.096 (ssmod.py:416)
        return np.stack([xp[0] - dt * xp[1] + q[0],
                         xp[1] - dt * np.exp(-(1 / 6.096) * xp[0]) * xp[1] ** 2 * xp[2] + q[1],
                         xp[2] + q[2]])
    if name == 'CoordinatedTurnTransition':  # ssmod.py:675-690 (no omega == 0 guard: NaN, Q11)
        om = xp[4]
        with np.errstate(all='ignore'):
            a, b = np.sin(om * dt), np.cos(om * dt)
            c, d = np.sin(om * dt) / om, (1 - np.cos(om * dt)) / om
            # row-by-row statement of mdyn.dot(x); terms with exact-zero coefficients are kept so
            # that NaN/inf propagate like in the dense product (0 * nan = nan)
            z = 0 * xp[0]
            r0 = 1 * xp[0] + c * xp[1] + z * xp[2] - d * xp[3] + z * xp[4]
            r1 = z * xp[0] + b * xp[1] + z * xp[2] - a * xp[3] + z * xp[4]
            r2 = z * xp[0] + d * xp[1] + 1 * xp[2] + c * xp[3] + z * xp[4]
            r3 = z * xp[0] + a * xp[1] + z * xp[2] + b * xp[3] + z * xp[4]
            r4 = z * xp[0] + z * xp[1] + z * xp[2] + z * xp[3] + 1 * xp[4]
        return np.stack([r0 + q[0], r1 + q[1], r2 + q[2], r3 + q[3], r4 + q[4]])
    if name == 'ConstantVelocity':  # ssmod.py:839-846, noise gain :833-836
        h = dt ** 2 / 2
        return np.stack([xp[0] + dt * xp[1] + h * q[0], xp[1] + dt * q[0], xp[2] + dt * xp[3] + h * q[1], xp[3] + dt * q[1]])
    if name == 'ConstantTurnRateSpeed':  # ssmod.py:755-774: non-additive noise; restated as written (heading += dt * x[3])
        with np.errstate(all='ignore'):
            zero = xp[4] == 0
            c = xp[2] / np.where(zero, 1.0, xp[4])
            f0 = np.where(zero, dt * xp[2] * np.cos(xp[3]),
                          c * (np.sin(xp[3] + xp[4] * dt) - np.sin(xp[3])) + 0.5 * dt ** 2 * np.cos(xp[3]) * q[0])
            f1 = np.where(zero, dt * xp[2] * np.sin(xp[3]),
                          c * (-np.cos(xp[3] + xp[4] * dt) + np.cos(xp[3])) + 0.5 * dt ** 2 * np.sin(xp[3]) * q[0])
        return np.stack([xp[0] + f0, xp[1] + f1, xp[2] + dt * q[0], xp[3] + (dt * xp[3] + 0.5 * dt ** 2 * q[1]),
                         xp[4] + dt * q[1]])
    raise NotImplementedError(name)


def dyn_fcn_cont(name, x, q, time):
    """Continuous-time drift used by Euler-Maruyama.                            ssmod.py:569-584"""
    if name == 'ReentryVehicle2DTransition':
        c = REENTRY
        b = c['b0'] * np.exp(x[4])
        R = np.sqrt(x[0] ** 2 + x[1] ** 2)
        V = np.sqrt(x[2] ** 2 + x[3] ** 2)
        D = b * np.exp((c['R0'] - R) / c['H0']) * V
        G = -c['Gm0'] / R ** 3
        return np.stack([x[2], x[3], D * x[2] + G * x[0] + q[0], D * x[3] + G * x[1] + q[1], q[2] + 0 * x[4]])
    if name == 'ReentryVehicle1DTransition':  # ssmod.py:423-426
        return np.stack([-x[1] + q[0], -np.exp(-(1 / 6.096) * x[0]) * x[1] ** 2 * x[2] + q[1], q[2] + 0 * x[2]])
    raise NotImplementedError(name)


def meas_fcn(name, x, r, time, radar_loc=(0.0, 0.0)):
    """Measurement function y_k = h(x_k[state_index], r_k); x is already index-selected."""
    if np.isscalar(r):
        r = [r] * 4
    if name == 'UNGMMeasurement':  # ssmod.py:1060-1061
        return (0.05 * x[0] ** 2 + r[0])[None]
    if name == 'UNGMNAMeasurement':  # ssmod.py:1085-1086: non-additive noise
        return (0.05 * r[0] * x[0] ** 2)[None]
    if name == 'Pendulum2DMeasurement':  # ssmod.py:1114-1115
        return (np.sin(x[0]) + r[0])[None]
    if name == 'RangeMeasurement':  # ssmod.py:1146-1148; (sx, sy) travels in radar_loc
        return (np.sqrt(radar_loc[0] ** 2 + (x[0] - radar_loc[1]) ** 2) + r[0])[None]
    if name == 'Radar2DMeasurement':  # ssmod.py:1227-1252
        rng = np.sqrt((x[0] - radar_loc[0]) ** 2 + (x[1] - radar_loc[1]) ** 2)
        theta = np.arctan2((x[1] - radar_loc[1]), (x[0] - radar_loc[0]))
        return np.stack([rng + r[0], theta + r[1]])
    if name == 'BearingMeasurement':  # ssmod.py:1189-1195; sensor_pos (4, 2) travels flattened in radar_loc
        sp = np.asarray(radar_loc, dtype=float).reshape(-1, 2)
        if np.isscalar(r) or len(r) < len(sp):
            r = [r if np.isscalar(r) else r[0]] * len(sp)
        return np.stack([np.arctan2(x[1] - sp[i, 1], x[0] - sp[i, 0]) + r[i] for i in range(len(sp))])
    raise NotImplementedError(name)


NONADDITIVE = {'UNGMNATransition', 'UNGMNAMeasurement', 'ConstantTurnRateSpeed'}   # noise_additive = False (ssmod.py:296, 751, 1081)


def _meas_eval(desc, x, time):
    """meas_eval: select state_index, zero additive noise; a non-additive model takes the augmented vector
    [x; r] apart (state_index is None for the only such model).              ssmod.py:960-1009"""
    if str(desc['obs_name']) in NONADDITIVE:
        dr = np.asarray(desc['r_cov']).shape[0]
        return meas_fcn(str(desc['obs_name']), x[:-dr], x[-dr:], time, desc['radar_loc'])
    si = np.asarray(desc['state_index']).astype(int)
    xs = x[si] if si.size else x
    return meas_fcn(str(desc['obs_name']), xs, 0.0, time, desc['radar_loc'])


def _dyn_eval(desc, x, time):
    """dyn_eval with zero additive noise; [x; q] taken apart for a non-additive model.  ssmod.py:129-166"""
    if str(desc['dyn_name']) in NONADDITIVE:
        dq = np.asarray(desc['q_cov']).shape[0]
        return dyn_fcn(str(desc['dyn_name']), x[:-dq], x[-dq:], time, float(desc['dyn_dt']))
    return dyn_fcn(str(desc['dyn_name']), x, 0.0, time, float(desc['dyn_dt']))


# ==============================================================================================
# simulators with injected noise                                   ssmod.py:168-244, 1011-1039
# ==============================================================================================
def simulate_discrete(desc, x0, q):
    """x[:, 0] = x0; x[:, k] = f(x[:, k-1], q[:, k-1], k-1).                    ssmod.py:168-199
    x0 (dx, M); q (dq, steps, M) -> x (dx, steps, M).  The noise passes straight through the
    model function (no gain matrix), as in the reference."""
    steps = q.shape[1]
    x = np.zeros((x0.shape[0], steps, x0.shape[1]))
    x[:, 0] = x0
    for k in range(1, steps):
        x[:, k] = dyn_fcn(str(desc['dyn_name']), x[:, k - 1], q[:, k - 1], k - 1, float(desc['dyn_dt']))
    return x


def simulate_continuous(desc, x0, q, dt):
    """Euler-Maruyama; q (dq, steps+1, M) is the raw noise sample, scaled by sqrt(dt)/dt here;
    returns x[:, 1:].                                                           ssmod.py:201-244"""
    steps = q.shape[1] - 1
    qs = (np.sqrt(dt) / dt) * q
    x = np.zeros((x0.shape[0], steps + 1, x0.shape[1]))
    x[:, 0] = x0
    for k in range(1, steps + 1):
        x[:, k] = x[:, k - 1] + dt * dyn_fcn_cont(str(desc['dyn_name']), x[:, k - 1], qs[:, k - 1], k - 1)
    return x[:, 1:]


def simulate_measurements(desc, x, r):
    """y[:, k] = h(x[state_index, k], r[:, k], k+1).                            ssmod.py:1011-1039"""
    si = np.asarray(desc['state_index']).astype(int)
    xs = x[si] if si.size else x
    return meas_fcn(str(desc['obs_name']), xs, r, None, desc['radar_loc'])


# ==============================================================================================
# small dense linear algebra, two back-ends
# ==============================================================================================
def _t(a):
    return np.swapaxes(a, -1, -2)


def chol_loops(A):
    """Lower Cholesky factor by explicit column loops, vectorised over leading axes, any float
    dtype.  Returns (L, ok): ok is False where a pivot is <= 0 (potrf's info > 0, which
    numpy.linalg.cholesky turns into LinAlgError at mtran.py:139, bqmtran.py:98).  A NaN pivot is not a
    failure: the OpenBLAS potrf behind numpy tests `ajj <= 0` only and lets NaNs propagate (measured,
    golden case c3_reentry_gpq_fail), and so does this function.
    Only the lower triangle of A is read, like LAPACK's 'L' variant."""
    n = A.shape[-1]
    L = np.zeros_like(A)
    ok = np.ones(A.shape[:-2], dtype=bool)
    with np.errstate(all='ignore'):
        for j in range(n):
            s = A[..., j, j] - np.sum(L[..., j, :j] ** 2, axis=-1)
            good = ~(s <= 0)  # True for NaN: propagates
            ok &= good
            s = np.where(good, s, 1.0)
            d = np.sqrt(s)
            L[..., j, j] = d
            for i in range(j + 1, n):
                L[..., i, j] = (A[..., i, j] - np.sum(L[..., i, :j] * L[..., j, :j], axis=-1)) / d
    return L, ok


def _solve_lower(L, B):
    """Forward substitution L X = B (B (..., n, m))."""
    n = L.shape[-1]
    X = np.zeros_like(B)
    for i in range(n):
        X[..., i, :] = (B[..., i, :] - np.sum(L[..., i, :i, None] * X[..., :i, :], axis=-2)) / L[..., i, i, None]
    return X


def _solve_upper_t(L, B):
    """Back substitution L^T X = B."""
    n = L.shape[-1]
    X = np.zeros_like(B)
    for i in range(n - 1, -1, -1):
        X[..., i, :] = (B[..., i, :] - np.sum(L[..., i + 1:, i, None] * X[..., i + 1:, :], axis=-2)) / L[..., i, i, None]
    return X


class _Fail(Exception):
    def __init__(self, code):
        self.code = code


# failure codes shared with the CUDA kernels (include/ssm_b200.h)
FAIL_CHOL_DYN, FAIL_CHOL_OBS, FAIL_CHOL_GAIN, FAIL_NONFINITE_GAIN, FAIL_CHOL_SMOOTH = 1, 2, 3, 4, 5


class _LA:
    """Linear-algebra back-end.  'lapack': per-trajectory, the reference's own library calls,
    raises _Fail(code).  'loops': batched, returns ok masks."""

    def __init__(self, backend):
        self.batched = backend == 'loops'

    def chol(self, A, code):
        if self.batched:
            return chol_loops(A)
        try:
            return np.linalg.cholesky(A), True
        except np.linalg.LinAlgError:
            raise _Fail(code)

    def gain(self, S, C, code):
        """(S^-1 C)^T for SPD S: cho_solve(cho_factor(S), C).T            ssinf.py:321, 342"""
        if self.batched:
            L, ok = chol_loops(S)
            fin = np.all(np.isfinite(S), axis=(-1, -2)) & np.all(np.isfinite(C), axis=(-1, -2))
            X = _solve_upper_t(L, _solve_lower(L, C))
            return _t(X), ok, fin
        try:
            return cho_solve(cho_factor(S), C).T, True, True
        except np.linalg.LinAlgError:
            raise _Fail(code)
        except ValueError:  # scipy's check_finite
            raise _Fail(FAIL_NONFINITE_GAIN)

    def whiten(self, S, v):
        """solve(cholesky(S), v)                                              ssinf.py:731"""
        if self.batched:
            L, _ = chol_loops(S)
            return _solve_lower(L, v[..., None])[..., 0]
        return np.linalg.solve(np.linalg.cholesky(S), v)


# ==============================================================================================
# moment transforms                                mtran.py:105-149, bq/bqmtran.py:60-223, 394-415
# ==============================================================================================
def _tf(desc, which):
    p = which + '_'
    tf = {k[len(p):]: desc[k] for k in desc if k.startswith(p)}
    tf['kind'] = str(tf['kind'])
    return tf


def transform_apply(la, tf, f, mean, cov, code):
    """One moment transform.  mean (..., D), cov (..., D, D) -> mean_f (..., E), cov_f (..., E, E),
    cov_fx (..., E, D), ok mask.  Sigma-point rules use the centred form with Wc = diag(wc)
    (mtran.py:137-149); BQ rules the un-centred form with dense Wc plus the expected model variance
    (bqmtran.py:97-109, 175, 198-199, 223); TPQ scales the model variance by the data
    (bqmtran.py:414-415, bqmod.py:1132-1160)."""
    dt = mean.dtype
    U = np.asarray(tf['points'], dtype=dt)
    wm = np.asarray(tf['wm'], dtype=dt)
    Wc = np.asarray(tf['Wc'], dtype=dt)
    L, ok = la.chol(cov, code)
    m = mean[..., :, None]
    x = m + L @ U                                   # (..., D, N)
    fx = np.moveaxis(f(np.moveaxis(x, -2, 0)), 0, -2)   # (..., E, N)
    mean_f = fx @ wm
    if tf['kind'] == 'sp':
        dfx = fx - mean_f[..., :, None]
        cov_f = dfx @ Wc @ _t(dfx)
        cov_fx = dfx @ Wc @ _t(x - m)
        return mean_f, cov_f, cov_fx, ok
    Wcc = np.asarray(tf['Wcc'], dtype=dt)
    mv = np.asarray(tf['model_var'], dtype=dt)
    e = fx.shape[-2]
    if tf['kind'] == 'tp':
        # I_out is 1x1 because StudentProcessKalman passes dim_out=1 (ssinf.py:550-551): the FULL
        # E x E matrix scale * model_var is added; nu is always the model default (Q7)
        iK = np.asarray(tf['iK'], dtype=dt)
        nu = float(tf['nu'])
        n = U.shape[1]
        scale = (nu - 2 + fx @ iK @ _t(fx)) / (nu - 2 + n)
        emv = scale * mv
        if int(tf['dim_out']) != 1:
            emv = emv * np.eye(e, dtype=dt)
    else:
        emv = mv * np.eye(e, dtype=dt)               # scalar or assigned (E, E) matrix, bqmtran.py:198
    outer = mean_f[..., :, None] * mean_f[..., None, :]
    cov_f = fx @ Wc @ _t(fx) - outer + emv
    cov_fx = fx @ _t(Wcc) @ _t(L)
    return mean_f, cov_f, cov_fx, ok


# ==============================================================================================
# filters and smoother                                                     ssinf.py:66-147, 254-344
# ==============================================================================================
def _noise_terms(desc, dt):
    G = np.asarray(desc['G'], dtype=dt)
    Qc = np.asarray(desc['q_cov'], dtype=dt)
    return G @ Qc @ G.T, np.asarray(desc['r_cov'], dtype=dt)   # ssinf.py:279, 291


def forward_pass(desc, y, backend='lapack', dtype=np.float64, init_mean=None, init_cov=None, t0=None):
    """Gaussian filter forward pass for every trajectory in y (dy, N, M).        ssinf.py:66-118
    Each trajectory starts from (m0, P0), i.e. reset() between trajectories, unless
    per-trajectory initial moments init_mean (dx, M) / init_cov (dx, dx, M) are given (the state a
    reference object carries over when reset() is NOT called, ssinf.py:93, 245); t0 (M,) shifts the
    time index per trajectory (one-step parity checks restart from stored filtered moments).
    Returns dict with fi_mean (dx,N,M), fi_cov (dx,dx,N,M), pr_mean/pr_cov/pr_xx_cov with N+1 time
    slots (slot 0 = initial moments, ssinf.py:93-96) and status (M,), 0 = ok, else
    (k << 8) | code with k the 1-based failing step."""
    M = y.shape[-1]
    if backend == 'lapack':
        outs = []
        for i in range(M):
            init = None if init_mean is None else (init_mean[..., i:i + 1], init_cov[..., i:i + 1])
            outs.append(_forward_batch(desc, y[..., i:i + 1], _LA('lapack'), np.float64, init,
                                       None if t0 is None else t0[i:i + 1]))
        return {k: np.concatenate([o[k] for o in outs], axis=-1) for k in outs[0]}
    init = None if init_mean is None else (init_mean, init_cov)
    return _forward_batch(desc, y, _LA('loops'), dtype, init, t0)


def _forward_batch(desc, y, la, dt, init=None, t0=None):
    dy, N, M = y.shape
    y = y.astype(dt)
    m0 = np.asarray(desc['m0'], dtype=dt)
    P0 = np.asarray(desc['P0'], dtype=dt)
    dx = m0.shape[0]
    GQG, R = _noise_terms(desc, dt)
    tf_dyn, tf_obs = _tf(desc, 'dyn'), _tf(desc, 'obs')
    # non-additive noise: the transform integrates over [x; noise] with blockdiag covariance (ssinf.py:271-272, 282-283)
    dyn_na, obs_na = str(desc['dyn_name']) in NONADDITIVE, str(desc['obs_name']) in NONADDITIVE
    q_cov = np.atleast_2d(np.asarray(desc['q_cov'], dtype=dt))
    q_mean = np.asarray(desc['q_mean'], dtype=dt).reshape(-1) if 'q_mean' in desc else np.zeros(q_cov.shape[0], dtype=dt)
    r_mean = np.asarray(desc['r_mean'], dtype=dt).reshape(-1) if 'r_mean' in desc else np.zeros(R.shape[0], dtype=dt)

    def augment(mean, cov, nm, nc):
        lead = mean.shape[:-1]
        d, dn = mean.shape[-1], nm.shape[0]
        ma = np.concatenate([mean, np.broadcast_to(nm, lead + (dn,))], axis=-1)
        Pa = np.zeros(lead + (d + dn, d + dn), dtype=dt)
        Pa[..., :d, :d] = cov
        Pa[..., d:, d:] = nc
        return ma, Pa
    batched = la.batched
    B = (M,) if batched else ()
    fi_mean = np.full((M, N + 1, dx), np.nan, dtype=dt)
    fi_cov = np.full((M, N + 1, dx, dx), np.nan, dtype=dt)
    if init is None:
        m = np.broadcast_to(m0, B + (dx,)).copy()
        P = np.broadcast_to(P0, B + (dx, dx)).copy()
        fi_mean[:, 0], fi_cov[:, 0] = m0, P0
    else:
        mi, Pi = np.moveaxis(init[0].astype(dt), -1, 0), np.moveaxis(init[1].astype(dt), -1, 0)
        m, P = (mi.copy(), Pi.copy()) if batched else (mi[0].copy(), Pi[0].copy())
        fi_mean[:, 0], fi_cov[:, 0] = mi, Pi
    toff = 0 if t0 is None else (np.asarray(t0)[:, None] if batched else int(np.asarray(t0)[0]))
    pr_mean, pr_cov, pr_xx = fi_mean.copy(), fi_cov.copy(), fi_cov.copy()
    status = np.zeros(M, dtype=np.int64)
    alive = np.ones(M, dtype=bool)
    eye_x = np.eye(dx, dtype=dt)

    def fail(mask_ok, k, code):
        nonlocal alive
        bad = alive & ~np.broadcast_to(mask_ok, (M,))
        status[bad] = (k << 8) | code
        alive = alive & ~bad

    with np.errstate(all='ignore'):
        for k in range(1, N + 1):
            t = toff + (k - 1)                                                 # ssinf.py:104, 277, 288
            try:
                ma, Pa = augment(m, P, q_mean, q_cov) if dyn_na else (m, P)      # ssinf.py:271-272
                mp, Pp, Pxx, ok = transform_apply(la, tf_dyn, lambda x: _dyn_eval(desc, x, t), ma, Pa, FAIL_CHOL_DYN)
                Pxx = Pxx[..., :dx]                                            # ssinf.py:294
                if batched:
                    fail(ok, k, FAIL_CHOL_DYN)
                if not dyn_na:
                    Pp = Pp + GQG                                              # ssinf.py:278-279
                if batched:  # keep dead trajectories harmless
                    Pp = np.where(alive[:, None, None], Pp, eye_x)
                ma, Pa = augment(mp, Pp, r_mean, R) if obs_na else (mp, Pp)      # ssinf.py:282-283
                my, Py, Pyx, ok = transform_apply(la, tf_obs, lambda x: _meas_eval(desc, x, t), ma, Pa, FAIL_CHOL_OBS)
                Pyx = Pyx[..., :dx]                                            # ssinf.py:293
                if batched:
                    fail(ok, k, FAIL_CHOL_OBS)
                if not obs_na:
                    Py = Py + R                                                # ssinf.py:290-291
                K, ok, fin = la.gain(Py, Pyx, FAIL_CHOL_GAIN)                  # ssinf.py:321
                if batched:
                    fail(fin, k, FAIL_NONFINITE_GAIN)
                    fail(ok, k, FAIL_CHOL_GAIN)
                yk = y[:, k - 1, :].T if batched else y[:, k - 1, 0]
                m = mp + (K @ (yk - my)[..., None])[..., 0]                    # ssinf.py:322
                P = Pp - K @ Py @ _t(K)                                        # ssinf.py:323
            except _Fail as e:
                status[0] = (k << 8) | e.code
                alive[:] = False
                break
            sel = alive if batched else slice(None)
            pr_mean[sel, k], pr_cov[sel, k], pr_xx[sel, k] = mp[sel] if batched else mp, \
                Pp[sel] if batched else Pp, Pxx[sel] if batched else Pxx
            fi_mean[sel, k], fi_cov[sel, k] = m[sel] if batched else m, P[sel] if batched else P
            if batched:
                m = np.where(alive[:, None], m, m0)
                P = np.where(alive[:, None, None], P, eye_x)
                if not alive.any():
                    break
    mv = lambda a: np.moveaxis(a, 0, -1)  # noqa: E731  (M, N, ...) -> (..., N, M)
    return dict(fi_mean=mv(np.swapaxes(fi_mean[:, 1:], 1, 2)),
                fi_cov=mv(np.moveaxis(fi_cov[:, 1:], 1, 3)),
                pr_mean=mv(np.swapaxes(pr_mean, 1, 2)),
                pr_cov=mv(np.moveaxis(pr_cov, 1, 3)),
                pr_xx_cov=mv(np.moveaxis(pr_xx, 1, 3)),
                status=status)


def backward_pass(desc, fwd, backend='lapack', dtype=np.float64):
    """RTS smoother over the stored forward-pass arrays.              ssinf.py:120-147, 325-344
    Reproduces the reference's index range k = N-2 .. 1 on arrays with N+1 slots:
    slots N and N-1 are never smoothed, and the recursion starts from the filtered moments of slot N
    (ssinf.py:117) combined with the predictive moments of slot N-1."""
    la = _LA(backend)
    dt = np.float64 if backend == 'lapack' else dtype
    fm = np.concatenate([fwd['pr_mean'][:, :1], fwd['fi_mean']], axis=1).astype(dt)       # slot 0 = m0
    fc = np.concatenate([fwd['pr_cov'][:, :, :1], fwd['fi_cov']], axis=2).astype(dt)
    pm, pc, px = fwd['pr_mean'].astype(dt), fwd['pr_cov'].astype(dt), fwd['pr_xx_cov'].astype(dt)
    dx, N1, M = fm.shape
    N = N1 - 1
    sm, sc = fm.copy(), fc.copy()
    status = np.array(fwd['status']).copy()
    tr = lambda a, k: np.moveaxis(a[..., k, :], -1, 0)  # noqa: E731  -> (M, ...)
    groups = [np.arange(M)] if la.batched else [np.array([i]) for i in range(M)]
    with np.errstate(all='ignore'):
        for g in groups:
            g = g[status[g] == 0]
            if g.size == 0:
                continue
            ms, Ps = tr(fm, N)[g], tr(fc, N)[g]
            if not la.batched:
                ms, Ps = ms[0], Ps[0]
            for k in range(N - 2, 0, -1):
                sq = (lambda a: a[g]) if la.batched else (lambda a: a[g][0])
                mp, Pp, Pxx = sq(tr(pm, k + 1)), sq(tr(pc, k + 1)), sq(tr(px, k + 1))
                mf, Pf = sq(tr(fm, k)), sq(tr(fc, k))
                try:
                    D, ok, _ = la.gain(Pp, Pxx, FAIL_CHOL_SMOOTH)               # ssinf.py:342
                except _Fail as e:
                    status[g] = (k << 8) | e.code
                    break
                ms = mf + (D @ (ms - mp)[..., None])[..., 0]                   # ssinf.py:343
                Ps = Pf + D @ (Ps - Pp) @ _t(D)                                # ssinf.py:344
                sm[:, k, g] = np.moveaxis(np.atleast_2d(ms), 0, -1)
                sc[:, :, k, g] = np.moveaxis(Ps.reshape((-1, dx, dx)), 0, -1)
    return dict(sm_mean=sm[:, 1:], sm_cov=sc[:, :, 1:], status=status)


def student_forward_pass(desc, y, backend='lapack', dtype=np.float64):
    """Studentian (scale-matrix) filter forward pass.                       ssinf.py:589-736
    desc carries dof, fixed_dof, x0_dof, q_dof, r_dof; P0 / q_cov / r_cov are the StudentRV
    *scale* matrices returned by get_stats() (utils.py:673-674)."""
    if backend == 'lapack':
        outs = [_student_batch(desc, y[..., i:i + 1], _LA('lapack'), np.float64) for i in range(y.shape[-1])]
        return {k: np.concatenate([o[k] for o in outs], axis=-1) for k in outs[0]}
    return _student_batch(desc, y, _LA('loops'), dtype)


def _student_batch(desc, y, la, dt):
    dy, N, M = y.shape
    y = y.astype(dt)
    m0 = np.asarray(desc['m0'], dtype=dt)
    P0 = np.asarray(desc['P0'], dtype=dt)
    dx = m0.shape[0]
    G = np.asarray(desc['G'], dtype=dt)
    q_cov, r_cov = np.asarray(desc['q_cov'], dtype=dt), np.asarray(desc['r_cov'], dtype=dt)
    dof, fixed = float(desc['dof']), bool(int(desc['fixed_dof']))
    q_dof, r_dof = float(desc['q_dof']), float(desc['r_dof'])
    dof_fi = float(desc['x0_dof'])
    s0 = (dof - 2) / dof                                                       # ssinf.py:615-618
    q_smat, r_smat = s0 * q_cov, s0 * r_cov
    GQG, GSG = G @ q_cov @ G.T, G @ q_smat @ G.T
    tf_dyn, tf_obs = _tf(desc, 'dyn'), _tf(desc, 'obs')
    batched = la.batched
    B = (M,) if batched else ()
    m = np.broadcast_to(m0, B + (dx,)).copy()
    S = np.broadcast_to(s0 * P0, B + (dx, dx)).copy()
    fi_mean = np.full((M, N + 1, dx), np.nan, dtype=dt)
    fi_cov = np.full((M, N + 1, dx, dx), np.nan, dtype=dt)
    fi_mean[:, 0], fi_cov[:, 0] = m0, P0
    pr_mean, pr_cov, pr_xx = fi_mean.copy(), fi_cov.copy(), fi_cov.copy()
    status = np.zeros(M, dtype=np.int64)
    alive = np.ones(M, dtype=bool)
    eye_x = np.eye(dx, dtype=dt)

    def fail(mask_ok, k, code):
        nonlocal alive
        bad = alive & ~np.broadcast_to(mask_ok, (M,))
        status[bad] = (k << 8) | code
        alive = alive & ~bad

    with np.errstate(all='ignore'):
        for k in range(1, N + 1):
            t = k - 1
            if fixed:                                                          # ssinf.py:650-660
                dof_pr = min(dof_fi, q_dof, r_dof)
                scale = (dof_pr - 2) / dof_pr
            else:
                scale = (dof - 2) / dof
            try:
                mp, Cp, Pxx, ok = transform_apply(la, tf_dyn, lambda x: _dyn_eval(desc, x, t), m, S, FAIL_CHOL_DYN)
                if batched:
                    fail(ok, k, FAIL_CHOL_DYN)
                Sp = scale * Cp                                                # ssinf.py:672
                Cp = Cp + GQG                                                  # ssinf.py:675
                Sp = Sp + GSG                                                  # ssinf.py:676
                if batched:
                    Sp = np.where(alive[:, None, None], Sp, eye_x)
                my, Cy, Cyx, ok = transform_apply(la, tf_obs, lambda x: _meas_eval(desc, x, t), mp, Sp, FAIL_CHOL_OBS)
                if batched:
                    fail(ok, k, FAIL_CHOL_OBS)
                Sy = scale * Cy + r_smat                                       # ssinf.py:687, 693
                Syx = scale * Cyx                                              # ssinf.py:688
                K, ok, fin = la.gain(Sy, Syx, FAIL_CHOL_GAIN)                  # ssinf.py:724
                if batched:
                    fail(fin, k, FAIL_NONFINITE_GAIN)
                    fail(ok, k, FAIL_CHOL_GAIN)
                yk = y[:, k - 1, :].T if batched else y[:, k - 1, 0]
                e = yk - my
                m = mp + (K @ e[..., None])[..., 0]                            # ssinf.py:725
                C = Sp - K @ Sy @ _t(K)                                        # ssinf.py:727
                delta = la.whiten(Sy, e)                                       # ssinf.py:731
                sc = (dof + np.sum(delta * delta, axis=-1)) / (dof + dy)       # ssinf.py:732
                S = (sc[..., None, None] if batched else sc) * C               # ssinf.py:733
                dof_fi += dy                                                   # ssinf.py:736
            except (_Fail, np.linalg.LinAlgError) as e:
                status[0] = (k << 8) | getattr(e, 'code', FAIL_CHOL_GAIN)
                alive[:] = False
                break
            sel = alive if batched else slice(None)
            pick = (lambda a: a[sel]) if batched else (lambda a: a)
            pr_mean[sel, k], pr_cov[sel, k], pr_xx[sel, k] = pick(mp), pick(Cp), pick(Pxx)
            fi_mean[sel, k], fi_cov[sel, k] = pick(m), pick(C)
            if batched:
                m = np.where(alive[:, None], m, m0)
                S = np.where(alive[:, None, None], S, eye_x)
    mv = lambda a: np.moveaxis(a, 0, -1)  # noqa: E731
    return dict(fi_mean=mv(np.swapaxes(fi_mean[:, 1:], 1, 2)),
                fi_cov=mv(np.moveaxis(fi_cov[:, 1:], 1, 3)),
                pr_mean=mv(np.swapaxes(pr_mean, 1, 2)),
                pr_cov=mv(np.moveaxis(pr_cov, 1, 3)),
                pr_xx_cov=mv(np.moveaxis(pr_xx, 1, 3)),
                status=status)


# ==============================================================================================
# scores                                   utils.py:18-148, research/gpq/icinco_demo.py:17-52
# ==============================================================================================
def squared_error(x, m):
    """utils.py:18-38"""
    return (x - m) ** 2


def mse_matrix(x, m):
    """Per-step sample MSE matrix over trajectories: x, m (dx, N, M) -> (dx, dx, N).  utils.py:41-64"""
    d = x - m
    return np.einsum('ikm,jkm->ijk', d, d) / x.shape[-1]


def _mat_sqrt(a):
    """Cholesky, SVD fallback u sqrt(s) when not PD.                           utils.py:412-433"""
    try:
        return np.linalg.cholesky(a)
    except np.linalg.LinAlgError:
        u, s, _ = np.linalg.svd(a)
        return u.dot(np.diag(np.sqrt(s)))


def neg_log_likelihood(x, m, P):
    """0.5 (log|P| + dx' P^-1 dx + d log 2pi), x, m (dx, N, M), P (dx, dx, N, M) -> (N, M).  utils.py:123-148"""
    d = np.moveaxis(x - m, 0, -1)                       # (N, M, dx)
    Pm = np.moveaxis(P, (0, 1), (-2, -1))               # (N, M, dx, dx)
    quad = np.einsum('...i,...ij,...j->...', d, np.linalg.inv(Pm), d)
    sign, logdet = np.linalg.slogdet(Pm)
    return 0.5 * (sign * logdet + quad + x.shape[0] * np.log(2 * np.pi))


def log_cred_ratio(x, m, P, mse):
    """10 (log10 dx'P^-1dx - log10 dx'MSE^-1dx) per (step, trajectory); mse (dx, dx, N).  utils.py:67-120"""
    dx, N, M = x.shape
    out = np.zeros((N, M))
    for k in range(N):
        sq_mse = _mat_sqrt(mse[..., k])
        for s in range(M):
            d = x[:, k, s] - m[:, k, s]
            a = sp_solve(_mat_sqrt(P[:, :, k, s]), d)
            b = sp_solve(sq_mse, d)
            out[k, s] = 10 * (np.log10(a.dot(a)) - np.log10(b.dot(b)))
    return out


def evaluate_performance(x, mean, cov):
    """RMSE / NCI / NLL exactly as aggregated by research/gpq/icinco_demo.py:17-52: RMSE is the
    trajectory-mean of sqrt(time-mean SE); NCI and NLL skip k = 0 but divide by N."""
    dx, N, M = x.shape
    rmse = np.sqrt(np.mean(squared_error(x, mean), axis=1)).mean(axis=1)       # (dx,)
    mse = mse_matrix(x, mean)
    nll = neg_log_likelihood(x, mean, cov)
    lcr = log_cred_ratio(x[:, 1:], mean[:, 1:], cov[:, :, 1:], mse[..., 1:])
    nll[0] = 0.0
    nci_t = np.concatenate([np.zeros((1, M)), lcr], axis=0)
    return dict(rmse=rmse, nci=nci_t.mean(axis=0).mean(), nll=nll.mean(axis=0).mean(), mse=mse, nll_km=nll, lcr_km=nci_t)
