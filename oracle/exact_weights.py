"""TEST INFRASTRUCTURE ONLY (see oracle/README.md): GaussianProcessModel.bq_weights (bq/bqmod.py:495-523) with the RBF
kernel expectations of bq/bqkern.py:329-424 evaluated in mpmath at `digits` decimal digits -- the arbiter for the
double-double weights kernel, and the source of the structured weights assigned to the reference in
gen_golden.gen_structured()."""
import numpy as np


def exact_gp_weights(par, x, digits=60):
    """par (1, D+1) = [alpha, l_1 .. l_D], x (D, N) unit points -> dict wm (N,), Wc (N, N), Wcc (D, N), model_var,
    integral_var, each rounded to float64 once."""
    import mpmath as mp
    mp.mp.dps = digits
    D, N = x.shape
    ell = [mp.mpf(float(v)) for v in np.asarray(par).ravel()[1:]]
    alpha = mp.mpf(float(np.asarray(par).ravel()[0]))
    X = [[mp.mpf(float(x[d, i])) for i in range(N)] for d in range(D)]
    K, Qm, qv, Rm = mp.matrix(N, N), mp.matrix(N, N), mp.matrix(1, N), mp.matrix(D, N)
    cdet = rdet = mp.mpf(1)
    for d in range(D):
        cdet *= 1 / ell[d] ** 2 + 1
        rdet *= 2 / ell[d] ** 2 + 1
    for i in range(N):
        s = sum(X[d][i] ** 2 / (ell[d] ** 2 + 1) for d in range(D))
        qv[i] = mp.exp(-s / 2) / mp.sqrt(cdet)
        for d in range(D):
            Rm[d, i] = qv[i] * X[d][i] / (ell[d] ** 2 + 1)
        for j in range(N):
            K[i, j] = mp.exp(-sum(((X[d][i] - X[d][j]) / ell[d]) ** 2 for d in range(D)) / 2) + (mp.mpf('1e-8') if i == j else 0)
            n = -sum((X[d][i] / ell[d]) ** 2 + (X[d][j] / ell[d]) ** 2 for d in range(D)) / 2 + \
                sum((X[d][i] / ell[d] ** 2 + X[d][j] / ell[d] ** 2) ** 2 / (2 / ell[d] ** 2 + 1) for d in range(D)) / 2
            Qm[i, j] = mp.exp(n) / mp.sqrt(rdet)
    iK = K ** -1
    tof = lambda M: np.array([[float(M[i, j]) for j in range(M.cols)] for i in range(M.rows)])  # noqa: E731
    QiK = Qm * iK
    return dict(wm=tof(qv * iK).ravel(), Wc=tof(iK * Qm * iK), Wcc=tof(Rm * iK), iK=tof(iK),
                model_var=float(alpha ** 2 * (1 - sum(QiK[i, i] for i in range(N)))),
                integral_var=float(alpha ** 2 / mp.sqrt(rdet) - (qv * iK * qv.T)[0, 0]))
