"""Bayesian-quadrature moment transforms (mirror of ssmtoybox/bq/bqmtran.py: BQTransform :11-282,
GaussianProcessTransform :285-310, BayesSardTransform :313-360, StudentTProcessTransform :363-415)."""
import numpy as np

from ..mtran import MomentTransform, _apply_device
from .bqmod import GaussianProcessModel, StudentTProcessModel, BayesSardModel


class BQTransform(MomentTransform):
    """Base class of BQ moment transforms."""
    _supported_models_ = ['gp', 'tp', 'bs']

    def __init__(self, dim_in, dim_out, kern_par, model, kern_str, point_str, point_par, estimate_par, **kwargs):
        self.model = BQTransform._get_model(dim_in, dim_out, model, kern_str, point_str, kern_par, point_par,
                                            estimate_par, **kwargs)
        self.I_out = np.eye(dim_out)

    def _tf_dict(self, prefix):
        """Current state of the transform as plain arrays.  Read at forward_pass time: research code
        overwrites wm / Wc / Wcc and model.model_var from outside (research/bsq/bsq_tracking.py:276-281,
        research/tpq/tpq_ungm.py:114-124)."""
        kind = 'tp' if isinstance(self.model, StudentTProcessModel) else 'bq'
        d = {prefix + 'kind': kind, prefix + 'points': self.model.points, prefix + 'wm': self.wm, prefix + 'Wc': self.Wc,
             prefix + 'Wcc': self.Wcc, prefix + 'model_var': np.asarray(self.model.model_var, dtype=np.float64),
             prefix + 'dim_out': self.I_out.shape[0]}
        if kind == 'tp':
            d[prefix + 'iK'] = self.model.iK
            d[prefix + 'nu'] = float(self.model.nu)
        return d

    def apply(self, f, mean, cov, fcn_par, kern_par=None):
        """(bqmtran.py:60-109)"""
        if kern_par is not None:
            self.wm, self.Wc, self.Wcc = self.weights(kern_par)
        return _apply_device(self._tf_dict('t_'), f, mean, cov, fcn_par)

    def weights(self, par, *args):
        wm, wc, wcc, emv, ivar = self.model.bq_weights(par, *args)
        return wm, wc, wcc

    @staticmethod
    def _get_model(dim_in, dim_out, model, kern_str, point_str, kern_par, point_par, estimate_par, **kwargs):
        """(bqmtran.py:226-279).  Note the 'tp' branch does not forward kwargs, so nu is always the model
        default 4.0."""
        model = model.lower()
        if model == 'gp':
            return GaussianProcessModel(dim_in, kern_par, kern_str, point_str, point_par, estimate_par)
        elif model == 'tp':
            return StudentTProcessModel(dim_in, kern_par, kern_str, point_str, point_par, estimate_par)
        elif model == 'bs':
            return BayesSardModel(dim_in, kern_par, point_str=point_str, point_par=point_par,
                                  estimate_par=estimate_par, **kwargs)
        raise NotImplementedError("integrand model '{}' has no device implementation".format(model))


class GaussianProcessTransform(BQTransform):
    def __init__(self, dim_in, dim_out, kern_par, kern_str='rbf', point_str='ut', point_par=None, estimate_par=False):
        super(GaussianProcessTransform, self).__init__(dim_in, dim_out, kern_par, 'gp', kern_str, point_str, point_par,
                                                       estimate_par)
        self.wm, self.Wc, self.Wcc = self.weights(kern_par)


class BayesSardTransform(BQTransform):
    def __init__(self, dim_in, dim_out, kern_par, multi_ind=2, point_str='ut', point_par=None, estimate_par=False):
        super(BayesSardTransform, self).__init__(dim_in, dim_out, kern_par, 'bs', 'rbf', point_str, point_par,
                                                 estimate_par, multi_ind=multi_ind)
        self.wm, self.Wc, self.Wcc = self.weights(kern_par, multi_ind)

    def weights(self, par, *args):
        multi_ind = args[0]
        wm, wc, wcc, emv, ivar = self.model.bq_weights(par, multi_ind)
        return wm, wc, wcc


class StudentTProcessTransform(BQTransform):
    def __init__(self, dim_in, dim_out, kern_par, kern_str='rbf', point_str='ut', point_par=None, estimate_par=False,
                 nu=3.0):
        super(StudentTProcessTransform, self).__init__(dim_in, dim_out, kern_par, 'tp', kern_str, point_str, point_par,
                                                       estimate_par, nu=nu)
        self.wm, self.Wc, self.Wcc = self.weights(kern_par)
