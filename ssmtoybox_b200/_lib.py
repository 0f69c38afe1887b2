"""ctypes binding of the C-ABI CUDA library (include/ssm_b200.h).

The library is the only compute back-end of this package: there is no CPU fallback.  Importing
this module fails loudly when libssmb200.so has not been built (python -m ssmtoybox_b200.build,
or __graft_entry__.build()).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('SSM_B200_LIB', os.path.join(_HERE, 'lib', 'libssmb200.so'))  # override: kernel experiments

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)

# constants of include/ssm_b200.h
SSM_OK, SSM_E_INVALID, SSM_E_UNSUPPORTED, SSM_E_CUDA = 0, -1, -2, -3
FAIL_CHOL_DYN, FAIL_CHOL_OBS, FAIL_CHOL_GAIN, FAIL_NONFINITE_GAIN, FAIL_CHOL_SMOOTH = 1, 2, 3, 4, 5
DYN_IDS = {'UNGMTransition': 1, 'Pendulum2DTransition': 2, 'ReentryVehicle2DTransition': 3,
           'CoordinatedTurnTransition': 4, 'ReentryVehicle1DTransition': 5, 'UNGMNATransition': 6,
           'ConstantVelocity': 7, 'ConstantTurnRateSpeed': 8}
OBS_IDS = {'UNGMMeasurement': 1, 'Pendulum2DMeasurement': 2, 'Radar2DMeasurement': 3, 'RangeMeasurement': 4,
           'UNGMNAMeasurement': 5, 'BearingMeasurement': 6}
NONADDITIVE = {'UNGMNATransition', 'UNGMNAMeasurement', 'ConstantTurnRateSpeed'}   # models integrated over the augmented vector [x; noise]
TF_SP, TF_BQ, TF_TP = 1, 2, 3
FAMILY_GAUSS, FAMILY_STUDENT = 1, 2
SIM_DISCRETE, SIM_CONTINUOUS, SIM_MEASURE = 1, 2, 3


class SsmTransform(C.Structure):
    _fields_ = [('kind', C.c_int32), ('dim_in', C.c_int32), ('dim_out', C.c_int32), ('n_pts', C.c_int32),
                ('points', c_double_p), ('wm', c_double_p), ('Wc', c_double_p), ('Wcc', c_double_p),
                ('model_var', c_double_p), ('iK', c_double_p), ('nu', C.c_double),
                ('tp_full_matrix', C.c_int32), ('reserved', C.c_int32)]


class SsmDesc(C.Structure):
    _fields_ = [('dyn_model', C.c_int32), ('obs_model', C.c_int32), ('dx', C.c_int32), ('dy', C.c_int32),
                ('dyn_par', C.c_double * 8), ('obs_par', C.c_double * 8),
                ('n_state_index', C.c_int32), ('state_index', C.c_int32 * 8),
                ('family', C.c_int32), ('reserved', C.c_int32),
                ('m0', c_double_p), ('P0', c_double_p), ('GQG', c_double_p), ('R', c_double_p),
                ('dof', C.c_double), ('x0_dof', C.c_double), ('q_dof', C.c_double), ('r_dof', C.c_double),
                ('fixed_dof', C.c_int32), ('reserved2', C.c_int32),
                ('tf_dyn', SsmTransform), ('tf_obs', SsmTransform),
                ('q_mean', c_double_p), ('q_cov', c_double_p), ('r_mean', c_double_p),
                ('dq', C.c_int32), ('reserved3', C.c_int32)]


class SsmRng(C.Structure):
    _fields_ = [('seed', C.c_uint64), ('traj_offset', C.c_int64),
                ('x0_mean', c_double_p), ('x0_factor', c_double_p), ('q_factor', c_double_p), ('r_factor', c_double_p),
                ('x0_dof', C.c_double), ('q_dof', C.c_double), ('r_dof', C.c_double),
                ('dq', C.c_int32), ('reserved', C.c_int32)]


class SsmError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            'ssmtoybox_b200: CUDA library {} not found. Build it with `python -m ssmtoybox_b200.build` '
            '(nvcc, sm_100a). There is no CPU fallback.'.format(LIB_PATH))
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    lib.ssm_abi_version.restype = C.c_int
    lib.ssm_last_error.restype = C.c_char_p
    lib.ssm_weights_reflective.restype = C.c_int
    lib.ssm_weights_reflective.argtypes = [C.POINTER(SsmTransform)]
    lib.ssm_filter.restype = C.c_int
    lib.ssm_filter.argtypes = [C.POINTER(SsmDesc), vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp, i64, i32, i64, vp]
    lib.ssm_filter_window.restype = C.c_int
    lib.ssm_filter_window.argtypes = [C.POINTER(SsmDesc), vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp, i64, i32, i32, i32, i64, vp]
    lib.ssm_filter_window_lower.restype = C.c_int
    lib.ssm_filter_window_lower.argtypes = lib.ssm_filter_window.argtypes
    lib.ssm_filter_scores.restype = C.c_int
    lib.ssm_filter_scores.argtypes = [C.POINTER(SsmDesc), vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp, i64, i32, i32, i32, i64, vp]
    lib.ssm_smooth_window.restype = C.c_int
    lib.ssm_smooth_window.argtypes = [i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, i32, i32, i32, i64, vp]
    lib.ssm_scores_phase1_window.restype = C.c_int
    lib.ssm_scores_phase1_window.argtypes = [i32, vp, vp, vp, vp, vp, vp, i64, i32, i32, i32, i64, vp]
    lib.ssm_scores_phase2_window.restype = C.c_int
    lib.ssm_scores_phase2_window.argtypes = [i32, vp, vp, vp, vp, vp, vp, i64, i32, i32, i32, i64, vp]
    lib.ssm_scores_phase1_traj.restype = C.c_int
    lib.ssm_scores_phase1_traj.argtypes = [i32, vp, vp, vp, vp, vp, vp, vp, i64, i32, i32, i32, i64, vp]
    lib.ssm_scores_phase1_quad.restype = C.c_int
    lib.ssm_scores_phase1_quad.argtypes = [i32, vp, vp, vp, vp, vp, vp, vp, vp, i64, i32, i32, i32, i64, vp]
    lib.ssm_smooth_quad.restype = C.c_int
    lib.ssm_smooth_quad.argtypes = [i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, i32, i32, i32, i64, vp]
    lib.ssm_smooth_scores.restype = C.c_int
    lib.ssm_smooth_scores.argtypes = [i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, i32, i32, i32, i64, vp]
    lib.ssm_scores_phase2_res.restype = C.c_int
    lib.ssm_scores_phase2_res.argtypes = [i32, vp, vp, vp, vp, vp, vp, i64, i32, i32, i32, i64, vp]
    lib.ssm_scores_phase2_quad.restype = C.c_int
    lib.ssm_scores_phase2_quad.argtypes = [i32, vp, vp, vp, vp, vp, vp, vp, i64, i32, i32, i32, i64, vp]
    lib.ssm_scores_phase2_traj.restype = C.c_int
    lib.ssm_scores_phase2_traj.argtypes = [i32, vp, vp, vp, vp, vp, vp, vp, i64, i32, i32, i32, i64, vp]
    lib.ssm_bootstrap_var.restype = C.c_int
    lib.ssm_bootstrap_var.argtypes = [vp, i64, i32, C.c_uint64, vp, vp, vp]
    lib.ssm_math_probe.restype = C.c_int
    lib.ssm_math_probe.argtypes = [i32, vp, vp, vp, i64, vp]
    lib.ssm_smooth.restype = C.c_int
    lib.ssm_smooth.argtypes = [i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, i32, i64, vp]
    lib.ssm_fp64_peak_kernel.restype = C.c_int
    lib.ssm_fp64_peak_kernel.argtypes = [i32, i32, vp, C.POINTER(dbl), vp]
    if True:
        lib.ssm_simulate.restype = C.c_int
        lib.ssm_simulate.argtypes = [C.POINTER(SsmDesc), C.POINTER(SsmRng), i32, dbl, i32, vp, vp, vp, vp, vp, i64, i32, i64, vp]
    if True:
        lib.ssm_bq_weights.restype = C.c_int
        lib.ssm_bq_weights.argtypes = [i32, i32, i32, c_double_p, c_double_p, c_int32_p, i32, vp, vp, vp, vp, vp, vp, i32, vp]
    if True:
        lib.ssm_scores_width.restype = i32
        lib.ssm_scores_width.argtypes = [i32]
        lib.ssm_scores_phase1.restype = C.c_int
        lib.ssm_scores_phase1.argtypes = [i32, vp, vp, vp, vp, vp, vp, i64, i32, i64, vp]
        lib.ssm_scores_phase2.restype = C.c_int
        lib.ssm_scores_phase2.argtypes = [i32, vp, vp, vp, vp, vp, vp, i64, i32, i64, vp]
    lib.ssm_transform_apply.restype = C.c_int
    lib.ssm_transform_apply.argtypes = [i32, i32, i32, i32, i32, c_double_p, C.POINTER(SsmTransform), dbl, vp, vp, vp, vp,
                                        vp, vp, i64, i64, vp]
    lib.ssm_transform_apply_batched.restype = C.c_int
    lib.ssm_transform_apply_batched.argtypes = [i32, i32, i32, i32, i32, c_double_p, i32, c_double_p, vp, vp, vp, vp, dbl, vp, vp, vp, vp,
                                                vp, vp, i64, i64, vp]
    lib.ssm_model_eval.restype = C.c_int
    lib.ssm_model_eval.argtypes = [i32, i32, i32, i32, i32, c_double_p, dbl, vp, vp, vp, i64, i64, vp]
    lib.ssm_rbf_eval.restype = C.c_int
    lib.ssm_rbf_eval.argtypes = [i32, i32, i32, c_double_p, c_double_p, c_double_p, i32, vp, vp]
    lib.ssm_rbf_expectations.restype = C.c_int
    lib.ssm_rbf_expectations.argtypes = [i32, i32, c_double_p, c_double_p, i32, vp, vp, vp, vp, vp]
    lib.ssm_rbf_student_expectations.restype = C.c_int
    lib.ssm_rbf_student_expectations.argtypes = [i32, i32, c_double_p, c_double_p, dbl, i64, C.c_uint64, vp, vp, vp, vp, vp]
    lib.ssm_sample_mixture.restype = C.c_int
    lib.ssm_sample_mixture.argtypes = [i32, i32, c_double_p, c_double_p, c_double_p, C.c_uint64, i64, vp, vp, i64, i64, vp]
    lib.ssm_gp_nlml.restype = C.c_int
    lib.ssm_gp_nlml.argtypes = [i32, i32, i32, i32, c_double_p, c_double_p, c_double_p, c_double_p, dbl, vp, vp, vp, vp]
    lib.ssm_sample.restype = C.c_int
    lib.ssm_sample.argtypes = [i32, c_double_p, c_double_p, dbl, C.c_uint64, i64, vp, i64, i64, vp]
    lib.ssm_memcpy2d.restype = C.c_int
    lib.ssm_memcpy2d.argtypes = [vp, C.c_uint64, vp, C.c_uint64, C.c_uint64, C.c_uint64, i32, vp]
    return lib


lib = _load()


def check(rc, what):
    if rc != SSM_OK:
        msg = lib.ssm_last_error().decode('utf-8', 'replace')
        if rc == SSM_E_UNSUPPORTED:
            raise NotImplementedError('{}: {}'.format(what, msg))
        if rc == SSM_E_INVALID:
            raise ValueError('{}: {}'.format(what, msg))
        raise SsmError('{}: {} (rc={})'.format(what, msg, rc))
