"""The C-ABI shared library without a GPU: it loads, exports every symbol the public header declares,
validates arguments before touching CUDA, and the Python lowering fills the descriptor as specified."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, golden

HEADER = os.path.join(ROOT, 'include', 'ssm_b200.h')


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(ssm_[a-z0-9_]+)\s*\(', src)))


def test_header_declares_entry_points():
    syms = declared_symbols()
    for s in ('ssm_filter', 'ssm_smooth', 'ssm_simulate', 'ssm_bq_weights', 'ssm_scores_phase1', 'ssm_scores_phase2',
              'ssm_transform_apply', 'ssm_model_eval', 'ssm_sample', 'ssm_abi_version', 'ssm_last_error'):
        assert s in syms


def test_library_exports_every_declared_symbol():
    from ssmtoybox_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH)
    lib = C.CDLL(_lib.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(lib, s), 'symbol {} declared in include/ssm_b200.h is not exported'.format(s)
    assert _lib.lib.ssm_abi_version() == 5


def test_struct_layout_matches_header(tmp_path):
    """sizeof / offsetof of the public structs as seen by a C compiler == the ctypes mirror."""
    import subprocess
    from ssmtoybox_b200 import _lib
    src = tmp_path / 'layout.c'
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "ssm_b200.h"\n'
        'int main(void) { printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(ssm_transform), sizeof(ssm_desc), '
        'sizeof(ssm_rng), offsetof(ssm_desc, m0), offsetof(ssm_desc, dof), offsetof(ssm_desc, tf_dyn), '
        'offsetof(ssm_desc, tf_obs), offsetof(ssm_transform, nu)); return 0; }\n')
    exe = tmp_path / 'layout'
    subprocess.run(['gcc', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [C.sizeof(_lib.SsmTransform), C.sizeof(_lib.SsmDesc), C.sizeof(_lib.SsmRng), _lib.SsmDesc.m0.offset,
            _lib.SsmDesc.dof.offset, _lib.SsmDesc.tf_dyn.offset, _lib.SsmDesc.tf_obs.offset, _lib.SsmTransform.nu.offset]
    assert got == want


def test_integration_md_struct_sketch_matches_header(tmp_path):
    """The ctypes structs a maintainer would copy out of INTEGRATION.md have the size and field offsets the C header
    gives them (an undersized sketch hands the library a short buffer)."""
    import subprocess
    md = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    block = re.search(r'```python\n(.*?)```', md, flags=re.S).group(1)
    # only the struct definitions of the sketch: no library load, no torch
    keep = ['import ctypes as C', 'P = C.POINTER(C.c_double)']
    keep += re.findall(r'^class Ssm\w+\(C\.Structure\):.*?\n(?=\S)', block, flags=re.S | re.M)
    ns = {}
    exec('\n'.join(keep), ns)
    src = tmp_path / 'layout.c'
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "ssm_b200.h"\n'
        'int main(void) { printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(ssm_transform), sizeof(ssm_desc), '
        'offsetof(ssm_desc, tf_obs), offsetof(ssm_desc, q_mean), offsetof(ssm_desc, r_mean), offsetof(ssm_desc, dq)); return 0; }\n')
    exe = tmp_path / 'layout'
    subprocess.run(['gcc', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    T, D = ns['SsmTransform'], ns['SsmDesc']
    assert got == [C.sizeof(T), C.sizeof(D), D.tf_obs.offset, D.q_mean.offset, D.r_mean.offset, D.dq.offset]


def test_argument_validation_needs_no_gpu():
    from ssmtoybox_b200 import _lib
    lib = _lib.lib
    assert lib.ssm_filter(None, None, None, None, None, None, None, None, None, None, None, None, 0, None, 1, 1, 1, None) == _lib.SSM_E_INVALID
    assert b'NULL' in lib.ssm_last_error()
    assert lib.ssm_smooth(5, None, None, None, None, None, None, None, None, None, None, None, 1, 1, 1, None) == _lib.SSM_E_INVALID
    assert lib.ssm_scores_width(5) == 5 + 25 + 3
    with pytest.raises(ValueError):
        _lib.check(_lib.SSM_E_INVALID, 'x')
    with pytest.raises(NotImplementedError):
        _lib.check(_lib.SSM_E_UNSUPPORTED, 'x')


def test_setup_kernels_validate_arguments_without_a_gpu():
    """ssm_rbf_student_expectations / ssm_gp_nlml: NULL pointers, sizes beyond the compiled capacities and bad
    degrees of freedom are rejected before any CUDA call."""
    import ctypes as C
    from ssmtoybox_b200 import _lib
    lib = _lib.lib
    par = (C.c_double * 9)(*([1.0] * 9))
    pts = (C.c_double * 512)()
    buf = C.c_void_p(8)      # never dereferenced: validation comes first
    assert lib.ssm_rbf_student_expectations(5, 11, None, pts, 4.0, 1000, 0, buf, buf, buf, buf, None) == _lib.SSM_E_INVALID
    assert lib.ssm_rbf_student_expectations(9, 11, par, pts, 4.0, 1000, 0, buf, buf, buf, buf, None) == _lib.SSM_E_UNSUPPORTED
    assert lib.ssm_rbf_student_expectations(5, 33, par, pts, 4.0, 1000, 0, buf, buf, buf, buf, None) == _lib.SSM_E_UNSUPPORTED
    assert lib.ssm_rbf_student_expectations(5, 11, par, pts, 0.0, 1000, 0, buf, buf, buf, buf, None) == _lib.SSM_E_INVALID
    assert lib.ssm_rbf_student_expectations(5, 11, par, pts, 4.0, 0, 0, buf, buf, buf, buf, None) == _lib.SSM_E_INVALID
    assert b'dof' in lib.ssm_last_error()
    assert lib.ssm_gp_nlml(5, 11, 5, 1, None, pts, pts, None, 0.0, buf, buf, buf, None) == _lib.SSM_E_INVALID
    assert lib.ssm_gp_nlml(9, 11, 5, 1, par, pts, pts, None, 0.0, buf, buf, buf, None) == _lib.SSM_E_UNSUPPORTED
    assert lib.ssm_gp_nlml(5, 33, 5, 1, par, pts, pts, None, 0.0, buf, buf, buf, None) == _lib.SSM_E_UNSUPPORTED
    assert lib.ssm_gp_nlml(5, 11, 9, 1, par, pts, pts, None, 0.0, buf, buf, buf, None) == _lib.SSM_E_UNSUPPORTED
    assert lib.ssm_gp_nlml(5, 11, 5, 1, par, pts, pts, None, 1.5, buf, buf, buf, None) == _lib.SSM_E_INVALID      # TP needs nu > 2
    assert lib.ssm_gp_nlml(5, 11, 5, 0, par, pts, pts, None, 0.0, buf, buf, buf, None) == _lib.SSM_OK             # empty batch


def test_lowering_from_golden_description():
    from ssmtoybox_b200 import _lib, device as dv
    g = golden('c3_reentry_gpq')
    low = dv.lower(g)
    d = low.desc
    assert (d.dyn_model, d.obs_model, d.dx, d.dy, d.family) == (3, 3, 5, 2, _lib.FAMILY_GAUSS)
    assert d.dyn_par[0] == 0.1 and d.obs_par[0] == 6374.0
    assert d.tf_dyn.kind == _lib.TF_BQ and d.tf_dyn.n_pts == 11 and d.tf_obs.dim_out == 2
    GQG = np.ctypeslib.as_array(d.GQG, shape=(5, 5))
    assert np.allclose(np.diag(GQG), [0, 0, 2.4e-5, 2.4e-5, 1e-6])  # G = [0; I3], ssmod.py:527
    mv = np.ctypeslib.as_array(d.tf_dyn.model_var, shape=(5, 5))
    assert np.allclose(mv, float(g['dyn_model_var']) * np.eye(5))    # model_var * I_out, bqmtran.py:198
    g = golden('c4_ct_tpq')
    low = dv.lower(g)
    assert low.desc.tf_dyn.kind == _lib.TF_TP and low.desc.tf_dyn.tp_full_matrix == 1 and low.desc.tf_dyn.nu == 4.0
    assert list(low.desc.state_index[:2]) == [0, 2]
    g = golden('c4_ct_fsstudent')
    low = dv.lower(g)
    assert low.desc.family == _lib.FAMILY_STUDENT and low.desc.dof == 6.0 and low.desc.fixed_dof == 1
    with pytest.raises(NotImplementedError):
        dv.lower(dict(g, dyn_name='NoSuchTransition'))
    g = golden('c9_ctb_ukf')                                             # 4 bearing sensors: positions in obs_par[0..7]
    low = dv.lower(g)
    assert (low.desc.dyn_model, low.desc.obs_model, low.desc.dx, low.desc.dy) == (4, 6, 5, 4)
    assert list(low.desc.obs_par) == [1000.0, 0, 0, 1000.0, -1000.0, 0, 0, -1000.0]
    g = golden('c10_ctrs_ukf')                                           # non-additive: transform over [x; q], 7-D
    low = dv.lower(g)
    assert (low.desc.dyn_model, low.desc.dq, low.desc.tf_dyn.dim_in, low.desc.tf_dyn.n_pts, low.desc.tf_obs.dim_in) == (8, 2, 7, 15, 5)


def test_product_package_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under ssmtoybox_b200/ may reference it."""
    pkg = os.path.join(ROOT, 'ssmtoybox_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh')):
                src = open(os.path.join(dirpath, f)).read()
                assert 'ssm_oracle' not in src and 'ref_shim' not in src and 'import oracle' not in src, f
                assert not re.search(r'sys\.path.*reference', src), f
