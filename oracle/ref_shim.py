"""Compatibility shim that makes the UNMODIFIED reference (/root/reference, jacobnzw/SSMToybox
v0.1.1a0) importable under the container's numpy 2.x / scipy 1.18 / no-matplotlib stack.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (ssmtoybox_b200/) may import this
module; it is used by oracle/gen_golden.py (run in the build container, where /root/reference
exists) to generate the golden vectors committed under tests/golden/.  /root/reference does
not exist on the GPU box, so nothing that runs there imports this file.

What is patched:
  * matplotlib is absent and bq/bqmod.py:3 imports it at module top -> stub modules;
  * removed numpy aliases np.int / np.float / np.asscalar / np.alltrue (utils.py:463-469,182);
  * scipy.log10 removed (utils.py:120);
  * scipy.special.factorial2(-1) must be 1 (reference relies on it in bqmod.py:656-661,694,727;
    scipy >= 1.11 returns 0).
"""
import sys
import types

import numpy as np
import scipy
import scipy.special

import os

# /root/reference in the build container; on the GPU box the pip-installed copy under baseline/_ref (git-ignored, ships
# with the gpurun snapshot: `pip install --no-index --no-deps --target baseline/_ref <copy of /root/reference>`), which
# holds the ssmtoybox package only (no research/ scripts)
_HERE = os.path.dirname(os.path.abspath(__file__))
_CANDIDATES = ('/root/reference', os.path.join(os.path.dirname(_HERE), 'baseline', '_ref'))
REFERENCE_PATH = next((p for p in _CANDIDATES if os.path.isdir(os.path.join(p, 'ssmtoybox'))), _CANDIDATES[0])


def available():
    """Path of an importable copy of the unmodified reference, or None."""
    return REFERENCE_PATH if os.path.isdir(os.path.join(REFERENCE_PATH, 'ssmtoybox')) else None


def install():
    """Install the shim and put the reference on sys.path. Idempotent."""
    if getattr(install, '_done', False):
        return
    # 1. matplotlib stubs
    for name in ('matplotlib', 'matplotlib.pyplot', 'matplotlib.gridspec', 'matplotlib.lines'):
        if name not in sys.modules:
            try:
                __import__(name)
            except ImportError:
                sys.modules[name] = types.ModuleType(name)
    mpl = sys.modules['matplotlib']
    for sub in ('pyplot', 'gridspec', 'lines'):
        if not hasattr(mpl, sub):
            setattr(mpl, sub, sys.modules['matplotlib.' + sub])
    # 2. numpy aliases
    for alias, target in (('int', int), ('float', float), ('bool', bool)):
        if alias not in np.__dict__:
            setattr(np, alias, target)
    if not hasattr(np, 'asscalar'):
        np.asscalar = lambda a: np.asarray(a).item()
    if not hasattr(np, 'alltrue'):
        np.alltrue = np.all
    # 3. scipy.log10
    if not hasattr(scipy, 'log10'):
        scipy.log10 = np.log10
    # 4. factorial2(-1) == 1 (and (-1)!! for array input is not needed by the reference)
    _f2 = scipy.special.factorial2

    def factorial2(n, exact=False, **kw):
        if np.isscalar(n) and n == -1:
            return 1 if exact else 1.0
        return _f2(n, exact=exact, **kw)

    if not getattr(_f2, '_shimmed', False):
        factorial2._shimmed = True
        scipy.special.factorial2 = factorial2
    if REFERENCE_PATH not in sys.path:
        sys.path.insert(0, REFERENCE_PATH)
    # 5. BearingMeasurement sets dim_noise = None at class level and MeasurementModel.__init__ calls
    #    np.zeros(self.dim_noise) before the subclass assigns it (ssmod.py:901, 1176, 1186-1187): numpy >= 1.20 rejects
    #    None as a shape.  The class default is set to the 4 sensors of the default set-up; the instance attribute is
    #    still assigned by the reference's own constructor.
    from ssmtoybox import ssmod
    if ssmod.BearingMeasurement.dim_noise is None:
        ssmod.BearingMeasurement.dim_noise = 4
    install._done = True
