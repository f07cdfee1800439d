"""Test helpers that talk to the oracle (oracle/_ref, the unmodified reference).
Only tests/, smoke() and bench.py's cpu_baseline/reference arm may import this."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from classpp_public_b200.modules import Inputs  # noqa: E402

from classpp_public_b200.upstream import (PR_KEYS, BA_KEYS, TH_IN_KEYS, PT_KEYS, TR_KEYS, BG_KEYS, TH_KEYS,  # noqa: E402,F401
                                          collect_inputs)


def inputs_from_reference(ref, with_nonlinear=True):
    """Collect the upstream quantities (what the C++ drop-in reads from the reference's
    Input/Background/Thermodynamics/Primordial/Nonlinear modules) into an `Inputs`."""
    return collect_inputs(ref.get, ref.scalar)


def perturb_info_from_reference(ref):
    from classpp_public_b200 import _capi as capi
    info = capi.PerturbInfo()
    info.k_size = ref.iscalar("pt.k_size")
    info.k_size_cl = ref.iscalar("pt.k_size_cl")
    info.k_size_cmb = ref.iscalar("pt.k_size_cmb")
    info.tau_size = ref.iscalar("pt.tau_size")
    info.tp_size = ref.iscalar("pt.tp_size")
    info.ln_tau_size = ref.iscalar("pt.ln_tau_size")
    for n, flag in (("t0", "t"), ("t1", "t"), ("t2", "t"), ("p", "p"), ("delta_m", "delta_m"),
                    ("delta_cb", "delta_cb"), ("phi_plus_psi", "phi_plus_psi")):
        has = ref.iscalar("pt.has_source_" + flag)
        setattr(info, "index_tp_" + n, ref.iscalar("pt.index_tp_" + n) if has else -1)
    info.k_min = ref.scalar("pt.k_min")
    info.k_max = ref.scalar("pt.k_max")
    return info


def reference_sources(ref):
    ntp = ref.iscalar("pt.tp_size")
    return np.stack([ref.get("pt.sources.%d" % tp) for tp in range(ntp)])
