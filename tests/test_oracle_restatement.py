"""Pin the numpy restatement (oracle/restate.py) against the reference's own outputs (golden fixtures)."""
import numpy as np
import pytest

from oracle import restate


@pytest.mark.parametrize("name", ["lcdm_coarse", "lcdm", "planck18"])
def test_spectra_restatement_matches_reference_cl(golden, name):
    a = golden(name).arrays
    sizes = a["ref.sizes"].astype(int)
    ct_size = sizes[8]
    rows = a["ref.l_rows"].astype(int)
    tr = a["ref.transfer_rows"]  # [tt, rows, q]
    # transfer type order of the reference for tCl,pCl,lCl: t2, e, t0, t1, lcmb
    idx = {"t2": 0, "e": 1, "t0": 2, "t1": 3, "lcmb": 4}
    cl = restate.spectra_cl(a["ref.kq"], a["pm.pk_at_transfer_k"], tr, idx)
    ref = a["ref.cl"].reshape(-1, ct_size)[rows]
    for c in range(ct_size):
        scale = np.max(np.abs(ref[:, c]))
        if scale > 0:
            assert np.max(np.abs(cl[:, c] - ref[:, c])) <= 1e-11 * scale


def test_spline_restatement_reproduces_private_table(reference):
    if reference is None:
        pytest.skip("oracle/_ref not built")
    ref = reference("lcdm_coarse", "thermodynamics")
    n, m = ref.iscalar("bg.bt_size"), ref.iscalar("bg.bg_size")
    tau = ref.get("bg.tau_table")
    y = ref.get("bg.background_table").reshape(n, m)
    d2 = ref.get("bg.d2background_dtau2_table").reshape(n, m)
    mine = restate.spline_est_deriv(tau, y)
    assert np.array_equal(mine, d2)  # same recurrences in the same order: bit-identical


def test_trapezoid_weights_integrate_linear_function_exactly():
    x = np.sort(np.random.default_rng(0).uniform(0, 10, 50))[::-1].copy()  # decreasing like tau0 - tau
    w = restate.trapezoidal_mweights(x)
    f = 3.0 * x + 1.0
    exact = 1.5 * (x[0] ** 2 - x[-1] ** 2) + (x[0] - x[-1])
    assert abs(np.sum(w * f) - exact) < 1e-10
