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


@pytest.mark.parametrize("name", ["lcdm_coarse", "planck18"])
def test_lensing_restatement_matches_reference_lensed_cl(golden, name):
    """oracle/restate.lensed_cl (numpy, generic Wigner-d recurrence) against the reference's lensing_cl_at_l at every
    l = 2..l_lensed_max, starting from the reference's own unlensed table: pins the restatement of the lensing stage."""
    a = golden(name).arrays
    ct = int(a["ref.sizes"][8])
    cl = a["ref.cl"].reshape(-1, ct)
    idx = {"tt": 0, "ee": 1, "te": 2, "bb": 3, "pp": 4}  # spectra_indices order for tCl,pCl,lCl (spectra_module.cpp:560-640)
    l_lens, cl_lens, l_lensed_max = restate.lensed_cl(a["ref.l"], cl, idx, 500, int(a["ref.l"][-1]))
    ref = a["ref.cl_lensed"].reshape(-1, ct)
    assert l_lensed_max == ref.shape[0] - 1
    ls = np.arange(2, l_lensed_max + 1)
    mine = restate.spline_eval(l_lens, cl_lens, restate.spline_est_deriv(l_lens, cl_lens), ls)
    r = ref[2:]
    assert np.max(np.abs(mine[:, 0] / r[:, 0] - 1.0)) < 1e-9
    assert np.max(np.abs(mine[:, 1] / r[:, 1] - 1.0)) < 1e-9
    assert np.max(np.abs(mine[:, 2] - r[:, 2]) / np.sqrt(r[:, 0] * r[:, 1])) < 1e-9
    assert np.max(np.abs(mine[:, 3] / r[:, 3] - 1.0)) < 1e-8
    assert np.max(np.abs(mine[:, 4] / r[:, 4] - 1.0)) < 1e-12


def test_wigner_d_restatement_closed_forms():
    mu = np.linspace(-0.9, 1.0, 7)
    d = restate.wigner_d(0, 0, mu, 3)
    assert np.allclose(d[:, 2], 0.5 * (3 * mu ** 2 - 1), atol=1e-14)         # Legendre P_2
    assert np.allclose(d[:, 3], 0.5 * (5 * mu ** 3 - 3 * mu), atol=1e-14)    # Legendre P_3
    d = restate.wigner_d(1, 1, mu, 2)
    assert np.allclose(d[:, 2], (1 + mu) / 2 * (2 * mu - 1), atol=1e-14)
    d = restate.wigner_d(2, -2, mu, 2)
    assert np.allclose(d[:, 2], (1 - mu) ** 2 / 4, atol=1e-14)
    d = restate.wigner_d(2, 0, mu, 2)
    assert np.allclose(d[:, 2], np.sqrt(6.0) / 4 * (1 - mu ** 2), atol=1e-14)
