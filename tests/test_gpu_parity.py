"""GPU parity tests (run on the B200 box): the CUDA path, called through the C ABI, against
 (1) the committed golden vectors generated from the reference,
 (2) the numpy restatement (oracle/restate.py), and
 (3) the live reference (oracle/_ref) for stage-isolated checks, when it is loadable.
Tolerances: the north-star bar is 1e-4 relative on C_l^{TT,EE,phiphi} (cross spectra normalised by
sqrt(C^XX C^YY), like the reference's own test, python/test_class.py:494-507)."""
import os

import numpy as np
import pytest

from classpp_public_b200 import modules as M
from classpp_public_b200 import _capi as capi
from oracle import restate

pytestmark = pytest.mark.gpu

CL_RTOL = 1e-4


class _NL:
    def __init__(self, arr):
        self.nl_corr_density_m = arr


def run_pipeline(inp):
    a = inp.arrays
    ctx = M.Context(0)
    bg = M.BackgroundModule(inp, ctx)
    th = M.ThermodynamicsModule(inp, bg)
    pt = M.PerturbationsModule(inp, bg, th)
    nl = _NL(a["nl.nl_corr_density_m"]) if "nl.nl_corr_density_m" in a else None
    tr = M.TransferModule(inp, bg, th, pt, nl)
    sp = M.SpectraModule(inp, pt, M.TabulatedPrimordial(a["pm.pk_at_transfer_k"]), nl, tr)
    return ctx, pt, tr, sp


def check_cl(sp, ref_cl, rtol=CL_RTOL):
    ct = sp.ct_size_
    cl = sp.cl_[0].reshape(-1, ct)
    ref = ref_cl.reshape(-1, ct)
    i = sp.info
    for name in ("tt", "ee", "pp"):
        c = getattr(i, "index_ct_" + name)
        if c >= 0:
            assert np.max(np.abs(cl[:, c] / ref[:, c] - 1.0)) < rtol, name
    for name, (x, y) in {"te": ("tt", "ee"), "tp": ("tt", "pp"), "ep": ("ee", "pp")}.items():
        c = getattr(i, "index_ct_" + name)
        if c >= 0:
            norm = np.sqrt(ref[:, getattr(i, "index_ct_" + x)] * ref[:, getattr(i, "index_ct_" + y)])
            assert np.max(np.abs(cl[:, c] - ref[:, c]) / norm) < rtol, name
    if i.index_ct_bb >= 0:
        assert np.all(cl[:, i.index_ct_bb] == 0.0)


# lcdm_coarse is not a BASELINE config: its coarse k/tau sampling amplifies the ODE tolerance (the reference's
# own C_l^TT moves by 5e-5 when tol_perturb_integration is halved there, vs 1e-5 for the real configs), so it
# gets 5e-4; the BASELINE configurations are held to the north-star 1e-4.
@pytest.mark.parametrize("name,rtol", [("lcdm_coarse", 5e-4), ("lcdm", CL_RTOL)])
def test_full_pipeline_cl_vs_golden(golden, name, rtol):
    inp = golden(name)
    ctx, pt, tr, sp = run_pipeline(inp)
    a = inp.arrays
    # grids coming back through the GPU context are still the bit-exact host grids
    assert np.array_equal(pt.k_[0], a["ref.k"]) and np.array_equal(pt.tau_sampling_, a["ref.tau"])
    check_cl(sp, a["ref.cl"], rtol)
    # sources: sub-sampled k columns. Each column is compared relative to its own maximum over tau;
    # the integration tolerance (tol_perturb_integration = 1e-5 on the state vector) is amplified by the
    # cancellations inside S_T0/S_T1, hence the looser 1e-2 here -- the C_l check above is the contract.
    nk, nt, ntp = pt.info.k_size, pt.info.tau_size, pt.info.tp_size
    mine = np.stack(pt.sources_[0]).reshape(ntp, nt, nk)[:, :, a["ref.k_cols"].astype(int)]
    ref = a["ref.sources_cols"]
    scale = np.max(np.abs(ref), axis=1, keepdims=True)
    assert np.max(np.abs(mine - ref) / np.where(scale > 0, scale, 1.0)) < 1e-2
    # delta_m and phi+psi (no cancellation) are at the integrator tolerance
    for tp in (pt.info.index_tp_delta_m, pt.info.index_tp_phi_plus_psi):
        assert np.max(np.abs(mine[tp] - ref[tp]) / scale[tp]) < 1e-4
    # transfer functions: sub-sampled l rows, relative to the row maximum
    ti = tr.info
    t = tr.transfer_[0].reshape(ti.tt_size, ti.l_size, ti.q_size)[:, a["ref.l_rows"].astype(int), :]
    rt = a["ref.transfer_rows"]
    s = np.max(np.abs(rt), axis=2, keepdims=True)
    assert np.max(np.abs(t - rt) / np.where(s > 0, s, 1.0)) < 5e-3
    # work counters: the device NDF15 takes the same decisions as the reference's evolver
    ks = pt.kstat_
    assert np.all(ks[:, 7] == 0)
    ctx.close()


def test_planck18_ncdm_halofit_pipeline_vs_golden(golden):
    """BASELINE config 2: massive neutrino hierarchy (neq up to 136) + halofit correction of phi+psi."""
    inp = golden("planck18")
    ctx, pt, tr, sp = run_pipeline(inp)
    check_cl(sp, inp.arrays["ref.cl"])
    assert pt.info.tp_size == 7 and pt.info.index_tp_delta_cb >= 0
    ctx.close()


def test_dense_precision_config_vs_golden(golden):
    """BASELINE config 3 stand-in (cl_permille.pre is not in the reference tree): denser k sampling (1200 modes), photon /
    ur hierarchies up to l = 30, tol_perturb_integration = 1e-6, half the time-sampling step, finer l and q grids.
    Unlensed C_l within 1e-4 and lensed TT/EE within 1e-4 of the reference."""
    inp = golden("lcdm_dense")
    a = inp.arrays
    ctx, pt, tr, sp = run_pipeline(inp)
    assert np.array_equal(pt.k_[0], a["ref.k"]) and np.array_equal(pt.tau_sampling_, a["ref.tau"])
    assert pt.info.k_size == 1200 and int(pt.kprofile_[:, 0, :].max()) > 80  # equations of the largest interval
    check_cl(sp, a["ref.cl"])
    le = M.LensingModule(inp, sp)
    ref = a["ref.cl_lensed"].reshape(-1, le.lt_size_)
    mine = np.array([le.lensing_cl_at_l(l) for l in range(2, le.l_lensed_max_ + 1)])
    for c in (le.index_lt_tt_, le.index_lt_ee_):
        assert np.max(np.abs(mine[:, c] / ref[2:, c] - 1.0)) < CL_RTOL
    ctx.close()


def test_massive_neutrinos_degenerate_pipeline_vs_golden(golden):
    """BASELINE config 4 (degenerate form): 3 degenerate massive neutrinos, m = 0.02 eV."""
    inp = golden("ncdm3_deg")
    ctx, pt, tr, sp = run_pipeline(inp)
    check_cl(sp, inp.arrays["ref.cl"])
    assert np.array_equal(pt.k_[0], inp.arrays["ref.k"]) and np.array_equal(pt.tau_sampling_, inp.arrays["ref.tau"])
    ctx.close()


def test_three_ncdm_species_large_system_vs_golden(golden):
    """Three separate ncdm species (316 equations, 58 hub variables, 18 chains) on the coarse grids: the
    large-system fallbacks of the device path (generic NDF, shared-memory Gauss-Jordan)."""
    inp = golden("ncdm3_coarse")
    ctx, pt, tr, sp = run_pipeline(inp)
    check_cl(sp, inp.arrays["ref.cl"], rtol=5e-4)  # coarse grids, see lcdm_coarse
    assert int(pt.kprofile_[:, 0, :].max()) == 316
    ctx.close()


@pytest.mark.parametrize("name", ["lcdm_coarse", "planck18"])
def test_spectra_stage_vs_numpy_restatement(golden, name):
    """Stage 3 in isolation on random transfer functions: CUDA quadrature == numpy restatement of
    array_spline + array_integrate_all_trapzd_or_spline (bit-level agreement up to summation order)."""
    inp = golden(name)
    ctx = M.Context(0)
    bg = M.BackgroundModule(inp, ctx)
    th = M.ThermodynamicsModule(inp, bg)
    pt = M.PerturbationsModule(inp, bg, th, solve=False)
    tr = M.TransferModule(inp, bg, th, pt, compute=False)
    ti = tr.info
    rng = np.random.default_rng(1)
    transfer = rng.standard_normal((ti.tt_size, ti.l_size, ti.q_size)) * 1e-3
    tr.set_transfer(transfer)
    pk = inp.arrays["pm.pk_at_transfer_k"]
    sp = M.SpectraModule(inp, pt, M.TabulatedPrimordial(pk), None, tr)
    idx = {n: getattr(ti, "index_tt_" + n) for n in ("t0", "t1", "t2", "e", "lcmb")}
    ref = restate.spectra_cl(tr.k_[0], pk, transfer, idx)
    cl = sp.cl_[0].reshape(-1, sp.ct_size_)
    scale = np.max(np.abs(ref), axis=0)
    assert np.max(np.abs(cl - ref) / np.where(scale > 0, scale, 1)) < 1e-11
    # linearity in P(k): doubling the primordial spectrum doubles every C_l (size-independent property)
    sp2 = M.SpectraModule(inp, pt, M.TabulatedPrimordial(2 * pk), None, tr)
    assert np.allclose(sp2.cl_[0], 2 * sp.cl_[0], rtol=1e-14, atol=0)
    # q-range partial sums add up to the full quadrature (multi-GPU partition)
    half = ti.q_size // 2
    p1 = M.SpectraModule(inp, pt, M.TabulatedPrimordial(pk), None, tr, q_range=(0, half))
    p2 = M.SpectraModule(inp, pt, M.TabulatedPrimordial(pk), None, tr, q_range=(half, ti.q_size))
    assert np.allclose(p1.cl_[0] + p2.cl_[0], sp.cl_[0], rtol=1e-12, atol=1e-30)
    ctx.close()


def test_bessel_table_vs_scipy(golden):
    """Device recurrence tables Phi_l(x)=j_l(x), Phi_l'(x) against scipy's spherical_jn."""
    inp = golden("lcdm_coarse")
    ctx, pt, tr, sp = run_pipeline(inp)
    x, phi, dphi, chi = tr.bessel_table()
    sel = np.unique(np.linspace(0, len(x) - 1, 400).astype(int))
    jl, djl = restate.bessel_table(tr.l_, x[sel])
    # the reference's tables are only accurate in absolute terms (|j_l| <= 1)
    assert np.max(np.abs(phi[:, sel] - jl)) < 1e-12
    assert np.max(np.abs(dphi[:, sel] - djl)) < 1e-12
    assert np.allclose(chi, restate.chi_at_phimin(tr.l_, inp.meta["pr.hyper_phi_min_abs"]), rtol=1e-13)
    ctx.close()


def test_stage_isolation_vs_live_reference(reference):
    """Transfer stage fed with the reference's S(k,tau) and spectra stage fed with the reference's
    Delta_l(q): isolates each kernel's error from the ODE tolerance."""
    if reference is None:
        pytest.skip("oracle/_ref not loadable on this box")
    from refutil import inputs_from_reference, perturb_info_from_reference, reference_sources
    ref = reference("lcdm_coarse", "lensing")
    inp = inputs_from_reference(ref)
    ctx = M.Context(0)
    bg = M.BackgroundModule(inp, ctx)
    th = M.ThermodynamicsModule(inp, bg)
    pt = M.PerturbationsModule.from_sources(inp, bg, ref.get("pt.k"), ref.get("pt.tau_sampling"),
                                            reference_sources(ref), perturb_info_from_reference(ref))
    tr = M.TransferModule(inp, bg, th, pt, None)
    ti = tr.info
    mine = tr.transfer_[0].reshape(ti.tt_size, ti.l_size, ti.q_size)
    r = ref.get("tr.transfer").reshape(mine.shape)
    for tt in range(ti.tt_size):
        assert np.max(np.abs(mine[tt] - r[tt])) < 1e-9 * np.max(np.abs(r[tt]))
    sp = M.SpectraModule(inp, pt, M.TabulatedPrimordial(ref.get("pm.pk_at_transfer_k")), None, tr)
    check_cl(sp, ref.get("sp.cl"), rtol=1e-9)
    # spectra_cl_at_l / cl_output: the l-spline of the table, at every integer l, against the reference's own
    lmax = 500
    ref_at_l = ref.get("sp.cl_at_l.%d" % lmax).reshape(lmax + 1, sp.ct_size_)
    out = sp.cl_output(lmax)
    for name in ("tt", "ee", "pp"):
        col = ref_at_l[:, getattr(sp.info, "index_ct_" + name)]
        assert out[name][0] == 0.0 and out[name][1] == 0.0
        assert np.max(np.abs(out[name][2:] / col[2:] - 1.0)) < 1e-8, name
    assert np.all(sp.spectra_cl_at_l(sp.l_max_tot_ + 1) == 0.0)
    with pytest.raises(M.CosmoComputationError):
        sp.cl_output(sp.l_max_tot_ + 1)
    ctx.close()


def test_sweep_of_different_cosmologies_in_one_batch_vs_live_reference():
    """BASELINE config 5 in miniature: a Latin-hypercube sweep of 6 DIFFERENT cosmologies (different tables,
    different k/tau/q grid sizes) pushed through ONE batched perturbation launch; every cosmology must match the
    reference run with the same parameters (coarse grids: 5e-4, see lcdm_coarse)."""
    from oracle import refprobe
    if not refprobe.available():
        pytest.skip("oracle/_ref not loadable on this box")
    import os
    from scipy.stats import qmc
    from classpp_public_b200.configs import CONFIGS
    from refutil import inputs_from_reference
    lo = np.array([0.020, 0.10, 0.60, 2.9, 0.92, 0.03])   # omega_b, omega_cdm, h, ln(1e10 A_s), n_s, tau_reio
    hi = np.array([0.024, 0.14, 0.75, 3.2, 1.00, 0.09])
    pts_lhs = qmc.scale(qmc.LatinHypercube(d=6, seed=0).random(6), lo, hi)
    refs, inps = [], []
    for x in pts_lhs:
        par = dict(CONFIGS["lcdm_coarse"], omega_b=x[0], omega_cdm=x[1], h=x[2], A_s=1e-10 * np.exp(x[3]), n_s=x[4],
                   tau_reio=x[5])
        ref = refprobe.RefCosmology(par, threads=os.cpu_count()).compute("spectra")
        refs.append(ref)
        inps.append(inputs_from_reference(ref))
    ctxs, mods, pts = [], [], []
    for inp in inps:
        c = M.Context(0)
        b = M.BackgroundModule(inp, c)
        t = M.ThermodynamicsModule(inp, b)
        ctxs.append(c)
        mods.append((b, t))
        pts.append(M.PerturbationsModule(inp, b, t, solve=False))
    assert len({p.info.k_size for p in pts} | {p.info.tau_size for p in pts}) > 2  # the grids really differ
    M.PerturbationsModule.solve_batch(pts)
    for inp, ref, (b, t), p in zip(inps, refs, mods, pts):
        assert np.array_equal(p.k_[0], ref.get("pt.k")) and np.array_equal(p.tau_sampling_, ref.get("pt.tau_sampling"))
        tr = M.TransferModule(inp, b, t, p, None)
        sp = M.SpectraModule(inp, p, M.TabulatedPrimordial(ref.get("pm.pk_at_transfer_k")), None, tr)
        check_cl(sp, ref.get("sp.cl"), rtol=5e-4)
    for c in ctxs:
        c.close()
    for r in refs:
        r.close()


def test_rk_evolver_vs_ndf15_golden(golden):
    """evolver = rk: the device Cash-Karp integrator.  The unmodified reference SEGFAULTS with
    `evolver = 0` in this container (oracle run), so this row has no reference output to pin against:
    the explicit integrator is checked against the reference's ndf15 result for the same inputs instead (both
    integrate the same equations to tol_perturb_integration)."""
    from classpp_public_b200.modules import Inputs
    base = golden("lcdm_coarse")
    inp = Inputs(dict(base.meta, **{"pr.evolver": 0}), base.arrays)
    ctx, pt, tr, sp = run_pipeline(inp)
    check_cl(sp, base.arrays["ref.cl"], rtol=5e-4)  # coarse grids, see lcdm_coarse
    ks = pt.kstat_
    assert np.all(ks[:, 7] == 0) and ks[:, 3].sum() == 0 and ks[:, 4].sum() == 0  # explicit: no Jacobian, no LU
    ctx.close()


def test_linear_matter_power_spectrum_vs_live_reference(reference):
    """P(k) (north star: within 1e-4 of the reference): 2 pi^2/k^3 P_R(k) delta_m^2 from the device-resident sources
    against NonlinearModule's linear P(k, z=0) at the nodes of the perturbation k grid."""
    if reference is None:
        pytest.skip("oracle/_ref not loadable on this box")
    from refutil import inputs_from_reference
    for name in ("lcdm_coarse", "ncdm3_deg"):
        ref = reference(name, "nonlinear")
        inp = inputs_from_reference(ref)
        ctx = M.Context(0)
        bg = M.BackgroundModule(inp, ctx)
        th = M.ThermodynamicsModule(inp, bg)
        pt = M.PerturbationsModule(inp, bg, th)
        pr = ref.get("pm.pk_at_pt_k")
        pk = pt.pk_linear(pr)
        rpk = ref.get("nl.pk_lin_m_at_pt_k")
        assert np.max(np.abs(pk / rpk - 1.0)) < 1e-4, name
        if pt.info.index_tp_delta_cb >= 0:
            rcb = ref.get("nl.pk_lin_cb_at_pt_k")
            assert np.max(np.abs(pt.pk_linear(pr, cb=True) / rcb - 1.0)) < 1e-4, name
        ctx.close()


def test_perturb_sources_at_tau(golden):
    """PerturbationsModule::perturb_sources_at_tau (z_max_pk = 0 branch: linear in tau,
    array_interpolate_two_bis). The unmodified reference cannot be run as the checker here: with z_max_pk = 0 it reads
    ln_tau_[0] of a never-allocated ln_tau_ (perturbations_module.cpp:95 vs :1555) and segfaults, so the check is the
    defining formula on our own table (bit-identical), whose columns are pinned to the reference by the golden test, plus
    the reference's out-of-range failure."""
    inp = golden("lcdm_coarse")
    a = inp.arrays
    ctx = M.Context(0)
    bg = M.BackgroundModule(inp, ctx)
    th = M.ThermodynamicsModule(inp, bg)
    pt = M.PerturbationsModule(inp, bg, th)
    tau_s, nk = pt.tau_sampling_, pt.info.k_size
    cols = a["ref.k_cols"].astype(int)
    for tp in (pt.info.index_tp_delta_m, pt.info.index_tp_phi_plus_psi):
        tab = pt.sources_[0][tp].reshape(-1, nk)
        for j, frac in ((3, 0.25), (len(tau_s) // 2, 0.5), (len(tau_s) - 2, 0.9)):
            tau = tau_s[j] + frac * (tau_s[j + 1] - tau_s[j])
            mine = pt.perturb_sources_at_tau(0, 0, tp, tau)
            w = (tau - tau_s[j]) / (tau_s[j + 1] - tau_s[j])
            assert np.array_equal(mine, tab[j] * (1.0 - w) + w * tab[j + 1])
        at_node = pt.perturb_sources_at_tau(0, 0, tp, tau_s[-1])
        assert np.array_equal(at_node, tab[-1])
        r = a["ref.sources_cols"][tp][-1]  # the reference's S(k, tau_0) at the sub-sampled k columns
        assert np.max(np.abs(at_node[cols] - r)) < 1e-3 * np.max(np.abs(r))
    with pytest.raises(M.CosmoComputationError, match="x_max"):
        pt.perturb_sources_at_tau(0, 0, 0, tau_s[-1] * 1.01)
    with pytest.raises(M.CosmoSevereError):
        pt.perturb_sources_at_tau(1, 0, 0, tau_s[-1])
    ctx.close()


def test_halofit_on_device_vs_golden(golden):
    """`non linear = halofit` (BASELINE config 2) with the NonlinearModule step on the device: the
    correction table R_NL(k,tau) against the reference's nl_corr_density_, and the C_l of the fully device-resident
    pipeline perturbations -> halofit -> transfer -> spectra against the reference's."""
    from classpp_public_b200.configs import CONFIGS
    inp = golden("planck18")
    a = inp.arrays
    par = CONFIGS["planck18"]
    prim_k = M.AnalyticPrimordial(par["A_s"], par["n_s"])  # the reference's analytic P_R(k), k_pivot = 0.05/Mpc
    ctx = M.Context(0)
    bg = M.BackgroundModule(inp, ctx)
    th = M.ThermodynamicsModule(inp, bg)
    pt = M.PerturbationsModule(inp, bg, th)
    nl = M.NonlinearModule(inp, bg, pt, prim_k, fetch=True)
    mine = nl.nl_corr_density_[0].reshape(pt.info.tau_size, pt.info.k_size)
    ref = a["nl.nl_corr_density_m"].reshape(pt.info.tau_size, pt.info.k_size)
    assert np.array_equal(mine == 1.0, ref == 1.0)  # same "not computable at this redshift" region, same linear k range
    assert np.max(np.abs(mine / ref - 1.0)) < 1e-4
    tr = M.TransferModule(inp, bg, th, pt, nl)  # correction applied from the device-resident table
    sp = M.SpectraModule(inp, pt, M.TabulatedPrimordial(a["pm.pk_at_transfer_k"]), nl, tr)
    check_cl(sp, a["ref.cl"])
    ctx.close()


@pytest.mark.parametrize("name,rtol", [("lcdm_coarse", 5e-4), ("lcdm", CL_RTOL), ("planck18", CL_RTOL)])
def test_lensed_cl_on_device_vs_golden(golden, name, rtol):
    """LensingModule on the device (clpp_lensing_compute + clpp_lensing_cl_at_l) against the
    reference's lensing_cl_at_l at every l = 2..l_lensed_max (fast mode, the default). The lensing operator itself
    (lensed/unlensed ratio, which cancels the 1e-4-level differences of the unlensed input) is held to 2e-5."""
    inp = golden(name)
    a = inp.arrays
    ctx, pt, tr, sp = run_pipeline(inp)
    le = M.LensingModule(inp, sp)
    lt = le.lt_size_
    ref = a["ref.cl_lensed"].reshape(-1, lt)
    assert le.l_lensed_max_ == ref.shape[0] - 1 and le.l_unlensed_max_ == sp.l_max_tot_
    mine = np.zeros_like(ref)
    for l in range(2, le.l_lensed_max_ + 1):
        mine[l] = le.lensing_cl_at_l(l)
    m, r = mine[2:], ref[2:]
    tt, ee, te, bb = le.index_lt_tt_, le.index_lt_ee_, le.index_lt_te_, le.index_lt_bb_
    assert np.max(np.abs(m[:, tt] / r[:, tt] - 1.0)) < rtol
    assert np.max(np.abs(m[:, ee] / r[:, ee] - 1.0)) < rtol
    assert np.max(np.abs(m[:, te] - r[:, te]) / np.sqrt(r[:, tt] * r[:, ee])) < rtol
    assert np.all(r[:, bb] > 0) and np.max(np.abs(m[:, bb] / r[:, bb] - 1.0)) < 2 * rtol
    # types lensing leaves alone: unlensed values on the (truncated) lensing l grid, splined there like the reference
    pp, tp, ep = le.index_lt_pp_, le.index_lt_tp_, le.index_lt_ep_
    assert np.max(np.abs(m[:, pp] / r[:, pp] - 1.0)) < rtol
    assert np.max(np.abs(m[:, tp] - r[:, tp]) / np.sqrt(r[:, tt] * r[:, pp])) < rtol
    assert np.max(np.abs(m[:, ep] - r[:, ep]) / np.sqrt(r[:, ee] * r[:, pp])) < rtol
    assert np.array_equal(le.cl_lens_.reshape(-1, lt)[:, pp], sp.cl_[0].reshape(-1, lt)[: le.l_size_, pp])
    # the operator: lensed/unlensed at the grid multipoles, ours vs the reference's
    lgrid = sp.l_.astype(int)
    ref_unl = a["ref.cl"].reshape(-1, lt)
    rows = np.nonzero((lgrid >= 2) & (lgrid <= le.l_lensed_max_))[0]
    sel = lgrid[rows]
    for c in (tt, ee):
        ratio_mine = mine[sel, c] / sp.cl_[0].reshape(-1, lt)[rows, c]
        ratio_ref = ref[sel, c] / ref_unl[rows, c]
        assert np.max(np.abs(ratio_mine / ratio_ref - 1.0)) < 2e-5
    # stage isolation: the numpy restatement of lensing_init fed with OUR unlensed table (oracle/restate.py, itself pinned
    # to the reference at 1e-9 by tests/test_oracle_restatement.py) -> the device kernels agree to rounding
    from oracle import restate
    idx = {n: getattr(le, "index_lt_%s_" % n) for n in ("tt", "ee", "te", "bb", "pp")}
    l_np, cl_np, _ = restate.lensed_cl(sp.l_, sp.cl_[0].reshape(-1, lt), idx, 500, int(inp.meta["pt.l_scalar_max"]))
    assert np.array_equal(l_np, le.l_)
    dev = le.cl_lens_.reshape(-1, lt)
    for c in (tt, ee, bb):
        assert np.max(np.abs(dev[:, c] / cl_np[:, c] - 1.0)) < 1e-9
    assert np.max(np.abs(dev[:, te] - cl_np[:, te]) / np.sqrt(cl_np[:, tt] * cl_np[:, ee])) < 1e-9
    with pytest.raises(M.CosmoComputationError, match="you asked for lensed Cls at l="):
        le.lensing_cl_at_l(le.l_lensed_max_ + 1)
    assert ctx.kernel_ms()["lensing"] > 0
    ctx.close()


def test_accurate_lensing_mode_vs_live_reference():
    """accurate_lensing = 1 (Gauss-Legendre nodes on [-1,1], no unlensed subtraction; lensing_module.cpp:237-248)
    against the unmodified reference run with the same switch, and with a non-default delta_l_max."""
    from oracle import refprobe
    if not refprobe.available():
        pytest.skip("oracle/_ref not loadable on this box")
    import os
    from classpp_public_b200.configs import CONFIGS
    from refutil import inputs_from_reference
    par = dict(CONFIGS["lcdm_coarse"], accurate_lensing=1, delta_l_max=300)
    ref = refprobe.RefCosmology(par, threads=os.cpu_count()).compute("lensing")
    inp = inputs_from_reference(ref)
    assert int(inp.meta["pr.accurate_lensing"]) == 1 and int(inp.meta["pr.delta_l_max"]) == 300
    ctx = M.Context(0)
    bg = M.BackgroundModule(inp, ctx)
    th = M.ThermodynamicsModule(inp, bg)
    pt = M.PerturbationsModule(inp, bg, th)
    tr = M.TransferModule(inp, bg, th, pt, None)
    sp = M.SpectraModule(inp, pt, M.TabulatedPrimordial(ref.get("pm.pk_at_transfer_k")), None, tr)
    le = M.LensingModule(inp, sp)
    lt = le.lt_size_
    r = ref.get("le.cl_lensed").reshape(-1, lt)[2:]
    assert le.l_lensed_max_ == ref.iscalar("le.l_lensed_max") == le.l_unlensed_max_ - 300
    m = np.array([le.lensing_cl_at_l(l) for l in range(2, le.l_lensed_max_ + 1)])
    tt, ee, te, bb = le.index_lt_tt_, le.index_lt_ee_, le.index_lt_te_, le.index_lt_bb_
    assert np.max(np.abs(m[:, tt] / r[:, tt] - 1.0)) < 5e-4
    assert np.max(np.abs(m[:, ee] / r[:, ee] - 1.0)) < 5e-4
    assert np.max(np.abs(m[:, te] - r[:, te]) / np.sqrt(r[:, tt] * r[:, ee])) < 5e-4
    assert np.max(np.abs(m[:, bb] / r[:, bb] - 1.0)) < 1e-3
    ctx.close()
    ref.close()


def test_k_range_partition_equals_full_solve(golden):
    """Multi-GPU partition property: integrating two k ranges separately fills the same source table."""
    inp = golden("lcdm_coarse")
    ctx = M.Context(0)
    bg = M.BackgroundModule(inp, ctx)
    th = M.ThermodynamicsModule(inp, bg)
    full = M.PerturbationsModule(inp, bg, th)
    s_full = np.stack(full.sources_[0])
    ctx2 = M.Context(0)
    bg2 = M.BackgroundModule(inp, ctx2)
    th2 = M.ThermodynamicsModule(inp, bg2)
    nk = full.info.k_size
    part = M.PerturbationsModule(inp, bg2, th2, k_range=(0, nk // 2))
    L = ctx2._lib
    import ctypes as C
    ctx2.check(L.clpp_perturb_solve(ctx2.handle, nk // 2, nk, ctx2.err))
    part._sources = None
    s_part = np.stack(part.sources_[0])
    assert np.array_equal(s_full, s_part)  # deterministic: bit-identical
    ctx.close(); ctx2.close()


def test_batched_solve_equals_individual_solves(golden):
    """clpp_perturb_solve_batch: several cosmologies in one launch give, per cosmology, bit-identical sources to
    one launch per cosmology (independent objects; the batch only changes the issue order of the modes)."""
    inp = golden("lcdm_coarse")
    single_ctx = M.Context(0)
    bg = M.BackgroundModule(inp, single_ctx)
    th = M.ThermodynamicsModule(inp, bg)
    single = M.PerturbationsModule(inp, bg, th)
    s_single = np.stack(single.sources_[0])
    ctxs, pts = [], []
    for _ in range(3):
        c = M.Context(0)
        b = M.BackgroundModule(inp, c)
        t = M.ThermodynamicsModule(inp, b)
        ctxs.append(c)
        pts.append(M.PerturbationsModule(inp, b, t, solve=False))
    M.PerturbationsModule.solve_batch(pts)
    for p in pts:
        assert np.array_equal(np.stack(p.sources_[0]), s_single)
        assert np.array_equal(p.kstat_[:, :6], single.kstat_[:, :6])
    for c in ctxs + [single_ctx]:
        c.close()


def test_cohorts_of_modes_per_cta_are_bit_identical(golden, monkeypatch):
    """Cohorts (several k modes per CTA, one warp each, CTA barrier at every step) only change WHEN a warp runs, never
    what it computes: sources and work counters are bit-identical to the one-mode-per-CTA launch, for a batch (same k
    of neighbouring cosmologies side by side) and for a single cosmology (adjacent k in one cohort, ragged last CTA)."""
    inp = golden("lcdm_coarse")
    out = {}
    for W in ("1", "4", "3"):
        monkeypatch.setenv("CLPP_COHORT", W)
        monkeypatch.setenv("CLPP_COHORT_LONG", W)
        ctxs, pts = [], []
        for _ in range(5):
            c = M.Context(0)
            b = M.BackgroundModule(inp, c)
            t = M.ThermodynamicsModule(inp, b)
            ctxs.append(c)
            pts.append(M.PerturbationsModule(inp, b, t, solve=False))
        M.PerturbationsModule.solve_batch(pts)
        single = M.PerturbationsModule(inp, *_tables(inp, ctxs[0]))  # one cosmology, cohorts of adjacent k
        out[W] = (np.stack(pts[3].sources_[0]), pts[3].kstat_[:, :6].copy(), np.stack(single.sources_[0]))
        for p in pts:
            assert np.array_equal(np.stack(p.sources_[0]), out[W][0])
        for c in ctxs:
            c.close()
    for W in ("4", "3"):
        assert np.array_equal(out[W][0], out["1"][0]) and np.array_equal(out[W][1], out["1"][1])
        assert np.array_equal(out[W][2], out["1"][2])
    assert np.array_equal(out["1"][0], out["1"][2])


def _tables(inp, ctx):
    bg = M.BackgroundModule(inp, ctx)
    return bg, M.ThermodynamicsModule(inp, bg)


def test_sweep_pipeline_equals_direct_calls(golden):
    """sweep.SweepPipeline (two context sets, per-cosmology stages of batch i under the launch of batch i+1) returns, for
    every batch, exactly what the direct module calls return."""
    from classpp_public_b200.sweep import SweepPipeline
    inp = golden("lcdm_coarse")
    a = inp.arrays
    ctx, pt, tr, sp = run_pipeline(inp)
    cl_direct = sp.cl_[0].copy()
    ctx.close()
    B, NSET = 3, 2
    sets = [[_tables(inp, M.Context(0)) for _ in range(B)] for _ in range(NSET)]

    def front(s, b):
        return M.PerturbationsModule(inp, sets[s][b][0], sets[s][b][1], solve=False)

    def back(s, b, p):
        t = M.TransferModule(inp, sets[s][b][0], sets[s][b][1], p, None)
        return M.SpectraModule(inp, p, M.TabulatedPrimordial(a["pm.pk_at_transfer_k"]), None, t).cl_[0].copy()

    pipe = SweepPipeline(B, front, back, n_sets=NSET)
    for _ in range(3):
        pipe.submit()
    res = pipe.drain()
    assert len(res) == NSET and len(pipe.solve_seconds) == 3
    for batch in res:
        assert len(batch) == B
        for cl in batch:
            assert np.array_equal(cl, cl_direct)
    pipe.close()
    for s in sets:
        for bg, th in s:
            bg.ctx.close()


def test_integrator_variants_agree(golden, monkeypatch):
    """Three implementations of the same NDF15 algorithm: the lane kernel (one thread per mode, the default), the
    warp-per-mode kernels with the register tail kernel (CLPP_WARP_PATH=1) and the generic shared-memory NDF
    (CLPP_GENERIC_ONLY=1).  They are not bit-identical (different summation orders flip an occasional accept/reject or
    order decision of the step controller): the two warp variants agree to 2e-5 in C_l, five times tighter than the
    parity bar; the lane kernel (different Jacobian probing and hub algebra) to 3e-4 on these coarse grids, which amplify
    the integration tolerance (the reference's own C_l moves by 5e-5 here when tol_perturb_integration is halved)."""
    inp = golden("lcdm_coarse")
    monkeypatch.setenv("CLPP_LANE", "1")
    ctx, pt, tr, sp = run_pipeline(inp)
    cl_lane = sp.cl_[0].copy()
    ctx.close()
    monkeypatch.delenv("CLPP_LANE")
    monkeypatch.setenv("CLPP_WARP_PATH", "1")
    ctx, pt, tr, sp = run_pipeline(inp)
    cl_tail = sp.cl_[0].copy()
    ctx.close()
    monkeypatch.delenv("CLPP_WARP_PATH")
    monkeypatch.setenv("CLPP_GENERIC_ONLY", "1")
    ctx, pt, tr, sp = run_pipeline(inp)
    cl_one = sp.cl_[0].copy()
    ctx.close()
    nz = cl_one != 0
    assert np.max(np.abs(cl_tail[nz] / cl_one[nz] - 1.0)) < 2e-5
    assert np.max(np.abs(cl_lane[nz] / cl_one[nz] - 1.0)) < 3e-4


def test_one_cosmology_over_two_ranks_equals_single_gpu(golden):
    """Multi-GPU path of ONE cosmology (cost-balanced k partition -> exchange of S -> q partition -> sum of the
    partial C_l), with the two ranks emulated as two contexts of this process and the collectives replaced by direct
    device copies: the merged result equals the single-context run (sources bit-identical, C_l to rounding)."""
    import torch
    from classpp_public_b200 import multigpu
    inp = golden("lcdm_coarse")
    a = inp.arrays
    prim = M.TabulatedPrimordial(a["pm.pk_at_transfer_k"])
    ctx, pt_full, tr_full, sp_full = run_pipeline(inp)
    s_full = np.stack(pt_full.sources_[0])
    world = 2
    ranks = []
    for r in range(world):
        c = M.Context(0)
        bg, th, pt, parts = multigpu.stage1_partition(inp, c, world)
        multigpu.solve_modes(pt, parts[r])
        ranks.append((c, bg, th, pt, parts))
    assert sorted(np.concatenate(ranks[0][4]).tolist()) == list(range(pt_full.info.k_size))
    # "all-gather": copy every rank's columns into every other rank's table
    views = [multigpu.device_sources(rk[3]) for rk in ranks]
    for r in range(world):
        idx = torch.as_tensor(ranks[r][4][r], device="cuda:0", dtype=torch.long)
        for o in range(world):
            if o != r:
                views[o].index_copy_(1, idx, views[r].index_select(1, idx))
    torch.cuda.synchronize()

    class LocalSum:  # "all-reduce": collect the partial tables, sum at the end
        parts = []

        def allreduce_sum(self, cl, device):
            self.parts.append(np.array(cl))
            return cl

    ex = LocalSum()
    for r, (c, bg, th, pt, parts) in enumerate(ranks):
        pt._sources = None
        assert np.array_equal(np.stack(pt.sources_[0]), s_full)  # every rank now holds the full, identical table
        multigpu.finish(inp, bg, th, pt, prim, None, r, world, ex)
    cl = np.sum(ex.parts, axis=0)
    full = sp_full.cl_[0]
    nz = full != 0
    assert np.max(np.abs(cl[nz] / full[nz] - 1.0)) < 1e-11
    for rk in ranks:
        rk[0].close()
    ctx.close()


def test_no_device_fails_loudly():
    with pytest.raises(M.CosmoComputationError):
        M.Context(device=9999)


# ---------------------------------------------------------------------------------------------------------------------
# round 2: parity holes named by the round-1 review

# NDF work counters of the unmodified reference (evolver rebuilt with verbose = 1, tools/evolver_ndf15.cpp:112; recorded in
# BASELINE.md section 2): accepted steps, failed steps, RHS evaluations, summed over all k modes
REF_STEPSTAT = {"lcdm": (660583, 29574, 1255883), "planck18": (2666773, 238119, 4874269)}


@pytest.mark.parametrize("name", ["lcdm", "planck18"])
def test_work_counters_vs_reference_stepstat(golden, name):
    """The device integrator follows the reference's step/order control: the number of accepted and failed steps summed
    over the k grid stays within 1 % / 5 % of the reference's stepstat (the roofline of bench.py is scored on the
    reference's counts, so extra device steps would be hidden cost).  RHS evaluations are FEWER than the reference's:
    the Jacobian is probed structurally (4 + block-size probes) instead of by finite differences."""
    inp = golden(name)
    ctx = M.Context(0)
    bg = M.BackgroundModule(inp, ctx)
    th = M.ThermodynamicsModule(inp, bg)
    pt = M.PerturbationsModule(inp, bg, th)
    ks = pt.kstat_
    steps, failed, fevals = int(ks[:, 0].sum()), int(ks[:, 1].sum()), int(ks[:, 2].sum())
    r_steps, r_failed, r_fevals = REF_STEPSTAT[name]
    assert np.all(ks[:, 7] == 0)
    assert abs(steps / r_steps - 1.0) < 0.01, (steps, r_steps)
    assert abs(failed / r_failed - 1.0) < 0.05, (failed, r_failed)
    assert fevals <= 1.02 * r_fevals, (fevals, r_fevals)
    ctx.close()


@pytest.mark.parametrize("name", ["lcdm", "planck18", "ncdm3_deg", "lcdm_dense"])
def test_matter_power_spectra_vs_golden(golden, name):
    """P(k) of the north star on the BASELINE configurations (golden file tests/golden/pk_z0.npz, generated from the
    unmodified reference by make_golden.py pk): linear P_m and P_cb at every node of the k grid within 1e-4, and for
    `non linear = halofit` the non-linear P_m = P_lin R_NL^2 with R_NL from the device halofit (nonlinear_pk_at_z)."""
    from classpp_public_b200.configs import CONFIGS
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "pk_z0.npz"))
    inp = golden(name)
    ctx = M.Context(0)
    bg = M.BackgroundModule(inp, ctx)
    th = M.ThermodynamicsModule(inp, bg)
    pt = M.PerturbationsModule(inp, bg, th)
    pr = z[name + "__pm.pk_at_pt_k"]
    pk = pt.pk_linear(pr)
    assert np.max(np.abs(pk / z[name + "__nl.pk_lin_m_at_pt_k"] - 1.0)) < 1e-4
    if name + "__nl.pk_lin_cb_at_pt_k" in z.files:
        assert np.max(np.abs(pt.pk_linear(pr, cb=True) / z[name + "__nl.pk_lin_cb_at_pt_k"] - 1.0)) < 1e-4
    if name + "__nl.pk_nl_m_at_pt_k" in z.files:
        par = CONFIGS[name]
        nl = M.NonlinearModule(inp, bg, pt, M.AnalyticPrimordial(par["A_s"], par["n_s"]), fetch=True)
        r_nl = nl.nl_corr_density_[0].reshape(pt.info.tau_size, pt.info.k_size)[-1]
        assert np.max(np.abs(pk * r_nl ** 2 / z[name + "__nl.pk_nl_m_at_pt_k"] - 1.0)) < 1e-4
    ctx.close()


def test_three_separate_ncdm_species_on_the_default_grids_vs_golden(golden):
    """BASELINE config 4 as written (N_ncdm = 3, m = 0.02 eV each: 316 equations, 58 hub variables) on the DEFAULT
    grids at the north-star tolerance (round 1 only had the coarse grids at 5e-4)."""
    inp = golden("ncdm3")
    ctx, pt, tr, sp = run_pipeline(inp)
    check_cl(sp, inp.arrays["ref.cl"])
    assert np.all(pt.kstat_[:, 7] == 0)
    ctx.close()


def test_dense_precision_lensed_te_bb_vs_golden(golden):
    """config 3 stand-in: lensed TE and BB as well (round 1 checked TT / EE only)."""
    inp = golden("lcdm_dense")
    ctx, pt, tr, sp = run_pipeline(inp)
    le = M.LensingModule(inp, sp)
    lref = inp.arrays["ref.cl_lensed"].reshape(-1, le.lt_size_)
    mine = np.array([le.lensing_cl_at_l(l) for l in range(2, le.l_lensed_max_ + 1)])
    ref = lref[2:le.l_lensed_max_ + 1]
    tt, ee, te, bb = le.index_lt_tt_, le.index_lt_ee_, le.index_lt_te_, le.index_lt_bb_
    assert np.max(np.abs(mine[:, te] - ref[:, te]) / np.sqrt(ref[:, tt] * ref[:, ee])) < CL_RTOL
    assert np.max(np.abs(mine[:, bb] / ref[:, bb] - 1.0)) < 2 * CL_RTOL
    ctx.close()


def test_lane_kernel_dense_and_structured_hub_solves_agree(golden, monkeypatch):
    """lane.cuh: the structured hub solve (block LU + 4x4 capacitance matrix of the metric coupling) against the dense
    LU of the same Newton matrix."""
    inp = golden("lcdm_coarse")
    monkeypatch.setenv("CLPP_LANE", "1")
    ctx, pt, tr, sp = run_pipeline(inp)
    cl_s, ks_s = sp.cl_[0].copy(), pt.kstat_[:, :2].copy()
    ctx.close()
    monkeypatch.setenv("CLPP_LANE_DENSE", "1")
    ctx, pt, tr, sp = run_pipeline(inp)
    cl_d, ks_d = sp.cl_[0].copy(), pt.kstat_[:, :2].copy()
    ctx.close()
    nz = cl_d != 0
    # not bit-identical (different operation order in the Newton solve flips an occasional step decision); on these coarse
    # grids that shows up at the 1e-4 level, like the other integrator variants (test_integrator_variants_agree)
    assert np.max(np.abs(cl_s[nz] / cl_d[nz] - 1.0)) < 3e-4
    assert abs(ks_s[:, 0].sum() / ks_d[:, 0].sum() - 1.0) < 0.01


@pytest.mark.parametrize("name", ["lcdm", "planck18", "ncdm3_deg"])
def test_lane_kernel_full_pipeline_vs_golden(golden, monkeypatch, name):
    """The thread-per-mode kernel (lane.cuh) alone, forced for a single cosmology (CLPP_LANE=1): C_l of the BASELINE
    configurations within the north-star 1e-4 of the reference."""
    monkeypatch.setenv("CLPP_LANE", "1")
    inp = golden(name)
    ctx, pt, tr, sp = run_pipeline(inp)
    check_cl(sp, inp.arrays["ref.cl"])
    assert np.all(pt.kstat_[:, 7] == 0)
    ctx.close()


def test_hybrid_launch_of_a_batch_vs_single_solves(golden, monkeypatch):
    """Hybrid launch (opt-in, CLPP_LANE=2): modes with k >= kcut in the warp-per-mode kernels, the bulk in the lane kernel
    on a second stream.  Every cosmology of the batch gets the C_l of its single-cosmology solve to the level
    the two integrator families agree at (3e-4 on these coarse grids), every mode is integrated exactly once."""
    inp = golden("lcdm_coarse")
    ctx, pt, tr, sp = run_pipeline(inp)
    cl_single = sp.cl_[0].copy()
    ctx.close()
    monkeypatch.setenv("CLPP_LANE", "2")
    monkeypatch.setenv("CLPP_LANE_KCUT", "0.05")  # coarse grid: 69 modes, both kernels get a share
    ctxs, pts, tabs = [], [], []
    for _ in range(8):
        c = M.Context(0)
        b = M.BackgroundModule(inp, c)
        t = M.ThermodynamicsModule(inp, b)
        ctxs.append(c)
        tabs.append((b, t))
        pts.append(M.PerturbationsModule(inp, b, t, solve=False))
    M.PerturbationsModule.solve_batch(pts)
    a = inp.arrays
    for (bg, th), p in zip(tabs, pts):
        ks = p.kstat_
        assert np.all(ks[:, 7] == 0) and np.all(ks[:, 0] > 0)
        tr = M.TransferModule(inp, bg, th, p, None)
        cl = M.SpectraModule(inp, p, M.TabulatedPrimordial(a["pm.pk_at_transfer_k"]), None, tr).cl_[0]
        nz = cl_single != 0
        assert np.max(np.abs(cl[nz] / cl_single[nz] - 1.0)) < 3e-4
    # the lane share and the warp share of one cosmology are both populated
    assert np.array_equal(pts[0].kstat_[:, 0], pts[5].kstat_[:, 0])
    for c in ctxs:
        c.close()


@pytest.mark.parametrize("kmax", ["1e9", "0.05"])
def test_lane_tails_of_a_batch_vs_warp_tails(golden, monkeypatch, kmax):
    """Batches of >= 8 cosmologies run the radiation-streaming tails of their modes below k = tail_lane_kmax in
    perturb_tail_lane_kernel (one thread per mode) instead of perturb_tail_kernel (one warp per mode).  Same integrator,
    different arithmetic order: every mode is integrated exactly once, the step counts of the two agree within 2 % in
    total, the sources of the cosmologies of a batch are identical to each other, and the C_l agree at the level the two
    integrator families agree at on these coarse grids (3e-4).  kmax = 1e9: every tail in lanes; 0.05: both kernels."""
    inp = golden("lcdm_coarse")
    a = inp.arrays
    out = {}
    for setting in ("0", kmax):
        monkeypatch.setenv("CLPP_TAIL_LANE_KMAX", setting)
        ctxs, pts, tabs = [], [], []
        for _ in range(8):
            c = M.Context(0)
            b = M.BackgroundModule(inp, c)
            t = M.ThermodynamicsModule(inp, b)
            ctxs.append(c)
            tabs.append((b, t))
            pts.append(M.PerturbationsModule(inp, b, t, solve=False))
        M.PerturbationsModule.solve_batch(pts)
        for p in pts:
            assert np.all(p.kstat_[:, 7] == 0) and np.all(p.kstat_[:, 0] > 0)
            assert np.array_equal(np.stack(p.sources_[0]), np.stack(pts[0].sources_[0]))
        bg, th = tabs[2]
        tr = M.TransferModule(inp, bg, th, pts[2], None)
        cl = M.SpectraModule(inp, pts[2], M.TabulatedPrimordial(a["pm.pk_at_transfer_k"]), None, tr).cl_[0].copy()
        out[setting] = (cl, pts[2].kstat_[:, :6].copy(), ctxs[0].launch_count)
        for c in ctxs:
            c.close()
    cl0, ks0, n0 = out["0"]
    cl1, ks1, n1 = out[kmax]
    nz = cl0 != 0
    assert np.max(np.abs(cl1[nz] / cl0[nz] - 1.0)) < 3e-4
    assert abs(ks1[:, 0].sum() / ks0[:, 0].sum() - 1.0) < 0.02
    if kmax == "0.05":
        assert n1 > n0  # the group holding both kinds of tails launched both tail kernels


def test_latin_hypercube_batch_at_full_resolution_vs_golden():
    """BASELINE config 5 in miniature, at FULL resolution and the north-star tolerance: the first 16 points of the seed-0
    Latin hypercube (Planck-18 settings: 1 ncdm species, halofit; omega_b, omega_cdm, h, A_s, n_s, tau_reio varied over the
    config-5 ranges) in ONE batched perturbation launch, upstream tables from the drop-in library.  Against the unmodified
    reference (tests/golden/lhs16.npz, make_golden.py lhs): unlensed C_l^{TT,EE,TE,pp}, lensed TT/EE/TE/BB, linear and
    non-linear P_m(k, z=0) of every cosmology."""
    import json
    from classpp_public_b200 import upstream
    if not upstream.available():
        pytest.skip("shim/_build/libclass_b200.so not built")
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "lhs16.npz"))
    pars = json.loads(str(z["params"]))
    inps = [upstream.inputs_for(p) for p in pars]
    tabs, pts = [], []
    for inp in inps:
        c = M.Context(0)
        c.set_option("lean_scratch", 1)
        b = M.BackgroundModule(inp, c)
        t = M.ThermodynamicsModule(inp, b)
        tabs.append((c, b, t))
        pts.append(M.PerturbationsModule(inp, b, t, solve=False))
    M.PerturbationsModule.solve_batch(pts)
    worst = {}
    for i, (inp, (c, bg, th), pt, par) in enumerate(zip(inps, tabs, pts, pars)):
        assert np.array_equal(pt.k_[0], z["%d__k" % i])  # grids bit-exact for every point of the sweep
        prim = M.AnalyticPrimordial(par["A_s"], par["n_s"])
        nl = M.NonlinearModule(inp, bg, pt, prim, fetch=True)
        tr = M.TransferModule(inp, bg, th, pt, nl)
        sp = M.SpectraModule(inp, pt, prim, nl, tr)
        check_cl(sp, z["%d__cl" % i])
        le = M.LensingModule(inp, sp)
        lref = z["%d__cl_lensed" % i].reshape(-1, le.lt_size_)
        ll = np.arange(2, le.l_lensed_max_ + 1)
        mine = np.array([le.lensing_cl_at_l(int(l)) for l in ll[::7]])
        ref = lref[ll[::7]]
        tt, ee, te, bb = le.index_lt_tt_, le.index_lt_ee_, le.index_lt_te_, le.index_lt_bb_
        e = {"lensed_tt": np.max(np.abs(mine[:, tt] / ref[:, tt] - 1)), "lensed_ee": np.max(np.abs(mine[:, ee] / ref[:, ee] - 1)),
             "lensed_te": np.max(np.abs(mine[:, te] - ref[:, te]) / np.sqrt(ref[:, tt] * ref[:, ee])),
             "lensed_bb": np.max(np.abs(mine[:, bb] / ref[:, bb] - 1))}
        pk = pt.pk_linear(prim.pk_at_k(pt.k_[0]))
        e["pk_lin"] = np.max(np.abs(pk / z["%d__pk_lin_m" % i] - 1))
        r_nl = nl.nl_corr_density_[0].reshape(pt.info.tau_size, pt.info.k_size)[-1]
        e["pk_nl"] = np.max(np.abs(pk * r_nl ** 2 / z["%d__pk_nl_m" % i] - 1))
        for k_, v in e.items():
            worst[k_] = max(worst.get(k_, 0.0), float(v))
        assert e["lensed_tt"] < CL_RTOL and e["lensed_ee"] < CL_RTOL and e["lensed_te"] < CL_RTOL and e["lensed_bb"] < 2 * CL_RTOL, (i, e)
        assert e["pk_lin"] < 1e-4 and e["pk_nl"] < 1e-4, (i, e)
    print("worst relative errors over the 16 cosmologies:", worst)
    for c, _, _ in tabs:
        c.close()
