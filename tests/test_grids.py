"""Host-side sampling grids must be BIT-EXACT with the reference (north_star): k, tau, q, k(q), l.
Checked against the committed golden fixtures (always) and the live oracle (when oracle/_ref is built)."""
import numpy as np
import pytest

from classpp_public_b200 import modules as M


def build_grids(inp):
    ctx = M.Context(device=-1)  # host-only context: grids need no GPU
    bg = M.BackgroundModule(inp, ctx)
    th = M.ThermodynamicsModule(inp, bg)
    pt = M.PerturbationsModule(inp, bg, th, solve=False)
    tr = M.TransferModule(inp, bg, th, pt, compute=False)
    return pt, tr


@pytest.mark.parametrize("name", ["lcdm_coarse", "lcdm", "planck18", "ncdm3_deg", "ncdm3_coarse", "lcdm_dense"])
def test_grids_bit_exact_vs_golden(golden, name):
    inp = golden(name)
    a = inp.arrays
    pt, tr = build_grids(inp)
    sizes = a["ref.sizes"].astype(int)
    assert [pt.info.k_size, pt.info.k_size_cl, pt.info.k_size_cmb, pt.info.tau_size, pt.info.tp_size] == list(sizes[:5])
    assert [tr.info.tt_size, tr.info.l_size, tr.info.q_size] == list(sizes[5:8])
    assert np.array_equal(pt.k_[0], a["ref.k"])            # bit-exact
    assert np.array_equal(pt.tau_sampling_, a["ref.tau"])  # bit-exact
    assert np.array_equal(tr.q_, a["ref.q"])
    assert np.array_equal(tr.k_[0], a["ref.kq"])
    assert np.array_equal(tr.l_, a["ref.l"].astype(np.int32))
    assert np.array_equal(tr.l_size_tt_[0], a["ref.l_size_tt"].astype(np.int32))
    tp = a["ref.tp_index"].astype(int)
    mine = [pt.info.index_tp_t0, pt.info.index_tp_t1, pt.info.index_tp_t2, pt.info.index_tp_p,
            pt.info.index_tp_delta_m, pt.info.index_tp_delta_cb, pt.info.index_tp_phi_plus_psi]
    assert mine == list(tp)


@pytest.mark.parametrize("name", ["lcdm_coarse", "ncdm3_deg"])
def test_grids_bit_exact_vs_live_reference(reference, name):
    if reference is None:
        pytest.skip("oracle/_ref not built")
    from refutil import inputs_from_reference
    ref = reference(name, "transfer")
    pt, tr = build_grids(inputs_from_reference(ref))
    assert np.array_equal(pt.k_[0], ref.get("pt.k"))
    assert np.array_equal(pt.tau_sampling_, ref.get("pt.tau_sampling"))
    assert np.array_equal(tr.q_, ref.get("tr.q"))
    assert np.array_equal(tr.l_, ref.get("tr.l").astype(np.int32))


def test_grids_bit_exact_for_a_latin_hypercube_sweep(reference):
    """BASELINE config 5: the grids stay bit-exact when the cosmology moves (k_min, k_rec, tau_rec, angular rescaling,
    q period ... all change): 5 points of the sweep's Latin hypercube on the coarse settings, against the live reference."""
    if reference is None:
        pytest.skip("oracle/_ref not built")
    import os
    from scipy.stats import qmc
    from oracle import refprobe
    from classpp_public_b200.configs import CONFIGS
    from refutil import inputs_from_reference
    lo = np.array([0.020, 0.10, 0.60, 2.9, 0.92, 0.03])   # omega_b, omega_cdm, h, ln(1e10 A_s), n_s, tau_reio
    hi = np.array([0.024, 0.14, 0.75, 3.2, 1.00, 0.09])
    sizes = set()
    for x in qmc.scale(qmc.LatinHypercube(d=6, seed=0).random(5), lo, hi):
        par = dict(CONFIGS["lcdm_coarse"], omega_b=x[0], omega_cdm=x[1], h=x[2], A_s=1e-10 * np.exp(x[3]), n_s=x[4],
                   tau_reio=x[5])
        ref = refprobe.RefCosmology(par, threads=os.cpu_count()).compute("transfer")
        pt, tr = build_grids(inputs_from_reference(ref))
        assert np.array_equal(pt.k_[0], ref.get("pt.k"))
        assert np.array_equal(pt.tau_sampling_, ref.get("pt.tau_sampling"))
        assert np.array_equal(tr.q_, ref.get("tr.q"))
        assert np.array_equal(tr.l_, ref.get("tr.l").astype(np.int32))
        sizes.add((pt.info.k_size, pt.info.tau_size, tr.info.q_size))
        ref.close()
    assert len(sizes) > 1  # the grids really differ from point to point


def test_spline_tables_match_private_reference_tables(reference, tmp_path):
    """csrc/host_tables.cpp rebuilds the second-derivative tables that the reference keeps private
    (background_module.h:177, thermodynamics_module.h:122-124): clpp_spline_table_lines on the reference's public tables must
    reproduce `d2background_dtau2_table_` and `d2thermodynamics_dz2_table_` BIT FOR BIT (same recurrences in the same order as
    array_spline_table_lines(..., _SPLINE_EST_DERIV_), tools/arrays.c:514-660) -- the interpolated background / thermodynamics
    of the device path, and through them the tau grid, hang on it."""
    if reference is None:
        pytest.skip("oracle/_ref not built")
    import os
    import subprocess
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    csrc = os.path.join(ROOT, "classpp_public_b200", "csrc")
    src = tmp_path / "spl.cpp"
    src.write_text("""
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "clpp_internal.h"
int main(int argc, char** argv) {
  const int n = atoi(argv[1]), m = atoi(argv[2]);
  std::vector<double> x(n), y((size_t)n * m), dd((size_t)n * m);
  FILE* f = fopen(argv[3], "rb");
  if (fread(x.data(), 8, n, f) != (size_t)n || fread(y.data(), 8, (size_t)n * m, f) != (size_t)n * m) return 2;
  fclose(f);
  clpp_spline_table_lines(x.data(), n, y.data(), m, dd.data());
  f = fopen(argv[4], "wb");
  fwrite(dd.data(), 8, (size_t)n * m, f);
  fclose(f);
  return 0;
}
""")
    exe = str(tmp_path / "spl")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-I", os.path.join(ROOT, "include"), "-I", csrc,
                           str(src), os.path.join(csrc, "host_tables.cpp"), "-o", exe])
    ref = reference("planck18", "thermodynamics")
    for x_key, y_key, d_key in (("bg.tau_table", "bg.background_table", "bg.d2background_dtau2_table"),
                                ("th.z_table", "th.thermodynamics_table", "th.d2thermodynamics_dz2_table")):
        x, y, d2 = ref.get(x_key), ref.get(y_key), ref.get(d_key)
        n, m = len(x), len(y) // len(x)
        fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
        with open(fin, "wb") as f:
            f.write(x.tobytes())
            f.write(y.tobytes())
        subprocess.check_call([exe, str(n), str(m), fin, fout])
        mine = np.fromfile(fout, dtype=np.float64)
        assert mine.shape == d2.shape
        assert np.array_equal(mine, d2), (d_key, float(np.max(np.abs(mine - d2))))


def test_unsupported_inputs_fail_loudly(golden):
    inp = golden("lcdm_coarse")
    bad = M.Inputs(dict(inp.meta), inp.arrays)
    bad.meta["pt.gauge"] = 0  # newtonian
    ctx = M.Context(device=-1)
    bg = M.BackgroundModule(bad, ctx)
    th = M.ThermodynamicsModule(bad, bg)
    with pytest.raises(M.CosmoComputationError, match="synchronous"):
        M.PerturbationsModule(bad, bg, th, solve=False)
    bad.meta["pt.gauge"] = 1
    bad.meta["ba.sgnK"] = 1
    bg = M.BackgroundModule(bad, M.Context(device=-1))
    th = M.ThermodynamicsModule(bad, bg)
    with pytest.raises(M.CosmoComputationError, match="flat"):
        M.PerturbationsModule(bad, bg, th, solve=False)


def test_gauss_legendre_host_routine_vs_numpy(tmp_path):
    """The host routine behind accurate_lensing = 1 (csrc/host_tables.cpp: clpp_gauss_legendre, restating
    tools/quadrature.c:752-788): nodes against numpy's leggauss, weights positive, symmetric and summing to 2."""
    import os
    import subprocess
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    csrc = os.path.join(ROOT, "classpp_public_b200", "csrc")
    src = tmp_path / "gl.cpp"
    src.write_text("""
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "clpp_internal.h"
int main(int argc, char** argv) {
  const int n = atoi(argv[1]);
  std::vector<double> mu(n), w(n);
  char err[2048];
  if (clpp_gauss_legendre(mu.data(), w.data(), n, 2.220446049250313e-16, err)) return 1;
  for (int i = 0; i < n; i++) printf("%.17g %.17g\\n", mu[i], w[i]);
  return 0;
}
""")
    exe = str(tmp_path / "gl")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-I", os.path.join(ROOT, "include"), "-I", csrc,
                           str(src), os.path.join(csrc, "host_tables.cpp"), "-o", exe])
    for n in (7, 64, 1169):
        out = np.array(subprocess.check_output([exe, str(n)]).split(), dtype=float).reshape(-1, 2)
        x, w = np.polynomial.legendre.leggauss(n)
        assert np.max(np.abs(out[:, 0] - x)) < 5e-16
        assert np.all(out[:, 1] > 0) and np.array_equal(out[:, 1], out[::-1, 1])
        assert abs(out[:, 1].sum() - 2.0) < 1e-14 and np.max(np.abs(out[:, 1] / w - 1.0)) < 1e-7
