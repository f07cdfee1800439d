"""Host-side mirror of the reference's module interface for the hot path.

The reference's "operator API" for this path is three C++ classes whose constructors do all
the work and which expose their results as public data members
(source/perturbations_module.h:7-178, source/transfer_module.h:7-54,
source/spectra_module.h:11-81).  The classes below keep the same names, the same
construction order and the same member names (trailing underscore included), and throw
`CosmoComputationError` where the reference constructors throw std::runtime_error
(classy maps that to CosmoComputationError, classy.pyx:90-101).

All numerics run in libclpp.so (CUDA, sm_100a) through the C ABI of include/clpp.h; this file
only moves descriptors and arrays.  Upstream stages that are out of scope (input parsing,
background, thermodynamics, primordial spectrum, halofit) are represented by `Inputs`: plain
tables + scalars, exactly what the C++ drop-in reads from the reference's upstream modules.
"""
import ctypes as C
import json

import numpy as np

from . import _capi as capi


class CosmoComputationError(RuntimeError):
    pass


class CosmoSevereError(ValueError):
    pass


def _fill(struct, meta, prefix_map):
    """Fill a ctypes struct from the flat `meta` dict (keys like 'pr.k_step_sub')."""
    for name, ctype in struct._fields_:
        key = prefix_map(name)
        if key is None:
            continue
        if key not in meta:
            raise CosmoSevereError("missing input '%s' for field %s" % (key, name))
        v = meta[key]
        setattr(struct, name, int(round(v)) if ctype in (C.c_int, C.c_long) else float(v))
    return struct


class Inputs:
    """Everything the three stages read from upstream: scalars (`meta`) and tables (`arrays`)."""

    def __init__(self, meta, arrays):
        self.meta = dict(meta)
        self.arrays = dict(arrays)

    # ---- persistence (tests/golden fixtures, bench inputs) ----
    def save(self, path, extra=None):
        d = {k.replace(".", "__"): v for k, v in self.arrays.items()}
        if extra:
            d.update({k.replace(".", "__"): v for k, v in extra.items()})
        np.savez_compressed(path, meta=np.array(json.dumps(self.meta)), **d)

    @classmethod
    def load(cls, path):
        z = np.load(path, allow_pickle=False)
        meta = json.loads(str(z["meta"]))
        arrays = {k.replace("__", "."): z[k] for k in z.files if k != "meta"}
        return cls(meta, arrays)

    # ---- descriptors ----
    def background_desc(self):
        m = self.meta
        d = capi.BackgroundDesc()

        def key(n):
            if n.startswith("index_bg_"):
                return "bg.index_" + n[len("index_bg_"):]
            if n in ("bt_size", "bg_size", "bg_size_short", "bg_size_normal", "conformal_age"):
                return "bg." + n
            return "ba." + n
        return _fill(d, m, key)

    def thermo_desc(self):
        d = capi.ThermoDesc()

        def key(n):
            if n.startswith("index_th_"):
                return "th.index_" + n[len("index_th_"):]
            return "th." + n
        return _fill(d, self.meta, key)

    def perturb_desc(self):
        d = capi.PerturbDesc()
        pt_names = {"has_cl_cmb_temperature", "has_cl_cmb_polarization", "has_cl_cmb_lensing_potential",
                    "has_pk_matter", "has_nl_corrections_based_on_delta_m", "gauge", "l_scalar_max",
                    "k_max_for_pk", "z_max_pk", "switch_sw", "switch_eisw", "switch_lisw", "switch_dop",
                    "switch_pol", "eisw_lisw_split_z", "three_ceff2_ur", "three_cvis2_ur"}
        self.meta.setdefault("pr.perturb_integration_stepsize", 0.5)  # precisions.h:226 (fixtures older than the rk evolver)
        return _fill(d, self.meta, lambda n: ("pt." if n in pt_names else "pr.") + n)

    def transfer_desc(self):
        d = capi.TransferDesc()
        pt_names = {"has_cl_cmb_temperature", "has_cl_cmb_polarization", "has_cl_cmb_lensing_potential",
                    "l_scalar_max"}
        tr_names = {"lcmb_rescale", "lcmb_tilt", "lcmb_pivot"}
        return _fill(d, self.meta,
                     lambda n: ("pt." if n in pt_names else "tr." if n in tr_names else "pr.") + n)


class Context:
    """One clpp_ctx: one cosmology on one CUDA device (device=-1: host-only, grids only)."""

    def __init__(self, device=0):
        self._lib = capi.lib()
        self._h = C.c_void_p()
        self._err = C.create_string_buffer(capi.ERRLEN)
        if self._lib.clpp_ctx_create(int(device), C.byref(self._h), self._err) != 0:
            raise CosmoComputationError(self._err.value.decode(errors="replace"))
        self.device = device

    def set_option(self, name, value):
        """clpp_ctx_set_option: 'lean_scratch' (sweeps: transfer work buffers from the device memory pool), 'lane_path'."""
        self.check(self._lib.clpp_ctx_set_option(self._h, name.encode(), float(value), self._err))

    def check(self, rc):
        if rc != 0:
            raise CosmoComputationError(self._err.value.decode(errors="replace"))

    @property
    def handle(self):
        return self._h

    @property
    def err(self):
        return self._err

    @property
    def launch_count(self):
        return int(self._lib.clpp_ctx_launch_count(self._h))

    @property
    def stream_ptr(self):
        p = C.c_void_p()
        self._lib.clpp_ctx_get_stream(self._h, C.byref(p))
        return p.value or 0

    def kernel_ms(self):
        out = np.zeros(8)
        self._lib.clpp_ctx_get_kernel_ms(self._h, capi.dptr(out))
        return dict(zip(("perturb", "k_spline", "bessel", "los", "spectra", "perturb_tail", "halofit", "lensing"),
                        out.tolist()))

    def fp64_peak_tflops(self):
        v = C.c_double()
        self.check(self._lib.clpp_measure_fp64_peak(self._h, C.byref(v), self._err))
        return v.value

    def close(self):
        if self._h:
            self._lib.clpp_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class BackgroundModule:
    """Stand-in for the reference BackgroundModule: owns tau_table_/background_table_ (upstream, out of scope)."""

    def __init__(self, inputs, ctx):
        self.inputs, self.ctx = inputs, ctx
        a = inputs.arrays
        self.desc = inputs.background_desc()
        self.tau_table_ = np.ascontiguousarray(a["bg.tau_table"], dtype=np.float64)
        self.background_table_ = np.ascontiguousarray(a["bg.background_table"], dtype=np.float64)
        self.bt_size_, self.bg_size_ = self.desc.bt_size, self.desc.bg_size
        self.conformal_age_ = self.desc.conformal_age
        L = ctx._lib
        ctx.check(L.clpp_set_background(ctx.handle, C.byref(self.desc), capi.dptr(self.tau_table_),
                                        capi.dptr(self.background_table_), ctx.err))
        if self.desc.has_ncdm:
            qs = np.ascontiguousarray(a["ncdm.q_size"], dtype=np.float64).astype(np.int32)
            arrs = [np.ascontiguousarray(a[k], dtype=np.float64)
                    for k in ("ncdm.q", "ncdm.w", "ncdm.dlnf0_dlnq", "ncdm.M", "ncdm.factor")]
            ctx.check(L.clpp_set_ncdm(ctx.handle, int(self.desc.N_ncdm), capi.iptr(qs),
                                      *[capi.dptr(x) for x in arrs], ctx.err))


class ThermodynamicsModule:
    """Stand-in for the reference ThermodynamicsModule (upstream, out of scope)."""

    def __init__(self, inputs, background_module):
        self.inputs, self.ctx = inputs, background_module.ctx
        a = inputs.arrays
        self.desc = inputs.thermo_desc()
        self.z_table_ = np.ascontiguousarray(a["th.z_table"], dtype=np.float64)
        self.thermodynamics_table_ = np.ascontiguousarray(a["th.thermodynamics_table"], dtype=np.float64)
        self.tau_rec_, self.rs_rec_ = self.desc.tau_rec, self.desc.rs_rec
        self.angular_rescaling_ = self.desc.angular_rescaling
        ctx = self.ctx
        ctx.check(ctx._lib.clpp_set_thermo(ctx.handle, C.byref(self.desc), capi.dptr(self.z_table_),
                                           capi.dptr(self.thermodynamics_table_), ctx.err))


class PerturbationsModule:
    """PerturbationsModule(input_module, background_module, thermodynamics_module)
    (reference: source/perturbations_module.h:9).  The constructor computes the k/tau grids on
    the host (bit-exact) and integrates every k mode on the GPU."""

    def __init__(self, inputs, background_module, thermodynamics_module, solve=True, k_range=None):
        ctx = background_module.ctx
        self.ctx, self.inputs = ctx, inputs
        L = ctx._lib
        self.desc = inputs.perturb_desc()
        self.info = capi.PerturbInfo()
        ctx.check(L.clpp_perturb_grids(ctx.handle, C.byref(self.desc), C.byref(self.info), ctx.err))
        i = self.info
        self.md_size_ = 1
        self.index_md_scalars_ = 0
        self.ic_size_ = [1]
        self.index_ic_ad_ = 0
        self.tp_size_ = [i.tp_size]
        self.k_size_, self.k_size_cl_, self.k_size_cmb_ = [i.k_size], [i.k_size_cl], [i.k_size_cmb]
        self.tau_size_, self.ln_tau_size_ = i.tau_size, i.ln_tau_size
        self.k_min_, self.k_max_ = i.k_min, i.k_max
        for n in ("t0", "t1", "t2", "p", "delta_m", "delta_cb", "phi_plus_psi"):
            setattr(self, "index_tp_%s_" % n, getattr(i, "index_tp_" + n))
        self.has_source_t_ = i.index_tp_t0 >= 0
        self.has_source_p_ = i.index_tp_p >= 0
        self.has_source_delta_m_ = i.index_tp_delta_m >= 0
        self.has_source_delta_cb_ = i.index_tp_delta_cb >= 0
        self.has_source_phi_plus_psi_ = i.index_tp_phi_plus_psi >= 0
        k = np.empty(i.k_size)
        tau = np.empty(i.tau_size)
        L.clpp_perturb_get_k(ctx.handle, capi.dptr(k))
        L.clpp_perturb_get_tau(ctx.handle, capi.dptr(tau))
        self.k_ = [k]
        self.tau_sampling_ = tau
        self._sources = None
        self._kstat = None
        if solve:
            lo, hi = k_range if k_range is not None else (0, i.k_size)
            ctx.check(L.clpp_perturb_solve(ctx.handle, int(lo), int(hi), ctx.err))
            self._fetch_kstat()

    def _fetch_kstat(self):
        """Work counters of the last solve are fetched lazily (first access of kstat_ / kprofile_ / ksections_)."""
        self._kstat = None
        self._sources = None

    def _kstat_table(self):
        if self._kstat is None:
            ks = (capi.KStat * self.info.k_size)()
            self.ctx._lib.clpp_perturb_get_kstat(self.ctx.handle, ks)
            self._kstat = np.frombuffer(ks, dtype=np.dtype(capi.KStat)).copy()
        return self._kstat

    @property
    def kstat_(self):
        """[k][steps, failed, fevals, jacobians, factorizations, solves, intervals, status] (evolver statistics)"""
        t = self._kstat_table()
        return np.stack([t[n] for n in ("steps", "failed", "fevals", "jacobians", "factorizations", "solves",
                                        "intervals", "status")], axis=1).astype(np.int64)

    @property
    def kprofile_(self):
        """[k][neq | steps | cycles][interval]: per approximation interval of every mode"""
        t = self._kstat_table()
        return np.stack([t["iv_neq"].astype(np.int64), t["iv_steps"].astype(np.int64), t["iv_cycles"].astype(np.int64)],
                        axis=1)

    @property
    def ksections_(self):
        """[k][72]: cycles per code section (only filled by a -DPT_PROF build)"""
        return self._kstat_table()["prof"].astype(np.int64)

    @staticmethod
    def solve_batch(modules):
        """Integrate every k mode of several PerturbationsModule objects (constructed with solve=False, one
        Context each, same device and precision settings) in ONE kernel launch (clpp_perturb_solve_batch)."""
        ctx0 = modules[0].ctx
        arr = (C.c_void_p * len(modules))(*[m.ctx.handle for m in modules])
        ctx0.check(ctx0._lib.clpp_perturb_solve_batch(arr, len(modules), ctx0.err))
        for m in modules:
            m._fetch_kstat()

    @classmethod
    def from_sources(cls, inputs, background_module, k, tau, sources, info):
        """Test/pipeline hook: a PerturbationsModule whose S(k,tau) is injected (e.g. the oracle's)."""
        self = cls.__new__(cls)
        ctx = background_module.ctx
        self.ctx, self.inputs, self.info = ctx, inputs, info
        k = np.ascontiguousarray(k, dtype=np.float64)
        tau = np.ascontiguousarray(tau, dtype=np.float64)
        src = np.ascontiguousarray(sources, dtype=np.float64)
        ctx.check(ctx._lib.clpp_perturb_set_sources(ctx.handle, C.byref(info), capi.dptr(k), capi.dptr(tau),
                                                    capi.dptr(src), ctx.err))
        self.k_, self.tau_sampling_ = [k], tau
        self.k_size_, self.k_size_cl_ = [info.k_size], [info.k_size_cl]
        self.tau_size_, self.tp_size_ = info.tau_size, [info.tp_size]
        self._sources = None
        return self

    def pk_linear(self, primordial_pk, index_tau=-1, cb=False):
        """Linear P(k) [Mpc^3] on the k grid of this module at sample index_tau (default: today), from delta_m
        (cb: delta_cb) and the primordial spectrum P_R(k_i) (reference: nonlinear_pk_linear, nonlinear_module.cpp:1886)."""
        pr = np.ascontiguousarray(primordial_pk, dtype=np.float64)
        assert len(pr) == self.info.k_size
        out = np.empty(self.info.k_size)
        self.ctx.check(self.ctx._lib.clpp_pk_linear(self.ctx.handle, capi.dptr(pr), int(index_tau), int(bool(cb)),
                                                    capi.dptr(out), self.ctx.err))
        return out

    def perturb_sources_at_tau(self, index_md, index_ic, index_tp, tau):
        """S^{tp}(k, tau) at all k, linear in tau between the sampling times (reference: perturbations_module.cpp:79;
        z_max_pk = 0 branch). index_md / index_ic must be 0 (scalars, adiabatic)."""
        if index_md != 0 or index_ic != 0:
            raise CosmoSevereError("only the scalar adiabatic mode is computed (index_md = index_ic = 0)")
        out = np.empty(self.info.k_size)
        self.ctx.check(self.ctx._lib.clpp_perturb_sources_at_tau(self.ctx.handle, int(index_tp), float(tau), capi.dptr(out),
                                                                 self.ctx.err))
        return out

    @property
    def sources_(self):
        """sources_[index_md][index_ic*tp_size+index_tp][index_tau*k_size+index_k] (perturbations.h:20)."""
        if self._sources is None:
            i = self.info
            out = np.empty((i.tp_size, i.tau_size * i.k_size))
            self.ctx.check(self.ctx._lib.clpp_perturb_get_sources(self.ctx.handle, capi.dptr(out), self.ctx.err))
            self._sources = [list(out)]
        return self._sources


class TransferModule:
    """TransferModule(input, background, thermodynamics, perturbations, nonlinear)
    (reference: source/transfer_module.h:9). `nonlinear_module` is None or an object with
    `nl_corr_density_m` = NonlinearModule::nl_corr_density_[index_pk_m] ([tau*k_size+k])."""

    def __init__(self, inputs, background_module, thermodynamics_module, perturbations_module,
                 nonlinear_module=None, compute=True, q_range=None):
        ctx = background_module.ctx
        self.ctx, self.inputs = ctx, inputs
        L = ctx._lib
        self.desc = inputs.transfer_desc()
        self.info = capi.TransferInfo()
        ctx.check(L.clpp_transfer_grids(ctx.handle, C.byref(self.desc), C.byref(self.info), ctx.err))
        i = self.info
        self.md_size_ = 1
        self.tt_size_ = [i.tt_size]
        self.l_size_max_, self.l_size_ = i.l_size_max, [i.l_size]
        self.q_size_ = i.q_size
        for n in ("t0", "t1", "t2", "e", "lcmb"):
            setattr(self, "index_tt_%s_" % n, getattr(i, "index_tt_" + n))
        l = np.empty(i.l_size_max, dtype=np.int32)
        lt = np.empty(i.tt_size, dtype=np.int32)
        L.clpp_transfer_get_l(ctx.handle, capi.iptr(l), capi.iptr(lt))
        q = np.empty(i.q_size)
        k = np.empty(i.q_size)
        L.clpp_transfer_get_q(ctx.handle, capi.dptr(q), capi.dptr(k))
        self.l_, self.l_size_tt_ = l, [lt]
        self.q_, self.k_ = q, [k]
        self.index_q_flat_approximation_ = 0
        self._transfer = None
        if compute:
            nl = None
            if nonlinear_module is not None and getattr(nonlinear_module, "nl_corr_density_m", None) is not None:
                nl = np.ascontiguousarray(nonlinear_module.nl_corr_density_m, dtype=np.float64)
            lo, hi = q_range if q_range is not None else (0, i.q_size)
            # the halofit table a device NonlinearModule left on this context is applied only when THAT module is passed
            # (the reference: no nonlinear module / method none means linear transfers)
            ctx.set_option("use_device_nl", 1 if (nonlinear_module is not None and getattr(nonlinear_module, "on_device", False)) else 0)
            ctx.check(L.clpp_transfer_compute(ctx.handle, capi.dptr(nl), int(lo), int(hi), ctx.err))
            # refresh counters
            self.n_integrals_, self.n_points_ = None, None

    def set_transfer(self, transfer):
        t = np.ascontiguousarray(transfer, dtype=np.float64)
        self.ctx.check(self.ctx._lib.clpp_transfer_set_transfer(self.ctx.handle, capi.dptr(t), self.ctx.err))
        self._transfer = None

    @property
    def transfer_(self):
        """transfer_[index_md][((index_ic*tt_size+index_tt)*l_size+index_l)*q_size+index_q]."""
        if self._transfer is None:
            i = self.info
            out = np.empty(i.tt_size * i.l_size * i.q_size)
            self.ctx.check(self.ctx._lib.clpp_transfer_get_transfer(self.ctx.handle, capi.dptr(out), self.ctx.err))
            self._transfer = [out]
        return self._transfer

    def bessel_table(self):
        i = self.info
        # x_size is known after compute
        L = self.ctx._lib
        nx = self._x_size()
        x = np.empty(nx)
        phi = np.empty((i.l_size_max, nx))
        dphi = np.empty((i.l_size_max, nx))
        chi = np.empty(i.l_size_max)
        self.ctx.check(L.clpp_transfer_get_bessel(self.ctx.handle, capi.dptr(x), capi.dptr(phi), capi.dptr(dphi),
                                                  capi.dptr(chi), self.ctx.err))
        return x, phi, dphi, chi

    def _x_size(self):
        m = self.inputs.meta
        xmax = self.q_[-1] * m["bg.conformal_age"]
        nx = int((xmax - m["pr.hyper_x_min"]) * m["pr.hyper_sampling_flat"] / (2 * np.pi))
        return max(nx, 2)


class SpectraModule:
    """SpectraModule(input, perturbations, primordial, nonlinear, transfer)
    (reference: source/spectra_module.h:13).  `primordial_module` must provide
    `pk_at_k(k) -> P_R(k)` (adiabatic scalar mode), the reference's primordial_spectrum_at_k."""

    def __init__(self, inputs, perturbations_module, primordial_module, nonlinear_module, transfer_module,
                 q_range=None):
        ctx = transfer_module.ctx
        self.ctx = ctx
        L = ctx._lib
        k = transfer_module.k_[0]
        pk = np.ascontiguousarray(primordial_module.pk_at_k(k), dtype=np.float64)
        self.info = capi.SpectraInfo()
        ti = transfer_module.info
        cl = np.zeros(ti.l_size * 7)
        if q_range is None:
            ctx.check(L.clpp_spectra_compute(ctx.handle, capi.dptr(pk), C.byref(self.info), capi.dptr(cl), ctx.err))
        else:
            ctx.check(L.clpp_spectra_compute_range(ctx.handle, capi.dptr(pk), int(q_range[0]), int(q_range[1]),
                                                   C.byref(self.info), capi.dptr(cl), ctx.err))
        i = self.info
        self.md_size_, self.ic_size_, self.ic_ic_size_ = 1, [1], [1]
        self.ct_size_ = i.ct_size
        for n in ("tt", "ee", "te", "bb", "pp", "tp", "ep"):
            idx = getattr(i, "index_ct_" + n)
            setattr(self, "index_ct_%s_" % n, idx)
            setattr(self, "has_%s_" % n, int(idx >= 0))
        self.l_size_ = [i.l_size]
        self.l_size_max_ = i.l_size
        self.l_ = transfer_module.l_[: i.l_size].astype(np.float64)
        self.cl_ = [cl[: i.l_size * i.ct_size].copy()]
        self.l_max_tot_ = int(self.l_[-1])

    def spectra_cl_at_l(self, l):
        """C_l^{ct} at (real) multipole l: spline in l, zero above l_max (reference: spectra_module.cpp:220)."""
        out = np.zeros(self.ct_size_)
        self.ctx.check(self.ctx._lib.clpp_spectra_cl_at_l(self.ctx.handle, float(l), capi.dptr(out), self.ctx.err))
        return out

    def cl_output(self, lmax):
        """dict name -> C_l[0..lmax] (dimensionless), like SpectraModule::cl_output (spectra_module.cpp:146)."""
        out = np.zeros((int(lmax) + 1) * self.ct_size_)
        self.ctx.check(self.ctx._lib.clpp_spectra_cl_output(self.ctx.handle, int(lmax), capi.dptr(out), self.ctx.err))
        tab = out.reshape(int(lmax) + 1, self.ct_size_)
        return {n: tab[:, getattr(self, "index_ct_%s_" % n)].copy() for n in ("tt", "ee", "te", "bb", "pp", "tp", "ep")
                if getattr(self, "has_%s_" % n)}


class LensingModule:
    """LensingModule(input, spectra) (reference: source/lensing_module.h): lensed TT, TE, EE, BB on the device
    (clpp_lensing_compute) from the C_l table of the SpectraModule. Members as in the reference: `l_`, `l_size_`,
    `lt_size_`, `cl_lens_` ([index_l*lt_size_+index_lt]), `l_unlensed_max_`, `l_lensed_max_`, `index_lt_*_`, `has_*_`."""

    def __init__(self, inputs, spectra_module):
        ctx = spectra_module.ctx
        self.ctx = ctx
        m = inputs.meta
        d = capi.LensingDesc()
        d.accurate_lensing = int(m.get("pr.accurate_lensing", 0))              # precisions.h:492
        d.delta_l_max = int(m.get("pr.delta_l_max", 500))                      # :494
        d.num_mu_minus_lmax = int(m.get("pr.num_mu_minus_lmax", 70))           # :493
        d.tol_gauss_legendre = float(m.get("pr.tol_gauss_legendre", np.finfo(np.float64).eps))  # :495
        self.info = capi.LensingInfo()
        l = np.zeros(spectra_module.l_size_max_)
        cl = np.zeros(spectra_module.l_size_max_ * spectra_module.ct_size_)
        ctx.check(ctx._lib.clpp_lensing_compute(ctx.handle, C.byref(d), C.byref(self.info), capi.dptr(l), capi.dptr(cl),
                                                ctx.err))
        i = self.info
        self.lt_size_, self.l_size_ = i.lt_size, i.l_size
        self.l_unlensed_max_, self.l_lensed_max_ = i.l_unlensed_max, i.l_lensed_max
        self.l_ = l[: i.l_size].copy()
        self.cl_lens_ = cl[: i.l_size * i.lt_size].copy()
        for n in ("tt", "ee", "te", "bb", "pp", "tp", "ep"):
            idx = getattr(i, "index_lt_" + n)
            setattr(self, "index_lt_%s_" % n, idx)
            setattr(self, "has_%s_" % n, int(idx >= 0))

    def lensing_cl_at_l(self, l):
        """lensed C_l^{lt} at integer l <= l_lensed_max_ (reference: lensing_module.cpp:111)."""
        out = np.zeros(self.lt_size_)
        self.ctx.check(self.ctx._lib.clpp_lensing_cl_at_l(self.ctx.handle, int(l), capi.dptr(out), self.ctx.err))
        return out

    def cl_output(self, lmax):
        """dict name -> lensed C_l[0..lmax] like LensingModule::cl_output (lensing_module.cpp:62-108)."""
        tab = np.zeros((int(lmax) + 1, self.lt_size_))
        for l in range(2, int(lmax) + 1):
            tab[l] = self.lensing_cl_at_l(l)
        return {n: tab[:, getattr(self, "index_lt_%s_" % n)].copy() for n in ("tt", "ee", "te", "bb", "pp", "tp", "ep")
                if getattr(self, "has_%s_" % n)}


class NonlinearModule:
    """NonlinearModule(input, background, perturbations, primordial) (reference: source/nonlinear_module.h) for
    `non linear = halofit`: R_NL(k,tau) = sqrt(P_NL/P_L) computed ON THE DEVICE from the resident delta_m sources
    (clpp_nonlinear_halofit).  The table stays on the device for the next TransferModule; `nl_corr_density_`
    fetches it in the reference layout [index_pk_m][tau*k_size+k]."""

    on_device = True

    def __init__(self, inputs, background_module, perturbations_module, primordial_module, fetch=False):
        ctx = perturbations_module.ctx
        self.ctx, self.info = ctx, perturbations_module.info
        m = inputs.meta
        d = capi.HalofitDesc()
        d.halofit_min_k_nonlinear = float(m.get("pr.halofit_min_k_nonlinear", 1.0e-4))   # precisions.h:432
        d.halofit_k_per_decade = float(m.get("pr.halofit_k_per_decade", 80.0))            # :436
        d.halofit_sigma_precision = float(m.get("pr.halofit_sigma_precision", 0.05))      # :441
        d.halofit_tol_sigma = float(m.get("pr.halofit_tol_sigma", 1.0e-6))                # :449
        pk = np.ascontiguousarray(primordial_module.pk_at_k(perturbations_module.k_[0]), dtype=np.float64)
        i = self.info
        out = np.empty(i.tau_size * i.k_size) if fetch else None
        imin = C.c_int()
        ctx.check(ctx._lib.clpp_nonlinear_halofit(ctx.handle, C.byref(d), capi.dptr(pk), capi.dptr(out), C.byref(imin),
                                                  ctx.err))
        self.index_tau_min_nl_ = imin.value
        self.nl_corr_density_ = [out] if fetch else None
        self.nl_corr_density_m = None  # TransferModule: None -> the device-resident table is used


class AnalyticPrimordial:
    """P_R(k) = A_s (k/k_pivot)^(n_s-1+...) -- stand-in for PrimordialModule (out of scope); tests use the
    tabulated values of the reference instead (Inputs.arrays['pm.pk_at_transfer_k'])."""

    def __init__(self, A_s, n_s, k_pivot=0.05, alpha_s=0.0):
        self.A_s, self.n_s, self.k_pivot, self.alpha_s = A_s, n_s, k_pivot, alpha_s

    def pk_at_k(self, k):
        lnk = np.log(np.asarray(k) / self.k_pivot)
        return self.A_s * np.exp((self.n_s - 1.0) * lnk + 0.5 * self.alpha_s * lnk ** 2)


class TabulatedPrimordial:
    def __init__(self, pk):
        self.pk = np.asarray(pk, dtype=np.float64)

    def pk_at_k(self, k):
        assert len(k) == len(self.pk)
        return self.pk
