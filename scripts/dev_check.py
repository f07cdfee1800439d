"""Developer check: stage-by-stage parity of the CUDA path against the live oracle (oracle/_ref)."""
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from oracle.refprobe import RefCosmology
from classpp_public_b200.configs import CONFIGS
from classpp_public_b200 import modules as M
from refutil import inputs_from_reference, perturb_info_from_reference, reference_sources


def relerr(a, b):
    s = np.max(np.abs(b))
    return np.max(np.abs(a - b)) / (s if s > 0 else 1.0)


def main(name, stages):
    t = time.time()
    ref = RefCosmology(CONFIGS[name], threads=os.cpu_count()).compute("lensing")
    print("[%s] reference total %.2fs perturb %.3f transfer %.3f spectra %.3f" % (
        name, time.time() - t, ref.scalar("time.perturb"), ref.scalar("time.transfer"), ref.scalar("time.spectra")))
    inp = inputs_from_reference(ref)
    rsrc = reference_sources(ref)
    nk, nt = ref.iscalar("pt.k_size"), ref.iscalar("pt.tau_size")
    rtr = ref.get("tr.transfer")
    rcl = ref.get("sp.cl")
    pk = ref.get("pm.pk_at_transfer_k")
    nl = ref.get("nl.nl_corr_density_m") if int(inp.meta["nl.method"]) != 0 else None

    class NL:
        nl_corr_density_m = nl

    if "spectra" in stages:
        ctx = M.Context(0)
        bg = M.BackgroundModule(inp, ctx); th = M.ThermodynamicsModule(inp, bg)
        pt = M.PerturbationsModule.from_sources(inp, bg, ref.get("pt.k"), ref.get("pt.tau_sampling"), rsrc,
                                                perturb_info_from_reference(ref))
        tr = M.TransferModule(inp, bg, th, pt, compute=False)
        tr.set_transfer(rtr)
        t = time.time()
        sp = M.SpectraModule(inp, pt, M.TabulatedPrimordial(pk), None, tr)
        print("  spectra: %.4fs  max rel err per ct:" % (time.time() - t),
              [float("%.2e" % relerr(sp.cl_[0].reshape(-1, sp.ct_size_)[:, c], rcl.reshape(-1, sp.ct_size_)[:, c]))
               for c in range(sp.ct_size_)])
        ctx.close()
    if "transfer" in stages:
        ctx = M.Context(0)
        bg = M.BackgroundModule(inp, ctx); th = M.ThermodynamicsModule(inp, bg)
        pt = M.PerturbationsModule.from_sources(inp, bg, ref.get("pt.k"), ref.get("pt.tau_sampling"), rsrc,
                                                perturb_info_from_reference(ref))
        t = time.time()
        tr = M.TransferModule(inp, bg, th, pt, NL if nl is not None else None)
        dt = time.time() - t
        mine = tr.transfer_[0].reshape(tr.info.tt_size, tr.info.l_size, tr.info.q_size)
        r = rtr.reshape(mine.shape)
        print("  transfer: %.4fs" % dt, "per-tt max|diff|/max|ref|:",
              [float("%.2e" % relerr(mine[i], r[i])) for i in range(mine.shape[0])])
        sp = M.SpectraModule(inp, pt, M.TabulatedPrimordial(pk), None, tr)
        print("  -> cl from own transfer, rel err per ct:",
              [float("%.2e" % np.max(np.abs(sp.cl_[0].reshape(-1, sp.ct_size_)[:, c] / np.where(rcl.reshape(-1, sp.ct_size_)[:, c] != 0, rcl.reshape(-1, sp.ct_size_)[:, c], 1) - 1) * (rcl.reshape(-1, sp.ct_size_)[:, c] != 0)))
               for c in range(sp.ct_size_)])
        ctx.close()
    if "perturb" in stages:
        ctx = M.Context(0)
        bg = M.BackgroundModule(inp, ctx); th = M.ThermodynamicsModule(inp, bg)
        t = time.time()
        pt = M.PerturbationsModule(inp, bg, th)
        dt = time.time() - t
        ks = pt.kstat_
        print("  perturb: %.4fs  steps %d failed %d fevals %d jac %d lu %d solves %d  max steps/k %d" % (
            dt, ks[:, 0].sum(), ks[:, 1].sum(), ks[:, 2].sum(), ks[:, 3].sum(), ks[:, 4].sum(), ks[:, 5].sum(), ks[:, 0].max()))
        mine = np.stack(pt.sources_[0]).reshape(-1, nt, nk)
        r = rsrc.reshape(-1, nt, nk)
        for tp in range(mine.shape[0]):
            d = np.abs(mine[tp] - r[tp])
            scale = np.max(np.abs(r[tp]), axis=0, keepdims=True)  # per-k scale
            rel = d / np.where(scale > 0, scale, 1)
            ik = np.unravel_index(np.argmax(rel), rel.shape)
            print("    tp %d: max|diff|/max_tau|ref| = %.2e at (tau idx %d, k idx %d)" % (tp, rel.max(), ik[0], ik[1]))
        t = time.time()
        tr = M.TransferModule(inp, bg, th, pt, NL if nl is not None else None)
        sp = M.SpectraModule(inp, pt, M.TabulatedPrimordial(pk), None, tr)
        print("  transfer+spectra %.4fs" % (time.time() - t))
        cl = sp.cl_[0].reshape(-1, sp.ct_size_); rc = rcl.reshape(-1, sp.ct_size_)
        for c, nm in enumerate(["tt", "ee", "te", "bb", "pp", "tp", "ep"][:sp.ct_size_]):
            if nm in ("tt", "ee", "pp"):
                print("    C_l %s full-pipeline max rel err %.2e" % (nm, np.max(np.abs(cl[:, c] / rc[:, c] - 1))))
        ctx.close()


if __name__ == "__main__":
    name = sys.argv[1] if len(sys.argv) > 1 else "lcdm_coarse"
    stages = sys.argv[2].split(",") if len(sys.argv) > 2 else ["spectra", "transfer", "perturb"]
    main(name, stages)
