"""The C-ABI library loads and exports every symbol that include/clpp.h declares (no GPU needed)."""
import ctypes
import os
import re

from classpp_public_b200 import _capi as capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "clpp.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(clpp_[a-z_0-9]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert header_symbols() == sorted(capi.SYMBOLS)


def test_library_exports_every_symbol():
    lib = ctypes.CDLL(capi.LIB_PATH)
    for s in header_symbols():
        assert hasattr(lib, s), "libclpp.so does not export " + s


def test_version_and_host_only_context():
    L = capi.lib()
    assert b"sm_100a" in L.clpp_version()
    h = ctypes.c_void_p()
    err = ctypes.create_string_buffer(capi.ERRLEN)
    assert L.clpp_ctx_create(-1, ctypes.byref(h), err) == 0
    # compute entry points must refuse to run without a device: no CPU fallback
    assert L.clpp_perturb_solve(h, 0, 0, err) != 0
    L.clpp_ctx_destroy(h)


def test_struct_sizes_match_header():
    """ctypes mirrors must have the C layout: compare with sizes compiled from the header."""
    import subprocess, tempfile, textwrap
    src = textwrap.dedent("""
        #include <stdio.h>
        #include "clpp.h"
        int main(void){ printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(clpp_background_desc), sizeof(clpp_thermo_desc),
          sizeof(clpp_perturb_desc), sizeof(clpp_perturb_info), sizeof(clpp_kstat), sizeof(clpp_transfer_desc),
          sizeof(clpp_transfer_info), sizeof(clpp_spectra_info), sizeof(clpp_halofit_desc), sizeof(clpp_lensing_desc),
          sizeof(clpp_lensing_info)); return 0; }""")
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    mine = [ctypes.sizeof(t) for t in (capi.BackgroundDesc, capi.ThermoDesc, capi.PerturbDesc, capi.PerturbInfo,
                                       capi.KStat, capi.TransferDesc, capi.TransferInfo, capi.SpectraInfo,
                                       capi.HalofitDesc, capi.LensingDesc, capi.LensingInfo)]
    assert sizes == mine
