"""Input parameter sets (the reference's .ini surface) for the BASELINE.json configs.

Keys/values are exactly what one would write in a CLASS++ .ini/.pre file
(reference: explanatory.ini, base_2018_plikHM_TTTEEE_lowl_lowE_lensing.ini).
"""

# config 1: minimal LambdaCDM, output=tCl,pCl,lCl,mPk, default precision, l_max_scalars=2500
LCDM = {
    "output": "tCl,pCl,lCl,mPk",
    "lensing": "yes",
    "l_max_scalars": 2500,
}

# config 2: Planck-2018 best fit (base_2018_plikHM_TTTEEE_lowl_lowE_lensing.ini), P(k) to k=1 h/Mpc
PLANCK18 = {
    "H0": 67.32117,
    "omega_b": 0.02238280,
    "N_ur": 2.03066666667,
    "omega_cdm": 0.1201075,
    "N_ncdm": 1,
    "omega_ncdm": 0.0006451439,
    "YHe": 0.2454006,
    "tau_reio": 0.05430842,
    "n_s": 0.9660499,
    "A_s": 2.100549e-09,
    "non linear": "halofit",
    "output": "tCl,pCl,lCl,mPk",
    "lensing": "yes",
    "P_k_max_h/Mpc": 1.0,
}

# config 2 without halofit / ncdm (massless), used as an intermediate parity case
PLANCK18_LINEAR = {k: v for k, v in PLANCK18.items() if k != "non linear"}

# config 4b: 3 degenerate massive neutrinos, m = 0.02 eV
NCDM3_DEG = dict(LCDM, **{"N_ur": 0.00641, "N_ncdm": 1, "deg_ncdm": 3, "m_ncdm": 0.02})

# a cheap, coarse variant of config 1 for fast CPU-side tests
# (same knobs the reference's own valgrind runs use, test_nightly.yml)
LCDM_COARSE = dict(LCDM, **{
    "l_max_scalars": 600,
    "k_step_sub": 0.2,
    "k_step_super": 0.02,
    "k_per_decade_for_pk": 4,
    "k_per_decade_for_bao": 10,
    "perturb_sampling_stepsize": 0.3,
    "q_linstep": 1.5,
})

# config 4 (three separate species, neq up to 316, 58 hub variables) on the coarse grids: exercises the large-system
# fallbacks of the device path (generic NDF, shared-memory Gauss-Jordan, 18 chains)
NCDM3_COARSE = dict(LCDM_COARSE, **{"N_ur": 0.00641, "N_ncdm": 3, "m_ncdm": "0.02,0.02,0.02"})

# config 4 as written in BASELINE.json: three separate species on the default grids (neq up to 316)
NCDM3 = dict(LCDM, **{"N_ur": 0.00641, "N_ncdm": 3, "m_ncdm": "0.02,0.02,0.02"})

# config 3 stand-in (cl_permille.pre is not in the reference tree): denser k sampling, larger photon / ur
# hierarchies, tighter integration tolerance, finer time sampling and l grid
LCDM_DENSE = dict(LCDM, **{
    "k_step_sub": 0.025,
    "k_step_super": 0.001,
    "l_max_g": 30,
    "l_max_pol_g": 20,
    "l_max_ur": 30,
    "tol_perturb_integration": 1.0e-6,
    "perturb_sampling_stepsize": 0.05,
    "l_logstep": 1.06,
    "l_linstep": 25,
    "q_linstep": 0.3,
})

CONFIGS = {
    "lcdm": LCDM,
    "planck18": PLANCK18,
    "planck18_linear": PLANCK18_LINEAR,
    "ncdm3_deg": NCDM3_DEG,
    "lcdm_coarse": LCDM_COARSE,
    "lcdm_dense": LCDM_DENSE,
    "ncdm3_coarse": NCDM3_COARSE,
    "ncdm3": NCDM3,
}
