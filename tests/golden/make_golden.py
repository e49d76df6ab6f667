"""Generates tests/golden/c0_ocp_golden.npz: oracle outputs for the reference VGP (C0) at the
committed seeded decision vector. The reference itself cannot run here (PSOPT/ADOL-C/IPOPT absent,
SURVEY.md section 8c), so these are ORACLE-generated regression vectors, pinned in turn by the
hand-derived known answers of appendix_b_kats.json. Run from the repo root:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_binding as ob  # noqa: E402
from etol_b200 import workloads as W  # noqa: E402

wl = W.reference_vgp("ocp")
o = ob.Oracle(wl)
fd = o.eval(wl.x, want=("f", "g", "jac", "grad"), jac_mode=1)
ex = o.eval(wl.x, want=("jac",), jac_mode=0)
irow, jcol, grp = o.structure()
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "c0_ocp_golden.npz"), x=wl.x, f=fd["f"], g=fd["g"],
                    jac_fd=fd["jac"], jac_exact=ex["jac"], grad=fd["grad"], irow=irow, jcol=jcol, group_of_col=grp)
print("wrote c0_ocp_golden.npz", o.nvars, o.ncons, o.nnz)

# ---- second fixture: Hessian of the Lagrangian, discretisation error, resampling (C0) and a user model ------
rng = np.random.default_rng(20261018)
lam = rng.normal(size=(1, o.ncons))
sigma = np.array([1.25])
hi, hj = ob.hess_structure(o)
from etol_b200 import tape as T  # noqa: E402
uw = W.unicycle(batch=1, nnodes=17, ncyl=2, ntracks=1)
uo = ob.Oracle(uw)
ufd = uo.eval(uw.x, want=("f", "g", "jac"), jac_mode=1)
uex = uo.eval(uw.x, want=("jac",), jac_mode=0)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "c0_ext_golden.npz"), x=wl.x, lam=lam, sigma=sigma,
                    hess=ob.eval_hess(o, wl.x, sigma, lam), hess_irow=hi, hess_jcol=hj,
                    ode_error=ob.ode_error(o, wl, wl.x), resample_41=ob.resample(o, wl, wl.x, [41]),
                    user_x=uw.x, user_f=ufd["f"], user_g=ufd["g"], user_jac_fd=ufd["jac"], user_jac_exact=uex["jac"])
print("wrote c0_ext_golden.npz")

# ---- third fixture: outputs of the REFERENCE'S OWN callbacks (oracle/_ref, `make -C oracle ref`) ----------------
# /root/reference/src/ePSOPT/ePSOPT.cpp + src/Examples/PSOPT/etol_psopt_example1.cpp compiled unmodified against
# oracle/refstub/psopt.h, run at every node of the seeded C0 decision vector. Needs /root/reference (this container).
import ref_binding as rb  # noqa: E402
if rb.available():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_oracle_vs_reference import _run_reference  # noqa: E402
    ref = rb.Reference(rb.REF_XML)
    out = _run_reference(ref, wl, wl.x)
    bd = ref.bounds()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "c0_ref_pernode.npz"), x=wl.x, t=out["t"], f=out["f"],
                        path=out["path"], df=out["df"], dpath=out["dpath"], L=out["L"], events=out["events"],
                        endpoint=out["endpoint"], bounds_events=bd["events"], bounds_path=bd["path"],
                        bounds_states=bd["states"], bounds_controls=bd["controls"])
    print("wrote c0_ref_pernode.npz from oracle/_ref (reference callbacks)")
else:
    print("oracle/_ref not available: c0_ref_pernode.npz left as committed")
