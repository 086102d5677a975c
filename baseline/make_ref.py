#!/usr/bin/env python
"""Stage the UNMODIFIED reference for `bench.py --impl reference`.

The reference (jaivardhankapoor/bayesian-ode) is pure Python without a setup.py / pyproject.toml, so there is nothing to
pip-install; what the hot path needs is three source trees:
    torchdiffeq/            the integrators (odeint, odeint_adjoint)
    samplers/               SGLD / pSGLD / aSGHMC / HAMCMC / RBFKernel
    optims/                 imported by gp.py:21 (the L-BFGS baseline itself is not run)
    scripts/vanderpol/gp.py the npde field (KernelRegression, K, sq_dist) and the VDP right-hand side
This script copies them byte for byte from /root/reference (or $BODE_REFERENCE) into baseline/_ref/, which is git-ignored (no
reference source enters the history) but NOT gpurun-ignored, so it travels to the GPU box like the built .so does.  Nothing is
patched; baseline/ref_arm.py stubs matplotlib / seaborn in sys.modules before importing gp.py (its heavy code sits under
`if __name__ == '__main__'`, gp.py:529) and aliases two torch functions that were removed after the reference was written
(torch.cholesky -> torch.linalg.cholesky); that is the whole adaptation.

    python baseline/make_ref.py        # idempotent; prints what it staged
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("BODE_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")
TREES = ["torchdiffeq", "samplers", "optims"]
FILES = ["scripts/vanderpol/gp.py", "__init__.py"]


def main():
    if not os.path.isdir(REF):
        print("make_ref: %s not present (GPU box?) -- keeping whatever baseline/_ref holds" % REF)
        return 0 if os.path.isdir(DST) else 1
    os.makedirs(DST, exist_ok=True)
    manifest = []
    for tree in TREES:
        dst = os.path.join(DST, tree)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(os.path.join(REF, tree), dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    for rel in FILES:
        src = os.path.join(REF, rel)
        if not os.path.exists(src):
            continue
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
    for pkg in ("scripts", "scripts/vanderpol"):                      # importable as scripts.vanderpol.gp (namespace markers only)
        init = os.path.join(DST, pkg, "__init__.py")
        if not os.path.exists(init):
            open(init, "w").close()
    for root, _, names in os.walk(DST):
        for n in sorted(names):
            if n.endswith(".py"):
                p = os.path.join(root, n)
                rel = os.path.relpath(p, DST)
                src = os.path.join(REF, rel)
                same = os.path.exists(src) and open(src, "rb").read() == open(p, "rb").read()
                manifest.append((rel, hashlib.sha256(open(p, "rb").read()).hexdigest()[:12], "identical" if same else "marker"))
    with open(os.path.join(DST, "MANIFEST.txt"), "w") as f:
        for rel, h, tag in sorted(manifest):
            f.write("%s  %s  %s\n" % (h, tag, rel))
    n_same = sum(1 for m in manifest if m[2] == "identical")
    print("make_ref: staged %d files (%d byte-identical to %s, %d empty package markers) in %s" % (
        len(manifest), n_same, REF, len(manifest) - n_same, DST))
    return 0


if __name__ == "__main__":
    sys.exit(main())
