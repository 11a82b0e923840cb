"""Install the UNMODIFIED reference package into baseline/_ref/ (git-ignored, shipped to the GPU box by gpurun).

The contract's recipe
    python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref /root/reference
fails in this image: the reference's build backend is hatchling + hatch-vcs (pyproject.toml), neither is installed and
there is no network (`ModuleNotFoundError: No module named 'hatchling'`, recorded in DESIGN.md).  The reference is a
pure-Python package, so what a wheel install would do is reproduced by hand: copy the package directory verbatim and
write the `_version.py` hatch-vcs generates at build time.  Nothing is patched.  Run from build() whenever
/root/reference is present; on the GPU box the prebuilt copy is used as is.

    python baseline/install_ref.py
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/fvgp"
DST = os.path.join(HERE, "_ref")


def install(force=False):
    pkg = os.path.join(DST, "fvgp")
    if not os.path.isdir(SRC):
        return os.path.isdir(pkg)
    if os.path.isdir(pkg) and not force:
        same = all(os.path.exists(os.path.join(pkg, f)) and
                   os.path.getsize(os.path.join(pkg, f)) == os.path.getsize(os.path.join(SRC, f))
                   for f in os.listdir(SRC) if f.endswith(".py"))
        if same:
            return True
    os.makedirs(DST, exist_ok=True)
    if os.path.isdir(pkg):
        shutil.rmtree(pkg)
    shutil.copytree(SRC, pkg, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    with open(os.path.join(pkg, "_version.py"), "w") as fh:
        fh.write("__version__ = '0+reference.unmodified'\n")
    return True


def import_reference():
    """Import the installed reference as module `fvgp` (with the scheduler / optimiser stubs registered).
    Raises ImportError when baseline/_ref is absent."""
    pkg = os.path.join(DST, "fvgp")
    if not os.path.isdir(pkg):
        raise ImportError("baseline/_ref/fvgp is missing: run `python baseline/install_ref.py` where /root/reference exists")
    if HERE not in sys.path:
        sys.path.insert(0, HERE)
    import ref_stubs
    ref_stubs.register()
    if DST not in sys.path:
        sys.path.insert(0, DST)
    import fvgp
    assert os.path.dirname(os.path.abspath(fvgp.__file__)) == pkg, "a different `fvgp` shadows baseline/_ref"
    return fvgp


if __name__ == "__main__":
    ok = install(force="--force" in sys.argv)
    print("baseline/_ref:", "installed" if ok else "reference tree not present")
