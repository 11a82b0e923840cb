"""Make the UNMODIFIED reference (/root/reference/fvgp) importable in the build container.

Only used to *generate* the committed golden fixtures and to pin the numpy oracle
(`oracle/fvgp_oracle.py`) against the real reference.  Nothing in the product, the
`-m gpu` tests, `smoke()` or `bench.py` imports this: /root/reference does not exist on
the GPU box.

The reference does not import as shipped: it needs a generated
`_version.py` and the third-party modules `distributed`, `dask.distributed`, `hgdl`
and `imate`, none of which is installed and none of which performs hot-path arithmetic
(dask only schedules the block tasks).  We therefore build a throw-away package
directory of symlinks to the reference files plus `_version.py`, and register
in-process synchronous stubs for the schedulers.  No reference source is copied into
this repository.
"""
import os
import sys
import tempfile
import types

REFERENCE = "/root/reference/fvgp"


class _Future:
    def __init__(self, value):
        self._value = value
        self.key = "stub-%d" % id(self)

    def result(self):
        return self._value

    def release(self):
        pass


class _Client:
    """Synchronous stand-in for distributed.Client: map() runs eagerly in-process."""
    id = "stub-client"

    def __init__(self, *a, **k):
        pass

    def scatter(self, data, **kwargs):
        return _Future(data)

    def submit(self, fn, *args, **kwargs):
        args = [a.result() if isinstance(a, _Future) else a for a in args]
        return _Future(fn(*args, **kwargs))

    def map(self, fn, *iterables, **kwargs):
        out = []
        for items in zip(*iterables):
            items = [a.result() if isinstance(a, _Future) else a for a in items]
            out.append(_Future(fn(*items)))
        return out

    def run(self, fn, *a, **k):
        return {"stub-worker": fn()}

    def scheduler_info(self, *a, **k):
        return {"workers": {"stub-worker": {}}}

    def cancel(self, *a, **k):
        pass

    def close(self):
        pass


def _as_completed(futures, with_results=False):
    for f in futures:
        yield (f, f.result()) if with_results else f


def install():
    """Returns the imported reference package (module object named `fvgp`)."""
    if not os.path.isdir(REFERENCE):
        raise RuntimeError("reference tree not present (expected in the build container only)")
    if "fvgp" in sys.modules and getattr(sys.modules["fvgp"], "__ref_shim__", False):
        return sys.modules["fvgp"]
    root = tempfile.mkdtemp(prefix="fvgp_ref_")
    pkg = os.path.join(root, "fvgp")
    os.mkdir(pkg)
    for name in os.listdir(REFERENCE):
        if name.endswith(".py"):
            os.symlink(os.path.join(REFERENCE, name), os.path.join(pkg, name))
    with open(os.path.join(pkg, "_version.py"), "w") as fh:
        fh.write("__version__ = '0+reference'\n")

    dist = types.ModuleType("distributed")
    dist.Client = _Client
    dist.Future = _Future
    dist.as_completed = _as_completed
    dist.get_worker = lambda: None
    dask = types.ModuleType("dask")
    dask.distributed = dist
    hgdl = types.ModuleType("hgdl")
    hgdl_hgdl = types.ModuleType("hgdl.hgdl")

    class HGDL:  # never reached by the golden generators
        def __init__(self, *a, **k):
            raise RuntimeError("hgdl stub")
    hgdl_hgdl.HGDL = HGDL
    hgdl.hgdl = hgdl_hgdl
    imate = types.ModuleType("imate")

    def _logdet(*a, **k):
        raise RuntimeError("imate stub: stochastic logdet is not available")
    imate.logdet = _logdet
    for name, mod in (("distributed", dist), ("dask", dask), ("dask.distributed", dist),
                      ("hgdl", hgdl), ("hgdl.hgdl", hgdl_hgdl), ("imate", imate)):
        sys.modules.setdefault(name, mod)
    sys.path.insert(0, root)
    import fvgp  # noqa: E402
    fvgp.__ref_shim__ = True
    fvgp.__stub_client__ = _Client
    return fvgp
