"""In-process stand-ins for the third-party modules the UNMODIFIED reference imports at module load
but that are not installed in this image: `distributed` / `dask.distributed`
(task scheduling of the gp2Scale block loop), `hgdl` (a hybrid optimiser) and `imate` (stochastic
log-determinant).  None of them performs arithmetic on the dense LML / gradient path or inside the
gp2Scale block kernels; the stub client runs `map` / `submit` eagerly in the calling process.

Used by `bench.py --impl reference` (reference copy under baseline/_ref, see install_ref.py) and by
tests/golden/ref_shim.py (reference tree under /root/reference, build container only).  This file is
ours; it contains no reference code."""
import sys
import types


class Future:
    def __init__(self, value):
        self._value = value
        self.key = "stub-%d" % id(self)

    def result(self):
        return self._value

    def release(self):
        pass


class Client:
    """Synchronous stand-in for distributed.Client: map() runs eagerly in-process."""
    id = "stub-client"

    def __init__(self, *a, **k):
        pass

    def scatter(self, data, **kwargs):
        return Future(data)

    def submit(self, fn, *args, **kwargs):
        args = [a.result() if isinstance(a, Future) else a for a in args]
        return Future(fn(*args, **kwargs))

    def map(self, fn, *iterables, **kwargs):
        out = []
        for items in zip(*iterables):
            items = [a.result() if isinstance(a, Future) else a for a in items]
            out.append(Future(fn(*items)))
        return out

    def run(self, fn, *a, **k):
        return {"stub-worker": fn()}

    def scheduler_info(self, *a, **k):
        return {"workers": {"stub-worker": {}}}

    def cancel(self, *a, **k):
        pass

    def close(self):
        pass


def as_completed(futures, with_results=False):
    for f in futures:
        yield (f, f.result()) if with_results else f


def register():
    """Put the stub modules into sys.modules (real installations, if present, win)."""
    dist = types.ModuleType("distributed")
    dist.Client = Client
    dist.Future = Future
    dist.as_completed = as_completed
    dist.get_worker = lambda: None
    dask = types.ModuleType("dask")
    dask.distributed = dist
    hgdl = types.ModuleType("hgdl")
    hgdl_hgdl = types.ModuleType("hgdl.hgdl")

    class HGDL:  # never reached on the timed path
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
    return Client
