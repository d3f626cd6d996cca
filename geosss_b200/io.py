"""
Pickle persistence of results with the reference's call signature (geosss/io.py:7-133):
``dump(obj, filename, gzip=False, lock=None, timeout=None)`` / ``load(filename, ...)``.

Files are interchangeable with upstream's (pickle protocol 2, optional gzip); the lock is a
``<filename>.lock`` directory as upstream.  Sample arrays keep the ``(n_chains, n_samples, d)`` layout the
scripts store (scripts/protein_reg3d3d.py:346-364) -- it is what ``Sampler.sample`` returns for N chains.
"""

import contextlib
import gzip as _gzip
import os
import pickle
import time


@contextlib.contextmanager
def _locked(filename, lock, timeout):
    if lock is None:
        yield
        return
    lockdir = filename + ".lock"
    deadline = None if not timeout or timeout <= 0 else time.time() + timeout
    while True:
        try:
            os.mkdir(lockdir)
            break
        except FileExistsError:
            if deadline is not None and time.time() > deadline:
                raise IOError("Failed to acquire Lock")
            time.sleep(0.01)
        except OSError:
            raise IOError("Failed to acquire Lock")
    try:
        yield
    finally:
        try:
            os.rmdir(lockdir)
        except OSError:
            raise IOError(f"missing lockfile {lockdir}")


def _open(filename, gzip, mode):
    return _gzip.GzipFile(filename, mode) if gzip else open(filename, mode)


def dump(this, filename, gzip=False, lock=None, timeout=None):
    """Write ``this`` to ``filename`` ('~' is expanded)."""
    filename = os.path.expanduser(filename)
    with _locked(filename, lock, timeout), _open(filename, gzip, "wb") as stream:
        pickle.dump(this, stream, protocol=2)


def load(filename, gzip=False, lock=None, timeout=None):
    """Read the object stored in ``filename``; a gzip file is recognised even without ``gzip=True``."""
    filename = os.path.expanduser(filename)
    with _locked(filename, lock, timeout):
        if not gzip:
            with open(filename, "rb") as probe:
                gzip = probe.read(2) == b"\x1f\x8b"
        with _open(filename, gzip, "rb") as stream:
            return pickle.load(stream)
