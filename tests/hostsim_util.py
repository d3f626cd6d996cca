"""
Builds and binds the TEST-ONLY host build of the device math (tests/hostsim/hostsim.cpp).
Lets the CPU test-suite check csrc/{philox,targets,transition}.cuh against the golden vectors
without a GPU.  Never imported by the product.
"""

import ctypes as C
import os
import subprocess

import numpy as np
from scipy.special import i0

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp")
OUT_DIR = os.path.join(ROOT, "tests", "_hostsim")
OUT = os.path.join(OUT_DIR, "libhostsim.so")
CSRC = os.path.join(ROOT, "geosss_b200", "csrc")

_pd = C.POINTER(C.c_double)


class Spec(C.Structure):
    _fields_ = [
        ("kind", C.c_int), ("d", C.c_int),
        ("K", C.c_int), ("single", C.c_int), ("mu", _pd), ("logw", _pd), ("logi0", _pd),
        ("A", _pd), ("b", _pd),
        ("M", C.c_int), ("kappa", C.c_double), ("knots", _pd), ("ang", _pd), ("cosang", _pd), ("sinang", _pd),
        ("nt", C.c_int), ("ns", C.c_int), ("k", C.c_int), ("outlier", C.c_int),
        ("sigma2", C.c_double), ("log_const", C.c_double), ("log_outlier", C.c_double), ("beta", C.c_double),
        ("tgt", _pd), ("wt", _pd), ("src", _pd), ("logws", _pd),
        ("shift", C.c_double * 3), ("dim", C.c_int),
    ]


def _newest_source():
    files = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    return max(os.path.getmtime(f) for f in files)


def load():
    os.makedirs(OUT_DIR, exist_ok=True)
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < _newest_source():
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-I", CSRC, SRC, "-o", OUT])
    lib = C.CDLL(OUT)
    assert lib.hs_spec_size() == C.sizeof(Spec)
    lib.hs_logp.argtypes = [C.POINTER(Spec), C.c_int, _pd, _pd, C.c_int64]
    lib.hs_run.argtypes = [C.POINTER(Spec), C.c_int, C.c_int, _pd, C.c_int64, C.c_uint64, C.c_uint64, _pd, C.c_int64,
                           C.c_uint64, _pd, C.POINTER(C.c_int32)]
    lib.hs_run.restype = C.c_int
    lib.hs_grad.argtypes = [C.POINTER(Spec), C.c_int, _pd, _pd, C.c_int64]
    lib.hs_run_mh.argtypes = [C.POINTER(Spec), C.c_int, C.c_int, _pd, C.c_int64, C.c_uint64, C.c_uint64, _pd, C.c_int64,
                              C.c_uint64, C.c_double, C.c_double, C.c_int, C.c_int64, _pd, _pd, _pd, C.POINTER(C.c_int32),
                              C.POINTER(C.c_int64)]
    lib.hs_run_mh.restype = C.c_int
    lib.hs_eigh.argtypes = [C.c_int, _pd, _pd, _pd]
    return lib


def _p(a):
    return a.ctypes.data_as(_pd)


class HostTarget:
    """Keeps the float64 parameter arrays alive next to the Spec that points into them."""

    def __init__(self, spec, keep):
        self.spec, self.keep = spec, keep


def vmf_mixture(mus, weights=None, single=False):
    mus = np.ascontiguousarray(np.atleast_2d(mus), dtype=float)
    K, d = mus.shape
    w = np.ones(K) if weights is None else np.asarray(weights, float)
    logw = np.log(w / w.sum())
    with np.errstate(over="ignore"):
        li0 = np.log(i0(np.linalg.norm(mus, axis=1)))
    s = Spec(kind=0, d=d, K=K, single=int(single), mu=_p(mus), logw=_p(logw), logi0=_p(li0))
    return HostTarget(s, (mus, logw, li0))


def bingham(A, b=None):
    A = np.ascontiguousarray(A, dtype=float)
    s = Spec(kind=1, d=len(A), A=_p(A))
    keep = [A]
    if b is not None:
        b = np.ascontiguousarray(b, dtype=float)
        s.b = _p(b)
        keep.append(b)
    return HostTarget(s, keep)


def curved(knots, kappa):
    knots = np.ascontiguousarray(knots, dtype=float)
    ang = np.arccos(np.clip(np.sum(knots[:-1] * knots[1:], axis=1), -1, 1))
    ca, sa = np.cos(ang), np.sin(ang)
    s = Spec(kind=2, d=knots.shape[1], M=len(knots), kappa=float(kappa), knots=_p(knots), ang=_p(ang), cosang=_p(ca), sinang=_p(sa))
    return HostTarget(s, (knots, ang, ca, sa))


def cpd(target, source, sigma, k, omega, beta=1.0, wt=None, ws=None, outlier=True, translation=None):
    tgt = np.ascontiguousarray(target, dtype=float)
    src = np.ascontiguousarray(source, dtype=float)
    dim = tgt.shape[1]
    if dim == 2:  # 3D -> 2D projection: (x, y, 0) rows
        tgt = np.ascontiguousarray(np.hstack([tgt, np.zeros((len(tgt), 1))]))
    wt = np.ones(len(tgt)) if wt is None else np.ascontiguousarray(wt, dtype=float)
    lws = np.zeros(len(src)) if ws is None else np.log(np.asarray(ws, float))
    gauss = dim / 2 * np.log(2 * np.pi * sigma**2)
    if outlier:
        lo = np.log(omega + 1e-308) - np.sum(np.log(np.ptp(tgt[:, :dim], 0)))
        with np.errstate(divide="ignore"):
            lc = np.log(1 - omega) - gauss
    else:
        lo, lc = -np.inf, -gauss
    s = Spec(kind=3, d=4, nt=len(tgt), ns=len(src), k=int(k), outlier=int(outlier), sigma2=float(sigma) ** 2,
             log_const=float(lc), log_outlier=float(lo), beta=float(beta), tgt=_p(tgt), wt=_p(wt), src=_p(src), logws=_p(lws), dim=dim)
    if translation is not None:
        for a, v in enumerate(translation):
            s.shift[a] = float(v)
    return HostTarget(s, (tgt, wt, src, lws))


def logp(lib, tgt, x, dtype):
    x = np.ascontiguousarray(np.atleast_2d(x), dtype=float)
    out = np.empty(len(x))
    lib.hs_logp(C.byref(tgt.spec), 1 if dtype == np.float64 else 0, _p(x), _p(out), len(x))
    return out


def run(lib, tgt, kind, x0, n_steps, dtype, tape=None, seed=0, chain=0, step0=0):
    x0 = np.ascontiguousarray(x0, dtype=float)
    samples = np.empty((n_steps, tgt.spec.d))
    nrej = np.empty(n_steps, dtype=np.int32)
    tp = None if tape is None else np.ascontiguousarray(tape, dtype=float)
    rc = lib.hs_run(C.byref(tgt.spec), 1 if dtype == np.float64 else 0, kind, None if tp is None else _p(tp),
                    0 if tp is None else len(tp), seed, chain, _p(x0), n_steps, step0, _p(samples),
                    nrej.ctypes.data_as(C.POINTER(C.c_int32)))
    return rc, samples, nrej


def grad(lib, tgt, x, dtype):
    x = np.ascontiguousarray(np.atleast_2d(x), dtype=float)
    out = np.empty_like(x)
    lib.hs_grad(C.byref(tgt.spec), 1 if dtype == np.float64 else 0, _p(x), _p(out), len(x))
    return out


def run_mh(lib, tgt, kind, x0, n_steps, dtype, stepsize, tape=None, seed=0, chain=0, step0=0, alpha=0.5, n_leap=10, adapt=0):
    """One chain of a baseline sampler (kind 2 RWMH, 3 independence, 4 mixture, 5 HMC).
    Returns rc, positions, momenta, step sizes, accepted flags, random-walk proposal count."""
    x0 = np.ascontiguousarray(x0, dtype=float)
    d = tgt.spec.d
    samples, momenta = np.empty((n_steps, d)), np.empty((n_steps, d))
    steps, acc = np.empty(n_steps), np.empty(n_steps, dtype=np.int32)
    n_walk = C.c_int64(0)
    tp = None if tape is None else np.ascontiguousarray(tape, dtype=float)
    rc = lib.hs_run_mh(C.byref(tgt.spec), 1 if dtype == np.float64 else 0, kind, None if tp is None else _p(tp),
                       0 if tp is None else len(tp), seed, chain, _p(x0), n_steps, step0, stepsize, alpha, n_leap, adapt,
                       _p(samples), _p(momenta), _p(steps), acc.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(n_walk))
    return rc, samples, momenta, steps, acc, n_walk.value
