"""oracle/_ref -- the reference's own C++ optimizer kernels -- TEST INFRASTRUCTURE.

``/root/reference/elasticdl/go/pkg/kernel/capi/kernel_api.cc`` compiled UNMODIFIED (oracle/Makefile
target ``ref``; its Eigen include is served by oracle/eigen_shim, see that header for what the
stand-in does and does not guarantee) into ``oracle/_ref/libkernel_api_ref.so`` and bound with the
signatures of ``kernel_api.h:10-37`` -- the same C ABI the Go PS binds through cgo
(``go/pkg/kernel/kernel.go:3-6``).

Used by tests/ to validate the restatement in ps_oracle.c and to generate the golden vectors in
tests/golden/ref_kernel_vectors.npz (tests/golden/gen_kernel_vectors.py).  Built only where
/root/reference exists; the .so is git-ignored and travels to the GPU box with the snapshot.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libkernel_api_ref.so")
REF_SRC = "/root/reference/elasticdl/go/pkg/kernel/capi/kernel_api.cc"


def build():
    """Compile oracle/_ref when the reference tree is present; returns the path or None."""
    if os.path.exists(REF_SRC):
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref"])
    return SO if os.path.exists(SO) else None


_lib = None


def lib():
    global _lib
    if _lib is None:
        try:
            so = build()
            if so is None:
                return None
            L = ctypes.CDLL(so)
        except (OSError, subprocess.CalledProcessError):  # no compiler / an unloadable prebuilt file: run without it
            return None
        f32p = ctypes.POINTER(ctypes.c_float)
        f, ll, b = ctypes.c_float, ctypes.c_longlong, ctypes.c_bool
        L.SGD.argtypes = [f32p, f32p, f, ll]  # kernel_api.h:10
        L.Momentum.argtypes = [f32p, f32p, f32p, f, b, f, ll]  # kernel_api.h:12-18
        L.Adam.argtypes = [f32p, f32p, f32p, f32p, f, ll, ll, f, f, f, f32p]  # kernel_api.h:20-30
        L.Adagrad.argtypes = [f32p, f32p, f32p, f, ll, f]  # kernel_api.h:32-37
        for fn in (L.SGD, L.Momentum, L.Adam, L.Adagrad):
            fn.restype = None
        _lib = L
    return _lib


def _p(a):
    assert a.dtype == np.float32 and a.flags.c_contiguous
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def sgd(g, p, lr):
    lib().SGD(_p(g), _p(p), lr, p.size)


def momentum(g, p, v, mu, nesterov, lr):
    lib().Momentum(_p(g), _p(p), _p(v), mu, bool(nesterov), lr, p.size)


def adam(g, p, m, v, lr, step, beta1, beta2, eps, max_square=None):
    ms = _p(max_square) if max_square is not None else ctypes.POINTER(ctypes.c_float)()
    lib().Adam(_p(g), _p(p), _p(m), _p(v), lr, p.size, step, beta1, beta2, eps, ms)


def adagrad(g, p, m, lr, eps):
    lib().Adagrad(_p(g), _p(p), _p(m), lr, p.size, eps)
