"""Generates tests/golden/ref_kernel_vectors.npz by RUNNING the reference's own C++ kernels
(/root/reference/elasticdl/go/pkg/kernel/capi/kernel_api.cc, compiled unmodified as oracle/_ref --
see oracle/Makefile `ref` and oracle/eigen_shim) on seeded inputs.  Run in the build container
(where /root/reference exists):  python tests/golden/gen_kernel_vectors.py

Cases: SGD, Momentum, Nesterov, Adam and AMSGrad at steps 1 / 5 / 1000 / 100000 (the bias correction is
evaluated in double and narrowed, kernel_api.cc:67), Adagrad; sizes 1, 7 (ragged), 10 (kernel_test.go's
size), 515; three successive applications each so slot state feeds back.  Inputs include negatives,
zeros and tiny / large magnitudes.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_kernels as R  # noqa: E402

F = np.float32


def inputs(rng, n):
    g = (rng.standard_normal(n) * 10.0 ** rng.integers(-4, 2, n)).astype(F)
    p = rng.standard_normal(n).astype(F)
    if n > 3:
        g[1] = 0.0
        p[2] = 0.0
        g[3] = F(1e-20)
    return g, p


def cases():
    """(name, kind, hyper-parameters, n) -- shared with the tests through the npz itself."""
    out = []
    for n in (1, 7, 10, 515):
        out.append(("sgd_n%d" % n, "sgd", dict(lr=0.1), n))
        out.append(("momentum_n%d" % n, "momentum", dict(mu=0.9, nesterov=0, lr=0.05), n))
        out.append(("nesterov_n%d" % n, "momentum", dict(mu=0.9, nesterov=1, lr=0.05), n))
        for step in (1, 5, 1000, 100000):
            out.append(("adam_s%d_n%d" % (step, n), "adam",
                        dict(lr=0.1 if step == 5 else 0.001, step=step, beta1=0.9, beta2=0.999, eps=1e-8, ams=0), n))
            out.append(("amsgrad_s%d_n%d" % (step, n), "adam",
                        dict(lr=0.001, step=step, beta1=0.9, beta2=0.999, eps=1e-8, ams=1), n))
        out.append(("adagrad_n%d" % n, "adagrad", dict(lr=0.05, eps=1e-7), n))
    return out


def run_ref(kind, hp, g3, p, s0, s1, s2):
    """Three applications in place with the reference kernels; gradient k of g3 in round k."""
    for k in range(3):
        g = np.ascontiguousarray(g3[k])
        if kind == "sgd":
            R.sgd(g, p, hp["lr"])
        elif kind == "momentum":
            R.momentum(g, p, s0, hp["mu"], hp["nesterov"], hp["lr"])
        elif kind == "adam":
            R.adam(g, p, s0, s1, hp["lr"], hp["step"] + k, hp["beta1"], hp["beta2"], hp["eps"],
                   s2 if hp["ams"] else None)
        elif kind == "adagrad":
            R.adagrad(g, p, s0, hp["lr"], hp["eps"])


def main():
    assert R.lib() is not None, "oracle/_ref not built (needs /root/reference)"
    rng = np.random.default_rng(20260921)
    out = {}
    names = []
    for name, kind, hp, n in cases():
        g3 = np.stack([inputs(rng, n)[0] for _ in range(3)])
        _, p = inputs(rng, n)
        s0 = np.abs(rng.standard_normal(n)).astype(F) * F(0.1)
        s1 = np.abs(rng.standard_normal(n)).astype(F) * F(0.01)
        s2 = np.abs(rng.standard_normal(n)).astype(F) * F(0.01)
        if kind == "adagrad":
            s0[:] = 0  # Q4: the Go PS starts the accumulator at 0
        out[name + "/g"] = g3
        for k, a in (("p", p), ("s0", s0), ("s1", s1), ("s2", s2)):
            out[name + "/in_" + k] = a.copy()
        run_ref(kind, hp, g3, p, s0, s1, s2)
        for k, a in (("p", p), ("s0", s0), ("s1", s1), ("s2", s2)):
            out[name + "/out_" + k] = a
        out[name + "/hp"] = np.array([kind] + ["%s=%r" % kv for kv in sorted(hp.items())])
        names.append(name)
    out["names"] = np.array(names)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_kernel_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(names), "cases", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
