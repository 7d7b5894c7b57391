"""The oracle against the reference's OWN compiled kernels (oracle/_ref) and against the vectors they produced.

* tests/golden/ref_kernel_vectors.npz was written by RUNNING /root/reference's kernel_api.cc (compiled unmodified,
  oracle/Makefile `ref`) -- tests/golden/gen_kernel_vectors.py.  The C restatement (ps_oracle.c) and its numpy twin
  must reproduce every output bit for bit.  Runs anywhere (the fixture is committed).
* where oracle/_ref itself is present (this container; the GPU box gets the prebuilt .so), the oracle is also
  compared with it live on fresh random inputs, and kernel_test.go's own vectors are replayed through it.
Bar: bit-exact (integer compare of the fp32 patterns).
"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import gen_kernel_vectors as G  # noqa: E402
from oracle import ps_oracle as O  # noqa: E402
from oracle import ref_kernels as R  # noqa: E402

F = np.float32
VEC = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_kernel_vectors.npz"))


def bits(a):
    return np.ascontiguousarray(a, dtype=F).view(np.uint32)


def run_oracle_c(kind, hp, g3, p, s0, s1, s2):
    for k in range(3):
        g = np.ascontiguousarray(g3[k])
        if kind == "sgd":
            O.lib.oracle_sgd(O._f32(g), O._f32(p), hp["lr"], p.size)
        elif kind == "momentum":
            O.lib.oracle_momentum(O._f32(g), O._f32(p), O._f32(s0), hp["mu"], hp["nesterov"], hp["lr"], p.size)
        elif kind == "adam":
            O.lib.oracle_adam(O._f32(g), O._f32(p), O._f32(s0), O._f32(s1), hp["lr"], p.size, hp["step"] + k,
                              hp["beta1"], hp["beta2"], hp["eps"], O._f32(s2) if hp["ams"] else O._null_f32())
        elif kind == "adagrad":
            O.lib.oracle_adagrad(O._f32(g), O._f32(p), O._f32(s0), hp["lr"], p.size, hp["eps"])


def run_oracle_np(kind, hp, g3, p, s0, s1, s2):
    for k in range(3):
        g = g3[k]
        if kind == "sgd":
            O.np_sgd(g, p, hp["lr"])
        elif kind == "momentum":
            O.np_momentum(g, p, s0, hp["mu"], hp["nesterov"], hp["lr"])
        elif kind == "adam":
            O.np_adam(g, p, s0, s1, hp["lr"], hp["step"] + k, hp["beta1"], hp["beta2"], hp["eps"],
                      s2 if hp["ams"] else None)
        elif kind == "adagrad":
            O.np_adagrad(g, p, s0, hp["lr"], hp["eps"])


CASES = {name: (kind, hp, n) for name, kind, hp, n in G.cases()}


def test_fixture_lists_the_generator_cases():
    assert list(VEC["names"]) == [c[0] for c in G.cases()]


@pytest.mark.parametrize("runner", [run_oracle_c, run_oracle_np], ids=["c", "numpy"])
@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_reference_kernel_vectors(name, runner):
    kind, hp, n = CASES[name]
    st = {k: VEC[name + "/in_" + k].copy() for k in ("p", "s0", "s1", "s2")}
    runner(kind, hp, VEC[name + "/g"], st["p"], st["s0"], st["s1"], st["s2"])
    for k in ("p", "s0", "s1", "s2"):
        assert np.array_equal(bits(st[k]), bits(VEC[name + "/out_" + k])), (name, k)


needs_ref = pytest.mark.skipif(R.lib() is None, reason="oracle/_ref not present (built only where /root/reference is)")


@needs_ref
@pytest.mark.parametrize("seed", range(6))
def test_oracle_equals_ref_live(seed):
    """Fresh random inputs each seed, sizes that exercise the SSE body and the scalar tail."""
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 3000))
    for name, kind, hp, _ in G.cases()[:13]:  # one of each kind / step
        g3 = np.stack([G.inputs(rng, n)[0] for _ in range(3)])
        base = [rng.standard_normal(n).astype(F)] + [np.abs(rng.standard_normal(n)).astype(F) for _ in range(3)]
        a = [x.copy() for x in base]
        b = [x.copy() for x in base]
        G.run_ref(kind, hp, g3, *a)
        run_oracle_c(kind, hp, g3, *b)
        for x, y in zip(a, b):
            assert np.array_equal(bits(x), bits(y)), (name, n)


@needs_ref
def test_ref_replays_kernel_test_go_vectors():
    """elasticdl/go/pkg/kernel/kernel_test.go:25-47 (SGD, exact) and :69-107 (Adam step 5, expected
    values written with the Go-side formula, compared there with tolerance 1e-4... we keep 1e-6)."""
    g = np.arange(10, dtype=F) * F(0.5) + F(0.25)
    p = np.arange(10, dtype=F) * F(-0.3) + F(1.0)
    want = p - F(0.1) * g
    R.sgd(g, p, 0.1)
    assert np.array_equal(bits(p), bits(want))
    rng = np.random.default_rng(3)
    g, p0, m0, v0 = [rng.random(10).astype(F) for _ in range(4)]
    p, m, v = p0.copy(), m0.copy(), v0.copy()
    R.adam(g, p, m, v, 0.1, 5, 0.9, 0.999, 1e-8)
    em = 0.9 * m0.astype(np.float64) + 0.1 * g
    ev = 0.999 * v0.astype(np.float64) + 0.001 * g.astype(np.float64) ** 2
    ep = p0 - 0.1 * np.sqrt(1 - 0.999 ** 5) / (1 - 0.9 ** 5) * em / (np.sqrt(ev) + 1e-8)
    assert np.allclose(m, em, rtol=1e-6) and np.allclose(v, ev, rtol=1e-6) and np.allclose(p, ep, rtol=1e-5, atol=1e-6)


@needs_ref
@pytest.mark.parametrize("dim", [1, 8])
def test_sparse_adam_rows_through_the_reference_kernel(dim):
    """The CPU arm of bench.py updates table rows with the reference's own compiled Adam (one call per row, as
    kernel.go:119-138 does through cgo): with and without the hook the tables end up bit-identical, duplicates
    applied sequentially, AMSGrad slot included."""
    import ctypes

    rng = np.random.default_rng(dim)
    ids = rng.integers(0, 50, 400).astype(np.int64)  # many duplicates: sequential application
    grads = (rng.standard_normal((400, dim)) * 0.1).astype(F)

    def run(hook):
        O.lib.oracle_set_ref_adam(ctypes.cast(R.lib().Adam, ctypes.c_void_p) if hook else None)
        try:
            tabs = [O.OracleTable(dim, "uniform", seed=3)] + [O.OracleTable(dim, "zero") for _ in range(3)]
            for step in (1, 2, 7):
                O.lib.oracle_sparse_adam(tabs[0]._h, tabs[1]._h, tabs[2]._h, tabs[3]._h, O._i64(ids), O._f32(grads),
                                         ids.size, 0.01, step, 0.9, 0.999, 1e-7)
            return [t.get(np.arange(50)) for t in tabs]
        finally:
            O.lib.oracle_set_ref_adam(None)

    for a, b in zip(run(True), run(False)):
        assert np.array_equal(bits(a), bits(b))
