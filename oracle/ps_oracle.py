"""CPU oracle for the ElasticDL parameter-server hot path -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module; nothing under elasticdl_b200/ does.

Three layers, each citing the reference file:line it follows (paths relative to
/root/reference/elasticdl/):

* ``lib``         ctypes binding of oracle/libps_oracle.so (ps_oracle.c): the
                  arithmetic (kernel_api.cc), the lazy row table
                  (embedding_table.go) and dedup (tensor_utils.py).
* ``np_*``        a numpy twin of the dense kernels (same op order, float32
                  IEEE, no FMA) used to cross-check the C restatement.
* ``OracleServer``/``OraclePSClient``  the control logic of
                  go/pkg/ps/{server,optimizer,model}.go and
                  python/worker/ps_client.py composed in-process (no gRPC).

Parity status: the dense kernels equal the reference's own kernel_api.cc
(compiled unmodified as oracle/_ref, see ref_kernels.py) bit for bit
(tests/test_oracle_vs_ref.py); everything is pinned by the reference's golden
vectors (tests/test_oracle_golden.py) for SGD/Momentum/Adam/AMSGrad/Adagrad, table semantics,
hashing, dedup, version/step logic.  FTRL is PARITY UNPINNED (TF arithmetic is
not under /root/reference; see ps_oracle.c).
"""
import ctypes
import hashlib
import os
import subprocess
from collections import namedtuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libps_oracle.so")


def build(force=False):
    src = os.path.join(_HERE, "ps_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def _load():
    build()
    lib = ctypes.CDLL(_SO)
    f32p = ctypes.POINTER(ctypes.c_float)
    i64p = ctypes.POINTER(ctypes.c_int64)
    i32p = ctypes.POINTER(ctypes.c_int32)
    vp = ctypes.c_void_p
    ll = ctypes.c_longlong
    f = ctypes.c_float
    sig = {
        "oracle_sgd": (None, [f32p, f32p, f, ll]),
        "oracle_momentum": (None, [f32p, f32p, f32p, f, ctypes.c_int, f, ll]),
        "oracle_adam": (None, [f32p, f32p, f32p, f32p, f, ll, ll, f, f, f, f32p]),
        "oracle_adagrad": (None, [f32p, f32p, f32p, f, ll, f]),
        "oracle_ftrl": (None, [f32p, f32p, f32p, f32p, f, ll, f, f, f]),
        "otable_new": (vp, [ctypes.c_int64, ctypes.c_int, f, ctypes.c_uint64]),
        "otable_free": (None, [vp]),
        "otable_size": (ctypes.c_int64, [vp]),
        "otable_dim": (ctypes.c_int64, [vp]),
        "otable_pull": (None, [vp, i64p, ctypes.c_int64, f32p]),
        "otable_set_rows": (None, [vp, i64p, ctypes.c_int64, f32p]),
        "otable_keys": (ctypes.c_int64, [vp, i64p, ctypes.c_int64]),
        "oracle_uniform_init": (f, [ctypes.c_uint64, ctypes.c_int64, ctypes.c_int64]),
        "oracle_sparse_sgd": (None, [vp, i64p, f32p, ctypes.c_int64, f]),
        "oracle_sparse_momentum": (None, [vp, vp, i64p, f32p, ctypes.c_int64, f, ctypes.c_int, f]),
        "oracle_sparse_adam": (None, [vp, vp, vp, vp, i64p, f32p, ctypes.c_int64, f, ll, f, f, f]),
        "oracle_set_ref_adam": (None, [vp]),
        "oracle_sparse_adagrad": (None, [vp, vp, i64p, f32p, ctypes.c_int64, f, f]),
        "oracle_sparse_ftrl": (None, [vp, vp, vp, i64p, f32p, ctypes.c_int64, f, f, f, f]),
        "oracle_indexed_apply": (None, [ctypes.c_int, f32p, f32p, f32p, f32p, ctypes.c_int64, i64p,
                                        f32p, ctypes.c_int64, f, ll, f, f, f, ctypes.c_int]),
        "oracle_dedup": (ctypes.c_int64, [f32p, i64p, ctypes.c_int64, ctypes.c_int64, f32p, i64p]),
        "oracle_unique": (ctypes.c_int64, [i64p, ctypes.c_int64, i64p, i32p]),
        "oracle_bench_ps": (ctypes.c_double, [ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                              ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(vp),
                                              i64p, f32p, ctypes.c_int64, ctypes.c_int64, f, f, f, f,
                                              ctypes.c_int, ctypes.POINTER(ctypes.c_double)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def _f32(a):
    assert a.dtype == np.float32 and a.flags.c_contiguous
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _i64(a):
    assert a.dtype == np.int64 and a.flags.c_contiguous
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))


def _null_f32():
    return ctypes.POINTER(ctypes.c_float)()


# --------------------------------------------------------------------------
# numpy twin of go/pkg/kernel/capi/kernel_api.cc (cross-check of the C port)
# --------------------------------------------------------------------------
F = np.float32


def np_sgd(g, p, lr):  # kernel_api.cc:6-14
    p[...] = p - F(lr) * g


def np_momentum(g, p, v, mu, nesterov, lr):  # kernel_api.cc:16-38
    v[...] = F(mu) * v + g
    if nesterov:
        p[...] = p - F(lr) * (g + F(mu) * v)
    else:
        p[...] = p - F(lr) * v


def adam_lr_t(lr, step, beta1, beta2):
    """kernel_api.cc:67 -- double math, narrowed to float by `lr *=`."""
    b1 = float(F(beta1))
    b2 = float(F(beta2))
    return F(float(F(lr)) * (np.sqrt(1.0 - b2 ** float(step)) / (1.0 - b1 ** float(step))))


def np_adam(g, p, m, v, lr, step, beta1, beta2, eps, ms=None):  # kernel_api.cc:40-77
    c1 = F(1.0 - float(F(beta1)))
    c2 = F(1.0 - float(F(beta2)))
    m[...] = F(beta1) * m + c1 * g
    v[...] = F(beta2) * v + c2 * (g * g)
    lr_t = adam_lr_t(lr, step, beta1, beta2)
    if ms is not None:
        ms[...] = np.maximum(ms, v)
        p[...] = p - (lr_t * m) / (np.sqrt(ms) + F(eps))
    else:
        p[...] = p - (lr_t * m) / (np.sqrt(v) + F(eps))


def np_adagrad(g, p, m, lr, eps):  # kernel_api.cc:79-96
    m[...] = m + g * g
    p[...] = p - (F(lr) * g) / (np.sqrt(m) + F(eps))


def np_ftrl(g, p, accum, linear, lr, l1, l2, l2s):  # see ps_oracle.c oracle_ftrl (UNPINNED)
    lr = F(lr)
    g_shr = g + (F(2.0) * F(l2s)) * p
    a_new = accum + g * g
    sigma = (np.sqrt(a_new) - np.sqrt(accum)) / lr
    lin = linear + (g_shr - sigma * p)
    quad = np.sqrt(a_new) / lr + F(2.0) * F(l2)
    out = np.where(np.abs(lin) > F(l1), (np.sign(lin) * F(l1) - lin) / quad, F(0)).astype(F)
    linear[...] = lin
    accum[...] = a_new
    p[...] = out


# --------------------------------------------------------------------------
# Hashing: python/common/hash_utils.py:17-62 ; go/pkg/ps/checkpoint.go:31-44
# --------------------------------------------------------------------------
def string_to_id(name, bucket_num):
    """sha256 hex digest parsed in RADIX 32 (sic), mod N -- hash_utils.py:17-19."""
    h = hashlib.sha256(name.encode("utf-8"))
    return int(h.hexdigest(), base=32) % bucket_num


def int_to_id(number, bucket_num):  # hash_utils.py:22-23
    return number % bucket_num


def scatter_embedding_vector(values, indices, bucket_num):  # hash_utils.py:26-62
    ps_ids = {}
    for i, item_id in enumerate(np.asarray(indices).tolist()):
        ps_ids.setdefault(int_to_id(item_id, bucket_num), []).append((i, item_id))
    return {
        ps_id: (values[[v[0] for v in pairs], :], [v[1] for v in pairs])
        for ps_id, pairs in ps_ids.items()
    }


def deduplicate_indexed_slices(values, indices):
    """tensor_utils.py:39-60 via the C restatement (first-occurrence order,
    occurrence-order accumulation)."""
    values = np.ascontiguousarray(values, dtype=np.float32)
    indices = np.ascontiguousarray(indices, dtype=np.int64)
    k = indices.shape[0]
    dim = int(np.prod(values.shape[1:])) if values.ndim > 1 else 1
    ov = np.empty((max(k, 1), dim), dtype=np.float32)
    oi = np.empty(max(k, 1), dtype=np.int64)
    u = lib.oracle_dedup(_f32(values), _i64(indices), k, dim, _f32(ov), _i64(oi))
    return ov[:u].reshape((u,) + values.shape[1:]), oi[:u]


def np_deduplicate_indexed_slices(values, indices):
    """Pure-python twin of tensor_utils.py:53-58 (dict, in-place +=)."""
    res = {}
    for index, i in enumerate(np.asarray(indices).tolist()):
        if i not in res:
            res[i] = np.array(values[index, :], dtype=np.float32)
        else:
            res[i] += values[index, :]
    return np.stack(list(res.values())), np.asarray(list(res.keys()), dtype=np.int64)


def unique_first_occurrence(ids):
    """tf.unique semantics (embedding_delegate.py:85)."""
    ids = np.ascontiguousarray(ids, dtype=np.int64).reshape(-1)
    k = ids.shape[0]
    ou = np.empty(max(k, 1), dtype=np.int64)
    oi = np.empty(max(k, 1), dtype=np.int32)
    u = lib.oracle_unique(_i64(ids), k, _i64(ou), oi.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
    return ou[:u].copy(), oi[:k].copy()


# --------------------------------------------------------------------------
# EmbeddingTable: go/pkg/common/embedding_table.go
# --------------------------------------------------------------------------
class OracleTable:
    def __init__(self, dim, initializer="zero", seed=0, init_constant=0.0):
        self.dim = int(dim)
        self.initializer = initializer
        # embedding_table.go:51 -- only the literal "uniform" randomises (quirk Q6)
        self._h = lib.otable_new(self.dim, 1 if initializer == "uniform" else 0,
                                 float(init_constant), int(seed))

    def __del__(self):
        try:
            lib.otable_free(self._h)
        except Exception:
            pass

    def __len__(self):
        return int(lib.otable_size(self._h))

    def get(self, ids):
        ids = np.ascontiguousarray(ids, dtype=np.int64).reshape(-1)
        out = np.empty((ids.shape[0], self.dim), dtype=np.float32)
        if ids.shape[0]:
            lib.otable_pull(self._h, _i64(ids), ids.shape[0], _f32(out))
        return out

    def set(self, ids, values):
        ids = np.ascontiguousarray(ids, dtype=np.int64).reshape(-1)
        values = np.ascontiguousarray(values, dtype=np.float32).reshape(ids.shape[0], self.dim)
        if ids.shape[0]:
            lib.otable_set_rows(self._h, _i64(ids), ids.shape[0], _f32(values))

    def keys(self):
        n = len(self)
        out = np.empty(max(n, 1), dtype=np.int64)
        n = lib.otable_keys(self._h, _i64(out), n)
        return out[:n]


# --------------------------------------------------------------------------
# Optimizers: go/pkg/ps/optimizer.go
# --------------------------------------------------------------------------
OPT_ARGS = {  # optimizer.go:296-301
    "SGD": ["learning_rate", "momentum", "nesterov"],
    "Adam": ["learning_rate", "beta_1", "beta_2", "epsilon", "amsgrad"],
    "Adagrad": ["learning_rate", "epsilon"],
    # not in the Go PS (python PS only); grammar extended the same way
    "Ftrl": ["learning_rate", "initial_accumulator_value", "l1_regularization_strength",
             "l2_regularization_strength", "l2_shrinkage_regularization_strength", "beta"],
}


def parse_opt_args(opt_type, opt_args):
    """optimizer.go:304-326: 'k=v;k=v;' -> dict; missing / redundant keys are errors."""
    args = {}
    for kv in opt_args.split(";"):
        if kv == "":
            continue
        arr = kv.split("=")
        args[arr[0]] = arr[1]
    if opt_type not in OPT_ARGS:
        raise ValueError("Unknown optimizer type %s" % opt_type)
    for name in OPT_ARGS[opt_type]:
        if name not in args:
            raise ValueError("Args passed to ps should contain %s" % name)
    if len(args) != len(OPT_ARGS[opt_type]):
        raise ValueError("Args passed to ps contain redundant items: %s" % args)
    return args


def parse_bool(s):
    """Go strconv.ParseBool accepted spellings (optimizer.go:349,375)."""
    if s in ("1", "t", "T", "TRUE", "true", "True"):
        return True
    if s in ("0", "f", "F", "FALSE", "false", "False"):
        return False
    raise ValueError("invalid bool %r" % s)


class OracleOptimizer:
    """BaseOptimizer + the four Go optimizers (optimizer.go:27-282) + FTRL.

    kind: sgd | momentum | adam | adagrad | ftrl.  ``step`` is the global
    per-shard call counter of optimizer.go:44 (quirk Q2).
    """

    def __init__(self, opt_type, opt_args, reproduce_q1=False):
        a = parse_opt_args(opt_type, opt_args)
        self.lr = F(float(a["learning_rate"]))
        self.step = 0
        self.reproduce_q1 = reproduce_q1
        self.slots = {}  # slot name -> OracleModel-like dict of dense arrays / tables
        if opt_type == "SGD":
            mu = float(a["momentum"])
            nesterov = parse_bool(a["nesterov"])
            if mu > 0.0:  # optimizer.go:353-356 (quirk Q5)
                self.kind, self.mu, self.nesterov = "momentum", F(mu), nesterov
                self.slot_names = ["v"]
            else:
                self.kind, self.slot_names = "sgd", []
        elif opt_type == "Adam":
            self.kind = "adam"
            self.beta1, self.beta2 = F(float(a["beta_1"])), F(float(a["beta_2"]))
            self.epsilon = F(float(a["epsilon"]))
            self.amsgrad = parse_bool(a["amsgrad"])
            self.slot_names = ["m", "v"] + (["max_square"] if self.amsgrad else [])
        elif opt_type == "Adagrad":
            self.kind = "adagrad"
            self.epsilon = F(float(a["epsilon"]))
            self.slot_names = ["m"]
        elif opt_type == "Ftrl":
            self.kind = "ftrl"
            self.init_accum = F(float(a["initial_accumulator_value"]))
            self.l1 = F(float(a["l1_regularization_strength"]))
            self.l2 = F(float(a["l2_regularization_strength"]))
            self.l2s = F(float(a["l2_shrinkage_regularization_strength"]))
            self.beta = F(float(a["beta"]))
            self.slot_names = ["accumulator", "linear"]
        self.dense_slots = {n: {} for n in self.slot_names}
        self.table_slots = {n: {} for n in self.slot_names}

    # optimizer.go:145-154,222-237,273-282 InitOptimizer
    def init_dense(self, name, shape):
        for n in self.slot_names:
            init = self.init_accum if (self.kind == "ftrl" and n == "accumulator") else F(0)
            self.dense_slots[n][name] = np.full(shape, init, dtype=np.float32)

    def init_table(self, name, dim):
        for n in self.slot_names:
            if name in self.table_slots[n]:
                continue  # model.go:57-63 idempotent
            c = float(self.init_accum) if (self.kind == "ftrl" and n == "accumulator") else 0.0
            self.table_slots[n][name] = OracleTable(dim, "zero", init_constant=c)

    def _ftrl_l2(self, lr):
        # keras Ftrl: l2 + beta / (2 * lr)
        return F(self.l2 + self.beta / (F(2.0) * F(lr))) if float(self.beta) != 0.0 else self.l2

    def dense_kernel(self, g, p, name, lr):
        g = np.ascontiguousarray(g, dtype=np.float32).reshape(-1)
        n = g.shape[0]
        pf = p.reshape(-1)
        s = [self.dense_slots[k][name].reshape(-1) for k in self.slot_names]
        if self.kind == "sgd":
            lib.oracle_sgd(_f32(g), _f32(pf), lr, n)
        elif self.kind == "momentum":
            lib.oracle_momentum(_f32(g), _f32(pf), _f32(s[0]), self.mu, int(self.nesterov), lr, n)
        elif self.kind == "adam":
            if self.amsgrad:
                lib.oracle_adam(_f32(g), _f32(pf), _f32(s[0]), _f32(s[1]), lr, n, self.step,
                                self.beta1, self.beta2, self.epsilon, _f32(s[2]))
                if not self.reproduce_q1:
                    return
                # quirk Q1 (optimizer.go:186-192): falls through and applies again
            lib.oracle_adam(_f32(g), _f32(pf), _f32(s[0]), _f32(s[1]), lr, n, self.step,
                            self.beta1, self.beta2, self.epsilon, _null_f32())
        elif self.kind == "adagrad":
            lib.oracle_adagrad(_f32(g), _f32(pf), _f32(s[0]), lr, n, self.epsilon)
        elif self.kind == "ftrl":
            lib.oracle_ftrl(_f32(g), _f32(pf), _f32(s[0]), _f32(s[1]), lr, n, self.l1,
                            self._ftrl_l2(lr), self.l2s)

    def sparse_kernel(self, ids, g, table, name, lr):
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        g = np.ascontiguousarray(g, dtype=np.float32)
        if g.shape[1] != table.dim:  # kernel.go:36-38
            raise ValueError("grad width is not equal to embedding dim")
        n = ids.shape[0]
        s = [self.table_slots[k][name]._h for k in self.slot_names]
        if self.kind == "sgd":
            lib.oracle_sparse_sgd(table._h, _i64(ids), _f32(g), n, lr)
        elif self.kind == "momentum":
            lib.oracle_sparse_momentum(table._h, s[0], _i64(ids), _f32(g), n, self.mu,
                                       int(self.nesterov), lr)
        elif self.kind == "adam":
            lib.oracle_sparse_adam(table._h, s[0], s[1], s[2] if self.amsgrad else None, _i64(ids),
                                   _f32(g), n, lr, self.step, self.beta1, self.beta2, self.epsilon)
        elif self.kind == "adagrad":
            lib.oracle_sparse_adagrad(table._h, s[0], _i64(ids), _f32(g), n, lr, self.epsilon)
        elif self.kind == "ftrl":
            lib.oracle_sparse_ftrl(table._h, s[0], s[1], _i64(ids), _f32(g), n, lr, self.l1,
                                   self._ftrl_l2(lr), self.l2s)

    def indexed_kernel(self, ids, g, p, name, lr):
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        g = np.ascontiguousarray(g, dtype=np.float32)
        dim = p.shape[1]
        if self.kind != "sgd" and g.shape[1] != dim:  # kernel.go:86-88 (IndexedSGD has no check)
            raise ValueError("grad width is not equal to embedding dim")
        s = [self.dense_slots[k][name] for k in self.slot_names] + [None, None, None]
        sp = [(_f32(x) if x is not None else _null_f32()) for x in s[:3]]
        kind = {"sgd": 0, "momentum": 1, "adam": 3 if getattr(self, "amsgrad", False) else 2,
                "adagrad": 4, "ftrl": 5}[self.kind]
        if self.kind == "momentum":
            h = (self.mu, 0, 0, int(self.nesterov))
        elif self.kind == "adam":
            h = (self.beta1, self.beta2, self.epsilon, 0)
        elif self.kind == "adagrad":
            h = (self.epsilon, 0, 0, 0)
        elif self.kind == "ftrl":
            h = (self.l1, self._ftrl_l2(lr), self.l2s, 0)
        else:
            h = (0, 0, 0, 0)
        lib.oracle_indexed_apply(kind, _f32(p), sp[0], sp[1], sp[2], dim, _i64(ids), _f32(g),
                                 ids.shape[0], lr, self.step, h[0], h[1], h[2], h[3])


Tensor = namedtuple("Tensor", ("name", "values", "indices"))  # tensor_utils.py:25
EmbeddingTableInfo = namedtuple("EmbeddingTableInfo", ("name", "dim", "initializer", "dtype"))


class OracleServer:
    """One PS shard: go/pkg/ps/server.go:54-230 + model.go:25-107 (no gRPC)."""

    def __init__(self, ps_id, opt_type, opt_args, num_ps=1, lr_staleness_modulation=False,
                 reproduce_q1=False, seed=0):
        self.id = ps_id
        self.num_ps = num_ps
        self.opt = OracleOptimizer(opt_type, opt_args, reproduce_q1)
        self.lr_staleness_modulation = lr_staleness_modulation
        self.dense = {}
        self.tables = {}
        self.version = 0
        self.initialized = False
        self.seed = seed

    # model.go:57-63
    def _set_table_info(self, info):
        if info.name in self.tables:
            return
        self.tables[info.name] = OracleTable(info.dim, info.initializer,
                                             seed=self.seed ^ (hash_name(info.name)))

    # model.go:66-88
    def _init_from_model(self, dense, infos, tables, version):
        for info in infos:
            self._set_table_info(info)
        for name, v in dense.items():
            self.dense[name] = np.array(v, dtype=np.float32)
        for name, (ids, vals) in tables.items():
            if name not in self.tables:
                raise ValueError("Embedding table %s is not created" % name)
            self.tables[name].set(ids, vals)
        if version >= 1:
            self.version = version

    def _init_optimizer(self, dense, infos):
        for name, v in dense.items():
            self.opt.init_dense(name, np.shape(v))
        for info in infos:
            self.opt.init_table(info.name, info.dim)

    # server.go:209-221
    def push_model(self, dense=None, infos=(), tables=None, version=0):
        if not self.initialized:
            self._init_from_model(dense or {}, infos, tables or {}, version)
            self._init_optimizer(dense or {}, infos)
            self.initialized = True

    # server.go:224-230
    def push_embedding_table_infos(self, infos):
        self._init_from_model({}, infos, {}, 0)
        self._init_optimizer({}, infos)

    # server.go:144-160 (Go: Version >= req, quirk Q8)
    def pull_dense_parameters(self, version):
        if not self.initialized:
            return False, self.version, {}
        out = {}
        if self.version >= version:
            out = {k: v.copy() for k, v in self.dense.items()}
        return True, self.version, out

    # server.go:163-173
    def pull_embedding_vectors(self, name, ids):
        if ids is None or len(ids) == 0:
            return np.zeros((0,), dtype=np.float32)
        if name not in self.tables:
            raise KeyError("Request embedding Table %s not found in Param" % name)
        return self.tables[name].get(ids)

    # optimizer.go:43-73
    def _apply_gradients(self, dense_grads, sparse_grads, lr):
        self.opt.step += 1
        for name, g in dense_grads.items():
            if name not in self.dense:
                raise KeyError("grad %s not in Parameter" % name)
            self.opt.dense_kernel(g, self.dense[name], name, lr)
        for name, (ids, g) in sparse_grads.items():
            if name in self.dense:
                self.opt.indexed_kernel(ids, g, self.dense[name], name, lr)
            elif name in self.tables:
                self.opt.sparse_kernel(ids, g, self.tables[name], name, lr)
            else:
                raise KeyError("grad %s not in Parameter" % name)

    # server.go:176-206
    def push_gradients(self, dense_grads, sparse_grads, learning_rate, version):
        lr = F(1.0)
        if self.lr_staleness_modulation and self.version > version:
            lr = F(lr / F(self.version - version))
        if learning_rate > 0.0:
            lr = F(lr * F(learning_rate))
        else:
            lr = F(lr * self.opt.lr)
        try:
            self._apply_gradients(dense_grads, sparse_grads, lr)
        except (KeyError, ValueError):
            return False, self.version
        self.version += 1
        return True, self.version


def hash_name(name):
    return int.from_bytes(hashlib.sha256(name.encode()).digest()[:8], "little")


class OraclePSClient:
    """python/worker/ps_client.py:87-301 over in-process OracleServers."""

    def __init__(self, servers):
        self.servers = servers
        self.ps_num = len(servers)
        self.parameter_to_ps = {}
        self.ps_to_parameter = {}

    def pull_embedding_vectors(self, layer_name, embedding_ids):  # ps_client.py:96-130
        ps_ids, ps_ids_index = {}, {}
        for idx, eid in enumerate(np.asarray(embedding_ids).tolist()):
            ps_id = int_to_id(eid, self.ps_num)
            ps_ids.setdefault(ps_id, []).append(eid)
            ps_ids_index.setdefault(ps_id, []).append(idx)
        embeddings, index = [], []
        for ps_id, ids in ps_ids.items():
            embeddings.append(self.servers[ps_id].pull_embedding_vectors(layer_name, ids))
            index.extend(ps_ids_index[ps_id])
        embeddings = np.concatenate(embeddings)
        new = np.empty_like(embeddings)
        new[index] = embeddings
        return new

    def partition_dense_parameters(self, param_names):  # ps_client.py:132-144
        for name in param_names:
            if name not in self.parameter_to_ps:
                ps_id = string_to_id(name, self.ps_num)
                self.parameter_to_ps[name] = ps_id
                self.ps_to_parameter.setdefault(ps_id, []).append(name)

    def push_dense_parameters(self, parameters, ps_id, version):  # ps_client.py:146-159
        dense = {p.name: p.values for p in parameters if self.parameter_to_ps[p.name] == ps_id}
        self.servers[ps_id].push_model(dense=dense, version=version)

    def pull_dense_parameters(self, ps_ids, model_versions):  # ps_client.py:161-188
        dense_params, uninit = {}, []
        for ps_id in ps_ids:
            if ps_id not in self.ps_to_parameter:
                continue
            ok, version, params = self.servers[ps_id].pull_dense_parameters(model_versions[ps_id])
            if not ok:
                uninit.append(ps_id)
            else:
                dense_params.update(params)
                model_versions[ps_id] = version
        return dense_params, uninit

    def push_gradients(self, grads, edl_grads, learning_rate, model_versions):  # ps_client.py:190-287
        dense_req = [dict() for _ in range(self.ps_num)]
        sparse_req = [dict() for _ in range(self.ps_num)]
        ps_grads = {}
        for grad in grads:
            ps_id = self.parameter_to_ps[grad.name]
            d = ps_grads.setdefault(ps_id, {})
            if grad.name not in d:
                d[grad.name] = grad
            elif grad.indices is not None:
                d[grad.name] = Tensor(None, np.concatenate([d[grad.name].values, grad.values]),
                                      np.concatenate([d[grad.name].indices, grad.indices]))
            else:
                d[grad.name] = Tensor(grad.name, d[grad.name].values + grad.values, None)
        for ps_id, pair in ps_grads.items():
            for name, grad in pair.items():
                if grad.indices is not None:
                    v, i = deduplicate_indexed_slices(grad.values, grad.indices)
                    sparse_req[ps_id][name] = (i, v)
                else:
                    dense_req[ps_id][name] = grad.values
        groups = {}
        for grad in edl_grads:
            if grad.name not in groups:
                groups[grad.name] = grad
            else:
                groups[grad.name] = Tensor(None, np.concatenate([groups[grad.name].values, grad.values]),
                                           np.concatenate([groups[grad.name].indices, grad.indices]))
        for name, grad in groups.items():
            v, i = deduplicate_indexed_slices(grad.values, grad.indices)
            for ps_id, (gv, gi) in scatter_embedding_vector(v, i, self.ps_num).items():
                sparse_req[ps_id][name] = (np.asarray(gi, dtype=np.int64), gv)
        accepted, max_version = False, -1
        for ps_id in range(self.ps_num):  # every shard, even empty (quirk Q7)
            ok, ver = self.servers[ps_id].push_gradients(dense_req[ps_id], sparse_req[ps_id],
                                                         learning_rate, model_versions[ps_id])
            accepted = accepted or ok
            max_version = max(max_version, ver)
        return accepted, max_version

    def push_embedding_table_infos(self, infos):  # ps_client.py:289-301
        for s in self.servers:
            s.push_embedding_table_infos(infos)
