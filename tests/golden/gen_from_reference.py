"""Generate golden vectors by EXECUTING the Python reference in this container.

Run here (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/gen_from_reference.py
Writes tests/golden/ref_python_vectors.json (committed).

What is executed, unmodified, from /root/reference:
  * elasticdl/python/common/hash_utils.py  (string_to_id, int_to_id,
    scatter_embedding_vector) -- imports only hashlib.
  * elasticdl/python/common/tensor_utils.py (merge_indexed_slices,
    deduplicate_indexed_slices) -- its tensorflow / generated-proto imports are
    not installed here, so stub modules are placed in sys.modules for import
    only (the two functions use numpy alone).  `np.stack(dict_values)` worked on
    the numpy the reference pins; numpy 2.x wants a sequence, so np.stack is
    wrapped to list() its argument -- same values, same order.
  * elasticdl/python/common/tensor_utils.py serializers (ndarray_to_pb, indexed_slices_to_pb, pb_to_ndarray,
    pb_to_indexed_slices) and elasticdl/python/common/dtypes.py -- executed in a SECOND import of both files
    (ref_wire_vectors()) with `tensor_pb2.TensorProto` / `elasticdl_pb2.IndexedSlicesProto` bound to message
    classes built at runtime by google.protobuf from the field numbers of TensorFlow's tensor.proto /
    tensor_shape.proto and elasticdl/proto/elasticdl.proto:12-15 (protoc is not installed), `odps.types` stubbed
    (three attribute reads at import) and `np.bool` aliased to np.bool_ (removed in numpy 2).  The bytes those
    functions serialise are what a reference worker puts on the wire and into checkpoint files.
  * elasticdl/python/common/save_utils.py CheckpointSaver (file naming, complete / latest version directory) on real
    temporary directory trees (ref_checkpoint_dir_vectors()).
  * elasticdl/python/worker/ps_client.py PSClient -- the worker-side boundary itself -- against recording fake stubs
    (ref_ps_client_vectors()): every request each PS receives, every value the client returns.
  * get_optimizer_info of elasticdl/python/common/model_utils.py (ref_optimizer_info_vectors()): the optimizer strings.
"""
import json
import os
import sys
import types

import numpy as np

REF = "/root/reference"
sys.path.insert(0, REF)

# --- stubs so that tensor_utils.py can be imported without TF / protoc output
for name in ["tensorflow", "tensorflow.core", "tensorflow.core.framework",
             "tensorflow.core.framework.tensor_pb2", "tensorflow.core.framework.types_pb2",
             "elasticdl.proto", "elasticdl.proto.elasticdl_pb2"]:
    if name not in sys.modules:
        sys.modules[name] = types.ModuleType(name)
sys.modules["tensorflow.core.framework"].tensor_pb2 = sys.modules["tensorflow.core.framework.tensor_pb2"]
sys.modules["tensorflow.core.framework"].types_pb2 = sys.modules["tensorflow.core.framework.types_pb2"]
tp = sys.modules["tensorflow.core.framework.types_pb2"]
for i, n in enumerate(["DT_INVALID", "DT_FLOAT", "DT_DOUBLE", "DT_INT32", "DT_UINT8", "DT_INT16",
                       "DT_INT8", "DT_STRING", "DT_COMPLEX64", "DT_INT64", "DT_BOOL"]):
    setattr(tp, n, i)
tp.DT_BFLOAT16, tp.DT_HALF, tp.DT_UINT16, tp.DT_UINT32, tp.DT_UINT64 = 14, 19, 17, 22, 23
sys.modules["elasticdl.proto"].elasticdl_pb2 = sys.modules["elasticdl.proto.elasticdl_pb2"]
# dtypes.py additionally imports `odps` and uses np.bool (removed in numpy 2); only
# the (unused here) proto serialisers need it, so it is stubbed as a whole.
_dt = types.ModuleType("elasticdl.python.common.dtypes")
_dt.dtype_numpy_to_tensor = lambda d: None
_dt.dtype_tensor_to_numpy = lambda d: None
sys.modules["elasticdl.python.common.dtypes"] = _dt

from elasticdl.python.common import hash_utils  # noqa: E402

_stack = np.stack
np.stack = lambda arrays, *a, **k: _stack(list(arrays), *a, **k)
_asarray = np.asarray
np.asarray = lambda a, *x, **k: _asarray(list(a) if isinstance(a, type({}.keys())) else a, *x, **k)
from elasticdl.python.common import tensor_utils  # noqa: E402

out = {}

# string_to_id / int_to_id
names = ["dense/kernel:0", "dense/bias:0", "dense_1/kernel:0", "dense_1/bias:0",
         "embedding/embeddings:0", "deepfm/linear:0", "t1", "t2", "conv2d/kernel:0", ""]
out["string_to_id"] = [
    {"name": n, "buckets": b, "id": hash_utils.string_to_id(n, b)}
    for n in names for b in (1, 2, 3, 4, 5, 8)
]
out["int_to_id"] = [
    {"id": i, "buckets": b, "ps": hash_utils.int_to_id(i, b)}
    for i in (0, 1, 7, 8, 1000003, 2 ** 40 + 5) for b in (1, 2, 3, 8)
]

# scatter_embedding_vector
rng = np.random.RandomState(5)
cases = []
for k, dim, nb in [(3, 2, 2), (11, 8, 2), (64, 4, 3), (257, 1, 8)]:
    ids = rng.randint(0, 50, size=k).astype(np.int64)
    vals = rng.randn(k, dim).astype(np.float32)
    res = hash_utils.scatter_embedding_vector(vals, ids, nb)
    cases.append({"ids": ids.tolist(), "values": vals.tolist(), "buckets": nb,
                  "result": {str(p): {"values": v.tolist(), "ids": list(map(int, i))}
                             for p, (v, i) in res.items()}})
out["scatter_embedding_vector"] = cases

# deduplicate_indexed_slices / merge_indexed_slices
cases = []
for k, dim, hi in [(3, 2, 4), (7, 8, 4), (64, 4, 10), (500, 8, 40), (1000, 1, 7)]:
    ids = rng.randint(0, hi, size=k).astype(np.int64)
    vals = rng.randn(k, dim).astype(np.float32)
    v, i = tensor_utils.deduplicate_indexed_slices(vals.copy(), ids)
    cases.append({"ids": ids.tolist(), "values": vals.tolist(),
                  "out_values": np.asarray(v, dtype=np.float32).tolist(),
                  "out_ids": [int(x) for x in i]})
out["deduplicate_indexed_slices"] = cases
a = tensor_utils.Tensor(None, rng.randn(3, 2).astype(np.float32), np.array([1, 3, 3], dtype=np.int64))
b = tensor_utils.Tensor(None, rng.randn(2, 2).astype(np.float32), np.array([5, 1], dtype=np.int64))
m = tensor_utils.merge_indexed_slices(a, b)
out["merge_indexed_slices"] = {"a": [a.values.tolist(), a.indices.tolist()],
                               "b": [b.values.tolist(), b.indices.tolist()],
                               "values": m.values.tolist(), "indices": m.indices.tolist()}



def ref_wire_vectors():
    """Bytes produced by the reference's own proto serialisers (see the module docstring)."""
    import importlib

    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

    fd = descriptor_pb2.FileDescriptorProto()
    fd.name, fd.package, fd.syntax = "edl_golden_wire.proto", "edlgolden", "proto3"
    F = descriptor_pb2.FieldDescriptorProto

    def msg(name, fields):
        m = fd.message_type.add()
        m.name = name
        for fname, num, typ, label, tname in fields:
            f = m.field.add()
            f.name, f.number, f.type, f.label = fname, num, typ, label
            if tname:
                f.type_name = tname

    O, R = F.LABEL_OPTIONAL, F.LABEL_REPEATED
    msg("Dim", [("size", 1, F.TYPE_INT64, O, None), ("name", 2, F.TYPE_STRING, O, None)])     # tensor_shape.proto
    msg("TensorShapeProto", [("dim", 2, F.TYPE_MESSAGE, R, ".edlgolden.Dim"), ("unknown_rank", 3, F.TYPE_BOOL, O, None)])
    msg("TensorProto", [("dtype", 1, F.TYPE_INT32, O, None),                                 # tensor.proto
                        ("tensor_shape", 2, F.TYPE_MESSAGE, O, ".edlgolden.TensorShapeProto"),
                        ("version_number", 3, F.TYPE_INT32, O, None), ("tensor_content", 4, F.TYPE_BYTES, O, None)])
    msg("IndexedSlicesProto", [("concat_tensors", 1, F.TYPE_MESSAGE, O, ".edlgolden.TensorProto"),  # elasticdl.proto:12-15
                               ("ids", 2, F.TYPE_INT64, R, None)])
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    cls = lambda n: message_factory.GetMessageClass(pool.FindMessageTypeByName("edlgolden." + n))  # noqa: E731
    sys.modules["tensorflow.core.framework.tensor_pb2"].TensorProto = cls("TensorProto")
    sys.modules["elasticdl.proto.elasticdl_pb2"].IndexedSlicesProto = cls("IndexedSlicesProto")
    odps = types.ModuleType("odps")
    odps.types = types.SimpleNamespace(bigint="bigint", double="double", string="string")
    sys.modules["odps"] = odps
    had_bool = hasattr(np, "bool")
    if not had_bool:
        np.bool = np.bool_
    try:
        del sys.modules["elasticdl.python.common.dtypes"]
        import elasticdl.python.common as common_pkg

        if hasattr(common_pkg, "dtypes"):
            delattr(common_pkg, "dtypes")
        real_dtypes = importlib.import_module("elasticdl.python.common.dtypes")  # the real file, unmodified
        tu = importlib.reload(tensor_utils)                                      # binds the real dtype functions
        assert tu.dtype_numpy_to_tensor is real_dtypes.dtype_numpy_to_tensor
        rng = np.random.RandomState(11)
        tensors = [np.float32(2.5).reshape(()), rng.randn(5).astype(np.float32), rng.randn(3, 4).astype(np.float32),
                   np.zeros((0, 8), np.float32), np.arange(6, dtype=np.int64).reshape(2, 3),
                   rng.randn(2, 1, 3).astype(np.float64)]
        out_t = []
        for a in tensors:
            pb = tu.ndarray_to_pb(a)
            back = tu.pb_to_ndarray(pb)
            assert back.shape == a.shape and np.array_equal(back, a)
            out_t.append({"dtype": a.dtype.name, "shape": list(a.shape), "data": a.reshape(-1).tolist(),
                          "hex": pb.SerializeToString().hex()})
        out_s = []
        for ids, dim in [([1, 3, 3, 5], 2), ([0], 8), ([7, 2 ** 40 + 5, 123456789], 4), (list(range(130, 150)), 1)]:
            vals = rng.randn(len(ids), dim).astype(np.float32)
            for as_array in (True, False):
                idx = np.asarray(ids, dtype=np.int64) if as_array else list(ids)
                pb = tu.indexed_slices_to_pb(tu.Tensor(None, vals, idx))
                back = tu.pb_to_indexed_slices(pb)
                assert np.array_equal(back.values, vals) and back.indices.tolist() == list(ids)
                out_s.append({"ids": list(ids), "values": vals.tolist(), "ids_as_array": as_array,
                              "hex": pb.SerializeToString().hex()})
        return {"tensor_proto": out_t, "indexed_slices_proto": out_s}
    finally:
        if not had_bool:
            del np.bool


out["wire"] = ref_wire_vectors()


CKPT_SCENARIOS = [  # {version: [shard files]} under one checkpoint directory
    {},
    {"3": ["variables-0-of-1.ckpt"]},
    {"3": ["variables-0-of-2.ckpt", "variables-1-of-2.ckpt"], "7": ["variables-0-of-2.ckpt"]},            # 7 incomplete
    {"10": ["variables-%d-of-3.ckpt" % i for i in range(3)], "9": ["variables-%d-of-3.ckpt" % i for i in range(3)],
     "100": ["variables-0-of-3.ckpt", "variables-2-of-3.ckpt"]},                                        # numeric, not lexical, order
    {"5": []},                                                                                           # empty version dir
    {"2": ["variables-0-of-1.ckpt"], "4": ["variables-0-of-2.ckpt", "variables-1-of-2.ckpt"]},
]


def ref_checkpoint_dir_vectors():
    """CheckpointSaver's directory logic (save_utils.py:124-141,192-227) executed on real directory trees: file naming,
    which version directories are complete, which is the latest complete one.  save_utils.py imports TensorFlow and the
    Python PS (for restore_params_from_checkpoint, not used here): stub modules stand in at import."""
    import tempfile

    for name in ["elasticdl.python.ps", "elasticdl.python.ps.embedding_table", "elasticdl.python.ps.parameters"]:
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["elasticdl.python.ps.embedding_table"].create_embedding_table = None
    sys.modules["elasticdl.python.ps.parameters"].Parameters = None
    from elasticdl.python.common import save_utils

    S = save_utils.CheckpointSaver
    res = []
    for sc in CKPT_SCENARIOS:
        with tempfile.TemporaryDirectory() as root:
            for v, files in sc.items():
                d = os.path.join(root, "version-" + v)
                os.makedirs(d)
                for fn in files:
                    open(os.path.join(d, fn), "wb").close()
            latest = S.get_valid_lastest_version_dir(root)
            res.append({"tree": sc,
                        "valid": {v: bool(S.check_checkpoint_valid(os.path.join(root, "version-" + v))) for v in sc},
                        "latest": None if latest is None else os.path.basename(latest)})
    with tempfile.TemporaryDirectory() as root:
        saver = S(root, 1, 0, False)
        names = [os.path.relpath(saver._get_checkpoint_file(v, False, i, n), root) for v, i, n in [(7, 0, 1), (120, 2, 3)]]
        missing = S.get_valid_lastest_version_dir(os.path.join(root, "nope"))
        missing_valid = bool(S.check_checkpoint_valid(os.path.join(root, "nope")))
    return {"scenarios": res, "file_names": names, "missing_dir_latest": missing, "missing_dir_valid": missing_valid}


out["checkpoint_dirs"] = ref_checkpoint_dir_vectors()



def ref_ps_client_vectors():
    """The reference's worker-side client, elasticdl/python/worker/ps_client.py:87-301, EXECUTED unmodified against
    recording fake stubs: which request reaches which PS (ids grouping and order, dedup / merge of gradients, versions,
    learning rate), what the client returns.  Message classes are built at runtime from elasticdl.proto:12-76;
    `elasticdl_pb2_grpc.PserverStub(channel)` is stubbed to return the "channel" itself (a fake stub object); the
    serialisers are the reference's (ref_wire_vectors() must have run: tensor_utils is already bound to real dtypes)."""
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

    fd = descriptor_pb2.FileDescriptorProto()
    fd.name, fd.package, fd.syntax = "edl_golden_client.proto", "edlclient", "proto3"
    F = descriptor_pb2.FieldDescriptorProto
    O, R = F.LABEL_OPTIONAL, F.LABEL_REPEATED

    def msg(name, fields, nested=()):
        m = fd.message_type.add()
        m.name = name
        for n in nested:
            m.nested_type.add().CopyFrom(n)
        for fname, num, typ, label, tname in fields:
            f = m.field.add()
            f.name, f.number, f.type, f.label = fname, num, typ, label
            if tname:
                f.type_name = tname

    def entry(name, vtype):
        e = descriptor_pb2.DescriptorProto()
        e.name = name
        e.options.map_entry = True
        k = e.field.add()
        k.name, k.number, k.type, k.label = "key", 1, F.TYPE_STRING, O
        v = e.field.add()
        v.name, v.number, v.type, v.label, v.type_name = "value", 2, F.TYPE_MESSAGE, O, vtype
        return e

    P = ".edlclient."
    msg("Dim", [("size", 1, F.TYPE_INT64, O, None)])
    msg("TensorShapeProto", [("dim", 2, F.TYPE_MESSAGE, R, P + "Dim")])
    msg("TensorProto", [("dtype", 1, F.TYPE_INT32, O, None), ("tensor_shape", 2, F.TYPE_MESSAGE, O, P + "TensorShapeProto"),
                        ("tensor_content", 4, F.TYPE_BYTES, O, None)])
    msg("IndexedSlicesProto", [("concat_tensors", 1, F.TYPE_MESSAGE, O, P + "TensorProto"), ("ids", 2, F.TYPE_INT64, R, None)])
    msg("EmbeddingTableInfo", [("name", 1, F.TYPE_STRING, O, None), ("dim", 2, F.TYPE_INT64, O, None),
                               ("initializer", 3, F.TYPE_STRING, O, None), ("dtype", 4, F.TYPE_INT32, O, None)])
    msg("Model", [("version", 1, F.TYPE_INT32, O, None),
                  ("embedding_table_infos", 2, F.TYPE_MESSAGE, R, P + "EmbeddingTableInfo"),
                  ("dense_parameters", 3, F.TYPE_MESSAGE, R, P + "Model.DenseParametersEntry"),
                  ("embedding_tables", 4, F.TYPE_MESSAGE, R, P + "Model.EmbeddingTablesEntry")],
        nested=[entry("DenseParametersEntry", P + "TensorProto"), entry("EmbeddingTablesEntry", P + "IndexedSlicesProto")])
    msg("PullEmbeddingVectorRequest", [("name", 1, F.TYPE_STRING, O, None), ("ids", 2, F.TYPE_INT64, R, None)])
    msg("PullDenseParametersRequest", [("version", 1, F.TYPE_INT32, O, None)])
    msg("PullDenseParametersResponse",
        [("initialized", 1, F.TYPE_BOOL, O, None), ("version", 2, F.TYPE_INT32, O, None),
         ("dense_parameters", 3, F.TYPE_MESSAGE, R, P + "PullDenseParametersResponse.DenseParametersEntry")],
        nested=[entry("DenseParametersEntry", P + "TensorProto")])
    msg("PushGradientsRequest", [("gradients", 1, F.TYPE_MESSAGE, O, P + "Model"), ("learning_rate", 2, F.TYPE_FLOAT, O, None)])
    msg("PushGradientsResponse", [("accepted", 1, F.TYPE_BOOL, O, None), ("version", 2, F.TYPE_INT32, O, None)])
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    pb2 = sys.modules["elasticdl.proto.elasticdl_pb2"]
    for n in ("TensorProto", "IndexedSlicesProto", "EmbeddingTableInfo", "Model", "PullEmbeddingVectorRequest",
              "PullDenseParametersRequest", "PullDenseParametersResponse", "PushGradientsRequest", "PushGradientsResponse"):
        setattr(pb2, n, message_factory.GetMessageClass(pool.FindMessageTypeByName("edlclient." + n)))
    sys.modules["tensorflow.core.framework.tensor_pb2"].TensorProto = pb2.TensorProto
    grpc_mod = types.ModuleType("elasticdl.proto.elasticdl_pb2_grpc")
    grpc_mod.PserverStub = lambda channel: channel
    sys.modules["elasticdl.proto.elasticdl_pb2_grpc"] = grpc_mod
    sys.modules["elasticdl.proto"].elasticdl_pb2_grpc = grpc_mod
    import importlib

    tu = importlib.reload(tensor_utils)  # rebinds tensor_pb2.TensorProto to the class Model's maps use
    from elasticdl.python.worker import ps_client as ref_client

    def tensor_dict(pb):
        a = tu.pb_to_ndarray(pb)
        return {"shape": list(a.shape), "data": np.asarray(a, dtype=np.float32).reshape(-1).tolist()}

    def model_dict(m):
        return {"version": m.version,
                "infos": [[i.name, i.dim, i.initializer, i.dtype] for i in m.embedding_table_infos],
                "dense": {k: tensor_dict(v) for k, v in m.dense_parameters.items()},
                "tables": {k: {"ids": [int(i) for i in v.ids], **tensor_dict(v.concat_tensors)}
                           for k, v in m.embedding_tables.items()}}

    class Future(object):
        def __init__(self, value):
            self._v = value

        def result(self):
            return self._v

    class Method(object):
        def __init__(self, fn):
            self._fn = fn

        def __call__(self, req):
            return self._fn(req)

        def future(self, req):
            return Future(self._fn(req))

    class FakeStub(object):
        """One PS: records every request, answers from a script."""

        def __init__(self, ps_id, log, initialized=True, version=5, accept=True, dense=None):
            self.ps_id, self.log = ps_id, log
            self.initialized, self.version, self.accept, self.dense = initialized, version, accept, dense or {}
            self.push_model = Method(lambda m: self.log.append(["push_model", ps_id, model_dict(m)]))
            self.push_embedding_table_infos = Method(
                lambda m: self.log.append(["push_embedding_table_infos", ps_id, model_dict(m)]))
            self.pull_dense_parameters = Method(self._pull_dense)
            self.pull_embedding_vectors = Method(self._pull_rows)
            self.push_gradients = Method(self._push_gradients)

        def _pull_dense(self, req):
            self.log.append(["pull_dense_parameters", self.ps_id, {"version": req.version}])
            res = pb2.PullDenseParametersResponse(initialized=self.initialized, version=self.version)
            if self.initialized:
                for k, a in self.dense.items():
                    tu.serialize_ndarray(a, res.dense_parameters[k])
            return res

        def _pull_rows(self, req):
            ids = [int(i) for i in req.ids]
            self.log.append(["pull_embedding_vectors", self.ps_id, {"name": req.name, "ids": ids}])
            rows = np.asarray([[i + 0.25 * c for c in range(4)] for i in ids], dtype=np.float32)
            return tu.ndarray_to_pb(rows)

        def _push_gradients(self, req):
            self.log.append(["push_gradients", self.ps_id,
                             {"learning_rate": req.learning_rate, "gradients": model_dict(req.gradients)}])
            self.version += 1
            return pb2.PushGradientsResponse(accepted=self.accept, version=self.version)

    T = tu.Tensor
    rng = np.random.RandomState(21)
    f32 = lambda *s: np.asarray(rng.randn(*s), dtype=np.float32)  # noqa: E731
    cases = []
    for ps_num in (1, 2, 3):
        log = []
        stubs = [FakeStub(p, log, initialized=(p != 1), version=5 + p, accept=(p != 0 or ps_num == 1),
                          dense={"from_ps_%d/kernel:0" % p: f32(2, 3)}) for p in range(ps_num)]
        client = ref_client.PSClient(stubs)
        names = ["dense/kernel:0", "dense/bias:0", "dense_1/kernel:0", "emb_keras/embeddings:0", "scalar:0"]
        client.partition_dense_parameters(names)
        infos = [tu.EmbeddingTableInfo("edl_emb", 4, "uniform", 1), tu.EmbeddingTableInfo("edl_emb2", 2, "zeros", 1)]
        client.push_embedding_table_infos(infos)
        params = [T(n, f32(3, 2) if "kernel" in n else f32(3), None) for n in names[:3]]
        for p in range(ps_num):
            client.push_dense_parameters(params, p, 3)
        versions = [1] * ps_num
        dense, uninit = client.pull_dense_parameters(list(range(ps_num)), versions)
        pull_ids = [3, 5, 1, 6, 10, 2, 1, 2, 4, 7, 9]
        rows = client.pull_embedding_vectors("edl_emb", pull_ids)
        grads = [T("dense/kernel:0", f32(3, 2), None), T("dense/bias:0", f32(3), None), T("scalar:0", f32(), None),
                 T("emb_keras/embeddings:0", f32(5, 2), np.array([4, 1, 4, 0, 1])),           # IndexedSlices of a dense param
                 T("emb_keras/embeddings:0", f32(2, 2), np.array([7, 4]))]                    # same name: merged, then dedup
        edl = [T("edl_emb", f32(6, 4), np.array([3, 1, 3, 8, 1, 6])), T("edl_emb2", f32(3, 2), np.array([5, 5, 2])),
               T("edl_emb", f32(2, 4), np.array([1, 9]))]
        g_in = [[g.name, g.values.tolist(), None if g.indices is None else g.indices.tolist()] for g in grads]
        e_in = [[g.name, g.values.tolist(), g.indices.tolist()] for g in edl]
        push_versions = [7 + p for p in range(ps_num)]
        accepted, max_version = client.push_gradients(grads, edl, 0.125, push_versions)
        # two DENSE gradients with one name: ps_client.py:217 does `namedtuple.values += ...` -- what happens is recorded
        try:
            n_before = len(log)
            client.push_gradients([T("dense/bias:0", f32(3), None), T("dense/bias:0", f32(3), None)], [], 0.5,
                                  list(push_versions))
            dup_dense = "ok"
        except Exception as err:  # noqa: BLE001
            dup_dense = type(err).__name__
        del log[n_before:]
        cases.append({"ps_num": ps_num, "param_names": names, "parameter_to_ps": client.parameter_to_ps,
                      "ps_to_parameter": {str(k): v for k, v in client.ps_to_parameter.items()},
                      "infos": [list(i) for i in infos],
                      "params": [[p.name, p.values.tolist()] for p in params],
                      "fake": [{"initialized": s.initialized, "version0": 5 + s.ps_id, "accept": s.accept,
                                "dense": {k: v.tolist() for k, v in s.dense.items()}} for s in stubs],
                      "pull_dense": {"versions_in": [1] * ps_num, "versions_out": versions, "uninit": uninit,
                                     "dense": {k: np.asarray(v).tolist() for k, v in dense.items()}},
                      "pull_ids": pull_ids, "pull_rows": rows.tolist(),
                      "grads": g_in, "edl_grads": e_in, "learning_rate": 0.125, "push_versions": push_versions,
                      "push_result": [bool(accepted), int(max_version)], "duplicate_dense_name": dup_dense, "log": log})
    return cases


out["ps_client"] = ref_ps_client_vectors()



def ref_optimizer_info_vectors():
    """get_optimizer_info (elasticdl/python/common/model_utils.py:227-254): the (opt_type, opt_args) strings a job hands
    the PS (`-opt_type=... -opt_args=...`).  model_utils.py imports TensorFlow, odps and the worker package at module
    level, so only THIS function is executed: its source segment is cut out of the file with `ast` and exec'd, unmodified,
    in a namespace whose `tf.keras.optimizers.{SGD,Adam,Adagrad}` are stand-in classes with the Keras `get_config()`
    contract (python floats / bools, or a callable for a schedule)."""
    import ast

    path = os.path.join(REF, "elasticdl/python/common/model_utils.py")
    src = open(path).read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "get_optimizer_info")
    code = ast.get_source_segment(src, fn)

    class _Opt(object):
        def __init__(self, **cfg):
            self._cfg = cfg

        def get_config(self):
            return dict(self._cfg)

    SGD, Adam, Adagrad = (type(n, (_Opt,), {}) for n in ("SGD", "Adam", "Adagrad"))
    ns = {"tf": types.SimpleNamespace(keras=types.SimpleNamespace(
        optimizers=types.SimpleNamespace(SGD=SGD, Adam=Adam, Adagrad=Adagrad)))}
    exec(compile(code, path, "exec"), ns)
    f = ns["get_optimizer_info"]
    opts = [SGD(learning_rate=0.01, momentum=0.0, nesterov=False, decay=0.0, name="SGD"),
            SGD(learning_rate=0.1, momentum=0.9, nesterov=True, decay=0.0, name="SGD"),
            SGD(learning_rate=(lambda: 0.05), momentum=0.5, nesterov=False),
            Adam(learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-07, amsgrad=False, decay=0.0, name="Adam"),
            Adam(learning_rate=3e-4, beta_1=0.8, beta_2=0.99, epsilon=1e-08, amsgrad=True),
            Adagrad(learning_rate=0.001, initial_accumulator_value=0.1, epsilon=1e-07),
            Adagrad(learning_rate=0.5, epsilon=1e-10)]
    return [{"opt_type": t, "opt_args": a} for t, a in (f(o) for o in opts)]


out["optimizer_info"] = ref_optimizer_info_vectors()

path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_python_vectors.json")
with open(path, "w") as f:
    json.dump(out, f)
print("wrote", path, {k: len(v) for k, v in out.items()})
