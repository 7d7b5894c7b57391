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

path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_python_vectors.json")
with open(path, "w") as f:
    json.dump(out, f)
print("wrote", path, {k: len(v) for k, v in out.items()})
