"""Value types of the PS boundary (elasticdl/python/common/tensor_utils.py:25-28)."""
from collections import namedtuple

Tensor = namedtuple("Tensor", ("name", "values", "indices"))



class UniqueTensor(Tensor):
    """A Tensor whose indices are already unique (e.g. a BET gradient keyed by the
    layer's unique ids): PSClient skips the dedup pass for it."""
    __slots__ = ()


# Unique ids whose COUNT stays on the device: `ids` is a padded int64 device array, `n_dev` a
# one-element int32 device tensor holding how many leading entries are live.  Lets the
# pull / push of a batch run without reading the count back to the host.
DeviceIds = namedtuple("DeviceIds", ("ids", "n_dev"))


# The reference's EmbeddingTableInfo is (name, dim, initializer, dtype).  The HBM
# tables are direct-indexed, so the number of ids (the layer's input_dim) travels
# with it as an optional fifth field.
EmbeddingTableInfo = namedtuple(
    "EmbeddingTableInfo", ("name", "dim", "initializer", "dtype", "capacity"),
    defaults=(None,),
)

DT_FLOAT = 1  # tensorflow types_pb2.DT_FLOAT, the only dtype the PS stores (ps_trainer.py:187)
