"""The hand-written codec of the gRPC `Pserver` facade (elasticdl_b200/ps/grpc_server.py) against google.protobuf,
message by message, with descriptors built at runtime from the field numbers of
/root/reference/elasticdl/proto/elasticdl.proto:48-76 (PullDenseParametersRequest / Response,
PullEmbeddingVectorsRequest, PushGradientsRequest / Response) and TensorFlow's tensor.proto: what this repo encodes,
the protobuf library parses to the same values; what the library serialises, this repo decodes -- both directions,
on fixed cases and on hypothesis-generated ones.  No GPU, no grpc channel: bytes only."""
import numpy as np
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

from elasticdl_b200.ps import checkpoint as ck
from elasticdl_b200.ps import grpc_server as gs


def _classes():
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

    fd = descriptor_pb2.FileDescriptorProto()
    fd.name = "edl_wire_test.proto"
    fd.package = "edlwire"
    fd.syntax = "proto3"
    F = descriptor_pb2.FieldDescriptorProto

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
        k.name, k.number, k.type, k.label = "key", 1, F.TYPE_STRING, F.LABEL_OPTIONAL
        v = e.field.add()
        v.name, v.number, v.type, v.label, v.type_name = "value", 2, F.TYPE_MESSAGE, F.LABEL_OPTIONAL, vtype
        return e

    O, R = F.LABEL_OPTIONAL, F.LABEL_REPEATED
    msg("Dim", [("size", 1, F.TYPE_INT64, O, None)])
    msg("TensorShapeProto", [("dim", 2, F.TYPE_MESSAGE, R, ".edlwire.Dim")])
    msg("TensorProto", [("dtype", 1, F.TYPE_INT32, O, None),
                        ("tensor_shape", 2, F.TYPE_MESSAGE, O, ".edlwire.TensorShapeProto"),
                        ("tensor_content", 4, F.TYPE_BYTES, O, None)])
    msg("Model", [("version", 1, F.TYPE_INT32, O, None)])  # elasticdl.proto:24-29 (only the field this test sets)
    msg("PullDenseParametersRequest", [("version", 1, F.TYPE_INT32, O, None)])  # :53-55
    msg("PullDenseParametersResponse",  # :57-61
        [("initialized", 1, F.TYPE_BOOL, O, None), ("version", 2, F.TYPE_INT32, O, None),
         ("dense_parameters", 3, F.TYPE_MESSAGE, R, ".edlwire.PullDenseParametersResponse.DenseParametersEntry")],
        nested=[entry("DenseParametersEntry", ".edlwire.TensorProto")])
    msg("PullEmbeddingVectorsRequest", [("name", 1, F.TYPE_STRING, O, None), ("ids", 2, F.TYPE_INT64, R, None)])  # :63-66
    msg("PushGradientsRequest", [("gradients", 1, F.TYPE_MESSAGE, O, ".edlwire.Model"),  # :68-71
                                 ("learning_rate", 2, F.TYPE_FLOAT, O, None)])
    msg("PushGradientsResponse", [("accepted", 1, F.TYPE_BOOL, O, None), ("version", 2, F.TYPE_INT32, O, None)])  # :73-76
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return {n: message_factory.GetMessageClass(pool.FindMessageTypeByName("edlwire." + n))
            for n in ("TensorProto", "Model", "PullDenseParametersRequest", "PullDenseParametersResponse",
                      "PullEmbeddingVectorsRequest", "PushGradientsRequest", "PushGradientsResponse")}


PB = _classes()
int32s = st.integers(min_value=-2 ** 31, max_value=2 ** 31 - 1)
int64s = st.integers(min_value=-2 ** 63, max_value=2 ** 63 - 1)


def _tensor_pb(arr):
    t = PB["TensorProto"]()
    t.dtype = 1  # DT_FLOAT
    for d in arr.shape:
        t.tensor_shape.dim.add().size = d
    t.tensor_content = np.ascontiguousarray(arr, dtype="<f4").tobytes()
    return t


@settings(max_examples=60, deadline=None)
@given(version=int32s)
def test_pull_dense_request_both_ways(version):
    m = PB["PullDenseParametersRequest"]()
    m.ParseFromString(gs.encode_pull_dense_request(version))
    assert m.version == version
    m = PB["PullDenseParametersRequest"](version=version)
    assert gs.decode_pull_dense_request(m.SerializeToString()) == version


@settings(max_examples=40, deadline=None)
@given(initialized=st.booleans(), version=int32s,
       shapes=st.lists(st.lists(st.integers(1, 5), min_size=0, max_size=3), min_size=0, max_size=4), seed=st.integers(0, 99))
def test_pull_dense_response_both_ways(initialized, version, shapes, seed):
    rng = np.random.default_rng(seed)
    dense = {"layer_%d/kernel:0" % i: rng.standard_normal(tuple(s)).astype(np.float32) for i, s in enumerate(shapes)}
    m = PB["PullDenseParametersResponse"]()
    m.ParseFromString(gs.encode_pull_dense_response(initialized, version, dense))
    assert m.initialized == initialized and m.version == version and set(m.dense_parameters) == set(dense)
    for k, a in dense.items():
        t = m.dense_parameters[k]
        assert t.dtype == 1 and [d.size for d in t.tensor_shape.dim] == list(a.shape)
        assert t.tensor_content == a.astype("<f4").tobytes()
    m = PB["PullDenseParametersResponse"](initialized=initialized, version=version)
    for k, a in dense.items():
        m.dense_parameters[k].CopyFrom(_tensor_pb(a))
    got_init, got_version, got = gs.decode_pull_dense_response(m.SerializeToString())
    assert got_init == initialized and got_version == version and set(got) == set(dense)
    for k, a in dense.items():
        assert got[k].shape == a.shape and np.array_equal(got[k], a)


@settings(max_examples=60, deadline=None)
@given(name=st.text(max_size=20), ids=st.lists(int64s, max_size=40))
def test_pull_embedding_request_both_ways(name, ids):
    m = PB["PullEmbeddingVectorsRequest"]()
    m.ParseFromString(gs.encode_pull_embedding_request(name, ids))
    assert m.name == name and list(m.ids) == ids
    m = PB["PullEmbeddingVectorsRequest"](name=name, ids=ids)  # proto3 packs repeated int64
    got_name, got_ids = gs.decode_pull_embedding_request(m.SerializeToString())
    assert got_name == name and got_ids.dtype == np.int64 and got_ids.tolist() == ids


def test_pull_embedding_request_accepts_unpacked_ids():
    """A proto2-style sender may write `repeated int64 ids` one varint field at a time."""
    buf = ck._ld(1, b"emb") + b"".join(ck._tag(2, 0) + ck._varint(i & (2 ** 64 - 1)) for i in (3, -1, 7))
    name, ids = gs.decode_pull_embedding_request(buf)
    assert name == "emb" and ids.tolist() == [3, -1, 7]


@settings(max_examples=60, deadline=None)
@given(version=int32s, lr=st.floats(width=32, allow_nan=False, allow_infinity=False))
def test_push_gradients_request_both_ways(version, lr):
    model_bytes = PB["Model"](version=version).SerializeToString()
    m = PB["PushGradientsRequest"]()
    m.ParseFromString(gs.encode_push_gradients_request(model_bytes, lr))
    assert m.gradients.version == version and np.float32(m.learning_rate) == np.float32(lr)
    m = PB["PushGradientsRequest"](learning_rate=lr)
    m.gradients.version = version
    got_model, got_lr = gs.decode_push_gradients_request(m.SerializeToString())
    assert np.float32(got_lr) == np.float32(lr)
    back = PB["Model"]()
    back.ParseFromString(got_model)
    assert back.version == version


@settings(max_examples=60, deadline=None)
@given(accepted=st.booleans(), version=int32s)
def test_push_gradients_response_both_ways(accepted, version):
    m = PB["PushGradientsResponse"]()
    m.ParseFromString(gs.encode_push_gradients_response(accepted, version))
    assert m.accepted == accepted and m.version == version
    m = PB["PushGradientsResponse"](accepted=accepted, version=version)
    assert gs.decode_push_gradients_response(m.SerializeToString()) == (accepted, version)


def test_tensor_codec_round_trip_and_empty():
    for shape in [(), (0,), (3,), (2, 0, 4), (5, 7)]:
        a = np.arange(int(np.prod(shape)), dtype=np.float32).reshape(shape)
        t = PB["TensorProto"]()
        t.ParseFromString(ck.encode_tensor(a))
        assert [d.size for d in t.tensor_shape.dim] == list(shape) and t.tensor_content == a.tobytes()
        b = ck.decode_tensor(_tensor_pb(a).SerializeToString())
        assert b.shape == a.shape and np.array_equal(a, b)
