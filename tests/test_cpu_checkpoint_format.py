"""The hand-written proto3 encoder/decoder of the checkpoint files against google.protobuf with
descriptors built at runtime from elasticdl/proto/elasticdl.proto:12-29 and TensorFlow's
tensor.proto / tensor_shape.proto field numbers (tensor_test.go:52-84 pins the byte layout:
little-endian fp32 tensor_content, dims [1,5])."""
import numpy as np
import pytest

from elasticdl_b200.ps import checkpoint as ck


def _pb_classes():
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

    fd = descriptor_pb2.FileDescriptorProto()
    fd.name = "edl_ckpt_test.proto"
    fd.package = "edltest"
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
        return m

    msg("Dim", [("size", 1, F.TYPE_INT64, F.LABEL_OPTIONAL, None), ("name", 2, F.TYPE_STRING, F.LABEL_OPTIONAL, None)])
    msg("TensorShapeProto", [("dim", 2, F.TYPE_MESSAGE, F.LABEL_REPEATED, ".edltest.Dim")])
    msg("TensorProto", [("dtype", 1, F.TYPE_INT32, F.LABEL_OPTIONAL, None),
                        ("tensor_shape", 2, F.TYPE_MESSAGE, F.LABEL_OPTIONAL, ".edltest.TensorShapeProto"),
                        ("tensor_content", 4, F.TYPE_BYTES, F.LABEL_OPTIONAL, None)])
    msg("IndexedSlicesProto", [("concat_tensors", 1, F.TYPE_MESSAGE, F.LABEL_OPTIONAL, ".edltest.TensorProto"),
                               ("ids", 2, F.TYPE_INT64, F.LABEL_REPEATED, None)])
    msg("EmbeddingTableInfo", [("name", 1, F.TYPE_STRING, F.LABEL_OPTIONAL, None),
                               ("dim", 2, F.TYPE_INT64, F.LABEL_OPTIONAL, None),
                               ("initializer", 3, F.TYPE_STRING, F.LABEL_OPTIONAL, None),
                               ("dtype", 4, F.TYPE_INT32, F.LABEL_OPTIONAL, None)])

    def entry(name, vtype):
        e = descriptor_pb2.DescriptorProto()
        e.name = name
        e.options.map_entry = True
        k = e.field.add()
        k.name, k.number, k.type, k.label = "key", 1, F.TYPE_STRING, F.LABEL_OPTIONAL
        v = e.field.add()
        v.name, v.number, v.type, v.label, v.type_name = "value", 2, F.TYPE_MESSAGE, F.LABEL_OPTIONAL, vtype
        return e

    msg("Model", [("version", 1, F.TYPE_INT32, F.LABEL_OPTIONAL, None),
                  ("embedding_table_infos", 2, F.TYPE_MESSAGE, F.LABEL_REPEATED, ".edltest.EmbeddingTableInfo"),
                  ("dense_parameters", 3, F.TYPE_MESSAGE, F.LABEL_REPEATED, ".edltest.Model.DenseParametersEntry"),
                  ("embedding_tables", 4, F.TYPE_MESSAGE, F.LABEL_REPEATED, ".edltest.Model.EmbeddingTablesEntry")],
        nested=[entry("DenseParametersEntry", ".edltest.TensorProto"),
                entry("EmbeddingTablesEntry", ".edltest.IndexedSlicesProto")])
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    get = lambda n: message_factory.GetMessageClass(pool.FindMessageTypeByName("edltest." + n))
    return get("Model"), get("TensorProto")


def test_tensor_proto_layout_tensor_test_go_52():
    _, TensorProto = _pb_classes()
    a = np.arange(5, dtype=np.float32).reshape(1, 5)
    raw = ck.encode_tensor(a)
    pb = TensorProto()
    pb.ParseFromString(raw)
    assert pb.dtype == 1 and [d.size for d in pb.tensor_shape.dim] == [1, 5]
    assert pb.tensor_content == a.astype("<f4").tobytes()
    assert np.array_equal(ck.decode_tensor(pb.SerializeToString()), a)


def test_model_roundtrip_against_protobuf_library():
    Model, _ = _pb_classes()
    rng = np.random.RandomState(0)
    infos = [("emb/embeddings:0", 8, "uniform", 1), ("wide", 1, "zero", 1)]
    dense = {"dense/kernel:0": rng.randn(3, 4).astype(np.float32), "dense/bias:0": rng.randn(4).astype(np.float32)}
    tables = {"emb/embeddings:0": (np.array([1, 3, 5, 2 ** 40], dtype=np.int64), rng.randn(4, 8).astype(np.float32)),
              "wide": (np.array([], dtype=np.int64), np.zeros((0, 1), np.float32))}
    raw = ck.encode_model(7, infos, dense, tables)
    pb = Model()
    pb.ParseFromString(raw)  # our bytes parse with the real library
    assert pb.version == 7
    assert [(i.name, i.dim, i.initializer, i.dtype) for i in pb.embedding_table_infos] == infos
    assert set(pb.dense_parameters) == set(dense) and set(pb.embedding_tables) == set(tables)
    assert list(pb.embedding_tables["emb/embeddings:0"].ids) == [1, 3, 5, 2 ** 40]
    assert pb.dense_parameters["dense/bias:0"].tensor_content == dense["dense/bias:0"].tobytes()
    # and the library's bytes parse with our decoder
    v, i2, d2, t2 = ck.decode_model(pb.SerializeToString())
    assert v == 7 and i2 == infos
    for k in dense:
        assert np.array_equal(d2[k], dense[k])
    for k in tables:
        assert np.array_equal(t2[k][0], tables[k][0])
        assert np.array_equal(t2[k][1].reshape(tables[k][1].shape), tables[k][1])
    # model_test.go:46-119: ids [1,3,5] rows [[1,2],[3,4],[5,6]] survive a save round trip
    raw = ck.encode_model(0, [("e1", 2, "zero", 1)], {}, {"e1": (np.array([1, 3, 5]), np.array([[1, 2], [3, 4], [5, 6]], np.float32))})
    _, _, _, t = ck.decode_model(raw)
    assert t["e1"][0].tolist() == [1, 3, 5] and t["e1"][1].tolist() == [[1, 2], [3, 4], [5, 6]]


def test_negative_version_and_varints():
    raw = ck.encode_model(-1, [], {}, {})
    assert ck.decode_model(raw)[0] == -1
    for n in (0, 1, 127, 128, 300, 2 ** 31 - 1, 2 ** 62):
        assert ck._read_varint(ck._varint(n), 0)[0] == n


def test_valid_version_dir(tmp_path):
    d = tmp_path / "version-5"
    d.mkdir()
    (d / "variables-0-of-2.ckpt").write_bytes(ck.encode_model(5, [], {}, {}))
    assert not ck.is_valid_version_dir(str(d))  # save_utils.py:212-227: file count must equal N
    (d / "variables-1-of-2.ckpt").write_bytes(ck.encode_model(5, [], {}, {}))
    assert ck.is_valid_version_dir(str(d))
    assert ck.latest_version_dir(str(tmp_path)) == str(d)


def _wire_vectors():
    import json
    import os

    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_python_vectors.json")) as f:
        return json.load(f)["wire"]


def test_tensor_proto_bytes_equal_the_reference_serialiser():
    """tests/golden `wire.tensor_proto`: bytes written by the reference's own ndarray_to_pb (tensor_utils.py:63-77,
    executed by tests/golden/gen_from_reference.py) + SerializeToString.  encode_tensor reproduces them byte for byte
    (0-d tensor: no tensor_shape; a dim of size 0: empty Dim; empty tensor: no tensor_content) and decode_tensor
    reads them back -- so checkpoint files and wire messages are interchangeable with a reference worker's."""
    vec = _wire_vectors()["tensor_proto"]
    assert {tuple(v["shape"]) for v in vec} >= {(), (5,), (3, 4), (0, 8)}
    for v in vec:
        a = np.asarray(v["data"], dtype=v["dtype"]).reshape(v["shape"])
        assert ck.encode_tensor(a).hex() == v["hex"], (v["dtype"], v["shape"])
        b = ck.decode_tensor(bytes.fromhex(v["hex"]))
        assert b.dtype == a.dtype and b.shape == a.shape and np.array_equal(a, b)


def test_indexed_slices_proto_bytes_equal_the_reference_serialiser():
    """`wire.indexed_slices_proto`: indexed_slices_to_pb (tensor_utils.py:103-122) with ids given as an array or a
    list.  The IndexedSlicesProto inside encode_model's embedding_tables entry is the same bytes; decode_model returns
    the same ids / values."""
    for v in _wire_vectors()["indexed_slices_proto"]:
        ids = np.asarray(v["ids"], dtype=np.int64)
        vals = np.asarray(v["values"], dtype=np.float32)
        model = ck.encode_model(0, [], {}, {"t": (ids, vals)})
        (f, _, entry), = list(ck._fields(model))
        assert f == 4
        parts = {f2: v2 for f2, _, v2 in ck._fields(entry)}
        assert parts[1] == b"t" and parts[2].hex() == v["hex"]
        _, _, _, tables = ck.decode_model(model)
        got_ids, got_vals = tables["t"]
        assert got_ids.tolist() == v["ids"] and np.array_equal(got_vals, vals)


def test_checkpoint_directory_logic_equals_the_reference(tmp_path):
    """tests/golden `checkpoint_dirs`: CheckpointSaver's directory logic (save_utils.py:124-141,192-227) executed by
    tests/golden/gen_from_reference.py on real directory trees -- same file names, same verdict on which version
    directories are complete, same latest complete version (numeric order), None for a missing directory."""
    import json
    import os

    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_python_vectors.json")) as f:
        vec = json.load(f)["checkpoint_dirs"]
    for n, sc in enumerate(vec["scenarios"]):
        root = tmp_path / ("case%d" % n)
        root.mkdir()
        for v, files in sc["tree"].items():
            d = root / ("version-" + v)
            d.mkdir()
            for fn in files:
                (d / fn).write_bytes(b"")
        for v, want in sc["valid"].items():
            assert ck.is_valid_version_dir(str(root / ("version-" + v))) == want, (n, v)
        latest = ck.latest_version_dir(str(root))
        assert (None if latest is None else os.path.basename(latest)) == sc["latest"], n
    got = [os.path.relpath(ck._file(str(tmp_path), v, i, k), str(tmp_path)) for v, i, k in [(7, 0, 1), (120, 2, 3)]]
    assert got == vec["file_names"]
    assert ck.latest_version_dir(str(tmp_path / "nope")) is vec["missing_dir_latest"] is None
    assert ck.is_valid_version_dir(str(tmp_path / "nope")) == vec["missing_dir_valid"]
