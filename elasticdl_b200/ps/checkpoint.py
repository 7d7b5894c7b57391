"""Checkpoint files byte-compatible with the reference PS.

Layout: ``<dir>/version-<v>/variables-<i>-of-<N>.ckpt`` holding one serialised
``proto.Model`` per shard (go/pkg/ps/checkpoint.go:136-141, python/common/save_utils.py:
124-141); a version directory is valid when it holds N files (save_utils.py:212-227);
restore re-shards to a different shard count by re-hashing dense names (string_to_id) and
row ids (id % N) (checkpoint.go:61-133).  Optimizer slots are NOT checkpointed, exactly like
the reference (quirk Q9).

`protoc` is not available here, so the proto3 wire format of the four messages involved
(elasticdl/proto/elasticdl.proto:12-29 Model / EmbeddingTableInfo / IndexedSlicesProto, and
tensorflow TensorProto / TensorShapeProto) is written and parsed by hand; tests cross-check it
against google.protobuf with runtime-built descriptors.
"""
import os
import re
import shutil

import numpy as np

DT_FLOAT = 1   # tensorflow/core/framework/types.proto
DT_INT64 = 9
_NP_OF_DT = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64, 6: np.int8, 5: np.int16}
_DT_OF_NP = {np.dtype(v): k for k, v in _NP_OF_DT.items()}


# ----------------------------------------------------------------------------- wire primitives
def _varint(n):
    n &= (1 << 64) - 1  # negative int32/int64 are sign-extended to 10 bytes
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _tag(field, wire):
    return _varint((field << 3) | wire)


def _ld(field, payload):  # length-delimited
    return _tag(field, 2) + _varint(len(payload)) + payload


def _read_varint(buf, pos):
    shift, val = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        val |= (b & 0x7F) << shift
        if not b & 0x80:
            return val, pos
        shift += 7


def _signed(v, bits=64):
    return v - (1 << bits) if v >> (bits - 1) else v


def _fields(buf):
    """Yield (field, wire_type, value) for every field of a serialised message."""
    pos, end = 0, len(buf)
    while pos < end:
        key, pos = _read_varint(buf, pos)
        field, wire = key >> 3, key & 7
        if wire == 0:
            v, pos = _read_varint(buf, pos)
        elif wire == 2:
            n, pos = _read_varint(buf, pos)
            v = bytes(buf[pos:pos + n])
            pos += n
        elif wire == 1:
            v = bytes(buf[pos:pos + 8])
            pos += 8
        elif wire == 5:
            v = bytes(buf[pos:pos + 4])
            pos += 4
        else:
            raise ValueError("unsupported wire type %d" % wire)
        yield field, wire, v


# ----------------------------------------------------------------------------- TensorProto
def encode_tensor(array):
    """tensorflow.TensorProto{dtype=1, tensor_shape=2{dim=2{size=1}}, tensor_content=4}: raw
    little-endian bytes (python/common/tensor_utils.py:63-71, go/pkg/common/tensor.go:185-200)."""
    array = np.asarray(array, order="C")  # (np.ascontiguousarray would turn a 0-d parameter into shape [1])
    dt = _DT_OF_NP[array.dtype]
    # canonical proto3, byte for byte what serialize_ndarray + SerializeToString give (tests/golden "wire" vectors,
    # produced by executing them): zero-valued scalars and empty bytes are not written -- a dim of size 0 is an empty
    # Dim message, a 0-d tensor has no tensor_shape field, an empty tensor no tensor_content field
    shape = b"".join(_ld(2, (_tag(1, 0) + _varint(int(d))) if d else b"") for d in array.shape)
    content = array.astype(array.dtype.newbyteorder("<"), copy=False).tobytes()
    return _tag(1, 0) + _varint(dt) + (_ld(2, shape) if array.ndim else b"") + (_ld(4, content) if content else b"")


def decode_tensor(buf):
    dt, dims, content = DT_FLOAT, [], b""
    for f, w, v in _fields(buf):
        if f == 1:
            dt = v
        elif f == 2:
            for f2, _, v2 in _fields(v):
                if f2 == 2:
                    size = 0
                    for f3, _, v3 in _fields(v2):
                        if f3 == 1:
                            size = _signed(v3)
                    dims.append(size)
        elif f == 4:
            content = v
    arr = np.frombuffer(content, dtype=np.dtype(_NP_OF_DT[dt]).newbyteorder("<")).astype(_NP_OF_DT[dt])
    # no dims + one element = a scalar (pb_to_ndarray, tensor_utils.py:80-95, builds shape [] from it)
    return arr.reshape(dims) if (dims or arr.size == 1) else arr


# ----------------------------------------------------------------------------- Model
def encode_model(version, infos, dense, tables):
    """infos: [(name, dim, initializer, dtype)]; dense: {name: ndarray};
    tables: {name: (ids int64[n], values float32[n, dim])}."""
    out = bytearray()
    if version:
        out += _tag(1, 0) + _varint(int(version))
    for name, dim, initializer, dtype in infos:
        msg = b""
        if name:
            msg += _ld(1, name.encode())
        if dim:
            msg += _tag(2, 0) + _varint(int(dim))
        if initializer:
            msg += _ld(3, str(initializer).encode())
        if dtype:
            msg += _tag(4, 0) + _varint(int(dtype))
        out += _ld(2, msg)
    for name, arr in dense.items():
        out += _ld(3, _ld(1, name.encode()) + _ld(2, encode_tensor(np.asarray(arr, dtype=np.float32))))
    for name, (ids, values) in tables.items():
        ids = np.asarray(ids, dtype=np.int64).reshape(-1)
        values = np.asarray(values, dtype=np.float32)
        if len(ids):
            values = values.reshape(len(ids), -1)
        sl = _ld(1, encode_tensor(values))
        if len(ids):
            sl += _ld(2, b"".join(_varint(int(i)) for i in ids))  # packed repeated int64
        out += _ld(4, _ld(1, name.encode()) + _ld(2, sl))
    return bytes(out)


def decode_model(buf):
    version, infos, dense, tables = 0, [], {}, {}
    for f, w, v in _fields(buf):
        if f == 1:
            version = _signed(v, 64)
        elif f == 2:
            name, dim, init, dtype = "", 0, "", 0
            for f2, _, v2 in _fields(v):
                if f2 == 1:
                    name = v2.decode()
                elif f2 == 2:
                    dim = _signed(v2)
                elif f2 == 3:
                    init = v2.decode()
                elif f2 == 4:
                    dtype = v2
            infos.append((name, dim, init, dtype))
        elif f in (3, 4):
            key, val = "", b""
            for f2, _, v2 in _fields(v):
                if f2 == 1:
                    key = v2.decode()
                elif f2 == 2:
                    val = v2
            if f == 3:
                dense[key] = decode_tensor(val)
            else:
                values, ids = np.zeros((0, 0), np.float32), []
                for f3, w3, v3 in _fields(val):
                    if f3 == 1:
                        values = decode_tensor(v3)
                    elif f3 == 2:
                        if w3 == 2:  # packed
                            pos = 0
                            while pos < len(v3):
                                x, pos = _read_varint(v3, pos)
                                ids.append(_signed(x))
                        else:
                            ids.append(_signed(v3))
                tables[key] = (np.asarray(ids, dtype=np.int64), values)
    return version, infos, dense, tables


# ----------------------------------------------------------------------------- files
def _file(directory, version, shard, n_shards):
    return os.path.join(directory, "version-%d" % version, "variables-%d-of-%d.ckpt" % (shard, n_shards))


def save(group, checkpoint_dir, version=None, dense_names=None, keep_checkpoint_max=0):
    """Write one file per LOCAL shard of `group` (every rank calls this in a multi-process
    group).  ≙ SaveModelToCheckpoint(dir/version-v, model, id, n) from saveCheckpointIfNeeded
    (server.go:128-141).  Returns the version directory."""
    state = group.snapshot()
    if version is None:
        version = max(s[0] for s in state)
    n = group.n_shards
    for shard in group.local_shards:
        infos, tables, dense = [], {}, {}
        for name, (tid, dim, is_dense, shape) in group.tables.items():
            if is_dense:
                continue
            infos.append((name, dim, group.table_initializers.get(name, "zero"), DT_FLOAT))
            ids = group.table_ids(name, shard)
            vals = group.pull_rows([(name, ids)])[0] if ids.numel() else np.zeros((0, dim), np.float32)
            tables[name] = (ids.cpu().numpy(), vals.cpu().numpy() if hasattr(vals, "cpu") else vals)
        for name, (tid, dim, is_dense, shape) in group.tables.items():
            if is_dense and group.dense_owner.get(name) == shard and (dense_names is None or name in dense_names):
                dense[name] = group.pull_dense([name])[name].cpu().numpy()
        path = _file(checkpoint_dir, version, shard, n)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "wb") as f:
            f.write(encode_model(state[shard][0], infos, dense, tables))
    if keep_checkpoint_max and 0 in group.local_shards:  # server.go:133-139: shard 0 prunes
        versions = sorted(int(m.group(1)) for d in os.listdir(checkpoint_dir)
                          for m in [re.match(r"version-(\d+)$", d)] if m)
        for v in versions[:-keep_checkpoint_max]:
            shutil.rmtree(os.path.join(checkpoint_dir, "version-%d" % v), ignore_errors=True)
    return os.path.join(checkpoint_dir, "version-%d" % version)


def is_valid_version_dir(path):
    """save_utils.py:212-227: all N shard files are present."""
    if not os.path.isdir(path):
        return False
    files = [f for f in os.listdir(path) if re.match(r"variables-\d+-of-\d+\.ckpt$", f)]
    if not files:
        return False
    n = int(re.match(r"variables-\d+-of-(\d+)\.ckpt$", files[0]).group(1))
    return len(files) == n


def latest_version_dir(checkpoint_dir):
    """CheckpointSaver.get_valid_lastest_version_dir (save_utils.py:192-209): the complete version directory with the
    highest version NUMBER, None when there is none (or no such directory)."""
    if not checkpoint_dir or not os.path.isdir(checkpoint_dir):
        return None
    best = None
    for d in os.listdir(checkpoint_dir):
        m = re.match(r"version-(\d+)$", d)
        if m and is_valid_version_dir(os.path.join(checkpoint_dir, d)):
            best = max(best or -1, int(m.group(1)))
    return None if best is None else os.path.join(checkpoint_dir, "version-%d" % best)


def load(group, client, version_dir):
    """≙ LoadModelFromCheckpoint for every local shard of `group` (checkpoint.go:98-133): read
    ALL files of the version, keep what hashes to the local shards under the CURRENT shard
    count, create missing tables / dense parameters, mark the shards initialised and adopt
    the saved version.  Returns the version."""
    if not is_valid_version_dir(version_dir):
        raise ValueError("%s is not a complete checkpoint" % version_dir)
    from elasticdl_b200.common.hash_utils import string_to_id

    n = group.n_shards
    version, all_infos, dense, rows = 0, {}, {}, {}
    for fname in sorted(os.listdir(version_dir)):
        with open(os.path.join(version_dir, fname), "rb") as f:
            v, infos, d, t = decode_model(f.read())
        version = max(version, v)
        for info in infos:
            all_infos[info[0]] = info
        dense.update(d)
        for name, (ids, vals) in t.items():
            rows.setdefault(name, []).append((ids, vals))
    # restored tables are hashed (the checkpoint does not store a capacity): size the slot pool from the
    # number of rows actually saved so that the restore itself can never overflow it
    saved_rows = {name: sum(len(p[0]) for p in parts) for name, parts in rows.items()}
    for name, dim, init, dtype in all_infos.values():
        if name not in group.tables:
            expected = max(2 * saved_rows.get(name, 0), 1 << 16)
            group.register_table(name, int(dim), init, None, expected_rows=expected)
    group.commit()
    if dense:
        client.partition_dense_parameters(dense.keys(), shapes={k: v.shape for k, v in dense.items()})
    for name, parts in rows.items():
        ids = np.concatenate([p[0] for p in parts])
        vals = np.concatenate([p[1] for p in parts]) if len(ids) else np.zeros((0, 1), np.float32)
        mine = np.isin(ids % n, group.local_shards)
        if mine.any():
            group.set_rows([(name, ids[mine], vals[mine])])
            # optimizer slots are not part of a checkpoint (quirk Q9): restored rows start from the slot's
            # initial value even when the group had trained before
            dim_ = vals.shape[1] if vals.ndim == 2 else 1
            for k in (1, 2, 3):
                init_v = 0.0
                if group.opt_type == "Ftrl" and k == 1:  # the "accumulator" slot (optimizer_wrapper.py:116-149)
                    init_v = float(dict(kv.split("=") for kv in group.opt_args.strip(";").split(";"))["initial_accumulator_value"])
                try:
                    group.slot_rows(name, ids[mine], k, np.full((int(mine.sum()), dim_), init_v, dtype=np.float32))
                except ValueError:  # the optimizer has no such slot
                    break
    by_shard = {}
    for name, v in dense.items():
        by_shard.setdefault(string_to_id(name, n), []).append((name, v))
    for shard in group.local_shards:
        if by_shard.get(shard):
            group.set_dense(by_shard[shard])
    group.check()  # a restore that overflowed a table or met a bad id fails HERE, before the shards read as initialised
    for shard in group.local_shards:
        group.set_shard_state(shard, version=version if version >= 1 else -1, initialized=1)
    return version
