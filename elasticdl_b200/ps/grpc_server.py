"""gRPC `Pserver` facade over the HBM shards: the wire surface of the reference's
parameter server (elasticdl/proto/elasticdl.proto:78-86, served by go/pkg/ps/server.go:144-230), so that an
UNMODIFIED remote ElasticDL worker -- python/worker/ps_client.py dialing one address per PS pod -- can train
against the GPU shards.  Off the fast path: requests are protobuf bytes over TCP like the reference's; the
arithmetic is the same CUDA kernels the in-process PSClient uses.

One servicer = one shard (`ps_id`), exactly like one `elasticdl_ps` process:
  push_model                    first writer wins (server.go:209-221): dense parameters of THIS shard, table infos,
                                optional embedding rows, version adopted if >= 1
  push_embedding_table_infos    Model.SetEmbeddingTableInfo + slot tables (server.go:224-230), idempotent
  pull_dense_parameters         initialized / version / all dense tensors when Version >= request (Go: server.go:144-160)
  pull_embedding_vectors        rows of the ids in request order, lazily created (server.go:163-173); empty ids -> empty proto
  push_gradients                ONE ApplyGradients on THIS shard (server.go:176-206): lr staleness modulation against
                                gradients.version, request learning_rate overrides, step++ even on failure (optimizer.go:44),
                                unknown name / width mismatch -> accepted=false + the error, Version++ on success

Messages are encoded / decoded by the hand-written proto3 codec of ps/checkpoint.py (no protoc here); the generic
gRPC handlers pass raw bytes.  grpcio is imported lazily: everything else in the package works without it.
"""
import threading
from concurrent import futures

import numpy as np
import torch

from elasticdl_b200 import _lib
from elasticdl_b200.ps import checkpoint as C

SERVICE = "proto.Pserver"
METHODS = ("push_model", "push_embedding_table_infos", "pull_dense_parameters", "pull_embedding_vectors",
           "push_gradients")


# ----------------------------------------------------------------------------- small messages
def encode_pull_dense_request(version):
    return (C._tag(1, 0) + C._varint(int(version))) if version else b""


def decode_pull_dense_request(buf):
    version = 0
    for f, _, v in C._fields(buf):
        if f == 1:
            version = C._signed(v, 64)
    return version


def encode_pull_dense_response(initialized, version, dense):
    out = bytearray()
    if initialized:
        out += C._tag(1, 0) + C._varint(1)
    if version:
        out += C._tag(2, 0) + C._varint(int(version))
    for name, arr in dense.items():
        out += C._ld(3, C._ld(1, name.encode()) + C._ld(2, C.encode_tensor(np.asarray(arr, dtype=np.float32))))
    return bytes(out)


def decode_pull_dense_response(buf):
    initialized, version, dense = False, 0, {}
    for f, _, v in C._fields(buf):
        if f == 1:
            initialized = bool(v)
        elif f == 2:
            version = C._signed(v, 64)
        elif f == 3:
            key, val = "", b""
            for f2, _, v2 in C._fields(v):
                if f2 == 1:
                    key = v2.decode()
                elif f2 == 2:
                    val = v2
            dense[key] = C.decode_tensor(val)
    return initialized, version, dense


def encode_pull_embedding_request(name, ids):
    ids = np.asarray(ids, dtype=np.int64).reshape(-1)
    out = C._ld(1, name.encode()) if name else b""
    if len(ids):
        out += C._ld(2, b"".join(C._varint(int(i)) for i in ids))
    return out


def decode_pull_embedding_request(buf):
    name, ids = "", []
    for f, w, v in C._fields(buf):
        if f == 1:
            name = v.decode()
        elif f == 2:
            if w == 2:
                pos = 0
                while pos < len(v):
                    x, pos = C._read_varint(v, pos)
                    ids.append(C._signed(x))
            else:
                ids.append(C._signed(v))
    return name, np.asarray(ids, dtype=np.int64)


def encode_push_gradients_request(model_bytes, learning_rate):
    out = C._ld(1, model_bytes)
    if learning_rate:
        out += C._tag(2, 5) + np.float32(learning_rate).tobytes()
    return out


def decode_push_gradients_request(buf):
    model, lr = b"", 0.0
    for f, w, v in C._fields(buf):
        if f == 1:
            model = v
        elif f == 2 and w == 5:
            lr = float(np.frombuffer(v, dtype="<f4")[0])
    return model, lr


def encode_push_gradients_response(accepted, version):
    out = bytearray()
    if accepted:
        out += C._tag(1, 0) + C._varint(1)
    if version:
        out += C._tag(2, 0) + C._varint(int(version))
    return bytes(out)


def decode_push_gradients_response(buf):
    accepted, version = False, 0
    for f, _, v in C._fields(buf):
        if f == 1:
            accepted = bool(v)
        elif f == 2:
            version = C._signed(v, 64)
    return accepted, version


# ----------------------------------------------------------------------------- the servicer
class PserverServicer(object):
    """One shard of a PSGroup behind the Pserver RPCs.  Every shard of the group may live in this process
    (one servicer and port per shard) -- the reference starts one process per PS pod."""

    def __init__(self, group, ps_id):
        self.group = group
        self.ps_id = int(ps_id)
        self.lock = threading.Lock()  # s.lock of server.go:60: one RPC at a time touches the model directory
        self.dense_names = []         # dense parameters of this shard, in arrival order

    # -- helpers
    def _register_dense(self, dense):
        changed = False
        for name, arr in dense.items():
            if name not in self.group.tables:
                self.group.register_dense(name, tuple(np.shape(arr)), self.ps_id)
                changed = True
            if name not in self.dense_names:
                self.dense_names.append(name)
        if changed:
            self.group.commit()

    def _register_tables(self, infos):
        new = False
        for name, dim, init, _dtype in infos:
            if name not in self.group.tables:
                # remote workers send no capacity: unbounded ids -> hashed table, lazily created rows
                self.group.register_table(name, int(dim), init or "zero", None)
                new = True
        if new:
            self.group.commit()

    # -- RPCs (request bytes in, response bytes out)
    def push_model(self, request, context=None):
        version, infos, dense, tables = C.decode_model(request)
        with self.lock:
            self._register_tables(infos)
            self._register_dense(dense)
            if self.group.try_init(self.ps_id):  # !s.Model.Initialized (server.go:212)
                if dense:
                    self.group.set_dense([(n, v) for n, v in dense.items()])
                for name, (ids, values) in tables.items():
                    mine = ids % self.group.n_shards == self.ps_id
                    if mine.any():
                        self.group.set_rows([(name, ids[mine], values[mine])])
                self.group.finish_init(self.ps_id, int(version))
        return b""

    def push_embedding_table_infos(self, request, context=None):
        _, infos, _, _ = C.decode_model(request)
        with self.lock:
            self._register_tables(infos)
        return b""

    def pull_dense_parameters(self, request, context=None):
        want = decode_pull_dense_request(request)
        with self.lock:
            version, _, initialized = self.group.snapshot()[self.ps_id]
            if not initialized:
                return encode_pull_dense_response(False, version, {})
            dense = {}
            if version >= want and self.dense_names:  # Go semantics (quirk Q8)
                pulled = self.group.pull_dense(self.dense_names)
                dense = {k: v.cpu().numpy() for k, v in pulled.items()}
            return encode_pull_dense_response(True, version, dense)

    def pull_embedding_vectors(self, request, context=None):
        name, ids = decode_pull_embedding_request(request)
        if len(ids) == 0:
            return b""  # empty TensorProto (server.go:164-166)
        if name not in self.group.tables:
            return self._fail(context, "Request embedding Table %s not found in Param" % name)
        (rows,) = self.group.pull_rows([(name, ids)])
        self.group.check()
        return C.encode_tensor(rows.cpu().numpy())

    def push_gradients(self, request, context=None):
        model, lr = decode_push_gradients_request(request)
        grad_version, _, dense, tables = C.decode_model(model)
        g, lib, h = self.group, self.group.lib, self.group._h
        with self.lock:
            versions = [0] * _lib.MAX_SHARDS
            versions[self.ps_id] = int(grad_version)
            st = g._stream()
            # effective lr (staleness modulation against gradients.version) + step++ on THIS shard only
            import ctypes

            mv = (ctypes.c_int32 * _lib.MAX_SHARDS)(*versions)
            _lib.check(lib.b200ps_push_begin_shard(h, self.ps_id, float(lr), mv, st))
            try:
                items = []
                for name, arr in dense.items():
                    if name not in g.tables or not g.tables[name][2]:
                        raise KeyError("grad %s not in Parameter" % name)
                    tid, _, _, shape = g.tables[name]
                    t = g._f32(arr)
                    if int(np.prod(shape)) != t.numel():
                        raise ValueError("grad %s has the wrong size" % name)
                    items.append((tid, 0, None, None, t))
                sparse = []
                for name, (ids, values) in tables.items():
                    if name not in g.tables:
                        raise KeyError("grad %s not in Parameter" % name)
                    tid, dim, _, _ = g.tables[name]
                    if len(ids) and values.reshape(len(ids), -1).shape[1] != dim:
                        raise ValueError("grad width is not equal to embedding dim")
                    if len(ids) == 0:
                        continue
                    # the Go kernels apply duplicate ids one after the other (kernel_test.go:49-66); the row
                    # kernels take unique ids, so duplicates are applied in successive launches
                    ids_np, vals_np = np.asarray(ids, dtype=np.int64), values.reshape(len(ids), -1)
                    seen_round = np.zeros(len(ids_np), dtype=np.int64)
                    counts = {}
                    for i, x in enumerate(ids_np.tolist()):
                        seen_round[i] = counts.get(x, 0)
                        counts[x] = seen_round[i] + 1
                    for r in range(int(seen_round.max()) + 1):
                        m = seen_round == r
                        sparse.append((r, tid, g._ids(ids_np[m]), g._f32(vals_np[m])))
                if items:
                    g.push_dense(items)
                for r in sorted({s[0] for s in sparse}):
                    g.push_rows([(tid, ids_t.numel(), ids_t, None, vals_t) for rr, tid, ids_t, vals_t in sparse if rr == r])
                g.check()
            except (KeyError, ValueError) as e:
                version = g.snapshot()[self.ps_id][0]
                self._fail(context, str(e), abort=False)
                return encode_push_gradients_response(False, version)
            _lib.check(lib.b200ps_push_end_shard(h, self.ps_id, st))
            version = g.snapshot()[self.ps_id][0]
        return encode_push_gradients_response(True, version)

    @staticmethod
    def _fail(context, msg, abort=True):
        if context is not None:
            import grpc

            context.set_code(grpc.StatusCode.INTERNAL)
            context.set_details(msg)
            return b""
        if abort:
            raise KeyError(msg)
        return b""


def add_pserver_to_server(servicer, server):
    """Register the five unary-unary methods of proto.Pserver with raw-bytes (de)serialisers."""
    import grpc

    def on_device(fn):  # gRPC worker threads start on device 0: run every RPC on the group's device
        def call(request, context):
            with torch.cuda.device(servicer.group.device):
                return fn(request, context)
        return call

    handlers = {m: grpc.unary_unary_rpc_method_handler(on_device(getattr(servicer, m)), request_deserializer=None,
                                                       response_serializer=None) for m in METHODS}
    server.add_generic_rpc_handlers((grpc.method_handlers_generic_handler(SERVICE, handlers),))


def serve(group, ps_id, port=0, max_workers=8):
    """Start a gRPC server for shard `ps_id` on 127.0.0.1:`port` (0 = pick one).  Returns (server, port, servicer)."""
    import grpc

    server = grpc.server(futures.ThreadPoolExecutor(max_workers=max_workers),
                         options=[("grpc.max_send_message_length", 256 << 20), ("grpc.max_receive_message_length", 256 << 20)])
    servicer = PserverServicer(group, ps_id)
    add_pserver_to_server(servicer, server)
    bound = server.add_insecure_port("127.0.0.1:%d" % port)
    server.start()
    return server, bound, servicer


class PserverStub(object):
    """What a reference worker's generated stub does (elasticdl_pb2_grpc.PserverStub): five unary calls."""

    def __init__(self, channel):
        for m in METHODS:
            setattr(self, m, channel.unary_unary("/%s/%s" % (SERVICE, m), request_serializer=None,
                                                 response_deserializer=None))
