"""gRPC Pserver facade (elasticdl_b200/ps/grpc_server.py) driven through real gRPC channels on 127.0.0.1 with the
vectors of go/pkg/ps/server_test.go:107-333: push_model -> pull_embedding_vectors / pull_dense_parameters ->
push_gradients with a request learning rate; plus the PS-side error contract (unknown gradient name)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
F = np.float32
SGD = ("SGD", "learning_rate=0.1;momentum=0.0;nesterov=false;")


def _codec_roundtrips():
    from elasticdl_b200.ps import grpc_server as S

    assert S.decode_pull_dense_request(S.encode_pull_dense_request(7)) == 7
    ini, ver, dense = S.decode_pull_dense_response(S.encode_pull_dense_response(True, 3, {"w": np.arange(6, dtype=F).reshape(2, 3)}))
    assert ini and ver == 3 and np.array_equal(dense["w"], np.arange(6, dtype=F).reshape(2, 3))
    name, ids = S.decode_pull_embedding_request(S.encode_pull_embedding_request("e1", [1, 3, 5]))
    assert name == "e1" and ids.tolist() == [1, 3, 5]
    assert S.decode_push_gradients_response(S.encode_push_gradients_response(True, 9)) == (True, 9)


def test_pserver_rpcs_server_test_go_107():
    import grpc

    from elasticdl_b200.ps import PSGroup
    from elasticdl_b200.ps import checkpoint as C
    from elasticdl_b200.ps import grpc_server as S

    _codec_roundtrips()
    n_shards = 2
    group = PSGroup(n_shards, *SGD, device=0)
    servers = [S.serve(group, i) for i in range(n_shards)]
    stubs = [S.PserverStub(grpc.insecure_channel("127.0.0.1:%d" % port)) for _, port, _ in servers]
    try:
        # dense parameters by string_to_id, two tables; every shard gets the infos (ps_client.py:289-301)
        from elasticdl_b200.common.hash_utils import string_to_id

        t1 = np.array([1, 2, 3, 4, 5, 6], dtype=F)
        t2 = np.array([[1, 2], [1.1, 2.2]], dtype=F)
        infos = [("e1", 2, "zero", 1)]
        dense_by = {0: {}, 1: {}}
        dense_by[string_to_id("t1", n_shards)]["t1"] = t1
        dense_by[string_to_id("t2", n_shards)]["t2"] = t2
        e1_ids = np.array([1, 3, 5, 6], dtype=np.int64)
        e1_vals = np.array([[1, 2], [3, 4], [5, 6], [7, 8]], dtype=F)
        for ps in range(n_shards):
            m = e1_ids % n_shards == ps
            stubs[ps].push_model(C.encode_model(0, infos, dense_by[ps], {"e1": (e1_ids[m], e1_vals[m])}))
        # a second writer is ignored (first writer wins, server.go:212)
        stubs[0].push_model(C.encode_model(0, infos, {k: v * 0 for k, v in dense_by[0].items()}, {}))
        # pull dense (Go: Version >= request)
        got = {}
        for ps in range(n_shards):
            ini, ver, dense = S.decode_pull_dense_response(stubs[ps].pull_dense_parameters(S.encode_pull_dense_request(0)))
            assert ini and ver == 0
            got.update(dense)
        assert np.array_equal(got["t1"], t1) and np.array_equal(got["t2"], t2)
        # pull embedding vectors per shard, request order kept; an id never set reads back as zeros (lazy create)
        for ps in range(n_shards):
            ids = np.array([i for i in (5, 1, 3, 6, 9, 4) if i % n_shards == ps], dtype=np.int64)
            rows = C.decode_tensor(stubs[ps].pull_embedding_vectors(S.encode_pull_embedding_request("e1", ids)))
            want = np.array([e1_vals[list(e1_ids).index(i)] if i in e1_ids else [0, 0] for i in ids], dtype=F)
            assert np.array_equal(rows.reshape(len(ids), 2), want)
        assert stubs[0].pull_embedding_vectors(S.encode_pull_embedding_request("e1", [])) == b""
        # push gradients = parameters with request lr 0.2 -> x - 0.2 x (server_test.go:292-333); every shard gets a push
        for ps in range(n_shards):
            m = e1_ids % n_shards == ps
            grads = C.encode_model(0, [], dense_by[ps], {"e1": (e1_ids[m], e1_vals[m])})
            acc, ver = S.decode_push_gradients_response(stubs[ps].push_gradients(S.encode_push_gradients_request(grads, 0.2)))
            assert acc and ver == 1
        for ps in range(n_shards):
            _, ver, dense = S.decode_pull_dense_response(stubs[ps].pull_dense_parameters(S.encode_pull_dense_request(0)))
            assert ver == 1
            for k, v in dense.items():
                assert np.allclose(v, dense_by[ps][k] - F(0.2) * dense_by[ps][k], rtol=1e-6)
            m = e1_ids % n_shards == ps
            rows = C.decode_tensor(stubs[ps].pull_embedding_vectors(S.encode_pull_embedding_request("e1", e1_ids[m])))
            assert np.allclose(rows.reshape(-1, 2), e1_vals[m] - F(0.2) * e1_vals[m], rtol=1e-6)
        # duplicate ids inside one request are applied one after the other (kernel_test.go:49-66): ids [1, 3, 3]
        g = np.full((3, 2), -1.0, dtype=F)
        before = C.decode_tensor(stubs[1].pull_embedding_vectors(S.encode_pull_embedding_request("e1", [1, 3]))).reshape(2, 2)
        acc, ver = S.decode_push_gradients_response(stubs[1].push_gradients(
            S.encode_push_gradients_request(C.encode_model(1, [], {}, {"e1": (np.array([1, 3, 3]), g)}), 0.1)))
        assert acc and ver == 2
        after = C.decode_tensor(stubs[1].pull_embedding_vectors(S.encode_pull_embedding_request("e1", [1, 3]))).reshape(2, 2)
        assert np.allclose(after - before, [[0.1, 0.1], [0.2, 0.2]], rtol=1e-5)
        # unknown gradient name: accepted = False, version unchanged, error details on the call (optimizer.go:49)
        bad = C.encode_model(2, [], {"nope": np.ones(3, dtype=F)}, {})
        with pytest.raises(grpc.RpcError) as e:
            stubs[1].push_gradients(S.encode_push_gradients_request(bad, 0.1))
        assert "not in Parameter" in e.value.details()
        assert group.snapshot()[1][0] == 2
    finally:
        for srv, _, _ in servers:
            srv.stop(0)
        group.close()
