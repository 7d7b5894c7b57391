"""DeepFM on Criteo-shaped data over the HBM parameter server: the workload of
BASELINE.json configs[1] (model_zoo/dac_ctr/deepfm_model.py:20-109, utils.py:17-41,
feature_config.py:61-195, feature_transform.py:33,102-110).

38 id groups (12 bucketised integer features + 26 hashed categorical features
capped at 1e6 buckets), one id per group per sample, two embedding tables per
group (wide dim 1, deep dim 8) -> 76 PS tables, 5 549 416 rows per family; dense
tower DNN[16,4] over 13 + 38*8 inputs, FM over the 38x8 deep embeddings, linear
part = sum of wide embeddings + Dense(1)(dense features); Adam lr 1e-3.

`DeepFMPSEngine.step` is one pass of the hot path: pull dense -> unique ids
(first-occurrence) -> pull rows (one launch per vector class for all 76 tables) ->
gather -> tower fwd/bwd (torch, library GEMMs) -> segment-sum of per-occurrence
gradients -> push (dense + 76 tables, fused Adam) -> version++.  No host sync
inside the step; nothing but ids/features/labels comes from the host.
"""
import math

import torch

from elasticdl_b200._lib import check

# model_zoo/dac_ctr/feature_config.py:61-75 (len(boundaries)+1 buckets; I4 is not in FEATURE_GROUPS)
INT_BUCKETS = [5, 10, 9, 10, 9, 9, 10, 10, 3, 6, 2, 9]
# feature_config.py:123-150 distinct counts, capped by MAX_HASHING_BUCKET_SIZE = 1e6 (feature_transform.py:33)
CAT_COUNTS = [1460, 582, 9264260, 2046299, 305, 24, 12506, 633, 3, 91211, 5670, 7659856, 3194, 27, 14876,
              5031503, 10, 5624, 2171, 4, 6477624, 18, 15, 272811, 105, 138075]
GROUP_ROWS = INT_BUCKETS + [min(c, 1000000) for c in CAT_COUNTS]
N_GROUPS = len(GROUP_ROWS)  # 38
N_DENSE = 13
DEEP_DIM = 8
assert N_GROUPS == 38 and sum(GROUP_ROWS) == 5549416


class DeepFMTower(torch.nn.Module):
    """deepfm_model.py:61-109 without the embedding lookups."""

    def __init__(self, n_groups=N_GROUPS, deep_dim=DEEP_DIM, n_dense=N_DENSE, hidden=(16, 4)):
        super().__init__()
        self.n_groups, self.deep_dim = n_groups, deep_dim
        self.dense_linear = torch.nn.Linear(n_dense, 1, bias=False)
        dims = [n_dense + n_groups * deep_dim] + list(hidden)
        self.dnn = torch.nn.ModuleList([torch.nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:])])
        self.dnn_logit = torch.nn.Linear(dims[-1], 1, bias=False)

    def forward(self, dense, wide, deep):
        """dense [B,13], wide [B,G], deep [B,G,D] -> logits [B]."""
        B = dense.shape[0]
        x = torch.cat([dense, deep.reshape(B, -1)], 1)
        for layer in self.dnn:
            x = torch.relu(layer(x))
        dnn_logit = self.dnn_logit(x).squeeze(1)
        linear_logit = wide.sum(1) + self.dense_linear(dense).squeeze(1)
        s = deep.sum(1)  # FM: 0.5 * sum_d[(sum_f e)^2 - sum_f e^2]  (deepfm_edl_embedding.py:50-56)
        fm = 0.5 * (s * s - (deep * deep).sum(1)).sum(1)
        return linear_logit + dnn_logit + fm


def synthetic_batch(batch, seed, device, dist="zipf", zipf_s=1.05, group_rows=GROUP_ROWS):
    """bucket ids uniform, categorical ids Zipf(s) over the
    table's rows (or uniform), dense features N(0,1), labels Bernoulli(0.25).
    Returns ids [G, B] int64 (group-major, like the reference's per-group id tensors)."""
    g = torch.Generator(device=device).manual_seed(seed)
    G = len(group_rows)
    u = torch.rand((G, batch), generator=g, device=device, dtype=torch.float64)
    rows = torch.tensor(group_rows, device=device, dtype=torch.float64).unsqueeze(1)
    if dist == "zipf":
        # bounded-Pareto inverse CDF: rank in [1, N], P(rank) ~ rank^-s
        a = 1.0 - zipf_s
        rank = torch.floor(((rows.pow(a) - 1.0) * u + 1.0).pow(1.0 / a))
        ids = (rank - 1.0).clamp_(min=0)
        ids = torch.minimum(ids, rows - 1)
        small = rows < 64  # the bucketised integer groups are uniform
        ids = torch.where(small.expand_as(ids), torch.floor(u * rows), ids)
    else:
        ids = torch.floor(u * rows)
    ids = ids.to(torch.int64)
    dense = torch.randn((batch, N_DENSE), generator=g, device=device)
    labels = (torch.rand(batch, generator=g, device=device) < 0.25).float()
    return ids, dense, labels


def id_widths(group_rows):
    """Bytes per id of every group on the wire: a table with at most 256 / 65536 rows needs one / two."""
    return [1 if r <= 256 else (2 if r <= 65536 else 4) for r in group_rows]


def _ids_layout(B, widths):
    """[(byte offset, width)] of the per-group id segments (each padded to 16 bytes) and the total."""
    off, out = 0, []
    for w in widths:
        out.append((off, w))
        off += (B * w + 15) // 16 * 16
    return out, off


def packed_nbytes(G, B, widths=None):
    """One training batch as ONE buffer: [ids, group by group at their own width | dense fp32 B x 13 | labels fp32 B].
    widths=None: int32 ids."""
    widths = widths or [4] * G
    return _ids_layout(B, widths)[1] + 4 * B * N_DENSE + 4 * B


def packed_views(buf, G, B, widths=None):
    """(ids, dense fp32 [B, 13], labels fp32 [B]) views of a packed uint8 buffer; ids is an int32 [G, B] view
    when every width is 4, else the uint8 id region itself (the dedup kernel reads it with the widths)."""
    widths = widths or [4] * G
    a = _ids_layout(B, widths)[1]
    b = a + 4 * B * N_DENSE
    ids = buf[:a].view(torch.int32).view(G, B) if all(w == 4 for w in widths) else buf[:a]
    return ids, buf[a:b].view(torch.float32).view(B, N_DENSE), buf[b:b + 4 * B].view(torch.float32)


_NP_UINT = {1: torch.uint8, 2: torch.int16, 4: torch.int32}


def pack_batch(ids, dense, labels, pin=False, widths=None):
    """Pack (ids int64/int32 [G, B], dense [B, 13], labels [B]) into one uint8 buffer on the same device
    (pin=True: pinned host memory).  ids are narrowed to `widths` bytes per group (default int32: every
    dac_ctr table has fewer than 2^31 rows); the device widens them again inside the dedup kernel
    (b200ps_unique_bounded_i32 / b200ps_unique_packed), so the PS sees the same int64 ids."""
    G, B = ids.shape
    widths = widths or [4] * G
    layout, a = _ids_layout(B, widths)
    buf = torch.zeros(packed_nbytes(G, B, widths), dtype=torch.uint8, device="cpu" if pin else ids.device)
    if pin:
        buf = buf.pin_memory()
    for g, (off, w) in enumerate(layout):
        # unsigned narrowing: two's complement of the low bytes (uint16 as int16 bits)
        seg = buf[off:off + B * w].view(_NP_UINT[w])
        v = ids[g].to(torch.int64)
        if w == 2:
            v = torch.where(v >= 32768, v - 65536, v)
        seg.copy_(v.to(_NP_UINT[w]))
    _, v_dense, v_labels = packed_views(buf, G, B, widths)
    v_dense.copy_(dense)
    v_labels.copy_(labels)
    return buf


def flatten_tower(tower, flat):
    """Re-point the tower's parameters at views of one flat buffer laid out as
    include/b200_deepfm.h documents: [w_dense 13 (+3 pad) | w1 | b1 | w2 | b2 | w3]."""
    order = [(tower.dense_linear.weight, 16), (tower.dnn[0].weight, None), (tower.dnn[0].bias, None),
             (tower.dnn[1].weight, None), (tower.dnn[1].bias, None), (tower.dnn_logit.weight, None)]
    off, views = 0, []
    with torch.no_grad():
        for p, padded in order:
            n = p.numel()
            flat[off:off + n].copy_(p.reshape(-1))
            p.data = flat[off:off + n].view(p.shape)
            views.append((off, n))
            off += padded or n
    return views


class HostFeeder:
    """Pinned-host batches -> staging slots on a copy stream -> the captured step.
    A batch travels as ONE packed buffer (pack_batch: int32 ids | dense | labels) = one H2D copy;
    submit() also takes the three separate tensors (three copies, ids narrowed on the device).
    lookahead=True drives engine.step_ahead_graph: slot k's dense/labels train while slot k+1's
    ids are deduplicated, so two batches must have landed before a step starts (depth >= 3)."""

    def __init__(self, engine, depth=2, lookahead=False):
        self.lookahead = bool(lookahead)
        if self.lookahead:
            depth = max(depth, 3)
            if engine.graphs_ahead is None:
                engine.capture_ahead()
        elif getattr(engine, "graph", None) is None:
            engine.capture()
        self.started = False
        self.e = engine
        dev, G, B = engine.device, engine.G, engine.B
        self.depth = depth
        self.slots = [torch.empty(packed_nbytes(G, B, engine.widths), dtype=torch.uint8, device=dev) for _ in range(depth)]
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.ready = [torch.cuda.Event() for _ in range(depth)]
        self.free = [torch.cuda.Event() for _ in range(depth)]
        self.head = self.tail = 0

    def submit(self, *host_batch):
        """Enqueue the H2D copy of one batch (pinned host memory) on the copy stream:
        submit(packed) or submit(ids, dense, labels)."""
        k = self.head % self.depth
        with torch.cuda.stream(self.copy_stream):
            if self.head >= self.depth:
                self.copy_stream.wait_event(self.free[k])
            if len(host_batch) == 1:
                self.slots[k].copy_(host_batch[0], non_blocking=True)
            else:
                self.e._fill(self.slots[k], host_batch)
            self.ready[k].record(self.copy_stream)
        self.head += 1

    def run_next(self):
        """Run one step on the oldest submitted batch; returns the loss (device scalar)."""
        k = self.tail % self.depth
        main = torch.cuda.current_stream(self.e.device)
        main.wait_event(self.ready[k])
        if not self.lookahead:
            loss = self.e.step_graph(self.slots[k])
        else:
            if not self.started:
                self.e.prepare_packed(self.slots[k])
                self.started = True
            k1 = k
            if self.tail + 1 < self.head:  # the next batch was submitted: its ids are deduplicated now
                k1 = (self.tail + 1) % self.depth
                main.wait_event(self.ready[k1])
            loss = self.e.step_ahead_graph(self.slots[k1])
        self.free[k].record(main)
        self.tail += 1
        return loss


class DeepFMPSEngine:
    def __init__(self, group, batch, lr=1e-3, init_std=0.01, seed=7, group_rows=GROUP_ROWS, deep_dim=DEEP_DIM,
                 init_rows=True, tower="tile", paired=None, exchange=None, id_transport="int32"):
        """tower="tile": rows of 32 samples gathered once into shared memory, forward / backward / parameter
        gradients from the tile (csrc/deepfm_tower2.cu);
        tower="fused": round 1's hand-written CUDA tower (csrc/deepfm_tower.cu, three row gathers);
        tower="mma": its tensor-core variant (csrc/deepfm_tower_mma.cu, rows gathered once, 3xTF32 mma.sync);
        tower="torch": torch autograd over library kernels (kept for A/B measurements and tests)."""
        assert tower in ("tile", "fused", "mma", "torch")
        if tower == "tile" and len(group_rows) > 38:
            tower = "fused"  # the tile tower holds at most 38 id groups in its shared-memory row
        self.tower_kind = tower
        # paired=True: the deep (dim 8) and wide (dim 1) tables of an id group share one record
        # per id, so one request per id serves both (ps_kernels.cuh "Paired tables")
        # (default: when rows live on peer GPUs -- remote reads are bounded by requests in flight)
        if paired is None:
            paired = False
        self.paired = bool(paired) and deep_dim == 8
        # exchange="owner": rank-per-GPU groups bucket ids by owner and move everything over NVLink in
        # contiguous runs, the owner serves / updates its own shard (csrc/ps_exchange.cuh);
        # exchange="direct": kernels dereference the peer shard row by row.
        if exchange is None:
            exchange = "owner" if (group.n_shards > 1 and len(group.local_shards) == 1 and deep_dim == 8) else "direct"
        assert exchange in ("owner", "direct")
        self.exchange = exchange
        if exchange == "owner":
            self.paired = False
        self.group = group
        self.B = int(batch)
        self.G = len(group_rows)
        self.D = deep_dim
        self.lr = lr
        dev = group.device
        self.device = dev
        G, B, D = self.G, self.B, self.D
        self.wide_names = ["group_%d_wide/embeddings:0" % i for i in range(G)]
        self.deep_names = ["group_%d_deep/embeddings:0" % i for i in range(G)]
        if self.paired:
            ids_ab = [group.register_pair(d, w, r) for d, w, r in zip(self.deep_names, self.wide_names, group_rows)]
            self.deep_ids = [a for a, _ in ids_ab]
            self.wide_ids = [b for _, b in ids_ab]
        else:
            self.wide_ids = [group.register_table(n, 1, "zero", r) for n, r in zip(self.wide_names, group_rows)]
            self.deep_ids = [group.register_table(n, D, "zero", r) for n, r in zip(self.deep_names, group_rows)]
        torch.manual_seed(seed)
        self.tower = DeepFMTower(G, D).to(dev)
        n_flat = int(group.lib.b200_deepfm_param_count(G))
        self.flat_params = torch.zeros(n_flat, dtype=torch.float32, device=dev)
        self.flat_grads = torch.zeros(n_flat, dtype=torch.float32, device=dev)
        self.flat_views = flatten_tower(self.tower, self.flat_params)
        # named_parameters order != flat order: keep both
        flat_order = [self.tower.dense_linear.weight, self.tower.dnn[0].weight, self.tower.dnn[0].bias,
                      self.tower.dnn[1].weight, self.tower.dnn[1].bias, self.tower.dnn_logit.weight]
        names = {id(p): n for n, p in self.tower.named_parameters()}
        self.params = [(names[id(p)], p) for p in flat_order]
        from elasticdl_b200.common.hash_utils import string_to_id

        self.dense_ids = []
        for n, p in self.params:
            self.dense_ids.append(group.register_dense(n, tuple(p.shape), string_to_id(n, group.n_shards)))
        group.commit()
        # explicit N(0, init_std) initial rows, seed 7: set through the PS (set_rows)
        if init_rows and init_std > 0:
            gen = torch.Generator(device=dev).manual_seed(seed)
            for names, dim in ((self.wide_names, 1), (self.deep_names, D)):
                for n, r in zip(names, group_rows):
                    if r % group.n_shards and group.n_shards > 1:
                        pass
                    ids = torch.arange(r, device=dev)
                    if len(group.local_shards) != group.n_shards:
                        # multi-process: every rank initialises the rows it owns
                        mine = torch.zeros(r, dtype=torch.bool, device=dev)
                        for s in group.local_shards:
                            mine |= (ids % group.n_shards) == s
                        vals = torch.randn((r, dim), generator=gen, device=dev) * init_std
                        ids, vals = ids[mine], vals[mine]
                    else:
                        vals = torch.randn((r, dim), generator=gen, device=dev) * init_std
                    if ids.numel():
                        group.set_rows([(n, ids, vals)])
        # first-writer-wins dense init (PushModel)
        for s in range(group.n_shards):
            if s in group.local_shards and group.try_init(s):
                mine = [(n, p.detach()) for (n, p) in self.params if string_to_id(n, group.n_shards) == s]
                if mine:
                    group.set_dense(mine)
                group.finish_init(s, 0)
        # persistent buffers
        f32 = dict(dtype=torch.float32, device=dev)
        import ctypes as _ctb

        self.bounds = (_ctb.c_int64 * G)(*[int(r) for r in group_rows])  # id ranges: dedup by direct address
        self.ws = torch.zeros(group.lib.b200ps_unique_bounded_workspace(G, B, self.bounds), dtype=torch.uint8, device=dev)
        self._predict_state = None  # predict() dedups into its own workspace / plan (lazy)
        # id transport of the packed batches (pack_batch / HostFeeder / captured graphs): "int32" or "narrow"
        # (1 / 2 / 4 bytes per id by table size -- 72 instead of 152 bytes of ids per dac_ctr sample)
        self.id_transport = id_transport
        self.widths = id_widths(group_rows) if id_transport == "narrow" else [4] * G
        self._c_widths = (_ctb.c_int32 * G)(*self.widths)
        import os as _os

        self.lookahead_blocks_per_sm = int(_os.environ.get("B200_LOOKAHEAD_UNIQUE_BLOCKS", "0"))
        self.lookahead_fork = _os.environ.get("B200_LOOKAHEAD_FORK", "push")  # "push" | "start" (tuning knob)
        self.bet_w = torch.zeros((G * B, 1), **f32)
        self.bet_d = torch.zeros((G * B, D), **f32)
        self.act_w = torch.empty((G * B, 1), **f32)
        self.act_d = torch.empty((G * B, D), **f32)
        self.gsum_w = torch.zeros((G * B, 1), **f32)
        self.gsum_d = torch.zeros((G * B, D), **f32)
        self.scratch = torch.empty(max(B * 44 + 16 * (N_DENSE + D * G), 320 * 16 + 16), **f32)  # b200_deepfm.h: backward state + W1^T | tile tower: W1 tile + counter
        self.loss_buf = torch.zeros(1, **f32)
        # the loss of every step also lands in a ring of pinned host memory, written by a 1-thread kernel at the end
        # of the step (b200_deepfm_publish_loss): loss_host(k) reads step k's value after a synchronize, no D2H copy
        self.LOSS_RING = 256
        self.loss_ring = torch.zeros(self.LOSS_RING, dtype=torch.float32).pin_memory()
        self.loss_cursor = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.logits_buf = torch.empty(B, **f32)
        self.zero_versions = [0] * group.n_shards

        self.loss_fn = torch.nn.BCEWithLogitsLoss()
        # everything addressed through one (uniq, inv, n_unique) triple is a "plan"; the lookahead
        # pipeline (prepare / step_ahead) alternates between two of them
        self.plans = [self._make_plan()]
        self.cur = 0
        self.side = None
        self.aux = None  # second stream of the branched step (see step())
        self.branch = _os.environ.get("B200_STEP_BRANCHES", "1") != "0"
        self.graph = None
        self.graphs_ahead = None
        if self.exchange == "owner":
            group.xchg_create(G, B, self.deep_ids, self.wide_ids)
        self.steps = 0

    _PLAN_ATTRS = ("uniq", "inv", "n_unique", "pull_segs", "push_segs", "pull_pair", "push_pair", "tower_args",
                   "pull_names", "push_names")

    def _make_plan(self):
        dev, G, B = self.device, self.G, self.B
        self.uniq = torch.empty(G * B, dtype=torch.int64, device=dev)
        self.inv = torch.empty(G * B, dtype=torch.int32, device=dev)
        self.n_unique = torch.empty(G, dtype=torch.int32, device=dev)
        self._build_segs()
        return {k: getattr(self, k) for k in self._PLAN_ATTRS}

    def _use(self, p):
        self.__dict__.update(self.plans[p])
        self.cur = p

    def _seg_items(self, ids_tab, rows, dim):
        B = self.B
        return [(ids_tab[t], B, self.uniq[t * B:(t + 1) * B], self.n_unique[t:t + 1], rows[t * B:(t + 1) * B])
                for t in range(self.G)]

    def _build_segs(self):
        g = self.group
        # all 2G tables in ONE flat launch when they fit (csrc/ps_flat.cuh), else one launch per family
        wide_pull, deep_pull = self._seg_items(self.wide_ids, self.bet_w, 1), self._seg_items(self.deep_ids, self.bet_d, self.D)
        wide_push, deep_push = self._seg_items(self.wide_ids, self.gsum_w, 1), self._seg_items(self.deep_ids, self.gsum_d, self.D)
        from elasticdl_b200 import _lib as _l0

        if 2 * self.G <= _l0.MAX_SEGS:
            self.pull_names, self.push_names = ("pull",), ("push",)
            self.pull_segs = [g.make_segs(deep_pull + wide_pull)]
            self.push_segs = [g.make_segs(deep_push + wide_push)]
        else:
            self.pull_names, self.push_names = ("pull_wide", "pull_deep"), ("push_wide", "push_deep")
            self.pull_segs = [g.make_segs(wide_pull), g.make_segs(deep_pull)]
            self.push_segs = [g.make_segs(wide_push), g.make_segs(deep_push)]
        self.pull_dense_segs = g.make_segs([(tid, 0, None, None, p) for tid, (_, p) in zip(self.dense_ids, self.params)])
        import ctypes as _ct
        from elasticdl_b200 import _lib as _l

        half, B = _l.MAX_SEGS // 2, self.B

        def pair_plan(rows_a, rows_b):
            items = self._seg_items(self.deep_ids, rows_a, self.D)
            plan = []
            for i in range(0, self.G, half):
                arr, n = g.make_segs(items[i:i + half])
                ptrs = (_ct.c_void_p * n)(*[rows_b[t * B:(t + 1) * B].data_ptr() for t in range(i, min(i + half, self.G))])
                plan.append((arr, ptrs, n))
            return plan

        self.pull_pair = pair_plan(self.bet_d, self.bet_w)
        self.push_pair = pair_plan(self.gsum_d, self.gsum_w)
        grad_views = [self.flat_grads[off:off + n] for off, n in self.flat_views]
        self.push_dense_segs = g.make_segs([(tid, 0, None, None, gv) for tid, gv in zip(self.dense_ids, grad_views)])
        from elasticdl_b200._lib import DeepFMArgs

        a = DeepFMArgs()
        a.G, a.B = self.G, self.B
        a.inv, a.n_unique = self.inv.data_ptr(), self.n_unique.data_ptr()
        a.bet_wide, a.bet_deep = self.bet_w.data_ptr(), self.bet_d.data_ptr()
        a.params, a.grads = self.flat_params.data_ptr(), self.flat_grads.data_ptr()
        a.gsum_wide, a.gsum_deep = self.gsum_w.data_ptr(), self.gsum_d.data_ptr()
        a.loss, a.logits, a.scratch = self.loss_buf.data_ptr(), self.logits_buf.data_ptr(), self.scratch.data_ptr()
        self.tower_args = a

    def step(self, ids, dense, labels, ev=None, after_tower=None):
        """ids int64 [G, B] (group-major), dense fp32 [B, 13], labels fp32 [B] -- all on the device.
        Returns the loss (device scalar).  ev: optional dict name -> list; CUDA-event pairs are
        recorded around the named PS kernels on the launching stream."""
        g, lib, h = self.group, self.group.lib, self.group._h
        G, B, D = self.G, self.B, self.D
        st = g._stream()

        def mark(name):
            if ev is None:
                return None
            e = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev.setdefault(name, []).append(e)
            e[0].record()
            return e

        def done(e):
            if e is not None:
                e[1].record()
        import ctypes as _ct

        # Branches: the step has six launch-bound kernels (pull_dense, tower prologue, push_begin, push_dense,
        # push_end; 4-8 us each, ncu profiles/r2_*) that do not depend on the row kernels beside them.  With
        # `branch` they run on a second stream -- inside a captured graph these become parallel branches --
        # while the main stream carries unique -> pull -> tower -> push rows:
        #   aux:  pull_dense -> push_begin -> tower prologue  | joined before the tower
        #   aux:  push_dense                                   | beside the row push, joined before push_end
        branch = self.branch and self.tower_kind == "tile" and ev is None
        main = torch.cuda.current_stream(self.device)
        if branch:
            if self.aux is None:
                self.aux = torch.cuda.Stream(device=self.device)
            a = self.tower_args
            a.dense, a.labels = dense.data_ptr(), labels.data_ptr()
        # (2) unique ids per group; wide and deep tables of a group share them
        if ids is not None:
            e = mark("unique")
            self._unique_into(ids)
            done(e)
        # (1) pull dense parameters straight into the tower's tensors
        if branch:
            self.aux.wait_stream(main)  # n_unique of this batch, parameters of the previous step
            with torch.cuda.stream(self.aux):
                arr, n = self.pull_dense_segs
                check(lib.b200ps_pull_dense(h, arr, n, g._stream()))
                rc = lib.b200_deepfm_tile_prologue(_ct.byref(a), g._stream())
                if rc:
                    raise RuntimeError("b200_deepfm_tile_prologue failed (%d)" % rc)
                ev_prep = torch.cuda.Event()
                ev_prep.record(self.aux)  # the tower waits for this, not for push_begin behind it
                g.push_begin(self.lr, self.zero_versions)
                ev_begin = torch.cuda.Event()
                ev_begin.record(self.aux)
        else:
            arr, n = self.pull_dense_segs
            check(lib.b200ps_pull_dense(h, arr, n, st))
        # (3) pull the unique rows of all 76 tables
        if self.exchange == "owner":
            e = mark("pull_exchange")
            check(lib.b200ps_xchg_pull(h, self.uniq.data_ptr(), self.n_unique.data_ptr(), self.bet_d.data_ptr(),
                                       self.bet_w.data_ptr(), st))
            done(e)
        elif self.paired:
            e = mark("pull_pair")
            for arr, ptrs, n in self.pull_pair:
                check(lib.b200ps_pull_rows_pair(h, arr, ptrs, n, st))
            done(e)
        else:
            for name, (arr, n) in zip(self.pull_names, self.pull_segs):
                e = mark(name)
                check(lib.b200ps_pull_rows(h, arr, n, st))
                done(e)
        if branch:
            main.wait_event(ev_prep)
            rc = lib.b200_deepfm_tile_main(_ct.byref(a), st)
            if rc:
                raise RuntimeError("b200_deepfm_tile_main failed (%d)" % rc)
            loss = self.loss_buf
            dense_segs = self.push_dense_segs
        elif self.tower_kind in ("tile", "fused", "mma"):
            # (4-6) gather + tower forward/backward + per-unique-id gradient sums: three launches
            e_t = mark("tower_fwd_bwd")
            a = self.tower_args
            a.dense, a.labels = dense.data_ptr(), labels.data_ptr()

            fn = {"mma": lib.b200_deepfm_fwd_bwd_mma, "tile": lib.b200_deepfm_fwd_bwd_tile}.get(self.tower_kind, lib.b200_deepfm_fwd_bwd)
            rc = fn(_ct.byref(a), st)
            if rc:
                raise RuntimeError("b200_deepfm_fwd_bwd failed (%d)" % rc)
            done(e_t)
            loss = self.loss_buf
            dense_segs = self.push_dense_segs
        else:
            loss, dense_segs = self._torch_tower(dense, labels, mark, done, st)
        if after_tower is not None:
            after_tower()  # lookahead pipeline: the next batch's dedup forks here (beside the push)
        # (7) push: one ApplyGradients per shard
        if branch:
            main.wait_event(ev_begin)  # this push's lr / Adam alpha / step (k_push_begin ran on the second stream)
            self.aux.wait_stream(main)
            with torch.cuda.stream(self.aux):
                arr, n = dense_segs
                check(lib.b200ps_push_dense(h, arr, n, g._stream()))
                self._publish_loss(loss, g._stream())
        else:
            g.push_begin(self.lr, self.zero_versions)
            arr, n = dense_segs
            check(lib.b200ps_push_dense(h, arr, n, st))
            self._publish_loss(loss, st)
        if self.exchange == "owner":
            e = mark("push_exchange")
            check(lib.b200ps_xchg_push(h, self.gsum_d.data_ptr(), self.gsum_w.data_ptr(), st))
            done(e)
        elif self.paired:
            e = mark("push_pair")
            for arr, ptrs, n in self.push_pair:
                check(lib.b200ps_push_rows_pair(h, arr, ptrs, n, st))
            done(e)
        else:
            for name, (arr, n) in zip(self.push_names, self.push_segs):
                e = mark(name)
                check(lib.b200ps_push_rows(h, arr, n, st))
                done(e)
        if branch:
            main.wait_stream(self.aux)
        g.push_end(sync=False)
        self.steps += 1
        return loss.detach().reshape(())

    def _publish_loss(self, loss, st):
        if loss is not self.loss_buf:  # torch tower: its loss tensor is not persistent
            self.loss_buf.copy_(loss.detach().reshape(1))
        rc = self.group.lib.b200_deepfm_publish_loss(self.loss_buf.data_ptr(), self.loss_ring.data_ptr(), self.LOSS_RING,
                                                     self.loss_cursor.data_ptr(), st)
        if rc:
            raise RuntimeError("b200_deepfm_publish_loss failed (%d)" % rc)

    def loss_host(self, step_index):
        """Loss of step `step_index` (0-based count of steps run by this engine) from the host ring; valid once
        the device has finished that step (synchronize / event) and for the last LOSS_RING steps."""
        return float(self.loss_ring[step_index % self.LOSS_RING])

    def _unique_into(self, ids, blocks_per_sm=0):
        """tf.unique per id group into the current plan, on the current stream.  blocks_per_sm > 0: a thin
        persistent grid (the lookahead pipeline runs the dedup beside the training kernels)."""
        g = self.group
        if ids.dtype == torch.uint8:  # the id region of a packed batch with per-group widths
            check(g.lib.b200ps_unique_packed(g._h, ids.data_ptr(), self._c_widths, self.G, self.B, self.bounds,
                                             self.uniq.data_ptr(), self.inv.data_ptr(), self.n_unique.data_ptr(),
                                             self.ws.data_ptr(), self.ws.numel(), int(blocks_per_sm), g._stream()))
            return
        check(g.lib.b200ps_unique_bounded_ex(g._h, ids.data_ptr(), 1 if ids.dtype == torch.int32 else 0, self.G, self.B,
                                             self.bounds, self.uniq.data_ptr(), self.inv.data_ptr(),
                                             self.n_unique.data_ptr(), self.ws.data_ptr(), self.ws.numel(),
                                             int(blocks_per_sm), g._stream()))

    # ------------------------------------------------------------------ lookahead pipeline
    def _ensure_plans(self):
        if len(self.plans) == 1:
            cur = self.cur
            self.plans.append(self._make_plan())
            self._use(cur)
            self.side = torch.cuda.Stream(device=self.device)

    def prepare(self, ids):
        """Start the lookahead pipeline: deduplicate the FIRST batch's ids.  Afterwards every
        step_ahead(dense_i, labels_i, ids_{i+1}) trains on batch i while the ids of batch i+1 are
        deduplicated on a second stream (the dedup only depends on the input batch, like the
        reference's dataset.prefetch(1) work, elasticdl/python/worker/worker.py:334)."""
        self._ensure_plans()
        self._unique_into(ids)

    def step_ahead(self, dense, labels, next_ids, ev=None):
        """One training step on the prepared batch + the dedup of `next_ids` (int64 or int32 [G, B],
        device) overlapped on the side stream.  Same kernels, same results as step()."""
        if len(self.plans) != 2:
            raise RuntimeError("call prepare(first_ids) before step_ahead")
        main = torch.cuda.current_stream(self.device)
        p = self.cur

        def fork():
            # The dedup of the next batch runs beside the PUSH of this one: the push kernels are light
            # (48 registers, < 1 KB of shared memory) and share the SMs with the persistent dedup blocks,
            # whereas the tower's CTAs need the whole register file of an SM and would simply wait for
            # the dedup to finish (measured: forking at the start of the step hid 11 of its 45 us).
            self.side.wait_stream(main)  # the other plan's buffers were last read by the previous step
            self._use(1 - p)
            with torch.cuda.stream(self.side):
                self._unique_into(next_ids, blocks_per_sm=self.lookahead_blocks_per_sm)
            self._use(p)

        if self.lookahead_fork != "push":
            fork()  # "start": beside the whole step
        loss = self.step(None, dense, labels, ev=ev, after_tower=fork if self.lookahead_fork == "push" else None)
        main.wait_stream(self.side)  # join
        self._use(1 - p)
        return loss

    def capture_ahead(self):
        """Two CUDA graphs (one per plan parity) of step_ahead over two static PACKED batch buffers
        S[0], S[1] (pack_batch layout): graph p trains on the dense / labels of S[p] (its ids were
        deduplicated into plan p by the previous step) while it deduplicates the ids of S[1 - p]."""
        if self.tower_kind == "torch":
            raise RuntimeError("graph capture needs the fused tower (torch autograd allocates)")
        dev, G, B = self.device, self.G, self.B
        self._ensure_plans()
        self.s_packed = [torch.zeros(packed_nbytes(G, B, self.widths), dtype=torch.uint8, device=dev) for _ in range(2)]
        views = [packed_views(b, G, B, self.widths) for b in self.s_packed]
        torch.cuda.synchronize(dev)
        start, steps = self.cur, self.steps
        self.graphs_ahead = [None, None]
        for p in (start, 1 - start):
            assert self.cur == p
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                self.step_ahead(views[p][1], views[p][2], views[1 - p][0])
            self.graphs_ahead[p] = gr
        self._use(start)
        self.steps = steps  # capture does not execute
        return self.graphs_ahead

    def _fill(self, dst, batch):
        """dst: static packed buffer; batch: (packed,) or (ids, dense, labels) -- device or pinned host."""
        if len(batch) == 1:
            dst.copy_(batch[0], non_blocking=True)
        elif all(w == 4 for w in self.widths):
            for d, src in zip(packed_views(dst, self.G, self.B), batch):
                d.copy_(src, non_blocking=True)
        else:  # separate tensors into a narrow layout: pack on the device (compatibility path)
            ids, dense, labels = batch
            dst.copy_(pack_batch(ids.to(self.device), dense.to(self.device), labels.to(self.device), widths=self.widths),
                      non_blocking=True)

    def prepare_packed(self, *batch):
        """Lookahead through the captured graphs: load the FIRST batch and deduplicate its ids."""
        if self.graphs_ahead is None:
            self.capture_ahead()
        self._fill(self.s_packed[self.cur], batch)
        self._unique_into(packed_views(self.s_packed[self.cur], self.G, self.B, self.widths)[0])

    def step_ahead_graph(self, *next_batch):
        """step_ahead through the captured graphs: trains on the batch loaded by the previous call (or
        prepare_packed) while `next_batch` -- (packed,) or (ids, dense, labels) -- is loaded and its ids
        are deduplicated.  Inputs may be pinned-host or device tensors."""
        self._fill(self.s_packed[1 - self.cur], next_batch)
        self.graphs_ahead[self.cur].replay()
        self._use(1 - self.cur)
        self.steps += 1
        return self.loss_buf.reshape(())

    # ------------------------------------------------------------------ CUDA graph
    def capture(self):
        """Capture one whole step into a CUDA graph.  Every buffer the step touches is persistent, so
        the graph replays on new inputs copied into the static packed input buffer.  lr is baked in:
        re-capture to change it."""
        if self.tower_kind == "torch":
            raise RuntimeError("graph capture needs the fused tower (torch autograd allocates)")
        dev, G, B = self.device, self.G, self.B
        self.s_one = torch.zeros(packed_nbytes(G, B, self.widths), dtype=torch.uint8, device=dev)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.step(*packed_views(self.s_one, G, B, self.widths))
        self.steps -= 1  # capture does not execute
        return self.graph

    def step_graph(self, *batch):
        """Same step through the captured graph; batch = (packed,) or (ids, dense, labels), pinned-host
        or device tensors."""
        self._fill(self.s_one, batch)
        self.graph.replay()
        self.steps += 1
        return self.loss_buf.reshape(())

    def host_feeder(self, depth=2, lookahead=False):
        """Input pipeline for host-resident batches: H2D copies run on a side stream into
        `depth` staging slots and overlap the previous step's kernels (the reference prefetches
        one batch too: dataset.prefetch(1), elasticdl/python/worker/worker.py:334)."""
        return HostFeeder(self, depth, lookahead)

    def _torch_tower(self, dense, labels, mark, done, st):
        """Steps (4)-(6) with torch autograd over library kernels (A/B reference for the fused tower)."""
        g, lib, h = self.group, self.group.lib, self.group._h
        G, B, D = self.G, self.B, self.D
        check(lib.b200ps_gather_rows(h, self.bet_w.data_ptr(), self.inv.data_ptr(), G, B, 1, self.act_w.data_ptr(), st))
        check(lib.b200ps_gather_rows(h, self.bet_d.data_ptr(), self.inv.data_ptr(), G, B, D, self.act_d.data_ptr(), st))
        e_t = mark("tower_fwd_bwd")
        act_w = self.act_w.detach().requires_grad_(True)
        act_d = self.act_d.detach().requires_grad_(True)
        wide = act_w.view(G, B).t()
        deep = act_d.view(G, B, D).permute(1, 0, 2)
        logits = self.tower(dense, wide, deep)
        loss = self.loss_fn(logits, labels)
        plist = [p for _, p in self.params]
        grads = torch.autograd.grad(loss, plist + [act_w, act_d])
        gw, gd = grads[-2].contiguous(), grads[-1].contiguous()
        done(e_t)
        check(lib.b200ps_segment_sum(h, gw.data_ptr(), self.inv.data_ptr(), G, B, 1, self.gsum_w.data_ptr(), st))
        e = mark("segment_sum_deep")
        check(lib.b200ps_segment_sum(h, gd.data_ptr(), self.inv.data_ptr(), G, B, D, self.gsum_d.data_ptr(), st))
        done(e)
        self._dense_grads = [gr.contiguous() for gr in grads[:-2]]
        return loss, g.make_segs([(tid, 0, None, None, gr) for tid, gr in zip(self.dense_ids, self._dense_grads)])

    def predict(self, ids, dense):
        """Forward only (logits) through the PS: unique -> pull -> fused tower forward.  Uses its own
        dedup workspace and (uniq, inv, n_unique) plan, so a prepared lookahead batch stays intact."""
        g, lib, h = self.group, self.group.lib, self.group._h
        st = g._stream()
        if self._predict_state is None:
            cur = self.cur
            ws = torch.zeros(lib.b200ps_unique_workspace(self.G, self.B), dtype=torch.uint8, device=self.device)
            plan = self._make_plan()
            self._use(cur)
            self._predict_state = (ws, plan)
        ws, plan = self._predict_state
        saved = {k: getattr(self, k) for k in self._PLAN_ATTRS}
        self.__dict__.update(plan)
        try:
            arr, n = self.pull_dense_segs
            check(lib.b200ps_pull_dense(h, arr, n, st))
            ids = ids.to(torch.int64)
            check(lib.b200ps_unique(h, ids.data_ptr(), self.G, self.B, self.uniq.data_ptr(), self.inv.data_ptr(),
                                    self.n_unique.data_ptr(), ws.data_ptr(), ws.numel(), st))
            for arr, n in self.pull_segs:
                check(lib.b200ps_pull_rows(h, arr, n, st))
            import ctypes as _ct

            a = self.tower_args
            a.dense = dense.data_ptr()
            fwd = lib.b200_deepfm_forward_tile if self.tower_kind == "tile" else lib.b200_deepfm_forward
            if fwd(_ct.byref(a), st):
                raise RuntimeError("b200_deepfm_forward failed")
        finally:
            self.__dict__.update(saved)
        return self.logits_buf

    def kernel_report(self, ev, uniq_per_step, opt_slots=2):
        """Average CUDA-event duration and algorithmic GB/s per named PS kernel.
        Algorithmic bytes: pull U*(8+8D); push U*(8+4D+(1+S)*8D); segment_sum
        k*(4+4D)+U*4D; unique k*(8+4)+U*8 (ids in, inverse out, unique ids out)."""
        D, k = self.D, self.G * self.B
        out = {}
        for name, pairs in ev.items():
            ms = [a.elapsed_time(b) for a, b in pairs]
            U = sum(uniq_per_step) / max(len(uniq_per_step), 1)
            nbytes = {
                "pull_wide": U * (8 + 8 * 1), "pull_deep": U * (8 + 8 * D),
                "pull": U * (8 + 8 * 1) + U * (8 + 8 * D),
                "push": U * (8 + 4 * 1 + (1 + opt_slots) * 8 * 1) + U * (8 + 4 * D + (1 + opt_slots) * 8 * D),
                # tower: per (sample, group) one rank read + one deep row + one wide value gathered, the
                # embedding gradient scattered back (reduced per unique id), + the dense inputs / labels
                "tower_fwd_bwd": k * (4 + 4 * D + 4) + U * (4 * D + 4) + self.B * 4 * (N_DENSE + 1),
                "pull_pair": U * (8 + 8 * D + 8 * 1), "pull_exchange": U * (8 + 8 * D + 8 * 1),
                "push_exchange": U * (8 + 4 * D + 4 * 1 + (1 + opt_slots) * 8 * (D + 1)),
                "push_pair": U * (8 + 4 * D + 4 * 1 + (1 + opt_slots) * 8 * (D + 1)),
                "push_wide": U * (8 + 4 * 1 + (1 + opt_slots) * 8 * 1),
                "push_deep": U * (8 + 4 * D + (1 + opt_slots) * 8 * D),
                "segment_sum_deep": k * (4 + 4 * D) + U * 4 * D,
                "unique": k * 12 + U * 8,
            }.get(name)
            avg = sum(ms) / len(ms)
            out[name] = {"ms": avg, "launches": len(ms)}
            if nbytes:
                out[name]["bytes"] = nbytes
                out[name]["gbs"] = nbytes / (avg * 1e-3) / 1e9
        return out

    @staticmethod
    def algorithmic_bytes(n_unique_total, opt_slots=2, deep_dim=DEEP_DIM):
        """SURVEY.md section 8d per-unit figures, fp32 rows, int64 ids:
        pull U*(8 + 4D + 4D); sparse push U*(8 + 4D + (1+S)*8D), for D in {1, deep_dim}."""
        pull = sum(n_unique_total * (8 + 8 * d) for d in (1, deep_dim))
        push = sum(n_unique_total * (8 + 4 * d + (1 + opt_slots) * 8 * d) for d in (1, deep_dim))
        return pull, push


class DeepFMLayersModel(torch.nn.Module):
    """The model a model_zoo job runs through the drop-in API (model_zoo/dac_ctr/deepfm_model.py:33-109):
    one elasticdl Embedding layer per feature group and family -- G deep layers (dim 8) and G wide layers
    (dim 1), each looked up with its own id tensor -- in front of DeepFMTower in eager torch.
    forward(features): features = {"dense": [B, 13] float32, "ids_0" .. "ids_{G-1}": [B] int64}.
    Carries `.optimizer` / `.loss` as ParameterServerTrainer expects (worker/ps_trainer.py)."""

    def __init__(self, group_rows=GROUP_ROWS, deep_dim=DEEP_DIM, lr=1e-3, initializer="zero"):
        super().__init__()
        from elasticdl_b200.layers import Embedding

        G = len(group_rows)
        self.deep = torch.nn.ModuleList(
            [Embedding(deep_dim, input_dim=r, embeddings_initializer=initializer, name="deep_%d" % g)
             for g, r in enumerate(group_rows)])
        self.wide = torch.nn.ModuleList(
            [Embedding(1, input_dim=r, embeddings_initializer=initializer, name="wide_%d" % g)
             for g, r in enumerate(group_rows)])
        self.tower = DeepFMTower(G, deep_dim)
        self.optimizer = torch.optim.Adam(self.tower.parameters(), lr=lr)
        bce = torch.nn.BCEWithLogitsLoss()
        self.loss = lambda labels, logits: bce(logits, labels)

    @staticmethod
    def features_of(ids, dense):
        """ids [G, B] int64 -> the per-feature dict (row views of the one array, as a dataset's
        feature columns are)."""
        f = {"dense": dense}
        for g in range(ids.shape[0]):
            f["ids_%d" % g] = ids[g]
        return f

    def forward(self, features):
        G = len(self.deep)
        deep = torch.stack([self.deep[g](features["ids_%d" % g]) for g in range(G)], 1)  # [B, G, D]
        wide = torch.cat([self.wide[g](features["ids_%d" % g]) for g in range(G)], 1)    # [B, G]
        return self.tower(features["dense"], wide, deep)
