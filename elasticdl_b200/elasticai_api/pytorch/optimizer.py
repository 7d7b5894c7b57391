"""DistributedOptimizer with fixed-global-batch gradient accumulation
(elasticai_api/pytorch/optimizer.py:22-296) over torch.distributed.

B200-first differences from the Horovod original:
  * every gradient is a VIEW of one flat fp32 bucket, so `synchronize()` issues ONE
    all-reduce for the whole model (the reference fires one Horovod collective per
    parameter, optimizer.py:162-168; Keras ResNet-50 has 214 tensors) -- on NVSwitch the cost
    of a collective is launch latency, not links, so one 102 MB NCCL all-reduce beats 214;
  * the pre/post-scale factors of optimizer.py:141-160 are folded into one scale.
  * fused=True (CUDA, world > 1, SGD / momentum / Adam without weight decay): no collective library on
    the data path at all.  The flat gradient bucket lives in a peer-mapped RAW buffer of an HBM
    parameter-server group (one shard per rank); the owner of parameter slice s runs ONE kernel that
    reads the N ranks' gradient slices over NVLink, sums, scales and applies the optimizer update in
    place (b200ps_push_dense_reduce: reduce-scatter + averaging + update fused), then every rank reads
    the updated slices back (b200ps_pull_dense: the all-gather) -- phases ordered by a device-side
    barrier (b200ps_barrier).  The wrapped optimizer's own step() is not used in this mode; its
    hyper-parameters (lr incl. schedulers, momentum, nesterov, betas, eps) travel to the kernels.
Semantics kept: Average / Sum ops, gradient_predivide_factor, backward_passes_per_step,
fixed_global_batch_size (the averaged gradient is invariant to the world size),
skip_synchronize(), set_backward_passes_per_step(), the zero_grad()/step() guards.
"""
import os
import warnings
from contextlib import contextmanager

import torch
import torch.distributed as dist

from elasticdl_b200.elasticai_api.common.base_controller import comm_size

Average = "Average"
Sum = "Sum"


class Compression(object):
    """Placeholder for horovod.torch.Compression (only `none` is meaningful on NVLink)."""

    class none(object):
        @staticmethod
        def compress(tensor):
            return tensor, None

        @staticmethod
        def decompress(tensor, ctx):
            return tensor


class _DevMem(object):
    """A raw device allocation exposed through __cuda_array_interface__ (zero-copy torch view)."""

    def __init__(self, ptr, n_floats, owner):
        self.__cuda_array_interface__ = {"shape": (int(n_floats),), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 3}
        self._owner = owner


def _ps_optimizer_args(opt):
    """(opt_type, opt_args) of the PS kernels for a torch optimizer, or None if it has no fused form."""
    g = opt.param_groups[0]
    if len(opt.param_groups) != 1 or g.get("weight_decay", 0) != 0:
        return None
    if isinstance(opt, torch.optim.SGD):
        if g.get("dampening", 0) != 0 or g.get("maximize", False):
            return None
        return "SGD", "learning_rate=%r;momentum=%r;nesterov=%s;" % (float(g["lr"]), float(g["momentum"]),
                                                                      "true" if g["nesterov"] else "false")
    if isinstance(opt, torch.optim.Adam) and not isinstance(opt, torch.optim.AdamW):
        if g.get("maximize", False):
            return None
        return "Adam", "learning_rate=%r;beta_1=%r;beta_2=%r;epsilon=%r;amsgrad=%s;" % (
            float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]),
            "true" if g.get("amsgrad", False) else "false")
    return None


class _FusedPSBackend(object):
    """The HBM parameter-server backing of DistributedOptimizer(fused=True): one shard per rank, parameter
    slice s (and its optimizer slots) on shard s, every rank's flat gradient bucket in a peer-mapped raw
    buffer.  See the module docstring; reference counterpart: optimizer.py:141-168 + the wrapped step."""

    def __init__(self, params, opt_type, opt_args):
        import ctypes

        from elasticdl_b200 import _lib
        from elasticdl_b200.ps import PSGroup

        self._ct, self._lib, self._check = ctypes, _lib.lib(), _lib.check
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        dev = params[0].device
        self.device = dev
        n = sum((p.numel() + 3) // 4 * 4 for p in params)
        per = ((n + self.world - 1) // self.world + 63) // 64 * 64  # 256 B aligned slices
        self.per, self.total = per, per * self.world
        g = PSGroup(self.world, opt_type, opt_args, device=dev.index, local_shards=[self.rank], track_rows=False)
        self.group = g
        self.slice_ids = [g.register_dense("allreduce/slice_%d" % s, (per,), s) for s in range(self.world)]
        self.raw_id = self._check(self._lib.b200ps_raw_register(g._h, b"allreduce/grads", self.total * 4))
        g.commit()  # exports this rank's shard, maps the peers' (CUDA IPC)
        self.grad_ptrs = []
        for r in range(self.world):
            ptr, nb = ctypes.c_void_p(), ctypes.c_size_t()
            self._check(self._lib.b200ps_raw_ptr(g._h, self.raw_id, r, ctypes.byref(ptr), ctypes.byref(nb)))
            self.grad_ptrs.append(ptr.value)
        self.flat_grads = torch.as_tensor(_DevMem(self.grad_ptrs[self.rank], self.total, self), device=dev)
        self.flat_params = torch.zeros(self.total, dtype=torch.float32, device=dev)
        # bind parameters and gradients to the flat buffers
        off = 0
        with torch.no_grad():
            for p in params:
                k = p.numel()
                self.flat_params[off:off + k].copy_(p.reshape(-1))
                p.data = self.flat_params[off:off + k].view_as(p)
                p.grad = self.flat_grads[off:off + k].view_as(p)
                off += (k + 3) // 4 * 4
        self.params = params
        self._reduce_ptrs = (ctypes.c_void_p * self.world)(*[q + self.rank * per * 4 for q in self.grad_ptrs])
        self._pull = g.make_segs([(tid, 0, None, None, self.flat_params[s * per:(s + 1) * per])
                                  for s, tid in enumerate(self.slice_ids)])
        self.load_params()

    def _stream(self):
        return self._ct.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def barrier(self):
        self._check(self._lib.b200ps_barrier(self.group._h, self._stream()))

    def load_params(self):
        """(Re)initialise the PS master copy of this rank's slice from the local parameters (call after a
        broadcast; every rank holds the same values then)."""
        s = self.rank
        self.group.set_dense([("allreduce/slice_%d" % s, self.flat_params[s * self.per:(s + 1) * self.per])])
        self.barrier()

    def rebind(self):
        off = 0
        for p in self.params:
            k = p.numel()
            view = self.flat_grads[off:off + k].view_as(p)
            if p.grad is None:
                view.zero_()
                p.grad = view
            elif p.grad.data_ptr() != view.data_ptr():
                view.copy_(p.grad)
                p.grad = view
            off += (k + 3) // 4 * 4

    def reduce_update_gather(self, lr, scale):
        lib, h, st = self._lib, self.group._h, self._stream()
        self._check(lib.b200ps_barrier(h, st))                      # every rank's gradients are complete
        self._check(lib.b200ps_push_begin_shard(h, self.rank, float(lr), None, st))
        self._check(lib.b200ps_push_dense_reduce(h, self.slice_ids[self.rank], self._reduce_ptrs, self.world,
                                                 float(scale), st))   # reduce + scale + update, one kernel
        self._check(lib.b200ps_push_end_shard(h, self.rank, st))
        self._check(lib.b200ps_barrier(h, st))                      # every slice is updated
        arr, n = self._pull
        self._check(lib.b200ps_pull_dense(h, arr, n, st))           # all-gather of the parameters


class _DistributedOptimizer(torch.optim.Optimizer):
    def __init__(self, params, named_parameters=None, compression=Compression.none, backward_passes_per_step=1,
                 op=Average, gradient_predivide_factor=1.0, global_batch_num_per_step=None,
                 fixed_global_batch_size=False, fused=False, reproduce_q10=False):
        super(self.__class__, self).__init__(params)
        self._fused_requested = fused
        self._reproduce_q10 = bool(reproduce_q10)
        self._ps = None
        self._compression = compression
        if named_parameters is not None:
            named_parameters = list(named_parameters)
        else:
            named_parameters = [("allreduce.noname.%s" % i, v)
                                for param_group in self.param_groups for i, v in enumerate(param_group["params"])]
        if any([not isinstance(p, tuple) for p in named_parameters]):
            raise ValueError("named_parameters should be a sequence of tuples (name, parameter), "
                             "usually produced by model.named_parameters().")
        dups = _DistributedOptimizer.find_duplicates([k for k, _ in named_parameters])
        if len(dups) > 0:
            raise ValueError("Parameter names in named_parameters must be unique. Found duplicates: %s"
                             % ", ".join(dups))
        all_param_ids = {id(v) for param_group in self.param_groups for v in param_group["params"]}
        named_param_ids = {id(v) for k, v in named_parameters}
        unnamed_param_ids = all_param_ids - named_param_ids
        if len(unnamed_param_ids):
            raise ValueError("named_parameters was specified, but one or more model parameters were not named. "
                             "Python object ids: %s" % ", ".join(str(i) for i in unnamed_param_ids))
        self._parameter_names = {v: k for k, v in sorted(named_parameters)}
        self.backward_passes_per_step = backward_passes_per_step
        self.op = op
        self.gradient_predivide_factor = gradient_predivide_factor
        self._synchronized = False
        self._should_synchronize = True
        self.fixed_global_batch_size = fixed_global_batch_size
        self._global_batch_num_per_step = global_batch_num_per_step
        self._backward_passes = 0
        self.update_gradients = True
        self._build_bucket()

    # ------------------------------------------------------------------ flat bucket
    def _build_bucket(self):
        ps = [p for g in self.param_groups for p in g["params"] if p.requires_grad]
        self._bucket_params = ps
        self._buckets = {}
        if self._fused_requested:
            mapped = _ps_optimizer_args(self)
            ok = (mapped is not None and dist.is_initialized() and dist.get_world_size() > 1 and len(ps) > 0
                  and all(p.is_cuda and p.dtype == torch.float32 for p in ps)
                  and len({p.device for p in ps}) == 1)
            if not ok:
                if self._fused_requested is True:
                    raise ValueError("fused=True needs CUDA fp32 parameters on one device, world_size > 1 and an "
                                     "SGD / Adam optimizer without weight decay (one param group)")
            else:
                self._ps = _FusedPSBackend(ps, *mapped)
                self._buckets[(ps[0].device, torch.float32)] = self._ps.flat_grads
                return
        groups = {}
        for p in ps:
            groups.setdefault((p.device, p.dtype), []).append(p)
        for key, plist in groups.items():
            n = sum((p.numel() + 3) // 4 * 4 for p in plist)  # 16 B aligned views
            flat = torch.zeros(n, device=key[0], dtype=key[1])
            off = 0
            for p in plist:
                p.grad = flat[off:off + p.numel()].view_as(p)  # optimizer.py:127 zero-initialised grads
                off += (p.numel() + 3) // 4 * 4
            self._buckets[key] = flat

    def _rebind_grads(self):
        """If user code replaced p.grad (zero_grad(set_to_none=True)), point it back at the bucket."""
        if self._ps is not None:
            self._ps.rebind()
            return
        for key, flat in self._buckets.items():
            off = 0
            for p in self._bucket_params:
                if (p.device, p.dtype) != key:
                    continue
                view = flat[off:off + p.numel()].view_as(p)
                if p.grad is None:
                    view.zero_()
                    p.grad = view
                elif p.grad.data_ptr() != view.data_ptr():
                    view.copy_(p.grad)
                    p.grad = view
                off += (p.numel() + 3) // 4 * 4

    def load_state_dict(self, *args, **kwargs):
        self._synchronized = False
        self._should_synchronize = True
        super(self.__class__, self).load_state_dict(*args, **kwargs)

    @staticmethod
    def find_duplicates(lst):
        seen, dups = set(), set()
        for el in lst:
            if el in seen:
                dups.add(el)
            seen.add(el)
        return dups

    def set_backward_passes_per_step(self, passes):
        self.backward_passes_per_step = passes

    def _scale_factor(self):
        """prescale * postscale / size of optimizer.py:141-160 as one factor applied to the SUM."""
        if self.op == Average:
            if self.fixed_global_batch_size:
                return 1.0 / self._global_batch_num_per_step
            # Quirk Q10.  Horovod's Average = sum * postscale / size, and its own optimizer passes postscale =
            # predivide (-> the mean).  The reference passes postscale = predivide * size() in BOTH branches
            # (optimizer.py:154,157: "Set size() to the multiplier because C++ backend ... will apply additional
            # 1 / size() factor"), which is what makes the fixed-global-batch mean come out right -- and makes the
            # plain mode return the SUM over ranks (executing the file shows it: tests/golden/gen_allreduce_reference.py).
            # Default here: the mean (Horovod's documented Average, what `op=Average` says); reproduce_q10=True gives
            # the reference's literal result.  The reference itself only ever runs the fixed-global-batch mode.
            return 1.0 if self._reproduce_q10 else 1.0 / comm_size()
        return 1.0

    def synchronize(self):
        self._rebind_grads()
        scale = self._scale_factor()
        if self._ps is not None:
            # reduce-scatter + averaging + optimizer update + all-gather on the PS shards: the parameters are
            # already updated when this returns (step() then has nothing left to do)
            self._ps.reduce_update_gather(self.param_groups[0]["lr"], scale)
            self._synchronized = True
            return
        for flat in self._buckets.values():
            if dist.is_initialized() and dist.get_world_size() > 1:
                if self.gradient_predivide_factor != 1.0:
                    flat.div_(self.gradient_predivide_factor)
                    scale_b = scale * self.gradient_predivide_factor
                else:
                    scale_b = scale
                dist.all_reduce(flat, op=dist.ReduceOp.SUM)
                if scale_b != 1.0:
                    flat.mul_(scale_b)
            elif scale != 1.0:
                flat.mul_(scale)
        self._synchronized = True

    @contextmanager
    def skip_synchronize(self):
        self._should_synchronize = False
        try:
            yield
        finally:
            self._should_synchronize = True

    def step(self, closure=None):
        self._backward_passes += 1
        if self.fixed_global_batch_size and self._backward_passes % self.backward_passes_per_step != 0:
            self.update_gradients = False
        else:
            self.update_gradients = True
            self._backward_passes = 0
        if not self.update_gradients:
            return
        if self._should_synchronize:
            if self._synchronized:
                warnings.warn("optimizer.step() called without optimizer.skip_synchronize() context after "
                              "optimizer.synchronize(). This can cause training slowdown.")
            self.synchronize()
        self._synchronized = False
        if self._ps is not None:
            return None  # the fused kernel applied the update inside synchronize()
        return super(self.__class__, self).step(closure)

    def zero_grad(self, set_to_none=False):
        if not self.update_gradients:
            return
        for flat in self._buckets.values():
            flat.zero_()
        self._rebind_grads()


def DistributedOptimizer(optimizer, named_parameters=None, compression=Compression.none, backward_passes_per_step=1,
                         op=Average, gradient_predivide_factor=1.0, global_batch_num_per_step=None,
                         fixed_global_batch_size=False, fused=False, reproduce_q10=False):
    """optimizer.py:266-296: returns an instance of a dynamically created subclass of the
    wrapped optimizer's class, sharing its param_groups."""
    global_batch_num_per_step = global_batch_num_per_step if global_batch_num_per_step else int(
        os.getenv("WORKER_NUM", 1))
    cls = type(optimizer.__class__.__name__, (optimizer.__class__,), dict(_DistributedOptimizer.__dict__))
    return cls(optimizer.param_groups, named_parameters, compression, backward_passes_per_step, op,
               gradient_predivide_factor, global_batch_num_per_step, fixed_global_batch_size, fused, reproduce_q10)
