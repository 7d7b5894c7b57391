"""DistributedOptimizer with fixed-global-batch gradient accumulation
(elasticai_api/pytorch/optimizer.py:22-296) over torch.distributed.

B200-first differences from the Horovod original:
  * every gradient is a VIEW of one flat fp32 bucket, so `synchronize()` issues ONE
    all-reduce for the whole model (the reference fires one Horovod collective per
    parameter, optimizer.py:162-168; Keras ResNet-50 has 214 tensors) -- on NVSwitch the cost
    of a collective is launch latency, not links, so one 102 MB NCCL all-reduce beats 214;
  * the pre/post-scale factors of optimizer.py:141-160 are folded into one scale.
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


class _DistributedOptimizer(torch.optim.Optimizer):
    def __init__(self, params, named_parameters=None, compression=Compression.none, backward_passes_per_step=1,
                 op=Average, gradient_predivide_factor=1.0, global_batch_num_per_step=None,
                 fixed_global_batch_size=False):
        super(self.__class__, self).__init__(params)
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
            return 1.0 / comm_size()
        return 1.0

    def synchronize(self):
        self._rebind_grads()
        scale = self._scale_factor()
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
        return super(self.__class__, self).step(closure)

    def zero_grad(self, set_to_none=False):
        if not self.update_gradients:
            return
        for flat in self._buckets.values():
            flat.zero_()
        self._rebind_grads()


def DistributedOptimizer(optimizer, named_parameters=None, compression=Compression.none, backward_passes_per_step=1,
                         op=Average, gradient_predivide_factor=1.0, global_batch_num_per_step=None,
                         fixed_global_batch_size=False):
    """optimizer.py:266-296: returns an instance of a dynamically created subclass of the
    wrapped optimizer's class, sharing its param_groups."""
    global_batch_num_per_step = global_batch_num_per_step if global_batch_num_per_step else int(
        os.getenv("WORKER_NUM", 1))
    cls = type(optimizer.__class__.__name__, (optimizer.__class__,), dict(_DistributedOptimizer.__dict__))
    return cls(optimizer.param_groups, named_parameters, compression, backward_passes_per_step, op,
               gradient_predivide_factor, global_batch_num_per_step, fixed_global_batch_size)
