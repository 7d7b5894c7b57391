"""AllReduceController (elasticai_api/common/base_controller.py:48-186) over
torch.distributed.  Horovod's shutdown()/init() on a rendezvous change becomes
destroy_process_group()/init_process_group(); the collective backend is NCCL when the
processes own GPUs (NVLink/NVSwitch, in-switch reduction when available) and gloo on CPU."""
import os
import time
from abc import abstractmethod
from contextlib import contextmanager
from functools import wraps

import torch
import torch.distributed as dist

DEFAULT_MAX_ALLREDUCE_RETRY_NUM = 5      # base_controller.py:41
DEFAULT_SECS_TO_CHECK_RENDEZVOUS = min(60, int(os.getenv("GLOO_TIMEOUT_SECONDS", 30)))  # :44-46
RETRY_ALLREDUCE_INTERVAL_SECS = float(os.getenv("ELASTICAI_RETRY_INTERVAL_SECS", 30))  # :47


class TrainingLoopStatus(object):
    START = 1
    END = 2
    PENDING = 3


def default_backend():
    return "nccl" if torch.cuda.is_available() else "gloo"


class RendevousManager(object):  # (sic) base_controller.py:50
    def __init__(self, master_client, backend=None):
        self.need_broadcast = True
        self._master_client = master_client
        self._rendezvous_id = None
        self._backend = backend

    def init_horovod_if_needed(self):
        """Name kept from the reference (base_controller.py:56-78): (re)build the
        communication group when the master reports a new rendezvous id."""
        rank_response = None
        for _ in range(DEFAULT_MAX_ALLREDUCE_RETRY_NUM):
            rank_response = self._master_client.get_comm_rank()
            if rank_response.rank_id < 0:
                time.sleep(RETRY_ALLREDUCE_INTERVAL_SECS)
            else:
                break
        if rank_response.rank_id < 0:
            raise ValueError("Invalid rank {}".format(rank_response.rank_id))
        if rank_response.rendezvous_id != self._rendezvous_id:
            self._restart(rank_response)

    init_if_needed = init_horovod_if_needed

    def _restart(self, r):
        # hvd.shutdown(); hvd.init()  (base_controller.py:80-93)
        need_new = (not dist.is_initialized()) or dist.get_world_size() != r.world_size or dist.get_rank() != r.rank_id
        if need_new:
            if dist.is_initialized():
                dist.destroy_process_group()
            if r.world_size > 1 or os.environ.get("MASTER_ADDR"):
                os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
                os.environ.setdefault("MASTER_PORT", str(r.rendezvous_port or 29500))
                kwargs = {}
                backend = self._backend or default_backend()
                if backend == "nccl":
                    kwargs["device_id"] = torch.device("cuda", torch.cuda.current_device())
                dist.init_process_group(backend, rank=r.rank_id, world_size=r.world_size, **kwargs)
        self._rendezvous_id = r.rendezvous_id
        self.need_broadcast = True

    def notify_training_loop_status(self, status):
        self._master_client.report_training_loop_status(status)


def comm_size():
    return dist.get_world_size() if dist.is_initialized() else 1


def comm_rank():
    return dist.get_rank() if dist.is_initialized() else 0


class AllReduceController(object):
    """Initialises the communication group and calls the function that runs forward and
    backward on one mini-batch.  If a collective raises, the controller re-initialises the
    group, broadcasts the variables from rank 0 and retries (base_controller.py:109-186)."""

    def __init__(self, master_client, data_shard_service, backend=None):
        self._rendezvous_manager = RendevousManager(master_client, backend)
        self.data_shard_service = data_shard_service
        self._last_init_time = 0
        self._first_call = True
        self._need_broadcast = True

    def elastic_run(self, func):
        @wraps(func)
        def wrapper(*args, **kwargs):
            self._init_variables_before_first_calling(func, *args, **kwargs)
            self._init_horovod_periodically()
            result = self.train_one_batch_with_retries(func, *args, **kwargs)
            self.data_shard_service.report_batch_done()
            return result

        return wrapper

    def init_horovod_locally(self):
        """base_controller.py:138-141: make collectives usable before the first rendezvous."""
        self._rendezvous_manager.init_horovod_if_needed()

    def _init_variables_before_first_calling(self, func, *args, **kwargs):
        if self._first_call:  # base_controller.py:143-147: run once so lazily-built variables exist
            func(*args, **kwargs)
            self._first_call = False

    def _init_horovod_periodically(self):
        cur_time = time.time()
        if cur_time - self._last_init_time > DEFAULT_SECS_TO_CHECK_RENDEZVOUS:
            self._rendezvous_manager.init_horovod_if_needed()
            self._last_init_time = cur_time

    def _broadcast_if_needed(self):
        if self._rendezvous_manager.need_broadcast:
            self.broadcast()
            self._rendezvous_manager.need_broadcast = False

    def notify_train_loop_start(self):
        self._rendezvous_manager.notify_training_loop_status(TrainingLoopStatus.START)

    def notify_train_loop_end(self):
        self._rendezvous_manager.notify_training_loop_status(TrainingLoopStatus.END)

    @contextmanager
    def scope(self):
        self.notify_train_loop_start()
        yield
        self.notify_train_loop_end()

    @abstractmethod
    def train_one_batch_with_retries(self, func, *args, **kwargs):
        pass

    @abstractmethod
    def broadcast(self):
        pass
