"""The slice of the master's RPC surface the allreduce controller calls
(elasticai_api/common/master_client.py:29-131).  The ElasticDL master (task sharding,
pod manager, rendezvous server) is cluster control plane and out of scope (SURVEY.md
section 2 rows 14-15); `LocalMasterClient` serves the same calls in-process from the
torchrun environment so the controller / data-shard API run standalone.  A real
master client only has to provide these four methods."""
import os
import threading
from collections import namedtuple

CommRank = namedtuple("CommRank", ("rank_id", "world_size", "rendezvous_id", "rendezvous_port"))
Shard = namedtuple("Shard", ("name", "start", "end"))
Task = namedtuple("Task", ("task_id", "shard", "type"))


class LocalMasterClient(object):
    def __init__(self, batch_size=1, num_epochs=1, dataset_size=0, shuffle=False, num_minibatches_per_shard=8):
        self._lock = threading.Lock()
        self.rendezvous_id = 1
        self._tasks = []
        self._next = 0
        self.reported = []
        self.training_loop_status = None
        records_per_task = max(batch_size * num_minibatches_per_shard, 1)
        tid = 0
        for _ in range(num_epochs or 1):
            for start in range(0, dataset_size or 0, records_per_task):
                self._tasks.append(Task(tid, Shard("", start, min(start + records_per_task, dataset_size)), "training"))
                tid += 1

    # elasticai_api/common/master_client.py:102-105 ; master/servicer.py:180-187
    def get_comm_rank(self):
        return CommRank(int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), self.rendezvous_id,
                        int(os.environ.get("MASTER_PORT", 0)))

    def report_training_loop_status(self, status):
        self.training_loop_status = status

    def get_task(self, task_type=None):
        with self._lock:
            # static round-robin split of the task list over ranks (the master does dynamic sharding)
            rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
            while self._next < len(self._tasks):
                t = self._tasks[self._next]
                self._next += 1
                if t.task_id % world == rank:
                    return t
            return Task(-1, Shard("", 0, 0), "none")

    def report_task_result(self, task_id, err_msg="", exec_counters=None):
        self.reported.append(task_id)

    def report_training_params(self, *a, **k):
        pass


def build_master_client(**kwargs):
    return LocalMasterClient(**kwargs)
