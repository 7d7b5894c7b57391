"""DataShardService / RecordIndexService (elasticai_api/common/data_shard_service.py:46-212):
workers fetch record-index shards from the master and report batches done."""
import threading
from collections import deque


class DataShardService(object):
    def __init__(self, batch_size, master_client=None, num_epochs=None, dataset_size=None, shuffle=False,
                 task_type="training"):
        self._mc = master_client
        self._batch_size = batch_size
        self._num_epochs = num_epochs
        self._dataset_size = dataset_size
        self._shuffle = shuffle
        self._task_type = task_type
        self._lock = threading.Lock()
        self._pending_tasks = deque()
        self._reported_record_count = 0
        self._current_task = None

    def get_minibatch_count_per_epoch(self):  # data_shard_service.py:97-100
        return self._dataset_size // self._batch_size if self._dataset_size else 0

    def get_current_task(self):
        return self._current_task

    def get_task(self, task_type=None):
        task = self._mc.get_task(task_type or self._task_type)
        if task.task_id < 0 or task.shard.end <= task.shard.start:
            return None
        with self._lock:
            self._pending_tasks.append(task)
            if len(self._pending_tasks) == 1:
                self._current_task = task
        return task

    def fetch_shard(self):
        task = self.get_task()
        return task.shard if task else None

    def report_batch_done(self, batch_size=None, err_msg=""):  # data_shard_service.py:111-148
        """Report a finished minibatch; completed tasks are reported to the master."""
        record_count = batch_size if batch_size else self._batch_size
        self._reported_record_count += record_count
        with self._lock:
            while self._pending_tasks:
                task = self._pending_tasks[0]
                total = task.shard.end - task.shard.start
                if self._reported_record_count < total:
                    break
                self._mc.report_task_result(task.task_id, err_msg)
                self._reported_record_count -= total
                self._pending_tasks.popleft()
            self._current_task = self._pending_tasks[0] if self._pending_tasks else None
        return True


class RecordIndexService(DataShardService):
    """Serves single record indices out of the fetched shards (data_shard_service.py:161-212)."""

    def __init__(self, master_client=None, batch_size=1, num_epochs=None, dataset_size=None, shuffle=False,
                 task_type="training"):
        super().__init__(batch_size, master_client, num_epochs, dataset_size, shuffle, task_type)
        self._indices = deque()

    def fetch_record_index(self):
        if not self._indices:
            shard = self.fetch_shard()
            if shard is None:
                return None
            self._indices.extend(range(shard.start, shard.end))
        return self._indices.popleft()
