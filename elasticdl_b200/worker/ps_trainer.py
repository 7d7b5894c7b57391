"""ParameterServerTrainer over torch: the caller of the PS hot path.

Mirrors elasticdl/python/worker/ps_trainer.py:36-440 method for method
(`train_minibatch` -> (accepted, version, loss), `_get_model`, `_report_gradient`,
`init_variables_if_need`, local updates between pulls) with the TensorFlow pieces
restated over torch autograd:

  * `model` is a torch.nn.Module with attributes `optimizer` (a torch optimizer on
    its non-embedding parameters; its lr travels to the PS on every push,
    ps_trainer.py:275) and `loss` (callable(labels, outputs));
  * non-embedding variables are `model.named_parameters()`; ElasticDL Embedding
    layers are found by type (ps_trainer.py:66-74) and wired to
    `ps_client.pull_embedding_vectors`.
"""
import numpy as np
import torch

from elasticdl_b200.common.tensor_utils import DT_FLOAT, EmbeddingTableInfo, Tensor, UniqueTensor
from elasticdl_b200.layers.embedding import Embedding


def find_layer(model, layer_class):
    return [m for m in model.modules() if isinstance(m, layer_class)]


class _NoTiming:
    def start_record_time(self, *_):
        pass

    def end_record_time(self, *_):
        pass


class ParameterServerTrainer(object):
    """Parameter Server Trainer"""

    def __init__(self, model, ps_client, timing=None, args=None):
        self._optimizer = model.optimizer
        self._loss = model.loss
        self._model = model
        self._ps_client = ps_client
        if self._ps_client is None:
            raise ValueError("PS channels are not set up under parameter server strategy")
        self._model_versions_from_ps = [-1 for _ in range(self._ps_client.ps_num)]
        self._timing = timing or _NoTiming()
        self._get_model_steps = getattr(args, "get_model_steps", 1) if args is not None else 1
        self._non_embed_grads = None
        self._non_embed_vars = {}
        self._evaluation_result = {}
        self._var_created = False
        self._model_version = -1
        self._init_embeddings()

    # ------------------------------------------------------------------ embeddings
    def _init_embeddings(self):
        self._embedding_layers = find_layer(self._model, Embedding)
        for layer in self._embedding_layers:
            layer.set_lookup_embedding_func(self._ps_client.pull_embedding_vectors)
        self._report_embedding_info()

    def _report_embedding_info(self):  # ps_trainer.py:186-214
        infos = [
            EmbeddingTableInfo(layer.embedding_weight_name, layer.output_dim, layer.embeddings_initializer,
                               DT_FLOAT, layer.input_dim)
            for layer in self._embedding_layers
        ]
        self._ps_client.push_embedding_table_infos(infos)

    def _set_tape_for_embedding(self, tape):
        for layer in self._embedding_layers:
            layer.set_tape(tape)

    def _reset_embedding(self):
        for layer in self._embedding_layers:
            layer.reset()

    # ------------------------------------------------------------------ model pull
    def _get_model(self):  # ps_trainer.py:149-184
        self._timing.start_record_time("get_model")
        dense_params, uninit_ps = self._ps_client.pull_dense_parameters(
            list(range(self._ps_client.ps_num)), self._model_versions_from_ps)
        if len(uninit_ps) > 0:
            for ps_id in uninit_ps:
                parameters = [Tensor(name, self._non_embed_vars[name].detach(), None)
                              for name in self._ps_client.ps_to_parameter[ps_id]]
                self._ps_client.push_dense_parameters(parameters, ps_id, self._model_versions_from_ps[ps_id])
            ps_params, uninit = self._ps_client.pull_dense_parameters(uninit_ps, self._model_versions_from_ps)
            if len(uninit) > 0:
                raise RuntimeError("PS initialization failed")
            dense_params.update(ps_params)
        with torch.no_grad():
            for k, v in dense_params.items():
                self._non_embed_vars[k].copy_(torch.as_tensor(v).reshape(self._non_embed_vars[k].shape))
        self._model_version = max(self._model_versions_from_ps)
        self._timing.end_record_time("get_model")

    def init_variables_if_need(self, features, labels=None):  # ps_trainer.py:304-342
        if self._var_created:
            return
        self._non_embed_vars = {name: p for name, p in self._model.named_parameters() if p.requires_grad}
        shapes = {name: tuple(p.shape) for name, p in self._non_embed_vars.items()}
        self._ps_client.partition_dense_parameters(self._non_embed_vars.keys(), shapes=shapes)
        self._var_created = True

    def get_trainable_items(self):
        bets = []
        for layer in self._embedding_layers:
            bets.extend(bet for (bet, _) in layer.embedding_and_ids)
        return list(self._non_embed_vars.values()) + bets

    # ------------------------------------------------------------------ train step
    def train_minibatch(self, features, labels, train_with_local_model=False):  # ps_trainer.py:371-385
        self.init_variables_if_need(features, labels)
        if not train_with_local_model:
            self._get_model()
        loss, grads = self._training_process_eagerly(features, labels)
        return (*self._update_global_model(grads), loss)

    def _training_process_eagerly(self, features, labels):  # ps_trainer.py:391-400
        self._set_tape_for_embedding(True)
        outputs = self._model(features)
        loss = self._loss(labels, outputs)
        grads = torch.autograd.grad(loss, self.get_trainable_items(), allow_unused=True)
        return loss.detach(), grads

    def _report_gradient(self, gradients):  # ps_trainer.py:239-280
        self._timing.start_record_time("report_gradient")
        grads = []
        names = list(self._non_embed_vars.keys())
        for i, name in enumerate(names):
            if gradients[i] is None:
                continue
            grads.append(Tensor(name, gradients[i], None))
        edl_grads = []
        bet_number = 0
        edl_embedding_grads = gradients[len(names):]
        for layer in self._embedding_layers:
            for i, (_, batch_ids) in enumerate(layer.embedding_and_ids):
                edl_grads.append(UniqueTensor(layer.embedding_weight_name,
                                              edl_embedding_grads[i + bet_number], batch_ids))
            bet_number += len(layer.embedding_and_ids)
        if len(edl_embedding_grads) != bet_number:
            raise ValueError("elasticdl.layers.embedding related gradient number %d does not match the "
                             "number of its output tensor %d." % (len(edl_embedding_grads), bet_number))
        learning_rate = float(self._optimizer.param_groups[0]["lr"])
        accepted, max_version = self._ps_client.push_gradients(
            grads, edl_grads, learning_rate, self._model_versions_from_ps)
        self._timing.end_record_time("report_gradient")
        return accepted, max_version

    def _update_global_model(self, grads):  # ps_trainer.py:408-414
        accepted, min_model_version = self._report_gradient(grads)
        if accepted and self._get_model_steps > 1:
            self._non_embed_grads = grads[: len(self._non_embed_vars)]
        self._reset_embedding()
        return accepted, min_model_version

    def _update_local_model(self):  # ps_trainer.py:139-147 (SSP local update of dense vars)
        if not self._non_embed_grads:
            return
        for p, g in zip(self._non_embed_vars.values(), self._non_embed_grads):
            p.grad = g
        self._optimizer.step()
        self._optimizer.zero_grad(set_to_none=True)
        self._non_embed_grads = None

    # ------------------------------------------------------------------ the worker's training loop
    def train_loop(self, batches, max_minibatch_retry_num=64, on_step=None):
        """The training loop of elasticdl/python/worker/worker.py:338-370 around train_minibatch, with its
        `get_model_steps` logic (SSP, docs/designs/async_sgd.md:129-157): the dense model is pulled from the
        PS only every `get_model_steps` minibatches (and after a failed one); in between the worker trains
        with its LOCAL model, which `_update_local_model` advances with the gradients it just pushed
        (dense variables only -- embedding rows are always pulled fresh).  A minibatch whose push is not
        accepted is retried up to `max_minibatch_retry_num` times (worker.py:181-234: "Worker got stuck").
        batches: iterable of (features, labels).  Returns [(version, loss)] per minibatch."""
        local_update_count = self._get_model_steps
        last_training_minibatch_failed = False
        out = []
        for features, labels in batches:
            if last_training_minibatch_failed or local_update_count >= self._get_model_steps:
                local_update_count = 0
                train_with_local_model = False
            else:
                train_with_local_model = True
            err_msg = ""
            try:  # _safe_process_minibatch (worker.py:274-300)
                for _ in range(max_minibatch_retry_num):
                    accepted, version, loss = self.train_minibatch(features, labels, train_with_local_model)
                    if accepted:
                        break
                else:
                    raise RuntimeError("Worker got stuck")
            except RuntimeError as err:
                err_msg = str(err)
            local_update_count += 1
            if err_msg:
                last_training_minibatch_failed = True
                out.append((None, None))
            else:
                last_training_minibatch_failed = False
                if local_update_count < self._get_model_steps:
                    self._update_local_model()
                out.append((version, loss))
            if on_step is not None:
                on_step(train_with_local_model, err_msg)
        return out

    # ------------------------------------------------------------------ eval / misc
    def get_model_version(self):
        return self._model_version

    def evaluate_minibatch(self, features, labels):
        with torch.no_grad():
            outputs = self._model(features)
        self._evaluation_result.setdefault("output", []).append(
            outputs if not isinstance(outputs, torch.Tensor) else outputs.cpu().numpy())
        self._evaluation_result.setdefault("label", []).append(np.asarray(torch.as_tensor(labels).cpu()))
        self._reset_embedding()

    def get_evaluation_result(self):
        return self._evaluation_result

    def reset_evaluation_result(self):
        self._evaluation_result = {}
