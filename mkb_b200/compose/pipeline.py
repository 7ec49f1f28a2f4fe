"""Training loop; mirror of mkb/compose/pipeline.py (Pipeline).

Same constructor, same ``learn(model, dataset, sampling, optimizer, loss, evaluation=None)``, same
evaluation cadence / early stopping / printed output.  The difference is inside the batch loop: when
the (model, loss) pair is one this package knows how to fuse — any mkb_b200 model with
``losses.Adversarial`` — the three reference calls

    score = model(sample); negative_score = model(sample, negative_sample, mode); loss(...)

(pipeline.py:211, :230-232, :234) run as ONE forward kernel and ``error.backward()`` as ONE backward
kernel.  Anything else takes the generic three-call route, which still runs on the CUDA kernels.
"""
from __future__ import annotations

import collections

import torch

from .. import ops
from ..losses import Adversarial
from ..models.base import BaseModel
from ..optim import DenseAdam
from ..sampling import NegativeSampling
from ..utils import Bar
from .trainer import DeviceTrainer

__all__ = ["Pipeline"]


class _RollingMean:
    """river.stats.RollingMean(window) stand-in (pipeline.py:189): mean of the last `window` values."""

    def __init__(self, window_size):
        self._d = collections.deque(maxlen=window_size)

    def update(self, x):
        self._d.append(x)
        return self

    def get(self):
        return sum(self._d) / len(self._d) if self._d else 0.0


class Pipeline:
    LOSS_LAG = 3  # device-resident route: the loss of step k is read back (and logged) after step k + LOSS_LAG was queued

    def __init__(self, epochs, eval_every=2000, early_stopping_rounds=3, device="cpu", fused=True,
                 loss_every=1, trainer_options=None, adopt_torch_adam=True):
        """``fused=False`` forces the generic three-call loop.  ``loss_every`` > 1 reads the loss back
        to the host only every that many steps (the reference's ``error.item()`` at pipeline.py:242 is
        a device sync per step); 1 keeps the reference behaviour.  ``trainer_options`` are keyword
        arguments for the device-resident ``DeviceTrainer`` (e.g. ``{"mode": "rowshard"}`` to row-shard
        the entity table over the GPUs of the process group).  ``adopt_torch_adam=False`` keeps a stock
        ``torch.optim.Adam`` stepping through autograd instead of being taken over by the device-resident step."""
        self.trainer_options = dict(trainer_options or {})
        self.adopt_torch_adam = bool(adopt_torch_adam)
        self.epochs = epochs
        self.eval_every = eval_every
        self.early_stopping_rounds = early_stopping_rounds
        self.device = device
        self.fused = fused
        self.loss_every = max(1, int(loss_every))
        self.metric_loss = _RollingMean(1000)
        self.round_without_improvement_valid = 0
        self.round_without_improvement_test = 0
        self.history_valid = collections.defaultdict(float)
        self.history_test = collections.defaultdict(float)
        self.valid_scores = {}
        self.test_scores = {}

    def _can_fuse(self, model, loss):
        return self.fused and isinstance(model, BaseModel) and type(loss) is Adversarial

    @staticmethod
    def _plain_torch_adam(optimizer):
        """torch.optim.Adam exactly as the reference's quick-start builds it (README.md:123-126): the update the
        device-resident step applies (kge_adam_step is torch.optim.Adam's rule, tested against it)."""
        if type(optimizer) is not torch.optim.Adam or len(optimizer.param_groups) != 1:
            return False
        g = optimizer.param_groups[0]
        return (not g.get("amsgrad", False) and g.get("weight_decay", 0) == 0 and not g.get("maximize", False)
                and not g.get("capturable", False) and not g.get("differentiable", False)
                and not g.get("decoupled_weight_decay", False) and not torch.is_tensor(g["lr"]))

    def _device_trainer(self, model, dataset, sampling, optimizer, loss):
        """The whole step can stay on the device when every piece is known: fused-capable (model, loss), this
        package's sampler, and dense Adam over the model's parameters in one param group — ``optim.DenseAdam``
        or a stock ``torch.optim.Adam`` (``adopt_torch_adam``), whose hyper-parameters and moments are taken
        over and kept current in ``optimizer.state``."""
        ours = isinstance(optimizer, DenseAdam) or (self.adopt_torch_adam and self._plain_torch_adam(optimizer))
        if not (self._can_fuse(model, loss) and ours
                and isinstance(sampling, NegativeSampling) and len(optimizer.param_groups) == 1
                and model.entity_embedding.is_cuda):
            return None
        # any OTHER trainable parameter with a gradient would be skipped by the device step
        listed = [p for p in optimizer.param_groups[0]["params"]
                  if p is not model.entity_embedding and p is not model.relation_embedding]
        if any(p.grad is not None for p in listed):
            return None
        if model.kernel_modulus is not None and not isinstance(optimizer, DenseAdam):
            return None  # pRotatE's trainable modulus keeps its moments inside the trainer
        # the fused forward keeps the query and the K scores of a positive in one CTA's shared memory
        if (model.entity_dim + 3 + int(sampling.size)) * 4 > 200 * 1024:
            return None
        params = {id(p) for p in optimizer.param_groups[0]["params"]}
        if not {id(model.entity_embedding), id(model.relation_embedding)} <= params:
            return None
        import torch.distributed as dist

        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        key = (id(model), id(sampling), id(optimizer), float(loss.alpha), int(dataset.batch_size), multi)
        if getattr(self, "_trainer_key", None) != key:  # keep buffers across learn() calls
            # multi-rank: every batch of the dataset is the GLOBAL batch; rank r trains on its r-th slice
            per_rank = -(-int(dataset.batch_size) // dist.get_world_size()) if multi else int(dataset.batch_size)
            self._trainer = DeviceTrainer.from_optimizer(model, sampling, optimizer, alpha=loss.alpha,
                                                         max_batch=per_rank, distributed=multi,
                                                         **self.trainer_options)
            self._trainer_key = key
        # hyper-parameters are re-read on every learn() call and every epoch (_learn_on_device): an LR scheduler
        # or an edit of optimizer.param_groups[0] between epochs takes effect like it does in the reference
        self._trainer.adopt_hyper_parameters(optimizer)
        return self._trainer

    @staticmethod
    def _rank_slice(sample, weight, trainer):
        """Rank r's share of a global batch: the r-th block of ceil(B / world) rows.  Short blocks (B not a
        multiple of world, or the epoch's last batch) are padded with copies of row 0 at weight 0 — such rows
        add nothing to the loss sums or to any gradient (every term carries the positive's weight), and all
        ranks keep the same local batch size, which the column-parallel record exchange relies on."""
        world, rank = trainer.world, trainer.rank
        B = sample.shape[0]
        per = -(-B // world)
        lo, hi = min(rank * per, B), min((rank + 1) * per, B)
        s, w = sample[lo:hi], weight[lo:hi]
        if hi - lo < per:
            pad = per - (hi - lo)
            s = torch.cat([s, sample[:1].expand(pad, -1)])
            w = torch.cat([w, torch.zeros(pad, dtype=weight.dtype, device=weight.device)])
        return s.contiguous(), w.contiguous()

    def _learn_on_device(self, trainer, dataset, epoch, optimizer=None):
        """Device-resident epoch: per batch two async H2D copies (skipped when the dataset already
        lives on the GPU), five kernel launches, and an async D2H copy of the loss sums into pinned
        memory that is read ``LOSS_LAG`` steps later (the reference's ``error.item()`` without the stall: the host
        may run a few steps ahead of the device, which absorbs its own jitter — the data loader, tqdm, the GC)."""
        dev = trainer.dev
        depth = self.LOSS_LAG + 1
        host = [torch.empty(4, dtype=torch.float32).pin_memory() for _ in range(depth)]
        done = [torch.cuda.Event() for _ in range(depth)]
        pending = collections.deque()
        if optimizer is not None:
            trainer.adopt_hyper_parameters(optimizer)
        bar = Bar(dataset=dataset, update_every=10)
        for k, data in enumerate(bar):
            sample, weight = data["sample"], data["weight"]
            if trainer.distributed:
                sample, weight = self._rank_slice(sample, weight, trainer)
            sample = sample.to(dev, non_blocking=True)
            weight = weight.to(dev, non_blocking=True)
            stats = trainer.step(sample, weight, data["mode"])
            slot = k % depth
            host[slot].copy_(stats, non_blocking=True)
            done[slot].record()
            pending.append(slot)
            if len(pending) == depth:  # the slot the NEXT step will overwrite must have been read
                self._record_loss(pending.popleft(), host, done, bar, epoch)
        while pending:
            self._record_loss(pending.popleft(), host, done, bar, epoch)
        trainer.sync_optimizer_state()
        trainer.sync_model()  # rowshard: gather the trained shards back into model.entity_embedding

    def _record_loss(self, slot, host, done, bar, epoch):
        done[slot].synchronize()
        s = host[slot]
        self.metric_loss.update(float(-(s[0] + s[1]) / (2 * s[2])))
        bar.set_description(f"Epoch: {epoch}, loss: {self.metric_loss.get():4f}")

    def learn(self, model, dataset, sampling, optimizer, loss, evaluation=None):
        fuse = self._can_fuse(model, loss)
        trainer = self._device_trainer(model, dataset, sampling, optimizer, loss)
        step = 0
        for epoch in range(self.epochs):
            if trainer is not None:
                self._learn_on_device(trainer, dataset, epoch, optimizer)
            bar = Bar(dataset=dataset if trainer is None else [], update_every=10)
            for data in bar:
                sample = data["sample"].to(self.device)
                mode = data["mode"]
                weight = data["weight"].to(self.device)
                negative_sample = sampling.generate(sample=sample, mode=mode).to(self.device)
                if fuse:
                    try:
                        error = ops.fused_adversarial_step(
                            model.spec, model.entity_embedding, model.relation_embedding, sample, negative_sample,
                            weight, mode, loss.alpha, modulus=model.kernel_modulus)
                    except ops.N.KgeError as e:
                        if e.code != ops.N.E_UNSUPPORTED:
                            raise
                        fuse = False  # K too large for the fused kernel's shared memory: three-call route
                if not fuse:
                    score = model(sample)
                    negative_score = model(sample=sample, negative_sample=negative_sample, mode=mode)
                    error = loss(score, negative_score, weight)
                error.backward()
                _ = optimizer.step()
                optimizer.zero_grad()
                step += 1
                if step % self.loss_every == 0:
                    self.metric_loss.update(error.item())
                    bar.set_description(f"Epoch: {epoch}, loss: {self.metric_loss.get():4f}")

            if evaluation is not None and (epoch + 1) % self.eval_every == 0:
                print(f"\n Epoch: {epoch}.")
                if dataset.valid:
                    self.valid_scores = evaluation.eval(model=model, dataset=dataset.valid)
                    self.valid_scores.update(evaluation.eval_relations(model=model, dataset=dataset.valid))
                    self.print_metrics(description="Validation:", metrics=self.valid_scores)
                if dataset.test:
                    self.test_scores = evaluation.eval(model=model, dataset=dataset.test)
                    self.test_scores.update(evaluation.eval_relations(model=model, dataset=dataset.test))
                    self.print_metrics(description="Test:", metrics=self.test_scores)
                    if (self.history_test["HITS@3"] > self.test_scores["HITS@3"]
                            and self.history_test["HITS@1"] > self.test_scores["HITS@1"]):
                        self.round_without_improvement_test += 1
                    else:
                        self.round_without_improvement_test = 0
                        self.history_test = self.test_scores
                else:
                    if (self.history_valid["HITS@3"] > self.valid_scores["HITS@3"]
                            and self.history_valid["HITS@1"] > self.valid_scores["HITS@1"]):
                        self.round_without_improvement_valid += 1
                    else:
                        self.round_without_improvement_valid = 0
                        self.history_valid = self.valid_scores
                if (self.round_without_improvement_valid == self.early_stopping_rounds
                        or self.round_without_improvement_test == self.early_stopping_rounds):
                    print(f"\n Early stopping at epoch {epoch}.")
                    self.print_metrics(description="Validation:", metrics=self.valid_scores)
                    self.print_metrics(description="Test:", metrics=self.test_scores)
                    return self

        if evaluation is not None:  # the reference dereferences evaluation unguarded here (App. B.12)
            print(f"\n Epoch: {self.epochs - 1}. \n")
            if dataset.valid:
                self.valid_scores = evaluation.eval(model=model, dataset=dataset.valid)
                self.valid_scores.update(evaluation.eval_relations(model=model, dataset=dataset.valid))
                self.print_metrics(description="Validation:", metrics=self.valid_scores)
            if dataset.test:
                self.test_scores = evaluation.eval(model=model, dataset=dataset.test)
                self.test_scores.update(evaluation.eval_relations(model=model, dataset=dataset.test))
                self.print_metrics(description="Test:", metrics=self.test_scores)
        return self

    @classmethod
    def print_metrics(cls, description, metrics):
        print(f"\t {description}")
        for metric, value in metrics.items():
            print(f"\t\t {metric}: {value}")
