"""``KdmkbModel`` — several knowledge bases trained together, each model distilling from the others; mirror of
mkb/distillation/kdmkb_model.py:152-560.

One step (``forward``, :286-360) per KB: the self-adversarial loss of its own batch scaled by ``1 - w_kl`` plus,
from every other KB taken as teacher, ``Distillation.distill`` on the teacher's batch scaled by ``w_kl``; then
``backward`` / ``optimizer.step`` per KB.  Every numeric piece runs on this package's kernels: the fused
sampler-to-loss step (``kge_sample_negatives`` / ``kge_filter_pool``, ``kge_fused_fwd`` / ``kge_fused_bwd``), the
3-D score path and the KL kernels inside ``distill``, the pre-computed top-k tables (``FastTopKSampling``:
``kge_score_fwd`` + ``kge_topk_rows``), ``Evaluation`` on the rank kernels.  The driver around them is host glue
with the reference's constructor, ``forward``, ``learn`` and printed metrics.

Not provided: ``classification`` datasets (BCE / ConvE mode, outside the KGE hot path).
"""
from __future__ import annotations

import collections
import os

import numpy as np
import torch

from .. import ops
from ..compose.pipeline import _RollingMean
from ..evaluation import Evaluation
from ..losses import Adversarial
from ..sampling import NegativeSampling
from ..utils import BarRange
from .distillation import Distillation
from .top_k_sampling import FastTopKSampling

__all__ = ["KdmkbModel"]


class KdmkbModel:
    def __init__(self, models, datasets, lr, alpha_kl, alpha_adv, negative_sampling_size, batch_size_entity,
                 batch_size_relation, n_random_entities, n_random_relations, update_distillation_every=500,
                 device="cuda", seed=None, warm_step=500, pool="independent"):
        """Arguments as the reference (:152-167), each a dict keyed like ``models`` / ``datasets``.  ``pool`` is this
        package's sampler switch ("reference" reproduces mkb's shared-pool draws)."""
        self.alpha_kl = alpha_kl
        self.batch_size_entity = batch_size_entity
        self.batch_size_relation = batch_size_relation
        self.n_random_entities = n_random_entities
        self.n_random_relations = n_random_relations
        self.update_distillation_every = update_distillation_every
        self.device = device
        self.seed = seed
        self._rng = np.random.RandomState(self.seed)
        self.warm_step = warm_step
        self.loss_function = collections.OrderedDict()
        for key, dataset in datasets.items():
            if dataset.classification:
                raise NotImplementedError("classification datasets (BCE) are outside the KGE hot path")
            self.loss_function[key] = Adversarial(alpha=alpha_adv[key])
        self.optimizers = collections.OrderedDict(
            (key, torch.optim.Adam(filter(lambda p: p.requires_grad, models[key].parameters()), lr=rate))
            for key, rate in lr.items())
        self.distillation = collections.OrderedDict()
        self._refresh_distillation(models, datasets)
        self.negative_sampling = collections.OrderedDict()
        self.validation = collections.OrderedDict()
        for key, dataset in datasets.items():
            self.negative_sampling[key] = NegativeSampling(
                size=negative_sampling_size[key], entities=dataset.entities, relations=dataset.relations,
                train_triples=dataset.train_triples, seed=seed, pool=pool)
            self.validation[key] = Evaluation(entities=dataset.entities, relations=dataset.relations, batch_size=2,
                                              true_triples=dataset.true_triples, device=device)
        self.metrics = {key: _RollingMean(1000) for key in datasets}

    def _refresh_distillation(self, models, datasets):
        """One ``Distillation`` per ordered (teacher, student) pair over ``FastTopKSampling`` tables of the CURRENT
        teacher (:200-219, rebuilt every ``update_distillation_every`` steps, :404-432)."""
        for t_key, t_data in datasets.items():
            for s_key, s_data in datasets.items():
                if t_key == s_key:
                    continue
                self.distillation[f"{t_key}_{s_key}"] = self._init_distillation(
                    sampling_method=FastTopKSampling, teacher=models[t_key], dataset_teacher=t_data,
                    dataset_student=s_data, batch_size_entity=self.batch_size_entity[t_key],
                    batch_size_relation=self.batch_size_relation[t_key],
                    n_random_entities=self.n_random_entities[t_key],
                    n_random_relations=self.n_random_relations[t_key], seed=self.seed, device=self.device)

    @classmethod
    def _init_distillation(cls, sampling_method, teacher, dataset_teacher, dataset_student, batch_size_entity,
                           batch_size_relation, n_random_entities, n_random_relations, seed, device):
        sampler = sampling_method(
            teacher=teacher, dataset_teacher=dataset_teacher, teacher_relations=dataset_teacher.relations,
            teacher_entities=dataset_teacher.entities, student_entities=dataset_student.entities,
            student_relations=dataset_student.relations, batch_size_entity=batch_size_entity,
            batch_size_relation=batch_size_relation, n_random_entities=n_random_entities,
            n_random_relations=n_random_relations, seed=seed, device=device)
        return Distillation(teacher_entities=dataset_teacher.entities, teacher_relations=dataset_teacher.relations,
                            student_entities=dataset_student.entities, student_relations=dataset_student.relations,
                            sampling=sampler, device=device)

    def forward(self, datasets, models, weight_kl):
        loss_models, samples = collections.OrderedDict(), collections.OrderedDict()
        key = None
        for key, dataset in datasets.items():
            data = next(dataset)
            model = models[key]
            dev = model.entity_embedding.device
            sample, mode = data["sample"].to(dev), data["mode"]
            weight = data["weight"].to(dev)
            negative_sample = self.negative_sampling[key].generate(sample=sample, mode=mode).to(dev)
            # model(sample), model(sample, negative_sample, mode) and the loss (:298-326) as the fused step
            loss = self.loss_function[key]
            try:
                error = ops.fused_adversarial_step(model.spec, model.entity_embedding, model.relation_embedding, sample,
                                                   negative_sample, weight, mode, loss.alpha,
                                                   modulus=model.kernel_modulus)
            except ops.N.KgeError as e:  # K beyond the fused kernel's shared memory: the three-call route
                if e.code != ops.N.E_UNSUPPORTED:
                    raise
                error = loss(model(sample), model(sample, negative_sample, mode=mode), weight)
            loss_models[key] = error * (1 - weight_kl[key])
            samples[key] = sample
        last = key  # the reference scales EVERY distillation term by the weight of the last dataset (:341-349)
        for t_key in datasets:
            for s_key in datasets:
                if t_key != s_key:
                    loss_models[s_key] = loss_models[s_key] + self.distillation[f"{t_key}_{s_key}"].distill(
                        teacher=models[t_key], student=models[s_key], sample=samples[t_key]) * weight_kl[last]
        for key in datasets:
            loss_models[key].backward()
            self.optimizers[key].step()
            self.optimizers[key].zero_grad()
            self.metrics[key].update(loss_models[key].item())
        return self.metrics

    def learn(self, models, datasets, max_step, eval_every=2000, update_every=10, log_dir=None, save_path=None):
        scores = []
        bar = BarRange(step=max_step, update_every=update_every)
        for step in bar:
            weight_kl = {k: 0 for k in datasets} if step < self.warm_step else dict(self.alpha_kl)
            metrics = self.forward(datasets, models, weight_kl)
            bar.set_description(text=", ".join(f"{k}: {v.get():4f}" for k, v in metrics.items()))
            if (step + 1) % self.update_distillation_every == 0:
                self._refresh_distillation(models, datasets)
            if (step + 1) % eval_every == 0:
                for key, dataset in datasets.items():
                    model = models[key].eval()
                    ev = self.validation[key]
                    valid = ev.eval(model=model, dataset=dataset.valid)
                    valid.update(ev.eval_relations(model=model, dataset=dataset.valid))
                    valid = collections.OrderedDict((f"valid_{m}", s) for m, s in valid.items())
                    test = ev.eval(model=model, dataset=dataset.test)
                    test.update(ev.eval_relations(model=model, dataset=dataset.test))
                    test = collections.OrderedDict((f"test_{m}", s) for m, s in test.items())
                    model.train()
                    print(f"\n Model: {key}, step {step}")
                    self.print_metrics(description="Validation:", metrics=valid)
                    self.print_metrics(description="Test:", metrics=test)
                    if log_dir is not None:
                        import pandas as pd

                        row = {"dataset": dataset.name, "id_dataset": key, "model_name": model.name, "step": step,
                               "alpha_kl": self.alpha_kl[key], "alpha_adv": self.loss_function[key].alpha,
                               "batch_size": dataset.batch_size,
                               "negative_sample_size": self.negative_sampling[key].size,
                               "hidden_dim": model.hidden_dim, "gamma": model.gamma.item(), **valid, **test}
                        scores.append(row)
                        pd.DataFrame(scores).to_csv(log_dir, index=False)
                    if save_path is not None:
                        name = f"kdmkb_{dataset.name}_{key}_{model.name}_{step}.pickle"
                        dev = model.entity_embedding.device
                        model.save(path=os.path.join(save_path, name))  # save() moves the model to the host
                        models[key] = model.to(dev).train()
        return self

    @classmethod
    def print_metrics(cls, description, metrics):
        print(f"\t {description}")
        for metric, value in metrics.items():
            print(f"\t\t {metric}: {value}")
