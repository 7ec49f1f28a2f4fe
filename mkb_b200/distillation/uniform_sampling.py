"""``UniformSampling`` — mirror of mkb/distillation/uniform_sampling.py:77-148: ONE random subset of the shared
entities and of the shared relations per call, repeated for every positive of the batch; supervised (the
distillation loop writes the ground truth into the last slot).  Pure index work: the same seeded
``numpy.RandomState`` calls as the reference, int64 tensors instead of its float ones."""
import numpy as np
import torch

__all__ = ["UniformSampling"]


class UniformSampling:
    def __init__(self, batch_size_entity, batch_size_relation, seed=None, **kwargs):
        self.batch_size_entity = batch_size_entity
        self.batch_size_relation = batch_size_relation
        self._rng = np.random.RandomState(seed)

    @property
    def supervised(self):
        """Include the ground truth."""
        return True

    def get(self, mapping_entities, mapping_relations, positive_sample_size, **kwargs):
        ent_t = self._rng.choice(a=list(mapping_entities.keys()), size=self.batch_size_entity, replace=False)
        rel_t = self._rng.choice(a=list(mapping_relations.keys()), size=self.batch_size_relation, replace=False)
        ent_s = [mapping_entities[e] for e in ent_t]
        rel_s = [mapping_relations[r] for r in rel_t]
        n = int(positive_sample_size)

        def rep(x):
            return torch.tensor(np.asarray(x, dtype=np.int64)).view(1, -1).repeat(n, 1)

        return rep(ent_t), rep(rel_t), rep(ent_t), rep(ent_s), rep(rel_s), rep(ent_s)
