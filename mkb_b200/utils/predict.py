"""Batch scoring of a list of triples; mirror of mkb/utils/predict.py (FetchToPredict, make_prediction)."""
import numpy as np
import torch

__all__ = ["FetchToPredict", "make_prediction"]


class FetchToPredict:
    """Iterate a list of triples as int64 ``[b, 3]`` batches in order (predict.py:11-58), without the
    reference's DataLoader worker process."""

    def __init__(self, dataset, batch_size, num_workers=1):
        self.dataset = dataset
        self.batch_size = batch_size
        self.num_workers = num_workers

    def __getitem__(self, idx):
        return torch.LongTensor(self.dataset[idx])

    def __len__(self):
        return len(self.dataset)

    def __iter__(self):
        arr = torch.from_numpy(np.asarray(self.dataset, dtype=np.int64).reshape(-1, 3))
        for lo in range(0, arr.shape[0], self.batch_size):
            yield arr[lo:lo + self.batch_size]

    @staticmethod
    def collate_fn(data):
        return torch.stack(data, dim=0)


def make_prediction(model, dataset, batch_size, num_workers=1, device="cuda"):
    """Scores of ``dataset`` (list of triples) under ``model`` as a flat tensor (predict.py:61-106).
    The kernels run on the model's CUDA device whatever ``device`` says."""
    dev = model.entity_embedding.device
    with torch.no_grad():
        parts = [model(x.to(dev)) for x in FetchToPredict(dataset=dataset, batch_size=batch_size,
                                                          num_workers=num_workers)]
        return torch.cat(parts).flatten() if parts else torch.zeros(0, device=dev)
