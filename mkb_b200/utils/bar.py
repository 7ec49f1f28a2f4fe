"""tqdm wrapper with the reference's interface (mkb/utils/bar.py:6-35)."""
import tqdm

__all__ = ["Bar"]


class Bar:
    def __init__(self, dataset, update_every=1, position=0):
        self.bar = tqdm.tqdm(dataset, position=position)
        self.update_every = update_every
        self.n = 0

    def set_description(self, text):
        if self.n % self.update_every == 0:
            self.bar.set_description(text)

    def __iter__(self):
        for x in self.bar:
            self.n += 1
            yield x
