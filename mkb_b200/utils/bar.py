"""tqdm wrapper with the reference's interface (mkb/utils/bar.py:6-35)."""
import tqdm

__all__ = ["Bar", "BarRange"]


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


class BarRange(Bar):
    """``BarRange(step, update_every, position=0)``: the same bar over ``range(step)`` (mkb/utils/bar.py:38-69)."""

    def __init__(self, step, update_every=1, position=0):
        super().__init__(range(step), update_every=update_every, position=position)
