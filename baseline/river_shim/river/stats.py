import collections


class Mean:
    """Running arithmetic mean."""

    def __init__(self):
        self.n, self.m = 0, 0.0

    def update(self, x):
        self.n += 1
        self.m += (x - self.m) / self.n
        return self

    def get(self):
        return self.m


class RollingMean:
    """Mean of the last ``window_size`` values."""

    def __init__(self, window_size):
        self.d = collections.deque(maxlen=window_size)

    def update(self, x):
        self.d.append(x)
        return self

    def get(self):
        return sum(self.d) / len(self.d) if self.d else 0.0
