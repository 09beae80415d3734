"""Running mean / std for observation and return normalisation (reference: xuance/common/statistic_tools.py:117-185).
Host-side NumPy - it lives on the rollout side of the path; float64 accumulators merged with the
parallel-variance (Chan) update, as the reference does."""
import numpy as np


class RunningMeanStd:
    def __init__(self, shape, epsilon=1e-4):
        self.shape = shape
        if isinstance(shape, dict):
            self.mean = {k: np.zeros(s, np.float32) for k, s in shape.items()}
            self.var = {k: np.ones(s, np.float32) for k, s in shape.items()}
            self.count = {k: epsilon for k in shape}
        else:
            self.mean = np.zeros(shape, np.float32)
            self.var = np.ones(shape, np.float32)
            self.count = epsilon

    @property
    def std(self):
        if isinstance(self.shape, dict):
            return {k: np.sqrt(self.var[k]) for k in self.shape}
        return np.sqrt(self.var)

    def update(self, x):
        if isinstance(x, dict):
            for k in self.shape:
                self._merge(k, np.mean(x[k], axis=0), np.square(np.std(x[k], axis=0)), x[k].shape[0])
        else:
            self._merge(None, np.mean(x, axis=0), np.square(np.std(x, axis=0)), x.shape[0])

    def _merge(self, key, b_mean, b_var, b_count):
        mean = self.mean if key is None else self.mean[key]
        var = self.var if key is None else self.var[key]
        count = self.count if key is None else self.count[key]
        delta = b_mean - mean
        tot = count + b_count
        new_mean = mean + delta * b_count / tot
        m2 = var * count + b_var * b_count + np.square(delta) * count * b_count / tot
        if key is None:
            self.mean, self.var, self.count = new_mean, m2 / tot, tot
        else:
            self.mean[key], self.var[key], self.count[key] = new_mean, m2 / tot, tot
