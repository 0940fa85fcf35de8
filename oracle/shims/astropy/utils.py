"""astropy.utils.NumpyRNGContext restated (analysis/mass_estimator.py:180)."""
from numpy import random


class NumpyRNGContext:
    def __init__(self, seed):
        self.seed = seed

    def __enter__(self):
        self.startstate = random.get_state()
        random.seed(self.seed)

    def __exit__(self, exc_type, exc_value, traceback):
        random.set_state(self.startstate)
