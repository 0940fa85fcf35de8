"""astropy.stats.bootstrap restated (analysis/mass_estimator.py:181-183).

bootnum draws of ``np.random.randint(0, n, size=samples)`` from the legacy global RNG, each
reduced by ``bootfunc``; ``bootfunc(data)`` is called once first only to size the result.
"""
import numpy as np


def bootstrap(data, bootnum=100, samples=None, bootfunc=None):
    if samples is None:
        samples = data.shape[0]
    assert samples > 0, "samples cannot be less than one"
    assert bootnum > 0, "bootnum cannot be less than one"
    if bootfunc is None:
        resultdims = (bootnum,) + (samples,) + data.shape[1:]
    else:
        try:
            resultdims = (bootnum, len(bootfunc(data)))
        except TypeError:
            resultdims = (bootnum,)
    boot = np.empty(resultdims)
    for i in range(bootnum):
        bootarr = np.random.randint(low=0, high=data.shape[0], size=samples)
        if bootfunc is None:
            boot[i] = data[bootarr]
        else:
            boot[i] = bootfunc(data[bootarr])
    return boot
