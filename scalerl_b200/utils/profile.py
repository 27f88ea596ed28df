"""Timings -- the wall-clock section timer ImpalaTrainer threads through get_batch / learn_process / get_action
(/root/reference scalerl/utils/profile.py:10-65; sections at impala_atari.py:245-266,375-386) -- plus NVTX ranges so the
same sections show up in Nsight timelines of the B200 learner.

Same surface as the reference class (``reset``, ``time(name)``, ``means``, ``vars``, ``stds``, ``summary(prefix)``): a
``Timings`` built here can be handed to reference code and vice versa.  Mean / variance are kept online with Welford's
update (count, mean, M2), so no per-sample list is stored.
"""
import contextlib
import timeit
from typing import Dict

try:                                    # NVTX is optional: only present with a CUDA build of torch
    from torch.cuda import nvtx as _nvtx
except Exception:                       # noqa: BLE001
    _nvtx = None

_NVTX_ON = False


def enable_nvtx(on: bool = True) -> None:
    """switch NVTX range emission of Timings.time / nvtx_range on or off (off by default: zero overhead)"""
    global _NVTX_ON
    _NVTX_ON = bool(on) and _nvtx is not None


@contextlib.contextmanager
def nvtx_range(name: str):
    """``with nvtx_range('learn'):`` -- an NVTX push/pop pair when enabled, otherwise nothing"""
    if _NVTX_ON:
        _nvtx.range_push(name)
        try:
            yield
        finally:
            _nvtx.range_pop()
    else:
        yield


class Timings:
    """Not thread-safe (as the reference's)."""

    def __init__(self):
        self._count: Dict[str, int] = {}
        self._mean: Dict[str, float] = {}
        self._m2: Dict[str, float] = {}
        self.reset()

    def reset(self):
        self.last_time = timeit.default_timer()

    def time(self, name: str):
        """close the section that started at the previous ``time``/``reset`` call and file it under ``name``"""
        now = timeit.default_timer()
        x = now - self.last_time
        self.last_time = now
        n = self._count.get(name, 0) + 1
        mean = self._mean.get(name, 0.0)
        delta = x - mean
        mean += delta / n
        self._count[name] = n
        self._mean[name] = mean
        self._m2[name] = self._m2.get(name, 0.0) + delta * (x - mean)
        if _NVTX_ON:
            _nvtx.mark(f'{name} {x * 1e3:.3f} ms')

    def means(self):
        return dict(self._mean)

    def vars(self):
        """population variance per section (what the reference's online update converges to)"""
        return {k: self._m2[k] / self._count[k] for k in self._mean}

    def stds(self):
        return {k: v ** 0.5 for k, v in self.vars().items()}

    def summary(self, prefix: str = ''):
        means, stds = self.means(), self.stds()
        total = sum(means.values()) or 1e-30
        out = prefix
        for k in sorted(means, key=means.get, reverse=True):
            out += f'\n    {k}: {1000 * means[k]:.6f}ms +- {1000 * stds[k]:.6f}ms ({100 * means[k] / total:.2f}%) '
        out += f'\nTotal: {1000 * total:.6f}ms'
        return out
