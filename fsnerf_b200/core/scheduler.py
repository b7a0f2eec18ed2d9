"""Learning-rate schedules of the train-step arithmetic.

Mirrors the interface of the reference's ``core.scheduler`` (``Constant`` and
``ExponentialDecay(optim, T, lro, r=...)`` with ``.step()`` / ``.lr``,
/root/reference/src/core/scheduler.py:6-80) so ``run-nerf.py:216-223,285`` works
unchanged, and exposes the closed form (``lr_at``) that the fused driver
(``engine.HotPath.train_step(lr=...)``) consumes without an optimizer object.
"""
from typing import Optional


def lr_at(t: int, T: int, lro: float, r: Optional[float] = None) -> float:
    """learning rate after ``t`` scheduler steps: lro * r**(t/T) for t < T, lro*r after;
    r=None is the constant schedule (reference: scheduler.py:42-50,74-80)."""
    if lro < 0:
        raise ValueError("lro must be a positive value.")
    if r is None:
        return lro
    return lro * (r ** (t / T)) if t < T else lro * r


class _Schedule:
    _r: Optional[float] = None

    def __init__(self, optim, T: int, lro: float, **kwargs):
        if lro < 0:
            raise ValueError("lro must be a positive value.")
        self.optim, self.T, self.lro, self.t = optim, T, lro, 0

    @property
    def lr(self) -> float:
        return lr_at(self.t, self.T, self.lro, self._r)

    def step(self) -> None:
        self.t += 1
        if self.optim is not None:
            for group in self.optim.param_groups:
                group["lr"] = self.lr


class Constant(_Schedule):
    pass


class ExponentialDecay(_Schedule):
    def __init__(self, optim, T: int, lro: float, **kwargs):
        super().__init__(optim, T, lro)
        self._r = self.r = kwargs["r"]
        self.lrf = lro * self.r
