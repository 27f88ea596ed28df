"""BaseAgent -- the agent interface of ScaleRL (/root/reference scalerl/algorithms/base.py:7-124).

When ScaleRL itself is importable the learner subclasses ITS ``BaseAgent`` (so ``isinstance(agent, BaseAgent)`` holds in
the host framework); otherwise an abstract class with the same methods and the same ``NotImplementedError`` defaults
stands in, so the package works stand-alone.
"""
from abc import ABCMeta
from typing import Any

try:
    from scalerl.algorithms.base import BaseAgent          # the host framework's own class
    HAVE_SCALERL = True
except Exception:                                           # noqa: BLE001 -- ScaleRL (or one of its imports) is absent
    HAVE_SCALERL = False

    class BaseAgent(metaclass=ABCMeta):
        """same surface as scalerl.algorithms.base.BaseAgent"""

        def __init__(self, args: Any) -> None:
            self.args = args

        def get_action(self, *args: Any, **kwargs: Any) -> Any:
            raise NotImplementedError

        def predict(self, *args: Any, **kwargs: Any) -> Any:
            raise NotImplementedError

        def get_value(self, *args: Any, **kwargs: Any) -> Any:
            raise NotImplementedError

        def learn(self, *args: Any, **kwargs: Any) -> Any:
            raise NotImplementedError

        def get_weights(self) -> dict:
            raise NotImplementedError('Subclasses should implement this method.')

        def set_weights(self, weights: dict) -> None:
            raise NotImplementedError('Subclasses should implement this method.')

        def save_checkpoint(self, path: str) -> None:
            raise NotImplementedError

        def load_checkpoint(self, path: str) -> None:
            raise NotImplementedError

        def name(self) -> str:
            return self.__class__.__name__.lower()
