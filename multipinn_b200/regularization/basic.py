from abc import ABC, abstractmethod


class BasicLosses(ABC):
    """``obj(trainer) -> (total_loss, losses[K])`` (reference ``regularization/basic.py:4-37``)."""

    @abstractmethod
    def __call__(self, trainer):
        raise NotImplementedError
