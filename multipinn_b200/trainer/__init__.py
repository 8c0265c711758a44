from .trainer import Trainer, TrainerMultiGPU, TrainerOneBatch

__all__ = ["Trainer", "TrainerOneBatch", "TrainerMultiGPU"]
