from .PINN import PINN, sum_squares

__all__ = ["PINN", "sum_squares"]
