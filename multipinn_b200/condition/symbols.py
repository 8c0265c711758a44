"""String DSL for derivative symbols, same surface as the reference's ``Symbols``
(reference ``multipinn/condition/symbols.py:9-257``): ``Symbols("x, t", "u")("u, u_x, u_t, u_xx")``
returns ``compute_gradients(model, arg)``.  In the automatic-differentiation mode (``"grad"``) the generated
function is a chain of ``unpack(grad(parent, arg))`` lines, i.e. exactly what a hand-written residual does, so
nothing else is needed for the fused kernels (and the residual tracer sees through it).  The finite-difference
mode (``"diff"``, reference symbols.py:160-234) generates calls of ``num_diff_random`` / ``num_diff_second_same`` /
``num_diff_second_cross``: value-only model calls at shifted points, each one fused forward launch.
"""
import re
from typing import Callable, List

import torch

from . import diff as _diff


class Symbols:
    def __init__(self, variables: str, functions: str, mode: str = "grad"):
        self.variables = self._split(variables)
        self.functions = self._split(functions)
        self.check_names()
        self.input_dim = len(self.variables)
        self.output_dim = len(self.functions)
        self.mode = mode

    @staticmethod
    def _split(text: str) -> List[str]:
        return text.replace(" ", "").rstrip(",").split(",")

    def check_names(self):
        names = self.variables + self.functions
        for name in names:
            if not re.fullmatch(r"[a-zA-Z]", name):
                raise ValueError(f"Names must be a single letter [a-zA-Z], but found {name}")
        if len(set(names)) != len(names):
            raise ValueError("Duplicate names in variables or functions")

    def parsing(self, output_format: str) -> List[str]:
        symbols = self._split(output_format)
        self.check_format(symbols)
        return symbols

    def check_format(self, derivatives: List[str]) -> None:
        symbols = derivatives
        if symbols == [""]:
            raise ValueError("No reason to call this function with zero symbols")
        for s in symbols:
            if len(s) == 1:
                if s not in self.variables and s not in self.functions:
                    raise ValueError(f"Symbol '{s}' is not a defined function or variable.")
                continue
            if s.count("_") != 1 or s[1] != "_":
                raise ValueError(f"Invalid symbol format: '{s}'. Expected format is 'u_xyz' or 'u' or 'x'.")
            fn, wrt = s.split("_")
            if not wrt:
                raise ValueError(f"Invalid symbol format: '{s}'. Variable part cannot be empty.")
            if fn not in self.functions:
                raise ValueError(f"Invalid symbol format: '{s}'. Function '{fn}' is not defined.")
            if any(c not in self.variables for c in wrt):
                raise ValueError(f"Invalid symbol format: '{s}'. Variable '{wrt}' contains undefined variables.")

    @staticmethod
    def _tuple(names: List[str]) -> str:
        return f"({names[0]},)" if len(names) == 1 else ", ".join(names)

    def _children(self, symbol: str) -> List[str]:
        sep = "" if "_" in symbol else "_"
        return [f"{symbol}{sep}{v}" for v in self.variables]

    def _one_hot(self, var: str) -> str:
        hot = ["1" if v == var else "0" for v in self.variables]
        return "torch.Tensor([[" + ", ".join(hot) + "]])"

    def generate_str_diff(self, derivatives: List[str]) -> str:
        """Finite-difference code: first derivatives by random-step forward differences along one axis, second
        derivatives by central / four-point differences (reference symbols.py:189-234)."""
        by_order = sorted(derivatives, key=lambda s: len(s.split("_")[1]) if "_" in s else 0)
        lines = ["def compute_gradients(model, arg):"]
        if any(s[0] in self.functions for s in by_order):
            pure = any(s in self.functions for s in by_order)
            non_cross = any(len(s) == 3 or (len(s) == 4 and s[-2] == s[-1]) for s in by_order)
            if pure or non_cross:
                lines.append("\tfns = model(arg)")
            if pure:
                lines.append("\t" + self._tuple(self.functions) + " = unpack(fns)")
        if any(s in self.variables for s in by_order):
            lines.append("\t" + self._tuple(self.variables) + " = unpack(arg)")
        known = set(self.variables + self.functions)
        for s in by_order:
            if s in known:
                continue
            if len(s) == 3:       # u_x: every function's derivative along x in one call
                kids = [f"{fn}_{s[-1]}" for fn in self.functions]
                lines.append(f"\t{self._tuple(kids)} = unpack(num_diff_random(model, arg, fns, {self._one_hot(s[-1])}))")
            elif len(s) == 4:     # u_xy
                a, b = s[-2], s[-1]
                kids = [f"{fn}_{a}{b}" for fn in self.functions]
                if a == b:
                    lines.append(f"\t{self._tuple(kids)} = unpack(num_diff_second_same(model, arg, fns, {self._one_hot(a)}))")
                else:
                    lines.append(f"\t{self._tuple(kids)} = unpack(num_diff_second_cross(model, arg, "
                                 f"{self._one_hot(a)}, {self._one_hot(b)}))")
            else:
                raise ValueError(f"'diff' mode doesn't support derivatives of order > 2, but {s} found")
            known.update(kids)
        lines.append("\treturn " + self._tuple(derivatives))
        return "\n".join(lines)

    def generate_str(self, derivatives: List[str]) -> str:
        if self.mode == "diff":
            return self.generate_str_diff(derivatives)
        if self.mode != "grad":
            raise ValueError(f"Unknown mode {self.mode}")
        by_order = sorted(derivatives, key=lambda s: len(s.split("_")[1]) if "_" in s else 0)
        lines = ["def compute_gradients(model, arg):"]
        if any(s[0] in self.functions for s in by_order):
            lines.append("\t" + self._tuple(self.functions) + " = unpack(model(arg))")
        if any(s in self.variables for s in by_order):
            lines.append("\t" + self._tuple(self.variables) + " = unpack(arg)")
        known = set(self.variables + self.functions)

        def ensure(symbol: str):
            if symbol in known:
                return
            parent = symbol[:-1].rstrip("_")
            ensure(parent)
            kids = self._children(parent)
            lines.append(f"\t{self._tuple(kids)} = unpack(grad({parent}, arg))")
            known.update(kids)

        for s in by_order:
            ensure(s)
        lines.append("\treturn " + self._tuple(derivatives))
        return "\n".join(lines)

    def __call__(self, output_format: str) -> Callable:
        derivatives = self.parsing(output_format)
        source = self.generate_str(derivatives)
        namespace = {"torch": torch, "unpack": _diff.unpack, "grad": _diff.grad,
                     "num_diff_random": _diff.num_diff_random, "num_diff_second_same": _diff.num_diff_second_same,
                     "num_diff_second_cross": _diff.num_diff_second_cross}
        exec(source, namespace)
        fn = namespace["compute_gradients"]
        fn.source = source
        return fn
