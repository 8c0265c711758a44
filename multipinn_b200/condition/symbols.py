"""String DSL for derivative symbols, same surface as the reference's ``Symbols``
(reference ``multipinn/condition/symbols.py:9-257``): ``Symbols("x, t", "u")("u, u_x, u_t, u_xx")``
returns ``compute_gradients(model, arg)``.  Only the automatic-differentiation mode (``"grad"``) is
lowered onto the jet path: the generated function is a chain of ``unpack(grad(parent, arg))`` lines,
i.e. exactly what a hand-written residual does, so nothing else is needed for the fused kernels.
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

    def generate_str(self, derivatives: List[str]) -> str:
        if self.mode != "grad":
            raise NotImplementedError(
                "Symbols(mode='diff') (finite differences) is not lowered onto the jet path; "
                "call the num_diff* helpers directly")
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
        namespace = {"torch": torch, "unpack": _diff.unpack, "grad": _diff.grad}
        exec(source, namespace)
        fn = namespace["compute_gradients"]
        fn.source = source
        return fn
