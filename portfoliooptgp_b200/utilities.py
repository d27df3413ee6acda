"""gpflow.utilities subset used by the reference: set_trainable, print_summary, deepcopy,
positive / triangular, to_default_float, parameter_dict, read_values."""
from __future__ import annotations

import copy
from typing import Dict

import numpy as np

from .base import Module, Parameter, positive, set_trainable, triangular  # noqa: F401


def deepcopy(obj, memo=None):
    """gpflow.utilities.deepcopy (Multi-Input_GPR/main.py:176): plain deep copy; models drop their
    engine handle and compiled kernel, which are rebuilt lazily."""
    return copy.deepcopy(obj, memo)


def to_default_float(x):
    return np.asarray(x, dtype=np.float64)


def parameter_dict(module: Module) -> Dict[str, Parameter]:
    return {"." + name: p for name, p in module.named_parameters()}


def read_values(module: Module) -> Dict[str, np.ndarray]:
    return {k: v.numpy() for k, v in parameter_dict(module).items()}


def _fmt_value(v: np.ndarray) -> str:
    v = np.asarray(v)
    if v.ndim == 0:
        return f"{float(v):.5g}"
    flat = v.reshape(-1)
    body = ", ".join(f"{x:.5g}" for x in flat[:3])
    return f"[{body}{'...' if flat.size > 3 else ''}]"


def tabulate_module_summary(module: Module, fmt: str = None) -> str:
    rows = []
    for name, p in module.named_parameters():
        rows.append([f"{type(module).__name__}.{name}", "Parameter", p.transform.name, "" if p.prior is None else str(p.prior),
                     str(p.trainable), str(tuple(p.shape)), "float64", _fmt_value(p.numpy())])
    headers = ["name", "class", "transform", "prior", "trainable", "shape", "dtype", "value"]
    try:
        from tabulate import tabulate
        tf = {"notebook": "html", None: "fancy_grid"}.get(fmt, fmt)
        return tabulate(rows, headers=headers, tablefmt=tf)
    except Exception:
        lines = ["  ".join(headers)] + ["  ".join(r) for r in rows]
        return "\n".join(lines)


def print_summary(module: Module, fmt: str = None) -> None:
    """gpflow.utilities.print_summary (GPR/main.py:40, models/model_trainer.py:22,51)."""
    print(tabulate_module_summary(module, "simple" if fmt == "notebook" else fmt))
