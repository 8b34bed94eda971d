"""Golden vectors for the dataset generator, produced by the REAL reference functions
(python-src/diffusion_training.py: generate_random_permittivity :54-93, generate_random_source :96-146) run in
the authoring container.  The module itself cannot be imported (it needs diffusers / matplotlib), so the two
function definitions are compiled from the reference source where it lies and executed with a `torch` shim whose
`rand` / `randint` hand out scripted draws -- the same draws the restatement is then fed.  Writes
tests/golden/dataset.npz.   usage: python -m oracle.make_golden_dataset
"""
from __future__ import annotations

import ast
import os
import sys
from typing import Tuple

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import dataset_oracle as do  # noqa: E402

REF = os.environ.get("FDTD2D_REFERENCE_ROOT", "/root/reference") + "/python-src/diffusion_training.py"


class TorchShim:
    """torch with scripted random draws: rand(shape...) pops from `fields` (full maps) or `scalars`."""

    def __init__(self, fields=(), scalars=(), ints=()):
        self.fields, self.scalars, self.ints = list(fields), list(scalars), list(ints)

    def rand(self, *shape, **kw):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        if shape == (1,):
            return torch.tensor([self.scalars.pop(0)], dtype=torch.float64)
        return torch.from_numpy(self.fields.pop(0).copy())

    def randint(self, lo, hi, size):
        v = self.ints.pop(0)
        assert lo <= v < hi, (lo, v, hi)
        return torch.tensor([v])

    def __getattr__(self, name):
        return getattr(torch, name)


def reference_functions(shim):
    tree = ast.parse(open(REF).read())
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("generate_random_permittivity", "generate_random_source")]
    ns = {"torch": shim, "F": F, "Tuple": Tuple, "device": torch.device("cpu")}
    exec(compile(ast.Module(body=keep, type_ignores=[]), REF, "exec"), ns)
    return ns["generate_random_permittivity"], ns["generate_random_source"]


def main():
    out = {}
    seed = 2026
    for g, (R, C) in enumerate([(64, 64), (60, 100), (256, 256)]):
        u = do.uniform_field(seed, g, R, C)
        sigma = do.sigma_of(seed, g)
        shim = TorchShim(fields=[u], scalars=[(sigma - 2.0) / 4.0])
        gen_eps, _ = reference_functions(shim)
        eps, mu = gen_eps((R, C), device=torch.device("cpu"))
        out[f"eps_{g}"], out[f"mu_{g}"] = eps.numpy(), mu.numpy()
        out[f"shape_{g}"] = np.array([R, C])
    # sources: scripted draws covering the three branches
    scripts = [((256, 256), [0.1, 0.2], [100, 40]), ((256, 256), [0.3, 0.9], [77, 150]), ((256, 256), [0.7], [200, 31]),
               ((60, 100), [0.4, 0.1], [30, 20]), ((60, 100), [0.2, 0.6], [50, 10]), ((64, 64), [0.99], [6, 57])]
    for i, (dim, scalars, ints) in enumerate(scripts):
        shim = TorchShim(scalars=list(scalars), ints=list(ints))
        _, gen_src = reference_functions(shim)
        src = gen_src(dim, device=torch.device("cpu"))
        out[f"src_{i}"] = src.numpy()
        out[f"src_{i}_dim"] = np.array(dim)
        out[f"src_{i}_scalars"] = np.array(scalars)
        out[f"src_{i}_ints"] = np.array(ints)
    out["seed"] = np.array(seed)
    path = os.path.join(ROOT, "tests", "golden", "dataset.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items() if k.startswith(("eps", "src_0"))})


if __name__ == "__main__":
    main()
