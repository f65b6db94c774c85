"""Minimal JAX-style pytree helpers (API parity with the reference module_a/torch_tree_util.py:
tree_flatten :69, tree_unflatten :130, tree_leaves :186, tree_structure :192, tree_map :198,
tree_reduce :229, Partial :268).  Containers: list, tuple, dict (keys in sorted order), None = empty."""
from __future__ import annotations

import functools
from typing import Any, Callable, List, Tuple

__all__ = ["PyTreeDef", "tree_flatten", "tree_unflatten", "tree_leaves", "tree_structure", "tree_map",
           "tree_reduce", "tree_all", "Partial"]

_LEAF = object()


class PyTreeDef:
    """Shape of a pytree: a nested spec of ('list'|'tuple'|'dict'|'none'|leaf)."""

    def __init__(self, spec: Any, num_leaves: int):
        self.spec = spec
        self.num_leaves = num_leaves

    def __repr__(self):
        return f"PyTreeDef(num_leaves={self.num_leaves})"

    def __eq__(self, other):
        return isinstance(other, PyTreeDef) and _spec_eq(self.spec, other.spec)

    def unflatten(self, leaves: List[Any]) -> Any:
        return tree_unflatten(self, leaves)


def _spec_eq(a, b) -> bool:
    if a is _LEAF or b is _LEAF:
        return a is b
    if a[0] != b[0]:
        return False
    if a[0] == "dict":
        return a[1] == b[1] and all(_spec_eq(x, y) for x, y in zip(a[2], b[2]))
    if a[0] == "none":
        return True
    return len(a[1]) == len(b[1]) and all(_spec_eq(x, y) for x, y in zip(a[1], b[1]))


def _flatten(tree: Any, out: List[Any]):
    if tree is None:
        return ("none",)
    if isinstance(tree, (list, tuple)) and not hasattr(tree, "_fields"):
        return ("list" if isinstance(tree, list) else "tuple", [_flatten(t, out) for t in tree])
    if isinstance(tree, dict):
        keys = sorted(tree.keys())
        return ("dict", keys, [_flatten(tree[k], out) for k in keys])
    out.append(tree)
    return _LEAF


def tree_flatten(tree: Any) -> Tuple[List[Any], PyTreeDef]:
    leaves: List[Any] = []
    spec = _flatten(tree, leaves)
    return leaves, PyTreeDef(spec, len(leaves))


def tree_unflatten(treedef: PyTreeDef, leaves: List[Any]) -> Any:
    leaves = list(leaves)
    if len(leaves) != treedef.num_leaves:
        raise ValueError(f"expected {treedef.num_leaves} leaves, got {len(leaves)}")
    it = iter(leaves)

    def build(spec):
        if spec is _LEAF:
            return next(it)
        kind = spec[0]
        if kind == "none":
            return None
        if kind == "dict":
            return {k: build(s) for k, s in zip(spec[1], spec[2])}
        seq = [build(s) for s in spec[1]]
        return seq if kind == "list" else tuple(seq)

    return build(treedef.spec)


def tree_leaves(tree: Any) -> List[Any]:
    return tree_flatten(tree)[0]


def tree_structure(tree: Any) -> PyTreeDef:
    return tree_flatten(tree)[1]


def tree_map(func: Callable, tree: Any, *rest: Any) -> Any:
    leaves, treedef = tree_flatten(tree)
    others = []
    for r in rest:
        rl, rd = tree_flatten(r)
        if rd != treedef:
            raise ValueError("tree_map: pytrees have different structures")
        others.append(rl)
    return tree_unflatten(treedef, [func(*xs) for xs in zip(leaves, *others)])


def tree_reduce(func: Callable, tree: Any, initializer: Any = None) -> Any:
    leaves = tree_leaves(tree)
    if initializer is None:
        if not leaves:
            raise TypeError("tree_reduce of an empty pytree with no initializer")
        return functools.reduce(func, leaves)
    return functools.reduce(func, leaves, initializer)


def tree_all(tree: Any) -> bool:
    return all(bool(x) for x in tree_leaves(tree))


class Partial:
    """functools.partial look-alike that stays callable and introspectable (reference :268-277)."""

    def __init__(self, func, *args, **kwargs):
        self.func, self.args, self.kwargs = func, args, kwargs

    def __call__(self, *more_args, **more_kwargs):
        return self.func(*self.args, *more_args, **{**self.kwargs, **more_kwargs})
