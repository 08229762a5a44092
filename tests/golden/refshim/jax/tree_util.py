"""pytrees: nested dict / list / tuple / NamedTuple with tensor (or None) leaves."""
import torch as _t


def _is_leaf(x):
    return x is None or isinstance(x, _t.Tensor) or not isinstance(x, (dict, list, tuple))


def tree_map(fn, tree, *rest):
    if _is_leaf(tree):
        return None if tree is None else fn(tree, *rest)
    if isinstance(tree, dict):
        return type(tree)((k, tree_map(fn, v, *[r[k] for r in rest])) for k, v in tree.items())
    if hasattr(tree, "_fields"):      # NamedTuple
        return type(tree)(*[tree_map(fn, v, *[r[i] for r in rest]) for i, v in enumerate(tree)])
    return type(tree)(tree_map(fn, v, *[r[i] for r in rest]) for i, v in enumerate(tree))


def tree_flatten(tree):
    leaves = []

    def walk(t):
        if _is_leaf(t):
            if t is not None:
                leaves.append(t)
            return ("leaf", t is None)
        if isinstance(t, dict):
            return ("dict", type(t), [(k, walk(v)) for k, v in t.items()])
        if hasattr(t, "_fields"):
            return ("nt", type(t), [walk(v) for v in t])
        return ("seq", type(t), [walk(v) for v in t])
    return leaves, walk(tree)


def tree_unflatten(treedef, leaves):
    it = iter(leaves)

    def build(d):
        if d[0] == "leaf":
            return None if d[1] else next(it)
        if d[0] == "dict":
            return d[1]((k, build(v)) for k, v in d[2])
        if d[0] == "nt":
            return d[1](*[build(v) for v in d[2]])
        return d[1](build(v) for v in d[2])
    return build(treedef)


def tree_leaves(tree):
    return tree_flatten(tree)[0]
