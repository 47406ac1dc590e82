"""Pytrees for the shim: dicts (keys in sorted order, as JAX flattens them), lists,
tuples, None; everything else is a leaf.  TEST INFRASTRUCTURE ONLY."""


def tree_flatten(tree):
    leaves = []

    def rec(t):
        if isinstance(t, dict):
            keys = sorted(t)
            return ("dict", keys, [rec(t[k]) for k in keys])
        if isinstance(t, (list, tuple)):
            return ("list" if isinstance(t, list) else "tuple", None, [rec(x) for x in t])
        if t is None:
            return ("none", None, [])
        leaves.append(t)
        return ("leaf", None, [])

    return leaves, rec(tree)


def tree_unflatten(treedef, leaves):
    it = iter(leaves)

    def rec(d):
        kind, keys, kids = d
        if kind == "dict":
            return {k: rec(c) for k, c in zip(keys, kids)}
        if kind == "list":
            return [rec(c) for c in kids]
        if kind == "tuple":
            return tuple(rec(c) for c in kids)
        if kind == "none":
            return None
        return next(it)

    return rec(treedef)


def tree_map(fn, tree, *rest):
    leaves, treedef = tree_flatten(tree)
    others = [tree_flatten(r)[0] for r in rest]
    return tree_unflatten(treedef, [fn(*xs) for xs in zip(leaves, *others)])
