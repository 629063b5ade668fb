"""Recipe for rebuilding an operator from the flat tuple of its leaf tensors (reference:
operators/linear_operator_representation_tree.py:7-44).  Autograd Functions receive ``(tree, *leaves)``, rebuild the
operator inside ``forward`` and never see Python objects that hold graph state."""
from __future__ import annotations


class LinearOperatorRepresentationTree(object):
    def __init__(self, linear_op):
        self._cls = linear_op.__class__
        self._kwarg_names = list(linear_op._differentiable_kwargs.keys())
        self._static_kwargs = linear_op._nondifferentiable_kwargs
        self._slots = []  # (start, stop, subtree) with subtree None for a plain tensor leaf
        pos = 0
        for arg in list(linear_op._args) + list(linear_op._differentiable_kwargs.values()):
            if hasattr(arg, "representation") and callable(arg.representation):
                width = len(arg.representation())
                self._slots.append((pos, pos + width, arg.representation_tree()))
                pos += width
            else:
                self._slots.append((pos, pos + 1, None))
                pos += 1
        self.num_leaves = pos

    def __call__(self, *leaves):
        rebuilt = []
        for start, stop, subtree in self._slots:
            rebuilt.append(leaves[start] if subtree is None else subtree(*leaves[start:stop]))
        nk = len(self._kwarg_names)
        if nk:
            args, kw_vals = rebuilt[:-nk], rebuilt[-nk:]
            return self._cls(*args, **dict(zip(self._kwarg_names, kw_vals)), **self._static_kwargs)
        return self._cls(*rebuilt, **self._static_kwargs)
