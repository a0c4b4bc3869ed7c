"""Stand-in for `opt_einsum` (absent here).  Test infrastructure for oracle/make_golden.py.

kronfluence only asks opt_einsum for a contraction ORDER (module/linear.py:90-121); the arithmetic is
torch's einsum.  `contract_path` below returns the flop-minimal pairwise order found by exhaustive
search (at most 4 operands are ever passed), which is what DynamicProgramming(minimize="flops") finds.
"""

import itertools

import torch


class DynamicProgramming:
    def __init__(self, *args, **kwargs):
        pass


def _parse(expr, shapes):
    lhs, out = expr.replace(" ", "").split("->")
    terms = lhs.split(",")
    expanded = []
    ell = None
    for term, shape in zip(terms, shapes):
        if "..." in term:
            n_named = len(term.replace("...", ""))
            n_ell = len(shape) - n_named
            letters = "".join(chr(ord("A") + i) for i in range(n_ell))
            ell = letters if ell is None or len(letters) > len(ell) else ell
            term = term.replace("...", letters)
        expanded.append(term)
    if "..." in out:
        out = out.replace("...", ell or "")
    sizes = {}
    for term, shape in zip(expanded, shapes):
        for ch, n in zip(term, shape):
            sizes[ch] = n
    return expanded, out, sizes


def contract_path(expr, *operands, optimize=None, **kwargs):
    shapes = [tuple(op.shape) for op in operands]
    terms, out, sizes = _parse(expr, shapes)

    def cost_of(order):
        live = list(terms)
        total = 0
        for i, j in order:
            a, b = live[i], live[j]
            rest = [t for k, t in enumerate(live) if k not in (i, j)]
            keep = set(out).union(*[set(t) for t in rest]) if rest else set(out)
            idx = set(a) | set(b)
            flops = 1
            for ch in idx:
                flops *= sizes[ch]
            total += flops
            new = "".join(ch for ch in sorted(idx) if ch in keep)
            live = rest + [new]
        return total

    def orders(n):
        if n == 1:
            yield []
            return
        for i, j in itertools.combinations(range(n), 2):
            for tail in orders(n - 1):
                yield [(i, j)] + tail

    best = min(orders(len(terms)), key=cost_of)
    return best, None


def contract(expr, *operands, **kwargs):
    return torch.einsum(expr, *operands)
