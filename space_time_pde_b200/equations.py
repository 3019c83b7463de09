"""Equation strings -> jet requirements + residual programs.

The reference turns every ``dif(y, x)`` in an equation string into a separate
``torch.autograd.grad`` call (src/pde.py:8-9,82).  Here the parsed sympy expression is
rewritten symbolically instead: output variables become functions of the input variables,
``dif`` becomes ``Derivative`` and sympy applies the chain / product rule, which leaves a plain
arithmetic expression over

    inputs  q_k,  outputs  y_i,  first partials  d y_i / d q_a,  second partials  d2 y_i / d q_a d q_b.

From that expression we derive (a) the *jet specification* the fused kernel must propagate and
(b) a postfix program for ``stpde_residuals`` (include/stpde.h) or, when the expression uses
functions outside {+, *, integer powers}, a torch-lambdified evaluator over the jet tensors.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import sympy
from sympy.core.function import AppliedUndef

OP_CONST, OP_Q, OP_Y, OP_JET, OP_ADD, OP_MUL, OP_NEG, OP_POWI, OP_END, OP_GRES = range(10)


class UnsupportedEquation(NotImplementedError):
    """The equation needs something the jet path cannot provide (e.g. third derivatives)."""


@dataclass(frozen=True)
class JetSpec:
    """Partials to propagate: first-order directions and (i <= j) second-order pairs."""
    first: Tuple[int, ...] = ()
    second: Tuple[Tuple[int, int], ...] = ()

    @property
    def n_jet(self) -> int:
        return len(self.first) + len(self.second)

    def plane(self, multi: Tuple[int, ...]) -> int:
        """Index of the derivative plane for a multi-index (sorted tuple of directions)."""
        if len(multi) == 1:
            return self.first.index(multi[0])
        return len(self.first) + self.second.index((multi[0], multi[1]))

    def union(self, other: "JetSpec") -> "JetSpec":
        second = tuple(sorted(set(self.second) | set(other.second)))
        first = set(self.first) | set(other.first)
        for a, b in second:
            first.update((a, b))
        return JetSpec(tuple(sorted(first)), second)

    def split(self, max_components: int) -> List["JetSpec"]:
        """Split into specs of at most ``max_components`` (1 + n_first + n_second) each."""
        room = max_components - 1 - len(self.first)
        if room >= len(self.second):
            return [self]
        if room < 1:
            raise UnsupportedEquation("too many first-order directions for one launch")
        return [JetSpec(self.first, self.second[i:i + room]) for i in range(0, len(self.second), room)]


@dataclass
class CompiledEquation:
    name: str
    expr: sympy.Expr                      # over input/output/jet symbols
    jet_symbols: Dict[sympy.Symbol, Tuple[int, Tuple[int, ...]]]   # symbol -> (output index, multi-index)
    spec: JetSpec
    program: Optional[Tuple[List[int], List[float]]] = None        # (words, consts) w/o plane resolution
    torch_fn: Optional[Callable] = None
    arg_symbols: List[sympy.Symbol] = field(default_factory=list)


def jet_symbol_name(out_name: str, in_names: Sequence[str], multi: Tuple[int, ...]) -> str:
    return f"{out_name}__" + "_".join(in_names[k] for k in multi)


def compile_equation(name: str, expr: sympy.Expr, in_vars: Sequence[sympy.Symbol],
                     out_vars: Sequence[sympy.Symbol]) -> CompiledEquation:
    """Rewrite ``dif`` into jet symbols (max order 2) and build the evaluators."""
    in_vars = list(in_vars)
    out_vars = list(out_vars)
    in_names = [s.name for s in in_vars]
    funcs = {v: sympy.Function("F_" + v.name)(*in_vars) for v in out_vars}
    e = expr.subs(funcs, simultaneous=True)

    def is_dif(node):
        return isinstance(node, AppliedUndef) and node.func.__name__ == "dif"

    def to_derivative(node):
        y, x = node.args
        if not (isinstance(x, sympy.Symbol) and x in in_vars):
            raise UnsupportedEquation(f"dif() second argument must be an input variable, got {x}")
        return sympy.Derivative(y, x)

    # bottom-up replacement handles nested dif(dif(.)) and products inside dif
    e = e.replace(is_dif, to_derivative).doit()
    leftover = [f for f in e.atoms(AppliedUndef) if f not in funcs.values()]
    if leftover:
        raise UnsupportedEquation(f"unknown functions in equation {name!r}: {leftover}")

    jet_symbols: Dict[sympy.Symbol, Tuple[int, Tuple[int, ...]]] = {}
    repl = {}
    for der in e.atoms(sympy.Derivative):
        f = der.expr
        if f not in funcs.values():
            raise UnsupportedEquation(f"cannot differentiate {f}")
        oi = list(funcs.values()).index(f)
        multi: List[int] = []
        for var, count in der.variable_count:
            multi += [in_vars.index(var)] * int(count)
        if len(multi) > 2:
            raise UnsupportedEquation(f"derivative order {len(multi)} > 2 in equation {name!r}")
        multi_t = tuple(sorted(multi))
        sym = sympy.Symbol(jet_symbol_name(out_vars[oi].name, in_names, multi_t))
        jet_symbols[sym] = (oi, multi_t)
        repl[der] = sym
    e = e.subs(repl, simultaneous=True).subs({f: v for v, f in funcs.items()}, simultaneous=True)

    first = set()
    second = set()
    for _, multi in jet_symbols.values():
        if len(multi) == 1:
            first.add(multi[0])
        else:
            second.add(multi)
            first.update(multi)
    spec = JetSpec(tuple(sorted(first)), tuple(sorted(second)))
    ce = CompiledEquation(name=name, expr=e, jet_symbols=jet_symbols, spec=spec)
    ce.arg_symbols = in_vars + out_vars + sorted(jet_symbols, key=lambda s: s.name)
    try:
        ce.program = _postfix(e, in_vars, out_vars, jet_symbols)
    except UnsupportedEquation:
        ce.program = None
    ce.torch_fn = sympy.lambdify(ce.arg_symbols, e, "math" if not ce.arg_symbols else _torch_namespace())
    return ce


def _torch_namespace():
    import torch
    return [{"sin": torch.sin, "cos": torch.cos, "exp": torch.exp, "log": torch.log, "sqrt": torch.sqrt,
             "tanh": torch.tanh, "Abs": torch.abs}, "math"]


def _postfix(e: sympy.Expr, in_vars, out_vars, jet_symbols, g_symbols=(), consts=None):
    """Postfix words with symbolic jet operands: (OP_JET, (output, multi)) resolved at bind time."""
    words: List = []
    consts = [] if consts is None else consts

    def const(v: float):
        v = float(v)
        if v not in consts:
            consts.append(v)
        words.extend((OP_CONST, consts.index(v)))

    def emit(node):
        if node.is_Symbol:
            if node in jet_symbols:
                words.extend((OP_JET, jet_symbols[node]))
            elif node in in_vars:
                words.extend((OP_Q, in_vars.index(node)))
            elif node in out_vars:
                words.extend((OP_Y, out_vars.index(node)))
            elif node in g_symbols:
                words.extend((OP_GRES, list(g_symbols).index(node)))
            else:
                raise UnsupportedEquation(f"free symbol {node}")
        elif node.is_Number:
            const(node)
        elif node.is_Add or node.is_Mul:
            op = OP_ADD if node.is_Add else OP_MUL
            args = list(node.args)
            if node.is_Mul and args[0] == -1 and len(args) > 1:
                emit(sympy.Mul(*args[1:]))
                words.extend((OP_NEG, 0))
                return
            emit(args[0])
            for a in args[1:]:
                emit(a)
                words.extend((op, 0))
        elif node.is_Pow and node.args[1].is_Integer:
            emit(node.args[0])
            words.extend((OP_POWI, int(node.args[1])))
        else:
            raise UnsupportedEquation(f"node {node.func} not supported by the residual kernel")

    emit(e)
    words.extend((OP_END, 0))
    if len(consts) > (256 if g_symbols else 128):
        raise UnsupportedEquation("too many constants")
    return words, consts


def bind_programs(equations: Sequence[CompiledEquation], spec: JetSpec, n_out: int):
    """Concatenate the per-equation programs against a concrete plane layout; None if any is missing."""
    words: List[int] = []
    consts: List[float] = []
    for ce in equations:
        if ce.program is None:
            return None
        w, c = ce.program
        base = len(consts)
        consts.extend(c)
        it = iter(w)
        for op in it:
            arg = next(it)
            if op == OP_CONST:
                arg += base
            elif op == OP_JET:
                oi, multi = arg
                arg = spec.plane(multi) * n_out + oi
            words.extend((op, arg))
    if len(words) > 640 or len(consts) > 128:
        return None
    return words, consts


def bind_adjoint_program(equations: Sequence[CompiledEquation], spec: JetSpec, in_vars, out_vars):
    """Postfix programs of the REVERSE sweep through the residual arithmetic (``stpde_residuals_backward``).

    One program per output symbol s - y_0..y_{o-1}, then every (jet plane, output) entry in plane-major order -
    evaluating  sum_e G_e * d residual_e / d s  with G_e = d loss / d residual_e (opcode OP_GRES).  The equations are
    differentiated symbolically; polynomial residuals stay polynomial.  None if a program cannot be built."""
    in_vars, out_vars = list(in_vars), list(out_vars)
    n_out = len(out_vars)
    if any(ce is None or ce.program is None for ce in equations):
        return None
    g_syms = [sympy.Symbol(f"__gres{e}") for e in range(len(equations))]
    jet_symbols: Dict[sympy.Symbol, Tuple[int, Tuple[int, ...]]] = {}
    for ce in equations:
        jet_symbols.update(ce.jet_symbols)
    by_target = {v: k for k, v in jet_symbols.items()}
    multis = [(k,) for k in spec.first] + [tuple(pair) for pair in spec.second]
    targets: List[Optional[sympy.Symbol]] = list(out_vars)
    for multi in multis:
        for oi in range(n_out):
            targets.append(by_target.get((oi, multi)))
    words: List[int] = []
    consts: List[float] = []
    try:
        for sym in targets:
            total = sympy.Integer(0)
            if sym is not None:
                for ge, ce in zip(g_syms, equations):
                    if ce.expr.has(sym):
                        total = total + ge * sympy.diff(ce.expr, sym)
            w, _ = _postfix(sympy.expand(total) if total != 0 else total, in_vars, out_vars, jet_symbols, g_syms, consts)
            it = iter(w)
            for op in it:
                arg = next(it)
                if op == OP_JET:
                    oi, multi = arg
                    arg = spec.plane(multi) * n_out + oi
                words.extend((op, arg))
    except UnsupportedEquation:
        return None
    if len(words) > 2048 or len(consts) > 256:
        return None
    return words, consts
