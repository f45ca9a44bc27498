"""Tracing of distribution-argument lambdas and moment functions into expression trees.

The reference evaluates ``lambda psi_z: psi_z.exp()`` or ``lambda z, x: z @ x`` eagerly on
first-class-dim tensors (reference: src/alan/dist.py:211-229), materialising every
intermediate.  Here each lambda is called ONCE, at plan time, with `Proxy` arguments that
record the arithmetic; the recorded tree is lowered to the register VM of csrc/vm.cuh and
fused with the log-density, so nothing but the factor cell is written at run time.
Anything the VM cannot express raises -- there is no eager fallback.
"""
from __future__ import annotations

import numbers
from typing import Optional

import torch

UNARY = {'neg', 'exp', 'log', 'sigmoid', 'square', 'sqrt', 'reciprocal', 'softplus', 'tanh', 'abs', 'log1p',
         'lgamma', 'cos', 'sin'}
BINARY = {'add', 'sub', 'mul', 'div', 'pow', 'lt'}


def _bshape(a, b):
    """right-aligned broadcast of positional shapes"""
    out = []
    for i in range(1, max(len(a), len(b)) + 1):
        x = a[-i] if i <= len(a) else 1
        y = b[-i] if i <= len(b) else 1
        if x != y and x != 1 and y != 1:
            raise Exception(f"positional shapes {a} and {b} do not broadcast")
        out.append(max(x, y))
    return tuple(reversed(out))


class Expr:
    """op in {'leaf','const'} | UNARY | BINARY | 'sumlast' | density ops (added by the planner)."""
    __slots__ = ('op', 'args', 'axes', 'pos_shape', 'ref', 'value', 'rename', 'mode', 'mdim')

    def __init__(self, op, args=(), axes=(), pos_shape=(), ref=None, value=None, rename=None, mode=0, mdim=None):
        self.op, self.args = op, tuple(args)
        self.axes, self.pos_shape = tuple(axes), tuple(pos_shape)
        self.ref, self.value = ref, value
        self.rename = rename or {}      # leaf only: axis name seen by the expression -> axis name of the tensor
        self.mode, self.mdim = mode, mdim   # leaf only: 0 plain, 1 shifted by one along mdim, 2 only at mdim == 0

    @staticmethod
    def leaf(ref, axes, pos_shape, rename=None, mode=0, mdim=None):
        return Expr('leaf', (), axes, pos_shape, ref=ref, rename=rename, mode=mode, mdim=mdim)

    @staticmethod
    def const(v):
        return Expr('const', (), (), (), value=float(v))

    @staticmethod
    def make(op, *args):
        args = [a if isinstance(a, Expr) else Expr.const(a) for a in args]
        axes = []
        for a in args:
            for x in a.axes:
                if x not in axes:
                    axes.append(x)
        if op == 'sumlast':
            (a,) = args
            if len(a.pos_shape) < 1:
                raise Exception("sum over the last positional dim of a tensor that has none")
            return Expr(op, args, axes, a.pos_shape[:-1])
        shape = ()
        for a in args:
            shape = _bshape(shape, a.pos_shape)
        return Expr(op, args, axes, shape)


class Proxy:
    """Stand-in for a scope tensor while a model lambda is being traced."""
    def __init__(self, expr: Expr):
        self.expr = expr

    # ---- python operators
    def __add__(self, o): return _mk('add', self, o)
    def __radd__(self, o): return _mk('add', o, self)
    def __sub__(self, o): return _mk('sub', self, o)
    def __rsub__(self, o): return _mk('sub', o, self)
    def __mul__(self, o): return _mk('mul', self, o)
    def __rmul__(self, o): return _mk('mul', o, self)
    def __truediv__(self, o): return _mk('div', self, o)
    def __rtruediv__(self, o): return _mk('div', o, self)
    def __neg__(self): return _mk('neg', self)
    def __pow__(self, o):
        if isinstance(o, numbers.Number) and float(o) == 2.0:
            return _mk('square', self)
        return _mk('pow', self, o)
    def __matmul__(self, o): return _matmul(self, o)
    def __rmatmul__(self, o): return _matmul(o, self)

    # ---- tensor methods used by model lambdas
    def exp(self): return _mk('exp', self)
    def log(self): return _mk('log', self)
    def sigmoid(self): return _mk('sigmoid', self)
    def square(self): return _mk('square', self)
    def sqrt(self): return _mk('sqrt', self)
    def reciprocal(self): return _mk('reciprocal', self)
    def tanh(self): return _mk('tanh', self)
    def cos(self): return _mk('cos', self)
    def sin(self): return _mk('sin', self)
    def abs(self): return _mk('abs', self)
    def log1p(self): return _mk('log1p', self)
    def lgamma(self): return _mk('lgamma', self)
    def neg(self): return _mk('neg', self)
    def pow(self, o): return self.__pow__(o)
    def add(self, o): return _mk('add', self, o)
    def sub(self, o): return _mk('sub', self, o)
    def mul(self, o): return _mk('mul', self, o)
    def div(self, o): return _mk('div', self, o)
    def matmul(self, o): return _matmul(self, o)

    def sum(self, dim=None):
        if dim in (-1, len(self.expr.pos_shape) - 1) and len(self.expr.pos_shape) >= 1:
            return Proxy(Expr.make('sumlast', self.expr))
        raise Exception("B200 engine: only .sum(-1) over the last positional dim can be traced")

    def __getattr__(self, name):
        raise Exception(f"B200 engine cannot trace tensor method `.{name}` inside a model lambda "
                        f"(supported: arithmetic, @, {sorted(UNARY)})")

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        name = _TORCH_FUNCS.get(func)
        if name is None:
            raise Exception(f"B200 engine cannot trace torch function {getattr(func, '__name__', func)} "
                            f"inside a model lambda")
        if kwargs:
            raise Exception(f"B200 engine: keyword arguments to {name} are not traced")
        if name == 'matmul':
            return _matmul(args[0], args[1])
        if name == 'pow':
            return Proxy.__pow__(_as_proxy(args[0]), args[1])
        return _mk(name, *args)


def _as_proxy(x):
    if isinstance(x, Proxy):
        return x
    if isinstance(x, numbers.Number):
        return Proxy(Expr.const(x))
    if isinstance(x, torch.Tensor) and x.ndim == 0:
        return Proxy(Expr.const(float(x)))
    raise Exception("B200 engine: only numbers and scope tensors can appear inside a traced lambda "
                    f"(got {type(x)}); pass tensors through `inputs`")


def _mk(op, *args):
    return Proxy(Expr.make(op, *[_as_proxy(a).expr for a in args]))


def _matmul(a, b):
    """``@`` on positional dims with named axes as batch: vector.vector and matrix.vector."""
    a, b = _as_proxy(a), _as_proxy(b)
    pa, pb = len(a.expr.pos_shape), len(b.expr.pos_shape)
    if pa >= 1 and pb == 1 and pa <= 2:
        if a.expr.pos_shape[-1] != b.expr.pos_shape[-1]:
            raise Exception(f"matmul shape mismatch {a.expr.pos_shape} @ {b.expr.pos_shape}")
        return Proxy(Expr.make('sumlast', Expr.make('mul', a.expr, b.expr)))
    raise Exception(f"B200 engine: `@` is traced for vector@vector and matrix@vector only "
                    f"(got positional ranks {pa} and {pb})")


_TORCH_FUNCS = {
    torch.exp: 'exp', torch.log: 'log', torch.sigmoid: 'sigmoid', torch.square: 'square', torch.sqrt: 'sqrt',
    torch.reciprocal: 'reciprocal', torch.tanh: 'tanh', torch.abs: 'abs', torch.log1p: 'log1p',
    torch.lgamma: 'lgamma', torch.neg: 'neg', torch.negative: 'neg', torch.cos: 'cos', torch.sin: 'sin',
    torch.add: 'add', torch.sub: 'sub', torch.mul: 'mul', torch.div: 'div', torch.true_divide: 'div',
    torch.matmul: 'matmul', torch.pow: 'pow', torch.nn.functional.softplus: 'softplus',
    torch.nn.functional.sigmoid: 'sigmoid',
}


def trace_function(f, arg_exprs):
    """Call `f` on proxies of its arguments; returns the result Expr."""
    out = f(*[Proxy(e) for e in arg_exprs])
    if isinstance(out, numbers.Number):
        return Expr.const(out)
    if not isinstance(out, Proxy):
        raise Exception("Lambda on a distribution returned a non-Tensor")
    return out.expr
