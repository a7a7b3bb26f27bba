"""Sliced operands and tensor-core products (csrc/umma_gemm.cu) on torch-owned CUDA memory.

fp64 matrices are split into int8 slices (exact int32 products on the tensor cores, recombined in fp64), fp32 matrices
into a TF32 hi/lo pair (3xTF32).  See include/pnode_b200.h "Tensor-core matrix products on SLICED operands"."""
import torch

from . import _lib
from .device import _stream
from .errors import Error

I8, TF32, I8X = 0, 1, 2


def kind_of(dtype, extended=False):
    """extended: one more int8 digit (55 instead of 48 bits) for products whose terms cancel by orders of magnitude."""
    if dtype == torch.float64:
        return I8X if extended else I8
    if dtype == torch.float32:
        return TF32
    raise Error(-10, "sliced operands exist for float64 (int8 slices) and float32 (TF32 pairs), not %s" % dtype)


class Sliced:
    """A row-major [rows][k] operand in sliced form; buffers are reused when `into` is given."""

    __slots__ = ("kind", "rows", "k", "buf", "exp")

    def __init__(self, kind, rows, k, device):
        lib = _lib.load()
        self.kind, self.rows, self.k = kind, rows, k
        self.buf = torch.empty(int(lib.pnode_sliced_bytes(kind, rows, k)), dtype=torch.uint8, device=device)
        self.exp = torch.zeros(rows, dtype=torch.int32, device=device)


def _check2d(x):
    if not x.is_cuda or x.dim() != 2 or x.stride(1) != 1:
        raise Error(-11, "sliced operand source must be a 2-d CUDA tensor with unit inner stride")


def slice_rows(x, into=None, extended=False):
    """Operand = x ([rows][k], reduction over the columns of x)."""
    _check2d(x)
    kind = kind_of(x.dtype, extended)
    rows, k = x.shape
    s = into if into is not None else Sliced(kind, rows, k, x.device)
    if (s.kind, s.rows, s.k) != (kind, rows, k):
        raise Error(-12, "sliced buffer does not match the source")
    _lib.check(_lib.load().pnode_slice_rows(kind, x.data_ptr(), x.stride(0), rows, k, s.buf.data_ptr(), s.exp.data_ptr(),
                                            _stream()))
    return s


def slice_cols(x, into=None, colsum=None, coef=0.0, extended=False):
    """Operand = x^T ([cols][rows], reduction over the rows of x); optionally colsum += coef * x.sum(0)."""
    _check2d(x)
    kind = kind_of(x.dtype, extended)
    rows, cols = x.shape
    s = into if into is not None else Sliced(kind, cols, rows, x.device)
    if (s.kind, s.rows, s.k) != (kind, cols, rows):
        raise Error(-12, "sliced buffer does not match the source")
    _lib.check(_lib.load().pnode_slice_cols(kind, x.data_ptr(), x.stride(0), rows, cols, s.buf.data_ptr(),
                                            s.exp.data_ptr(), None if colsum is None else colsum.data_ptr(), float(coef),
                                            _stream()))
    return s


def gemm(a, b, out=None, alpha=1.0, bias=None, relu=False, mask=None, accumulate=False):
    """out[m][n] (+)= mask(relu(alpha * (sum_k a[m][k] b[n][k] + bias[n])))."""
    if a.kind != b.kind or a.k != b.k:
        raise Error(-13, "sliced gemm: operand kinds / reduction lengths differ")
    dtype = torch.float32 if a.kind == TF32 else torch.float64
    if out is None:
        out = torch.empty(a.rows, b.rows, dtype=dtype, device=a.buf.device)
    if out.dtype != dtype or out.stride(1) != 1 or out.shape != (a.rows, b.rows):
        raise Error(-14, "sliced gemm: bad output tensor")
    _lib.check(_lib.load().pnode_sliced_gemm(a.kind, a.buf.data_ptr(), a.exp.data_ptr(), b.buf.data_ptr(), b.exp.data_ptr(),
                                             a.rows, b.rows, a.k, out.data_ptr(), out.stride(0), float(alpha),
                                             None if bias is None else bias.data_ptr(), int(bool(relu)),
                                             None if mask is None else mask.data_ptr(),
                                             0 if mask is None else mask.stride(0), int(bool(accumulate)), _stream()))
    return out
