"""Thin tensor-level wrappers over the C ABI (shape checks + pointer extraction only).

No arithmetic happens in Python/torch here: torch provides device memory and the stream.
"""
import torch

from . import lib as _l

ACT_NONE, ACT_GELU, ACT_QUICK_GELU = 0, 1, 2
IGNORE_INDEX = -100
IMAGE_TOKEN_INDEX = -200


def _chk(t, dtype=torch.bfloat16):
    assert t.is_cuda and t.dtype == dtype, (t.device, t.dtype)
    return t


def gemm(a, b, out=None, *, a_mn=False, b_mn=False, bias=None, residual=None, act=ACT_NONE,
         accumulate=False, tile_n=0):
    """out[M,N] (+)= op(a) @ op(b)^T (+bias)(act)(+residual).

    a_mn=False: a is [M,K] (K contiguous);  a_mn=True: a is stored [K,M].
    b_mn=False: b is [N,K] (K contiguous);  b_mn=True: b is stored [K,N].
    Row strides may exceed the row length (column-block views of a wider buffer).
    """
    _chk(a), _chk(b)
    assert a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1
    if a_mn:
        K, M = a.shape
    else:
        M, K = a.shape
    if b_mn:
        Kb, N = b.shape
    else:
        N, Kb = b.shape
    assert K == Kb, (a.shape, b.shape, a_mn, b_mn)
    if out is None:
        assert not accumulate
        out = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
    _chk(out)
    assert out.shape == (M, N) and out.stride(1) == 1
    if residual is not None:
        _chk(residual)
        assert residual.shape == (M, N) and residual.stride(1) == 1
    if bias is not None:
        _chk(bias)
        assert bias.numel() == N and bias.is_contiguous()
    _l.call("rlaifv_gemm_bf16", _l.ptr(a), a.stride(0), int(a_mn), _l.ptr(b), b.stride(0), int(b_mn),
            _l.ptr(out), out.stride(0), M, N, K, _l.ptr(bias), _l.ptr(residual),
            residual.stride(0) if residual is not None else 0, int(act), int(accumulate), int(tile_n),
            _l.stream_ptr())
    return out
