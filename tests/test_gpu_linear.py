"""The reranker stage's dense kernels against a plain PyTorch fp32 reference of the same op.

Floating point: inputs are bf16 on both sides, the reference computes in fp32; the kernel accumulates in fp32 on the
tensor cores and rounds its OUTPUT to bf16, so the tolerance is one bf16 rounding of the result (relative 2^-8) plus
a small absolute term for cancellation: |y - ref| <= 2^-7 * |ref| + 2e-3.

Every test here needs a GPU:  python -m pytest tests -m gpu
"""

import pytest
import torch

from tensor_truth_b200 import _lib

pytestmark = pytest.mark.gpu
RTOL, ATOL = 2.0 ** -7, 2e-3


def _linear(x, w, bias=None, residual=None, act=0):
    y = torch.empty((x.shape[0], w.shape[0]), dtype=torch.bfloat16, device=x.device)
    _lib.check(_lib.lib().tt_linear_bf16(_lib.ptr(x), x.shape[0], x.shape[1], _lib.ptr(w), w.shape[0], _lib.ptr(bias),
                                         _lib.ptr(residual), act, _lib.ptr(y), torch.cuda.current_stream().cuda_stream))
    return y


def _close(y, ref):
    err = (y.float() - ref).abs()
    bound = RTOL * ref.abs() + ATOL
    assert bool((err <= bound).all()), float((err - bound).max())


@pytest.mark.parametrize("t", [1, 100, 257, 4096, 7001])
@pytest.mark.parametrize("k_in,n_out", [(1024, 1024), (1024, 3072), (1024, 4096), (4096, 1024), (128, 256)])
def test_linear_matches_fp32_reference(t, k_in, n_out):
    g = torch.Generator(device="cuda").manual_seed(t * 131 + k_in + n_out)
    x = torch.randn((t, k_in), generator=g, device="cuda").to(torch.bfloat16)
    w = (torch.randn((n_out, k_in), generator=g, device="cuda") / k_in ** 0.5).to(torch.bfloat16)
    bias = torch.randn((n_out,), generator=g, device="cuda")
    res = torch.randn((t, n_out), generator=g, device="cuda").to(torch.bfloat16)
    base = x.float() @ w.float().T
    _close(_linear(x, w), base)
    _close(_linear(x, w, bias), base + bias)
    _close(_linear(x, w, bias, act=1), torch.nn.functional.gelu(base + bias))
    _close(_linear(x, w, bias, residual=res), base + bias + res.float())
    torch.cuda.synchronize()


def test_linear_argument_errors():
    x = torch.zeros((4, 1024), dtype=torch.bfloat16, device="cuda")
    w = torch.zeros((1000, 1024), dtype=torch.bfloat16, device="cuda")  # n_out not a multiple of 256
    with pytest.raises(_lib.TTError):
        _linear(x, w)


@pytest.mark.parametrize("dim", [64, 1024, 2048])
def test_layernorm_and_embedding_layernorm(dim):
    L = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device="cuda").manual_seed(dim)
    t = 1003
    x = (3.0 * torch.randn((t, dim), generator=g, device="cuda") + 0.5).to(torch.bfloat16)
    gamma = 1.0 + 0.1 * torch.randn((dim,), generator=g, device="cuda")
    beta = 0.1 * torch.randn((dim,), generator=g, device="cuda")
    y = torch.empty_like(x)
    _lib.check(L.tt_layernorm_bf16(_lib.ptr(x), t, dim, _lib.ptr(gamma), _lib.ptr(beta), 1e-5, _lib.ptr(y), st))
    _close(y, torch.nn.functional.layer_norm(x.float(), (dim,), gamma, beta, 1e-5))
    # the encoder's input layer: word + position + token-type embeddings, then LayerNorm
    vocab, n_pos = 5000, 514
    we = torch.randn((vocab, dim), generator=g, device="cuda").to(torch.bfloat16)
    pe = torch.randn((n_pos, dim), generator=g, device="cuda").to(torch.bfloat16)
    te = torch.randn((1, dim), generator=g, device="cuda").to(torch.bfloat16)
    ids = torch.randint(0, vocab, (t,), generator=g, device="cuda", dtype=torch.int32)
    pos = torch.randint(0, n_pos, (t,), generator=g, device="cuda", dtype=torch.int32)
    _lib.check(L.tt_embed_layernorm_bf16(_lib.ptr(ids), _lib.ptr(pos), t, dim, _lib.ptr(we), _lib.ptr(pe), _lib.ptr(te),
                                         _lib.ptr(gamma), _lib.ptr(beta), 1e-5, _lib.ptr(y), st))
    ref = torch.nn.functional.layer_norm(we[ids.long()].float() + pe[pos.long()].float() + te.float(), (dim,), gamma, beta, 1e-5)
    _close(y, ref)
    torch.cuda.synchronize()
