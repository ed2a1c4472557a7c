"""``tt_attention_varlen_bf16`` (packed variable-length attention on tcgen05, csrc/attention.cu) and ``tt_cls_head_f32``
against plain PyTorch fp32 references of the same ops on the same (bf16-rounded) inputs.  Floating point: the kernel
rounds the softmax numerators and its output to bf16 (fp32 accumulation); tolerance 2e-2 absolute on outputs of O(1)."""

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _reference(qkv, cu, n_heads):
    """fp32 softmax(q k^T / sqrt(d)) v per sequence and head, from the bf16 values the kernel reads."""
    t, three_h = qkv.shape
    h = three_h // 3
    d = h // n_heads
    out = torch.zeros((t, h), dtype=torch.float32, device=qkv.device)
    x = qkv.float()
    for i in range(len(cu) - 1):
        a, b = int(cu[i]), int(cu[i + 1])
        q = x[a:b, :h].view(b - a, n_heads, d).transpose(0, 1)
        k = x[a:b, h:2 * h].view(b - a, n_heads, d).transpose(0, 1)
        v = x[a:b, 2 * h:].view(b - a, n_heads, d).transpose(0, 1)
        p = torch.softmax(q @ k.transpose(1, 2) * d ** -0.5, dim=-1)
        out[a:b] = (p @ v).transpose(0, 1).reshape(b - a, h)
    return out


@pytest.mark.parametrize("lens", [[128], [1, 2, 3], [300, 77, 129, 512, 1, 64, 255, 257], [512] * 3, [17] * 40])
@pytest.mark.parametrize("n_heads", [16, 2])
def test_varlen_attention_matches_fp32_reference(lens, n_heads):
    from tensor_truth_b200 import _lib

    L = _lib.lib()
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(sum(lens) * 31 + n_heads)
    total, h = sum(lens), n_heads * 64
    qkv = (torch.randn((total, 3 * h), generator=g) * 1.5).to(torch.bfloat16).to(dev)
    cu = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32, device=dev)
    out = torch.full((total, h), float("nan"), dtype=torch.bfloat16, device=dev)
    _lib.check(L.tt_attention_varlen_bf16(qkv.data_ptr(), total, n_heads, 64, cu.data_ptr(), len(lens), max(lens),
                                          total // 128 + len(lens), 0.125, out.data_ptr(), None))
    torch.cuda.synchronize()
    _lib.check_status(0)
    ref = _reference(qkv, cu.cpu().numpy(), n_heads)
    got = out.float()
    assert torch.isfinite(got).all()
    err = (got - ref).abs().max().item()
    assert err < 2e-2, err


def test_attention_sharp_softmax_and_large_scores():
    """Scores far from zero (one dominant key per row): the two-pass maximum keeps exp2 in range."""
    from tensor_truth_b200 import _lib

    L = _lib.lib()
    dev = torch.device("cuda:0")
    lens = [200, 333]
    total, n_heads, h = sum(lens), 4, 256
    g = torch.Generator(device="cpu").manual_seed(5)
    qkv = (torch.randn((total, 3 * h), generator=g) * 6.0).to(torch.bfloat16).to(dev)
    cu = torch.tensor([0, 200, 533], dtype=torch.int32, device=dev)
    out = torch.empty((total, h), dtype=torch.bfloat16, device=dev)
    _lib.check(L.tt_attention_varlen_bf16(qkv.data_ptr(), total, n_heads, 64, cu.data_ptr(), 2, 512, total // 128 + 2, 0.125,
                                          out.data_ptr(), None))
    torch.cuda.synchronize()
    ref = _reference(qkv, [0, 200, 533], n_heads)
    err = (out.float() - ref).abs().max().item()
    assert err < 0.12, err  # outputs of magnitude ~6-20 here: bf16 output rounding alone is up to 0.06


def test_attention_rejects_unsupported_shapes():
    from tensor_truth_b200 import _lib

    L = _lib.lib()
    x = torch.zeros((8, 3 * 128), dtype=torch.bfloat16, device="cuda")
    cu = torch.tensor([0, 8], dtype=torch.int32, device="cuda")
    out = torch.zeros((8, 128), dtype=torch.bfloat16, device="cuda")
    assert L.tt_attention_varlen_bf16(x.data_ptr(), 8, 4, 32, cu.data_ptr(), 1, 8, 1, 0.1, out.data_ptr(), None) == -3  # head_dim
    assert L.tt_attention_varlen_bf16(x.data_ptr(), 8, 2, 64, cu.data_ptr(), 1, 600, 1, 0.1, out.data_ptr(), None) == -3  # too long


def test_cls_head_matches_torch():
    from tensor_truth_b200 import _lib

    L = _lib.lib()
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(9)
    hdim, n = 1024, 7
    lens = [5, 1, 300, 12, 64, 2, 9]
    x = torch.randn((sum(lens), hdim), generator=g).to(torch.bfloat16).to(dev)
    cu = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32, device=dev)
    w1 = (torch.randn((hdim, hdim), generator=g) * 0.03).to(dev)
    b1 = torch.randn((hdim,), generator=g).to(dev) * 0.1
    w2 = (torch.randn((1, hdim), generator=g) * 0.05).to(dev)
    b2 = torch.randn((1,), generator=g).to(dev)
    out = torch.empty((n,), dtype=torch.float32, device=dev)
    _lib.check(L.tt_cls_head_f32(x.data_ptr(), cu.data_ptr(), n, hdim, w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(),
                                 out.data_ptr(), None))
    torch.cuda.synchronize()
    cls = x[cu[:-1].long()].float()
    ref = (torch.tanh(cls @ w1.T + b1) @ w2.T + b2).squeeze(-1)
    assert (out - ref).abs().max().item() < 1e-4
