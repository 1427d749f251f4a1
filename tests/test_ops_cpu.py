"""torch.ops.tavk.* (multi_modal_emotion_b200/ops.py): registration, schemas, fake-tensor shape propagation through forward
AND backward — all without a GPU — and the no-CPU-fallback rule."""
import pytest
import torch


def test_operators_are_registered_with_schemas():
    from multi_modal_emotion_b200 import ops

    for name in ops.OPS:
        op = getattr(torch.ops.tavk, name)
        assert "tavk::" + name in str(op.default._schema)
    s = str(torch.ops.tavk.attention.default._schema)
    assert "Tensor qkv" in s and "Int heads" in s.replace("int heads", "Int heads") and "Tensor? key_bias" in s


def test_fake_tensor_shapes_forward_and_backward():
    from torch._subclasses.fake_tensor import FakeTensorMode

    from multi_modal_emotion_b200 import ops

    with FakeTensorMode():
        x = torch.empty(2, 185, 768, requires_grad=True)
        w, b = torch.empty(768, requires_grad=True), torch.empty(768, requires_grad=True)
        y = ops.layer_norm(x, w, b, 1e-12)
        pooled = ops.mean_pool(y)
        lw, lb = torch.empty(7, 768, requires_grad=True), torch.empty(7, requires_grad=True)
        logits = ops.small_linear(pooled, lw, lb)
        loss = ops.cross_entropy(logits, torch.empty(2, dtype=torch.long), torch.empty(7))
        assert y.shape == (2, 185, 768) and pooled.shape == (2, 768) and logits.shape == (2, 7) and loss.shape == ()
        loss.backward()
        assert x.grad.shape == x.shape and w.grad.shape == (768,) and lw.grad.shape == (7, 768) and lb.grad.shape == (7,)
        qkv = torch.empty(2, 323, 3 * 768, dtype=torch.bfloat16, requires_grad=True)
        o, lse = torch.ops.tavk.attention(qkv, 12, None)
        assert o.shape == (2, 323, 768) and o.dtype == torch.bfloat16 and lse.shape == (2, 12, 323)
        o.float().sum().backward()
        assert qkv.grad.shape == qkv.shape
        # linear (tcgen05 GEMM) -> embed_add -> dropout: shapes through forward and backward
        xl = torch.empty(2, 149, 1024, requires_grad=True)
        wl, bl = torch.empty(768, 1024, requires_grad=True), torch.empty(768, requires_grad=True)
        table = torch.empty(3, 768, requires_grad=True)
        h = torch.ops.tavk.linear(xl, wl, bl)
        h = torch.ops.tavk.embed_add(h, torch.empty(2, 149, dtype=torch.long), table)
        h, keep = torch.ops.tavk.dropout(h, 0.4, 123, torch.empty(1, dtype=torch.long))
        assert h.shape == (2, 149, 768) and keep.dtype == torch.uint8
        h.sum().backward()
        assert xl.grad.shape == xl.shape and wl.grad.shape == wl.shape and bl.grad.shape == (768,) and table.grad.shape == (3, 768)
        assert torch.ops.tavk.linear(xl, wl, None).shape == (2, 149, 768)


def test_cpu_tensors_are_refused():
    from multi_modal_emotion_b200 import ops

    with pytest.raises(RuntimeError, match="no CPU"):
        ops.mean_pool(torch.zeros(1, 4, 768))
    with pytest.raises(RuntimeError, match="no CPU"):
        ops.layer_norm(torch.zeros(2, 768), torch.ones(768), torch.zeros(768))
