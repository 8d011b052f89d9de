"""Weight packing kernels (fp32 nn.Parameter -> fp16 tensor-core operand): vectorised and scalar paths against torch."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _lib():
    from fastervit_b200 import lib
    lib.load()
    return lib


@pytest.mark.parametrize("rows,cols", [(37, 784), (16, 27), (5, 12), (128, 3136)])
def test_cast_pad(rows, cols):
    L = _lib()
    g = torch.Generator(device="cuda").manual_seed(rows)
    src = torch.randn(rows, cols, device="cuda", generator=g)
    ld = (cols + 7) // 8 * 8
    dst = torch.full((rows, ld), 7.0, device="cuda", dtype=torch.float16)
    L.call("fvit_cast_pad_f16", src.data_ptr(), cols, dst.data_ptr(), ld, rows, cols, ld)
    assert torch.equal(dst[:, :cols], src.half())
    assert (dst[:, cols:] == 0).all()


@pytest.mark.parametrize("h,hd,hdp,C", [(16, 49, 64, 784), (3, 20, 32, 60), (4, 32, 32, 128)])
def test_cast_headpad_rows_and_cols(h, hd, hdp, C):
    L = _lib()
    g = torch.Generator(device="cuda").manual_seed(h * hd)
    # qkv weight [3*h*hd, C] -> [3*h*hdp, C]: rows of every head zero padded
    w = torch.randn(3 * h * hd, C, device="cuda", generator=g)
    ld = (C + 7) // 8 * 8
    dst = torch.full((3 * h * hdp, ld), 7.0, device="cuda", dtype=torch.float16)
    L.call("fvit_cast_headpad_f16", w.data_ptr(), C, dst.data_ptr(), ld, 3 * h * hdp, C, hd, hdp, 1, 0)
    ref = torch.zeros(3 * h, hdp, C, device="cuda")
    ref[:, :hd] = w.view(3 * h, hd, C)
    assert torch.equal(dst[:, :C], ref.view(-1, C).half())
    # proj weight [C, h*hd] -> [C, h*hdp]: columns of every head zero padded
    p = torch.randn(C, h * hd, device="cuda", generator=g)
    dst2 = torch.full((C, h * hdp), 7.0, device="cuda", dtype=torch.float16)
    L.call("fvit_cast_headpad_f16", p.data_ptr(), h * hd, dst2.data_ptr(), h * hdp, C, h * hdp, hd, hdp, 0, 1)
    ref2 = torch.zeros(C, h, hdp, device="cuda")
    ref2[:, :, :hd] = p.view(C, h, hd)
    assert torch.equal(dst2, ref2.view(C, -1).half())
