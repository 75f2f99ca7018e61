"""tcgen05 / TMA path: the 3xTF32 GEMM building block and the SVGD tensor path against float64 references."""
import numpy as np
import pytest
import torch

import ocbnn_oracle as orc
from ocbnn_b200 import engine as E

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("m,n,k", [(128, 128, 32), (128, 64, 64), (300, 200, 100), (512, 64, 256), (257, 31, 40), (1000, 1000, 31), (130, 22464 // 16, 513)])
def test_gemm_tf32x3(m, n, k):
    g = torch.Generator().manual_seed(m * 7 + n * 3 + k)
    A = torch.randn(m, k, generator=g)
    B = torch.randn(n, k, generator=g) * 3 + 0.5
    C = E.gemm_tf32x3(A.cuda(), B.cuda()).cpu().double()
    ref = A.double() @ B.double().T
    scale = (A.double().abs() @ B.double().abs().T)                 # error relative to sum |a||b| (no cancellation credit)
    err = ((C - ref).abs() / scale).max().item()
    assert err < 2e-6, err
    plain_tf32 = 5e-4
    assert err < plain_tf32 / 50
