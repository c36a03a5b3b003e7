"""The ordering kernel against torch's CUDA sort (the un-vendored arithmetic behind libs/ops/csrc/nms.cpp:51)."""
import pytest
import torch

from oracle import oracle
from phnet_b200.ops import sort_order

pytestmark = pytest.mark.gpu

SIZES = list(range(1, 40)) + [63, 64, 65, 100, 127, 128, 129, 255, 256, 257, 1000, 1024, 4095, 4096, 4097, 8192, 20000]


def nasty(N, variant, g):
    s = torch.floor(torch.rand(N, generator=g) * 8.0) / 8.0
    if variant >= 1 and N > 3:
        s[torch.randint(0, N, (max(1, N // 10),), generator=g)] = 1.0
        s[torch.randint(0, N, (max(1, N // 16),), generator=g)] = -0.0
    if variant >= 2 and N > 3:
        s[torch.randint(0, N, (max(1, N // 12),), generator=g)] = float("nan")
        s[torch.randint(0, N, (max(1, N // 20),), generator=g)] = -float("inf")
    return s


@pytest.mark.parametrize("variant", [0, 1, 2])
def test_order_matches_torch_cuda_sort(cuda_device, variant):
    g = torch.Generator().manual_seed(77 + variant)
    for N in SIZES:
        s = nasty(N, variant, g)
        sd = s.to(cuda_device)
        want = torch.sort(sd, 0, True)[1]
        got = sort_order(sd)
        assert torch.equal(got, want), f"N={N} variant={variant}: first diff at {int((got != want).nonzero()[0])}"
        # and the CPU oracle models the same permutation
        assert (oracle.order(s.numpy()) == want.cpu().numpy()).all(), f"oracle order N={N} variant={variant}"


def test_order_distinct_scores_batched(cuda_device):
    g = torch.Generator().manual_seed(5)
    s = torch.rand(64, 1000, generator=g).to(cuda_device)
    got = sort_order(s)
    want = torch.sort(s, 1, True)[1]
    assert torch.equal(got, want)


def test_stable_model_matches_torch_stable(cuda_device):
    g = torch.Generator().manual_seed(9)
    for N in (2, 17, 32, 33, 128, 1000, 5000):
        s = nasty(N, 1, g).to(cuda_device)
        assert torch.equal(sort_order(s, sort_model=2), torch.sort(s, dim=0, descending=True, stable=True)[1]), N
