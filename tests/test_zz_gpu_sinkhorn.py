"""Sinkhorn-Knopp on the GPU (crw_sinkhorn_knopp, utils/__init__.py:615-641 + model.py:83-87): same number of sweeps and the
same matrix as the reference's rule; the simulator twin of these tests compares against the reference's own function."""
import pytest
import torch

from oracle import crw_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("R,N,M,tol,max_iter", [(3, 12, 12, 0.01, 100), (2, 49, 49, 1e-4, 1000), (8, 196, 196, 0.01, 100)])
def test_sinkhorn_knopp_matches_reference_rule(R, N, M, tol, max_iter):
    from sapienza_video_contrastive_b200 import ops
    g = torch.Generator().manual_seed(R * 100 + N)
    A = torch.randn(R, N, M, generator=g) * 0.3
    A[0, 1, 2] = -1e20
    ref, its = O.sinkhorn_knopp((A / 0.07).exp(), tol, max_iter)
    out, n_it = ops.sinkhorn_knopp(A.to(DEV), tol, max_iter, exp_temperature=0.07)
    assert n_it == its
    torch.testing.assert_close(out.cpu(), ref, rtol=2e-5, atol=1e-9)


def test_stoch_mat_sinkhorn_branch():
    """CRW.stoch_mat(do_sinkhorn=True), model.py:83-87."""
    import argparse
    from sapienza_video_contrastive_b200 import CRW
    args = argparse.Namespace(device=DEV, dropout=0.0, featdrop=0.0, temp=0.07, head_depth=0, model_type="scratch", remove_layers=[],
                              dilate_superpixels=False, flip=False, sk_targets=False)
    crw = CRW(args).to(DEV)
    A = torch.randn(2, 9, 9, generator=torch.Generator().manual_seed(1)) * 0.2
    out = crw.stoch_mat(A.to(DEV), do_dropout=False, do_sinkhorn=True)
    ref, _ = O.sinkhorn_knopp((A / 0.07).exp(), 0.01, 100)
    torch.testing.assert_close(out.cpu(), ref, rtol=2e-5, atol=1e-9)
