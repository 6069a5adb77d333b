"""tcgen05 / TMEM Gram kernel (3xTF32) against float64 torch and against the SIMT fp32 kernel."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def W():
    import wdgh_b200
    wdgh_b200._lib.require_device()
    return wdgh_b200


@pytest.mark.timeout(120)
@pytest.mark.parametrize("m,d", [(1, 1), (63, 7), (128, 32), (129, 33), (500, 1433), (1000, 10), (777, 130),
                                 (3000, 40), (2708, 7)])
def test_gram_tensor_cores_match_fp64(W, m, d):
    gen = torch.Generator(device="cuda").manual_seed(m * 131 + d)
    z = torch.randn(m, d, device="cuda", generator=gen) * (1 + torch.rand(m, 1, device="cuda", generator=gen) * 3)
    ref = z.double() @ z.double().T
    got = W.graph.gram(z, use_tensor_cores=True)
    torch.cuda.synchronize()
    scale = ref.abs().max().item()
    err = (got.double() - ref).abs().max().item()
    assert err <= 2e-5 * scale, (err, scale)     # north_star tolerance is 1e-4 relative
    # off-diagonal tiles are mirrored exactly; inside a diagonal tile (i,j)/(j,i) differ by summation order only
    assert (got - got.T).abs().max().item() <= 1e-5 * scale
    simt = W.graph.gram(z, use_tensor_cores=False)
    assert (got - simt).abs().max().item() <= 3e-5 * scale


@pytest.mark.timeout(120)
def test_similarity_same_with_tensor_cores(W):
    """hm.similarity / gntk_homophily_ through the tensor-core Gram give the golden answers too."""
    import _golden as G
    z = G.load("syn_4000_0.2_1")
    n = int(z["in_n"])
    ei = z["in_edge_index"].astype(np.int64)
    labels = z["in_labels"]
    idx = torch.from_numpy(ei)
    A = torch.sparse_coo_tensor(idx, torch.ones(idx.shape[1]), (n, n)).coalesce().cuda()
    oh = torch.eye(int(labels.max()) + 1)[torch.from_numpy(labels)]
    old = (W.graph.USE_TENSOR_CORES, W.graph.KR_USE_TENSOR_CORES)
    try:
        W.graph.USE_TENSOR_CORES = W.graph.KR_USE_TENSOR_CORES = True
        got = float(W.homophily_metrics.similarity(oh, A, oh, hard=None, LP=1))
        got_h = float(W.homophily_metrics.similarity(oh, A, oh, hard=1, LP=1))
        kg, kx = W.homophily_metrics.gntk_homophily_(torch.from_numpy(z["in_features"]), A, z["in_gntk_sample"], 1)
    finally:
        W.graph.USE_TENSOR_CORES, W.graph.KR_USE_TENSOR_CORES = old
    assert abs(got - float(z["out_soft_las"])) <= 1.5 / n
    assert abs(got_h - float(z["out_hard_las"])) <= 1.5 / n
    for a, key in ((kg, "out_gntk_KG_l1"), (kx, "out_gntk_KX_l1")):
        ref = z[key]
        np.testing.assert_allclose(a.cpu().numpy(), ref, rtol=1e-4, atol=2e-5 * max(1.0, float(np.abs(ref).max())))
