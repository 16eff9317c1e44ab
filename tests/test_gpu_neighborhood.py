"""exact grid k-NN + density / occlusion mapping features on the GPU vs the
reference's NeighborhoodBasedMappingFeatures (tests/golden/neighborhood_features.npz) and the
brute-force oracle."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _setting(g, with_feat):
    from deepviewagg_b200.core.multimodal.image import ImageMapping, SameSettingImageData
    W, H, n_img = [int(v) for v in g["size"]]
    im = SameSettingImageData(pos=torch.zeros(n_img, 3), opk=torch.zeros(n_img, 3), ref_size=(W, H),
                              proj_upscale=1, downscale=1)
    im.mappings = ImageMapping.from_dense(g["pid"], g["iid"], g["pix"], g["feat"] if with_feat else None,
                                          num_points=g["pos"].shape[0])
    return im


@pytest.mark.parametrize("cell_size", [None, 0.05, 0.4, 3.0])
def test_knn_grid_matches_reference_neighbors(cell_size):
    from deepviewagg_b200.core.multimodal.mapping import knn_grid
    from oracle.neighborhood_oracle import knn_bruteforce
    g = load_golden("neighborhood_features")
    pos = g["pos"].cuda()
    nbr, d2 = knn_grid(pos, 20, cell_size=cell_size, return_dist2=True)
    assert torch.equal(nbr.cpu(), g["neighbors_k20"])            # incl. the duplicated point's index ties
    _, d2_ref = knn_bruteforce(g["pos"].numpy(), 20)
    assert np.array_equal(d2.cpu().numpy(), d2_ref)
    for k in (1, 7, 64):
        n2 = knn_grid(pos, k, cell_size=cell_size)
        want, _ = knn_bruteforce(g["pos"].numpy(), k)
        assert np.array_equal(n2.cpu().numpy(), want), k


@pytest.mark.parametrize("device_maps", ["cpu", "cuda"])
def test_neighborhood_features_vs_reference(device_maps):
    from deepviewagg_b200.core.multimodal.mapping import NeighborhoodBasedMappingFeatures
    g = load_golden("neighborhood_features")
    cases = (("klist", True, dict(k=[20, 5], voxel=0.05)), ("k7", False, dict(k=7)),
             ("density_only", False, dict(k=[4, 16], voxel=0.1, occlusion=False)),
             ("occlusion_only", True, dict(k=10, density=False)))
    for tag, with_feat, kw in cases:
        im = _setting(g, with_feat).to(device_maps)
        out = NeighborhoodBasedMappingFeatures(**kw)(g["pos"], im)
        got, want = out.mappings.features.cpu().numpy(), g[f"{tag}_features"].numpy()
        assert got.shape == want.shape and out.mappings.features.device.type == device_maps, tag
        fin = np.isfinite(want)
        assert (np.isfinite(got) == fin).all(), tag
        assert np.abs(got[fin] - want[fin]).max() <= 2e-7 * np.abs(want[fin]).max(), tag
        if "occlusion" in tag or tag == "k7":                     # occlusion columns are exact
            assert np.array_equal(got[:, -1], want[:, -1]), tag


def test_knn_grid_large_cloud_sampled_against_bruteforce():
    """300 k points on noisy surfaces + a dense clump + far outliers: 700 sampled queries against
    the brute-force order, all rows sorted, self first."""
    from deepviewagg_b200.core.multimodal.mapping import knn_grid
    gen = torch.Generator().manual_seed(3)
    n = 300_000
    uv = torch.rand(n, 2, generator=gen) * torch.tensor([40.0, 25.0])
    z = torch.where(torch.rand(n, generator=gen) < 0.6, 0.03 * torch.randn(n, generator=gen),
                    3.0 + 0.5 * torch.sin(uv[:, 0]) + 0.03 * torch.randn(n, generator=gen))
    pos = torch.cat([uv, z[:, None]], 1)
    pos[:5000] = torch.tensor([5.0, 5.0, 1.0]) + 0.01 * torch.randn(5000, 3, generator=gen)   # clump
    pos[5000:5010] = 500.0 + 100 * torch.rand(10, 3, generator=gen)                           # outliers
    k = 20
    pos_d = pos.cuda()
    knn_grid(pos_d, k)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    nbr, d2 = knn_grid(pos_d, k, return_dist2=True)
    t1.record()
    torch.cuda.synchronize()
    print(f"knn_grid: {n} points, k={k}: {t0.elapsed_time(t1):.1f} ms (grid build + search)")
    nbr, d2 = nbr.cpu(), d2.cpu()
    assert (nbr[:, 0] == torch.arange(n)).all() and (d2[:, 1:] >= d2[:, :-1]).all()
    q = torch.cat([torch.arange(0, 5010, 50), torch.randint(0, n, (600,), generator=gen)])
    p = pos.numpy().astype(np.float32)
    for i in q.tolist():
        d = p[i] - p
        dd = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        cand = np.argpartition(dd, 4 * k)[:4 * k]               # superset of the k nearest, then exact order
        cand = np.concatenate([cand, np.nonzero(dd <= dd[cand].max())[0]])
        cand = np.unique(cand)
        want = cand[np.lexsort((cand, dd[cand]))][:k]
        assert np.array_equal(nbr[i].numpy(), want), i
