"""Pin the oracle: oracle/pooling_oracle.py vs fixtures produced by the EXECUTED reference
(tests/golden/*.npz, oracle/make_golden.py) and vs the reference's own known-answer snippets
(pooling.py:913-921 etc., values recorded in SURVEY.md 8c). CPU only."""
import pytest
import torch

from conftest import load_golden, rel_err
from oracle import pooling_oracle as O

TOL = 1e-5  # same algorithm, same fp32 ops, same order -> only BLAS / reduction-order noise


def test_kat_softmax_values():
    g = load_golden("kat_softmax")
    exp = torch.tensor([0.011656231, 0.031684920, 0.086128540, 0.234121665, 0.636408627])
    exp_s = torch.tensor([0.067486435, 0.105545297, 0.165067390, 0.258156866, 0.403743982])
    out = O.segment_softmax_csr(g["src"], g["csr"])
    out_s = O.segment_softmax_csr(g["src"], g["csr"], scaling=True)
    assert torch.allclose(out[:5, 0], exp, atol=1e-7) and torch.allclose(out, g["out"], atol=1e-7)
    assert torch.allclose(out_s[:5, 0], exp_s, atol=1e-7) and torch.allclose(out_s, g["out_scaled"], atol=1e-7)
    em = O.segment_softmax_csr(torch.tensor([[1.], [2.], [3.]]), torch.tensor([0, 2, 2, 3]))
    assert torch.allclose(em.view(-1), torch.tensor([0.268941432, 0.731058598, 1.0]), atol=1e-7)
    assert torch.equal(O.gather_csr(torch.tensor([[1.], [2.], [3.]]), torch.tensor([0, 2, 2, 5])), g["gather"])
    gt = O.gating(torch.tensor([[-1., .5], [2., 0.]]), torch.ones(1, 2), torch.zeros(1, 2), 2)
    assert torch.allclose(gt, g["gating"], atol=1e-7)
    assert torch.allclose(gt, torch.tensor([[0, 0.462117165], [0.964027584, 0]]), atol=1e-7)
    assert O.group_sizes(10, 4) == [3, 3, 2, 2] == g["group_sizes_10_4"].tolist()
    assert O.group_sizes(512, 4) == [128] * 4 == g["group_sizes_512_4"].tolist()
    assert [O.nearest_power_of_2(x, 64) for x in (48, 96, 160, 288)] == [64, 128, 128, 256] == g["npow2"].tolist()


@pytest.mark.parametrize("tag", ["k7", "k32"])
def test_segment_primitives(tag):
    g = load_golden("segment_ops_" + tag)
    x, ptr = g["src"], g["ptr"]
    for red in ("sum", "mean", "max", "min"):
        xr = x.clone().requires_grad_(True)
        o = O.segment_csr(xr, ptr, reduce=red)
        assert rel_err(o, g[f"out_{red}"]) < TOL
        gr = torch.autograd.grad((o * g["w"]).sum(), xr)[0]
        assert rel_err(gr, g[f"grad_{red}"]) < TOL
        assert rel_err(O.segment_gather_csr(x, ptr, reduce=red), g[f"seg_gather_{red}"]) < TOL
    # empty segments reduce to 0 for every mode (pooling.py:870)
    empty = (ptr[1:] == ptr[:-1])
    assert empty.any()
    for red in ("sum", "mean", "max", "min"):
        assert (g[f"out_{red}"][empty] == 0).all()
    for s in (0, 1):
        xr = x.clone().requires_grad_(True)
        o = O.segment_softmax_csr(xr, ptr, scaling=bool(s))
        assert rel_err(o, g[f"softmax_{s}"]) < TOL
        gr = torch.autograd.grad((o * g["wv"]).sum(), xr)[0]
        assert (gr - g[f"softmax_grad_{s}"]).abs().max() < 1e-6
    assert torch.equal(O.gather_csr(g["gather_src"], ptr), g["gather_out"])


def _params(sd):
    """state_dict tensors; parameters (not BN buffers) require grad."""
    return {k: v.clone().requires_grad_(v.is_floating_point() and "running_" not in k)
            for k, v in sd.items()}


def _check_group(name, eval_mode=False):
    g = load_golden(name)
    kw = dict(g["kw"])
    G = kw["num_groups"]
    sd = _params(g["sd"])
    x_mod = g["x_mod"].clone().requires_grad_(True)
    x_map = g["x_map"].clone().requires_grad_(True)
    enc_kw = {k: kw[k] for k in ("pool", "fusion", "use_num", "use_min", "use_max") if k in kw}
    r = O.group_pool(sd, x_mod, x_map, g["ptr"], G, use_mod=kw.get("use_mod", False),
                     gating_on=kw.get("gating", True), group_scaling=kw.get("group_scaling", True),
                     map_encoder_name=kw.get("map_encoder", "DeepSetFeat"), training=not eval_mode, **enc_kw)
    assert rel_err(r["out"], g["out"]) < 2e-5, name
    if eval_mode:
        return
    assert rel_err(r["C"], g["last_C"]) < 2e-5 and rel_err(r["A"], g["last_A"]) < 2e-5
    names = ["x_mod", "x_map"] + [k for k in sd if sd[k].requires_grad and "param/" + k in g["grad"]]
    tensors = [x_mod, x_map] + [sd[k] for k in names[2:]]
    grads = torch.autograd.grad((r["out"] * g["w"]).sum(), tensors, allow_unused=True)
    for n, gr in zip(names, grads):
        ref = g["grad"][n if n in ("x_mod", "x_map") else "param/" + n]
        gr = torch.zeros_like(ref) if gr is None else gr
        assert (gr - ref).abs().max() <= 5e-5 * max(1.0, float(ref.abs().max())), (name, n)


@pytest.mark.parametrize("name", ["group_pool_toy", "group_pool_c64", "group_pool_usemod",
                                  "group_pool_g1_nogate", "group_pool_oddgroups", "group_pool_minmax"])
def test_group_pool_matches_reference(name):
    _check_group(name)
    _check_group(name + "_eval", eval_mode=True)


@pytest.mark.parametrize("name", ["qkv_pool_base", "qkv_pool_modqk"])
def test_qkv_pool_matches_reference(name):
    g = load_golden(name)
    kw = dict(g["kw"])
    sd = _params(g["sd"])
    x_main = g["x_main"].clone().requires_grad_(True)
    x_mod = g["x_mod"].clone().requires_grad_(True)
    x_map = g["x_map"].clone().requires_grad_(True)
    r = O.qkv_pool(sd, x_main, x_mod, x_map, g["ptr"], kw["num_groups"], nc_qk=kw["nc_qk"],
                   use_mod_q=kw.get("use_mod_q", False), use_mod_k=kw.get("use_mod_k", False),
                   group_scaling=kw.get("group_scaling", False), use_num=kw.get("use_num", False))
    assert rel_err(r["out"], g["out"]) < 2e-5
    assert rel_err(r["C"], g["last_C"]) < 2e-5 and rel_err(r["A"], g["last_A"]) < 2e-5
    grads = torch.autograd.grad((r["out"] * g["w"]).sum(), [x_main, x_mod, x_map])
    for n, gr in zip(("x_main", "x_mod", "x_map"), grads):
        ref = g["grad"][n]
        assert (gr - ref).abs().max() <= 5e-5 * max(1.0, float(ref.abs().max())), n


def test_simple_pools_and_fusion():
    g = load_golden("simple_pools")
    for mode in ("max", "mean", "min", "sum"):
        assert rel_err(O.bimodal_csr_pool(g["x_mod"], g["ptr"], mode), g["bimodal_" + mode]) < TOL
    for mode in ("max", "min"):
        for feat in (0, 5):
            assert torch.equal(O.heuristic_pool(g["x_mod"], g["x_map"], g["ptr"], feat, mode),
                               g[f"heuristic_{mode}_{feat}"])
    for mode in ("residual", "concatenation", "both", "modality"):
        assert torch.equal(O.bimodal_fusion(g["fusion_a"], g["fusion_b"], mode), g["fusion_" + mode])


@pytest.mark.parametrize("tag", ["half", "quarter"])
def test_sparse_interpolation_restatement_is_bit_exact(tag):
    """oracle/image_oracle.py vs the reference's sparse_interpolation (image.py:105-170) executed by
    make_golden; the torch restatement in core/multimodal/image.py (CPU tensors) as well."""
    import numpy as np
    from oracle.image_oracle import sparse_interpolation_pixels
    from deepviewagg_b200.core.multimodal.image import sparse_interpolation
    g = load_golden("sparse_interpolation")
    W, H, ds = [int(v) for v in g[f"{tag}_size"]]
    x, pix, batch = g[f"{tag}_x"], g[f"{tag}_pix"], g[f"{tag}_batch"]
    out = sparse_interpolation_pixels(x.numpy(), pix.numpy(), batch.numpy(), (W, H))
    assert np.array_equal(out, g[f"{tag}_out"].numpy())
    coords = (pix / (torch.tensor([[W, H]], dtype=torch.float) - 1))[:, [1, 0]]
    xt = x.clone().requires_grad_(True)
    o2 = sparse_interpolation(xt, coords, batch)
    assert torch.equal(o2.detach(), g[f"{tag}_out"])
    (gx,) = torch.autograd.grad((o2 * g[f"{tag}_w"]).sum(), [xt])
    assert rel_err(gx, g[f"{tag}_gx"]) < 1e-6


def test_neighborhood_features_restatement():
    """oracle/neighborhood_oracle.py vs the reference's NeighborhoodBasedMappingFeatures executed by
    make_golden (image.py:483-612): neighbours identical, occlusion bit-exact, density within one
    ulp of the scalar division (the reference divides by a Python scalar)."""
    import numpy as np
    from oracle.neighborhood_oracle import knn_bruteforce, neighborhood_features
    from deepviewagg_b200.core.multimodal.image import ImageMapping
    g = load_golden("neighborhood_features")
    pos = g["pos"].numpy()
    nbr, d2 = knn_bruteforce(pos, 20)
    assert np.array_equal(nbr, g["neighbors_k20"].numpy())
    assert (d2[:, 1:] >= d2[:, :-1]).all() and (d2[[10, 11, 12], :3] == 0).all()      # the duplicated point
    m = ImageMapping.from_dense(g["pid"], g["iid"], g["pix"], g["feat"], num_points=len(pos))
    cases = (("klist", dict(k_list=[20, 5], voxel=0.05), 8), ("k7", dict(k_list=[7]), 0),
             ("density_only", dict(k_list=[4, 16], voxel=0.1, occlusion=False), 0),
             ("occlusion_only", dict(k_list=[10], density=False), 8))
    for tag, kw, n_old in cases:
        got = neighborhood_features(pos, nbr, m.pointers.numpy(), m.images.numpy(), **kw)
        want = g[f"{tag}_features"].numpy()
        if n_old:                                                   # existing columns are kept in front
            assert np.array_equal(want[:, :n_old], m.features.numpy()), tag
        want = want[:, n_old:]
        assert got.shape == want.shape, tag
        fin = np.isfinite(want)
        assert (np.isfinite(got) == fin).all(), tag
        assert np.abs(got[fin] - want[fin]).max() <= 2e-7 * np.abs(want[fin]).max(), tag
