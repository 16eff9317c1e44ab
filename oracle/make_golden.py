"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE (oracle; test infrastructure).

Run in the build container only (needs /root/reference):
    PYTORCH_JIT=0 python -m oracle.make_golden
The reference's own modules (loaded by file path, oracle/ref_loader.py) are run on seeded inputs;
inputs, parameters (state_dict), outputs and autograd gradients are committed as small fixtures.
torch_scatter is the documented-semantics stand-in (oracle/scatter_standin.py) -- stated in the
fixture metadata.  The GPU box has no /root/reference: tests only read the committed fixtures.
"""
import os
import sys
import types

os.environ.setdefault("PYTORCH_JIT", "0")

import numpy as np  # noqa: E402
import torch  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
META = "reference=DeepViewAgg@41543bc; torch_scatter=oracle/scatter_standin.py (documented semantics); torch=%s" % torch.__version__


def save(name, **arrays):
    conv = {}
    for k, v in arrays.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        conv[k] = np.asarray(v)
    conv["__meta__"] = np.array(META)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **conv)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def ragged_ptr(gen, n, mean, p_empty=0.15, max_mult=4):
    counts = torch.poisson(torch.full((n,), float(mean)), generator=gen).clamp(0, max_mult * mean).long()
    counts[torch.rand(n, generator=gen) < p_empty] = 0
    return torch.cat([torch.zeros(1, dtype=torch.long), counts.cumsum(0)])


def sd_arrays(module, prefix="sd/"):
    return {prefix + k: v.clone() for k, v in module.state_dict().items()}


def grads_of(out_scalar, tensors, names, prefix):
    gs = torch.autograd.grad(out_scalar, tensors, allow_unused=True)
    return {prefix + n: (g if g is not None else torch.zeros_like(t)) for n, g, t in zip(names, gs, tensors)}


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = ref_loader.load_reference()
    P = ref.pooling

    # ---- known-answer snippets the reference itself holds (pooling.py:913-921) -----------------
    src = torch.arange(15).float().view(-1, 1).repeat_interleave(2, dim=1)
    csr = torch.LongTensor([0, 5, 10, 15])
    save("kat_softmax", src=src, csr=csr, out=P.segment_softmax_csr(src, csr),
         out_scaled=P.segment_softmax_csr(src.clone(), csr, scaling=True),
         empty_mid=P.segment_softmax_csr(torch.tensor([[1.], [2.], [3.]]), torch.LongTensor([0, 2, 2, 3])),
         gather=P.gather_csr(torch.tensor([[1.], [2.], [3.]]), torch.LongTensor([0, 2, 2, 5])),
         gating=P.Gating(2)(torch.tensor([[-1., .5], [2., 0.]])),
         group_sizes_10_4=P.group_sizes(10, 4), group_sizes_512_4=P.group_sizes(512, 4),
         npow2=np.array([P.nearest_power_of_2(x, 64) for x in (48, 96, 160, 288)]))

    # ---- segment primitives on ragged data (ties included: post-ReLU style zeros) --------------
    gen = torch.Generator().manual_seed(1234)
    n = 129
    ptr = ragged_ptr(gen, n, 5)
    V = int(ptr[-1])
    for K, tag in ((7, "k7"), (32, "k32")):
        x = torch.randn(V, K, generator=gen)
        x = torch.where(torch.rand(V, K, generator=gen) < 0.3, torch.zeros(()), x).clamp(min=-0.5)
        arrays = dict(src=x, ptr=ptr)
        w = torch.randn(n, K, generator=gen)
        for red in ("sum", "mean", "max", "min"):
            xr = x.clone().requires_grad_(True)
            o = P.segment_csr(xr, ptr, reduce=red)
            arrays[f"out_{red}"] = o
            arrays[f"grad_{red}"] = torch.autograd.grad((o * w).sum(), xr)[0]
            arrays[f"seg_gather_{red}"] = P.segment_gather_csr(x, ptr, reduce=red)
        arrays["w"] = w
        wv = torch.randn(V, K, generator=gen)
        for scaling in (False, True):
            xr = x.clone().requires_grad_(True)
            o = P.segment_softmax_csr(xr, ptr, scaling=scaling)
            arrays[f"softmax_{int(scaling)}"] = o
            arrays[f"softmax_grad_{int(scaling)}"] = torch.autograd.grad((o * wv).sum(), xr)[0]
        arrays["wv"] = wv
        sr = torch.randn(n, K, generator=gen).requires_grad_(True)
        o = P.gather_csr(sr, ptr)
        arrays["gather_src"] = sr
        arrays["gather_out"] = o
        arrays["gather_grad"] = torch.autograd.grad((o * wv).sum(), sr)[0]
        save(f"segment_ops_{tag}", **arrays)

    # ---- pools -----------------------------------------------------------------------------------
    def run_group(name, N, mean_v, in_mod, out_mod, G, seed, **kw):
        gen = torch.Generator().manual_seed(seed)
        torch.manual_seed(seed)
        ptr = ragged_ptr(gen, N, mean_v)
        V = int(ptr[-1])
        m = P.GroupBimodalCSRPool(in_map=8, in_mod=in_mod, out_mod=out_mod, num_groups=G,
                                  save_last=True, **kw)
        with torch.no_grad():  # non-trivial BN affine + gating so that every gradient is exercised
            for k, p in m.named_parameters():
                if "batch_norm" in k or k.startswith("G."):
                    p.add_(0.3 * torch.randn(p.shape, generator=gen))
        m.train()
        x_mod = torch.randn(V, in_mod, generator=gen).relu().requires_grad_(True)
        x_map = torch.rand(V, 8, generator=gen).requires_grad_(True)
        w = torch.randn(N, m.out_mod, generator=gen)
        sd0 = sd_arrays(m)
        out = m(None, x_mod, x_map, ptr)
        params = dict(m.named_parameters())
        g = grads_of((out * w).sum(), [x_mod, x_map] + list(params.values()),
                     ["x_mod", "x_map"] + ["param/" + k for k in params], "grad/")
        save(name, ptr=ptr, x_mod=x_mod, x_map=x_map, w=w, out=out, last_C=m._last_C, last_A=m._last_A,
             last_G=m._last_G if m.G else np.zeros(0), kw=np.array(repr(dict(
                 in_map=8, in_mod=in_mod, out_mod=out_mod, num_groups=G, **kw))), **sd0, **g)
        # eval-mode forward with the post-step running stats
        m.eval()
        with torch.no_grad():
            out_eval = m(None, x_mod, x_map, ptr)
        save(name + "_eval", ptr=ptr, x_mod=x_mod, x_map=x_map, out=out_eval, **sd_arrays(m),
             kw=np.array(repr(dict(in_map=8, in_mod=in_mod, out_mod=out_mod, num_groups=G, **kw))))

    run_group("group_pool_toy", 1000, 4, 8, 8, 4, 11, use_num=True)             # config #0 shape
    run_group("group_pool_c64", 160, 8, 64, 64, 4, 12, use_num=True)             # S3DIS-like
    run_group("group_pool_usemod", 200, 6, 24, 32, 4, 13, use_num=True, use_mod=True)
    run_group("group_pool_g1_nogate", 150, 5, 20, 20, 1, 14, gating=False, group_scaling=False,
              map_encoder="MLPSetFeat")
    run_group("group_pool_oddgroups", 150, 5, 10, 10, 4, 15, use_num=True, pool="max_mean",
              fusion="both")                                                      # group sizes 3,3,2,2
    run_group("group_pool_minmax", 120, 5, 12, 12, 3, 16, map_encoder="MinMaxDiffSetFeat",
              use_num=True)                                                       # G=3: unfused path

    def run_qkv(name, N, mean_v, in_main, in_mod, G, D, seed, **kw):
        gen = torch.Generator().manual_seed(seed)
        torch.manual_seed(seed)
        ptr = ragged_ptr(gen, N, mean_v)
        V = int(ptr[-1])
        m = P.QKVBimodalCSRPool(in_main=in_main, in_map=8, in_mod=in_mod, num_groups=G, nc_qk=D,
                                save_last=True, **kw)
        with torch.no_grad():
            for k, p in m.named_parameters():
                if "batch_norm" in k or k.startswith("G."):
                    p.add_(0.3 * torch.randn(p.shape, generator=gen))
        m.train()
        x_main = torch.randn(N, in_main, generator=gen).requires_grad_(True)
        x_mod = torch.randn(V, in_mod, generator=gen).relu().requires_grad_(True)
        x_map = torch.rand(V, 8, generator=gen).requires_grad_(True)
        w = torch.randn(N, m.out_mod, generator=gen)
        sd0 = sd_arrays(m)
        out = m(x_main, x_mod, x_map, ptr)
        params = dict(m.named_parameters())
        g = grads_of((out * w).sum(), [x_main, x_mod, x_map] + list(params.values()),
                     ["x_main", "x_mod", "x_map"] + ["param/" + k for k in params], "grad/")
        save(name, ptr=ptr, x_main=x_main, x_mod=x_mod, x_map=x_map, w=w, out=out, last_C=m._last_C,
             last_A=m._last_A, last_G=m._last_G if m.G else np.zeros(0),
             kw=np.array(repr(dict(in_main=in_main, in_map=8, in_mod=in_mod, num_groups=G, nc_qk=D, **kw))),
             **sd0, **g)

    run_qkv("qkv_pool_base", 300, 6, 10, 32, 4, 8, 21, use_num=True)
    run_qkv("qkv_pool_modqk", 150, 5, 12, 16, 4, 4, 22, use_num=True, use_mod_q=True, use_mod_k=True,
            group_scaling=True)

    # ---- BimodalCSRPool / Heuristic / fusion ------------------------------------------------------
    gen = torch.Generator().manual_seed(31)
    ptr = ragged_ptr(gen, 200, 4)
    V = int(ptr[-1])
    x_mod = torch.randn(V, 24, generator=gen).relu()
    x_map = torch.rand(V, 8, generator=gen)
    arrays = dict(ptr=ptr, x_mod=x_mod, x_map=x_map)
    for mode in ("max", "mean", "min", "sum"):
        arrays["bimodal_" + mode] = P.BimodalCSRPool(mode=mode)(None, x_mod, None, ptr)
    for mode in ("max", "min"):
        for feat in (0, 5):
            arrays[f"heuristic_{mode}_{feat}"] = P.HeuristicBimodalCSRPool(mode=mode, feat=feat)(
                None, x_mod, x_map, ptr)
    a = torch.randn(200, 24, generator=gen)
    b = torch.randn(200, 24, generator=gen)
    arrays["fusion_a"], arrays["fusion_b"] = a, b
    for mode in ref.fusion.BimodalFusion.MODES:
        arrays["fusion_" + mode] = ref.fusion.BimodalFusion(mode)(a, b)
    save("simple_pools", **arrays)

    make_integer_golden(ref)
    make_branch_golden(ref)
    make_branch_golden(ref, interpolate=True, name="unimodal_branch_interp")
    make_interp_golden(ref)
    make_neighborhood_golden(ref)
    make_camera_golden(ref)
    make_visibility_model_golden(ref)
    make_block_down_golden(ref)
    make_image_ops_golden(ref)


def make_camera_golden(ref):
    """numba camera_projection_cpu for the pinhole (scannet, kitti360_perspective) and fisheye
    (kitti360_fisheye) cameras (visibility.py:219-339, 478-538)."""
    import numpy as np
    vis = ref.visibility
    gen = torch.Generator().manual_seed(17)
    n = 6000
    xyz = (torch.rand(n, 3, generator=gen) - 0.5) * torch.tensor([10., 10., 4.])

    def rot(ax, ay, az):
        cx, sx, cy, sy, cz, sz = np.cos(ax), np.sin(ax), np.cos(ay), np.sin(ay), np.cos(az), np.sin(az)
        rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
        ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
        rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
        return rz @ ry @ rx
    cam_pos = np.array([0.4, -0.3, 0.2])
    c2w = np.eye(4)
    c2w[:3, :3] = rot(-1.4, 0.1, 0.5)
    c2w[:3, 3] = cam_pos
    intr = np.eye(4, dtype=np.float32)
    intr[0, 0], intr[1, 1], intr[0, 2], intr[1, 2] = 250.0, 250.0, 159.5, 119.5
    fish = np.array([2.2134, 0.016798, 0.7572, 1336.3, 1335.7, 716.94, 705.76], dtype=np.float32)
    cases = {
        # scannet: img_extrinsic is world->camera (inverted inside), kitti360: camera->world
        "scannet": dict(ext=torch.from_numpy(np.linalg.inv(c2w)).float(), size=(320, 240), crop=(0, 0),
                        pin=torch.from_numpy(intr), fish=None),
        "kitti360_perspective": dict(ext=torch.from_numpy(c2w).float(), size=(320, 240), crop=(10, 20),
                                     pin=torch.from_numpy(intr), fish=None),
        "kitti360_fisheye": dict(ext=torch.from_numpy(c2w).float(), size=(1400, 1400), crop=(0, 0), pin=None,
                                 fish=torch.from_numpy(fish)),
    }
    for cam, c in cases.items():
        idx, dist, xp, yp = vis.camera_projection_cpu(
            xyz, torch.from_numpy(cam_pos).float(), img_intrinsic_pinhole=c["pin"],
            img_intrinsic_fisheye=c["fish"], img_extrinsic=c["ext"], img_size=c["size"], crop_top=c["crop"][0],
            crop_bottom=c["crop"][1], r_max=8, r_min=0.3, camera=cam)
        extra = {}
        if cam == "kitti360_fisheye":   # fisheye splat + z-buffer (visibility.py:876-953, 1126-1130)
            sp = vis.fisheye_splat_cpu(xp.numpy(), yp.numpy(), xyz[idx].numpy(), c["ext"].numpy(), c["fish"].numpy(),
                                       img_size=c["size"], crop_top=0, crop_bottom=0, voxel=0.05, k_swell=1.0,
                                       d_swell=1000)
            extra["splat"] = sp
            for exact in (False, True):
                i2, xpix, ypix = vis.visibility_from_splatting_cpu(
                    xp, yp, dist, xyz[idx], img_extrinsic=c["ext"], img_intrinsic_fisheye=c["fish"],
                    img_size=c["size"], voxel=0.05, k_swell=1.0, d_swell=1000, exact=exact, camera=cam)
                extra[f"vis_idx_{int(exact)}"], extra[f"vis_x_{int(exact)}"], extra[f"vis_y_{int(exact)}"] = i2, xpix, ypix
        if cam == "scannet":
            extra["c2w"] = np.linalg.inv(np.ascontiguousarray(c["ext"].numpy()))  # numba's own inverse (f32)
        save(f"camera_{cam}", xyz=xyz, img_xyz=torch.from_numpy(cam_pos).float(), ext=c["ext"],
             pin=c["pin"] if c["pin"] is not None else np.zeros(0), fish=c["fish"] if c["fish"] is not None else np.zeros(0),
             size=np.array(c["size"]), crop=np.array(c["crop"]), r=np.array([0.3, 8.0]),
             proj_idx=idx, dist=dist, x_proj=xp, y_proj=yp, **extra)


def make_visibility_model_golden(ref):
    """Z4 / Z5: the assembled dict of SplattingVisibility.__call__ (visibility.py:1677-1776) --
    idx, x, y, depth and the [k, 6] viewing-condition features of postprocess_features
    (:1548-1582: normalised depth, linearity, planarity, scattering, |cos(view, normal)|,
    normalised pixel height) -- executed on the CPU (numba) path for the three camera families."""
    vis = ref.visibility
    gen = torch.Generator().manual_seed(29)
    n = 7000
    xyz = (torch.rand(n, 3, generator=gen) - 0.5) * torch.tensor([12., 12., 4.])
    geo = torch.rand(n, 3, generator=gen)
    normals = torch.nn.functional.normalize(torch.randn(n, 3, generator=gen), dim=1)
    img_xyz = torch.tensor([0.3, -0.2, 0.1])
    c2w = np.eye(4)
    a, b, c = -1.4, 0.1, 0.5
    rx = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
    ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
    rz = np.array([[np.cos(c), -np.sin(c), 0], [np.sin(c), np.cos(c), 0], [0, 0, 1]])
    c2w[:3, :3] = rz @ ry @ rx
    c2w[:3, 3] = img_xyz.numpy()
    intr = np.eye(4, dtype=np.float32)
    intr[0, 0], intr[1, 1], intr[0, 2], intr[1, 2] = 250.0, 250.0, 159.5, 119.5
    fish = torch.tensor([2.2134, 0.016798, 0.7572, 1336.3, 1335.7, 716.94, 705.76])
    cases = {
        "equirect_exact": (dict(voxel=0.05, exact=True, img_size=(512, 256), crop_top=16, crop_bottom=24, r_max=8,
                                r_min=0.5, camera="s3dis_equirectangular"),
                           dict(img_opk=torch.tensor([0.05, -0.1, 0.7]))),
        "equirect_splat": (dict(voxel=0.05, exact=False, img_size=(512, 256), crop_top=0, crop_bottom=0, r_max=8,
                                r_min=0.5, camera="s3dis_equirectangular"),
                           dict(img_opk=torch.tensor([0.05, -0.1, 0.7]))),
        "scannet": (dict(voxel=0.03, exact=True, img_size=(320, 240), r_max=8, r_min=0.3, camera="scannet"),
                    dict(img_extrinsic=torch.from_numpy(np.linalg.inv(c2w)).float(),
                         img_intrinsic_pinhole=torch.from_numpy(intr))),
        "kitti360_fisheye": (dict(voxel=0.05, exact=True, img_size=(1400, 1400), r_max=8, r_min=0.3,
                                  camera="kitti360_fisheye"),
                             dict(img_extrinsic=torch.from_numpy(c2w).float(), img_intrinsic_fisheye=fish)),
    }
    for tag, (ctor, call) in cases.items():
        model = vis.SplattingVisibility(**ctor)
        out = model(xyz, img_xyz, linearity=geo[:, 0], planarity=geo[:, 1], scattering=geo[:, 2],
                    normals=normals, **call)
        arrays = dict(xyz=xyz, img_xyz=img_xyz, geo=geo, normals=normals,
                      ctor_keys=np.array(list(ctor.keys())), **{"ctor/" + k: np.asarray(v) for k, v in ctor.items()},
                      **{"call/" + k: v for k, v in call.items()},
                      **{"out/" + k: v for k, v in out.items()})
        print(tag, {k: tuple(v.shape) for k, v in out.items()})
        save("visibility_model_" + tag, **arrays)


def toy_settings(gen, n_points, specs, F=8):
    """Dense (point, image, pixel, feature) mappings per setting: one pixel per view ('exact'
    splatting, the only mode the shipped configs use)."""
    out = []
    for (W, H, n_img, mean_v) in specs:
        counts = torch.poisson(torch.full((n_points,), float(mean_v)), generator=gen).clamp(0, n_img).long()
        pid = torch.arange(n_points).repeat_interleave(counts)
        iid = torch.cat([torch.randperm(n_img, generator=gen)[:int(c)] for c in counts]) if pid.numel() else pid
        pix = torch.stack([torch.randint(0, W, (pid.numel(),), generator=gen),
                           torch.randint(0, H, (pid.numel(),), generator=gen)], 1).short()
        feat = torch.rand(pid.numel(), F, generator=gen)
        out.append(dict(W=W, H=H, n_img=n_img, pid=pid, iid=iid, pix=pix, feat=feat))
    return out


def make_branch_golden(ref, interpolate=False, name="unimodal_branch_toy"):
    """config #0 ("toy: 1k points x 4 views, CPU forward through the DeepViewAgg module"): the
    reference's UnimodalBranch (modules.py:249-566) on a two-setting ImageData, atomic max pool,
    GroupBimodalCSRPool view pool, concatenation fusion; forward + gradients."""
    import numpy as np
    P, I, M = ref.pooling, ref.image, ref.modules
    gen = torch.Generator().manual_seed(4242)
    torch.manual_seed(4242)
    N, C3, C = 1000, 12, 16
    specs = [(64, 32, 3, 2.2), (48, 48, 2, 1.6)]
    xs_shape = [(3, C, 16, 32), (2, C, 48, 48)]          # setting 0 at half resolution (downscale 2)
    settings = toy_settings(gen, N, specs)
    ims, xs, arrays = [], [], {}
    for s, (st, shp) in enumerate(zip(settings, xs_shape)):
        im = I.SameSettingImageData(path=np.array([f"img_{s}_{i}" for i in range(st["n_img"])]),
                                    pos=torch.zeros(st["n_img"], 3), opk=torch.zeros(st["n_img"], 3),
                                    ref_size=(st["W"], st["H"]), proj_upscale=1, downscale=1)
        im.mappings = I.ImageMapping.from_dense(st["pid"], st["iid"], st["pix"], st["feat"], num_points=N)
        x = torch.randn(shp, generator=gen).relu().requires_grad_(True)
        im.x = x
        ims.append(im)
        xs.append(x)
        for k in ("pid", "iid", "pix", "feat"):
            arrays[f"s{s}_{k}"] = st[k]
        arrays[f"s{s}_x"] = x
        arrays[f"s{s}_size"] = np.array([st["W"], st["H"], st["n_img"]])
    mod = I.ImageData(ims)
    view_pool = P.GroupBimodalCSRPool(in_map=8, in_mod=C, num_groups=4, use_num=True)
    with torch.no_grad():
        for k, p in view_pool.named_parameters():
            if "batch_norm" in k or k.startswith("G."):
                p.add_(0.3 * torch.randn(p.shape, generator=gen))
    branch = M.UnimodalBranch(None, P.BimodalCSRPool(mode="max"), view_pool, ref.fusion.BimodalFusion("concatenation"),
                              interpolate=interpolate)
    branch.train()
    x_3d = torch.randn(N, C3, generator=gen).requires_grad_(True)
    w = torch.randn(N, C3 + C, generator=gen)
    sd0 = {"sd/" + k: v.clone() for k, v in view_pool.state_dict().items()}
    out = branch({"x_3d": x_3d, "x_seen": None, "modalities": {"image": mod}}, "image")
    params = dict(view_pool.named_parameters())
    gs = torch.autograd.grad((out["x_3d"] * w).sum(), [x_3d] + xs + list(params.values()), allow_unused=True)
    names = ["x_3d", "s0_x", "s1_x"] + ["param/" + k for k in params]
    grads = {"grad/" + n: (g if g is not None else torch.zeros(1)) for n, g in zip(names, gs)}
    save(name, x_3d=x_3d, w=w, out=out["x_3d"], x_seen=out["x_seen"],
         csr=mod.view_cat_csr_indexing, n_points=np.array(N), **arrays, **sd0, **grads)


class _StubSampler:
    last_idx = None


class _PickBlock(torch.nn.Module):
    """Dense 3D block with a `.sampler` (modules.py:131-141): keeps the rows `idx` of its input."""

    def __init__(self, idx):
        super().__init__()
        self.sampler = _StubSampler()
        self.idx = idx

    def forward(self, x):
        self.sampler.last_idx = self.idx
        return x[self.idx]


class _FakeCoordsManager:
    def __init__(self, src, target):
        self.src, self.target = src, target

    def get_coords_map(self, stride_in, stride_out):
        return self.src, self.target


class _FakeSparseTensor:
    """The three attributes forward_3d_block_down reads from a MinkowskiEngine tensor
    (modules.py:146-158): F, tensor_stride, coords_man.get_coords_map."""

    def __init__(self, F, stride, coords_man):
        self.F, self.tensor_stride, self.coords_man = F, [stride], coords_man


class _StridedBlock(torch.nn.Module):
    """Fake strided sparse conv: parent feature = mean of its children's, stride doubles."""

    def __init__(self, parent):
        super().__init__()
        self.parent = parent

    def forward(self, x):
        n_out = int(self.parent.max()) + 1
        F = torch.zeros(n_out, x.F.shape[1]).index_add_(0, self.parent, x.F)
        cnt = torch.zeros(n_out).index_add_(0, self.parent, torch.ones(self.parent.numel()))
        return _FakeSparseTensor(F / cnt.clamp(min=1).unsqueeze(1), x.tensor_stride[0] * 2, x.coords_man)


def _dump_mappings(mod, prefix):
    out = {}
    for s, im in enumerate(mod):
        m = im.mappings
        out[f"{prefix}s{s}_pointers"] = m.pointers
        out[f"{prefix}s{s}_images"] = m.images
        out[f"{prefix}s{s}_atomic_pointers"] = m.values[1].pointers
        out[f"{prefix}s{s}_pixels"] = m.pixels
        out[f"{prefix}s{s}_features"] = m.features
        out[f"{prefix}s{s}_num_views"] = np.array(im.num_views)
    return out


def make_block_down_golden(ref):
    """U2: MultimodalBlockDown.forward_3d_block_down (modules.py:101-236) EXECUTED on a non-Identity
    block: (a) a dense block with a sampler -> 'pick' re-indexing of x_seen and of every setting's
    mappings; (b) a strided sparse block -> child->parent index from the coordinate map, x_seen
    scatter (:225) and ImageData.select_points(idx, 'merge') (image.py:2211-2273)."""
    I, M = ref.image, ref.modules
    gen = torch.Generator().manual_seed(9090)
    N, C3 = 600, 6
    specs = [(64, 32, 3, 2.2), (48, 48, 2, 1.6)]
    settings = toy_settings(gen, N, specs)

    def build():
        ims = []
        for s, st in enumerate(settings):
            im = I.SameSettingImageData(path=np.array([f"img_{s}_{i}" for i in range(st["n_img"])]),
                                        pos=torch.zeros(st["n_img"], 3), opk=torch.zeros(st["n_img"], 3),
                                        ref_size=(st["W"], st["H"]), proj_upscale=1, downscale=1)
            im.mappings = I.ImageMapping.from_dense(st["pid"], st["iid"], st["pix"], st["feat"], num_points=N)
            ims.append(im)
        return I.ImageData(ims)

    arrays = {"n_points": np.array(N)}
    for s, st in enumerate(settings):
        for k in ("pid", "iid", "pix", "feat"):
            arrays[f"s{s}_{k}"] = st[k]
        arrays[f"s{s}_size"] = np.array([st["W"], st["H"], st["n_img"]])
    x_3d = torch.randn(N, C3, generator=gen)
    x_seen = torch.rand(N, generator=gen) < 0.6
    arrays["x_3d"], arrays["x_seen"] = x_3d, x_seen

    # (a) 'pick': N indices with repetitions (the reference compares against arange(N), so the
    # sampler must return as many indices as there are input points, modules.py:133-140)
    pick_idx = torch.randint(0, N, (N,), generator=gen)
    d = M.MultimodalBlockDown.forward_3d_block_down(
        {"x_3d": x_3d.clone(), "x_seen": x_seen.clone(), "modalities": {"image": build()}}, _PickBlock(pick_idx))
    arrays["pick_idx"] = pick_idx
    arrays["pick_x_3d"], arrays["pick_x_seen"] = d["x_3d"], d["x_seen"]
    arrays.update(_dump_mappings(d["modalities"]["image"], "pick_"))

    # (b) 'merge' through a fake MinkowskiEngine tensor: children in shuffled order in the coords map
    n_out = 170
    parent = torch.randint(0, n_out, (N,), generator=gen)
    parent[:n_out] = torch.arange(n_out)                      # every parent voxel has a child
    src = torch.randperm(N, generator=gen)
    cm = _FakeCoordsManager(src, parent[src])
    M.me = types.SimpleNamespace(SparseTensor=_FakeSparseTensor)
    try:
        d = M.MultimodalBlockDown.forward_3d_block_down(
            {"x_3d": _FakeSparseTensor(x_3d.clone(), 1, cm), "x_seen": x_seen.clone(),
             "modalities": {"image": build()}}, _StridedBlock(parent))
    finally:
        M.me = None
    arrays["merge_parent"], arrays["merge_src"] = parent, src
    arrays["merge_x_3d"] = d["x_3d"].F
    arrays["merge_x_seen"] = d["x_seen"] > 0 if d["x_seen"].dtype != torch.bool else d["x_seen"]
    arrays.update(_dump_mappings(d["modalities"]["image"], "merge_"))
    save("block_down", **arrays)


def make_image_ops_golden(ref):
    """I3 / I4: ImageMapping.select_images (image.py:2029-2093), select_views (:2095-2165), crop (:2279-2342),
    downscale_images / upscale_images (:1916-2027) executed on the mapping of the image_mapping fixture."""
    lex, image = ref.lex, ref.image
    gen = torch.Generator().manual_seed(77)                  # same inputs as make_integer_golden
    n_points, n_items = 500, 4000
    pid = torch.randint(0, n_points, (n_items,), generator=gen)
    iid = torch.randint(0, 5, (n_items,), generator=gen)
    pix = torch.randint(0, 64, (n_items, 2), generator=gen).short()
    feat = torch.rand(n_items, 3, generator=gen)
    keep = lex.lexargunique(pid, iid, pix[:, 0].long(), pix[:, 1].long())
    pid, iid, pix, feat = pid[keep], iid[keep], pix[keep], feat[keep]
    # a FRESH mapping per operation: the reference's clone() is shallow (csr.py:147-156) and upscale_images
    # writes into the nested pixel CSR it shares with the original (image.py:2024), so results would depend on
    # the order of the calls
    fresh = lambda: image.ImageMapping.from_dense(pid.clone(), iid.clone(), pix.clone(), feat.clone(),  # noqa: E731
                                                  num_points=n_points + 7)
    m = fresh()
    gen2 = torch.Generator().manual_seed(99)
    arrays = dict(point_ids=pid, image_ids=iid, pixels=pix, features=feat, num_points=np.array(n_points + 7))

    def dump(tag, mm):
        arrays[f"{tag}_pointers"] = mm.pointers
        arrays[f"{tag}_images"] = mm.images
        arrays[f"{tag}_atomic_pointers"] = mm.values[1].pointers
        arrays[f"{tag}_pixels"] = mm.pixels
        arrays[f"{tag}_features"] = mm.features

    img_idx = torch.tensor([3, 0, 4])
    arrays["img_idx"] = img_idx
    dump("select_images", fresh().select_images(img_idx))
    view_mask = (torch.rand(m.num_items, generator=gen2) < 0.6) & (m.images != 2)   # image 2 disappears: renumbering
    arrays["view_mask"] = view_mask
    mv, seen_images = fresh().select_views(view_mask)               # (mapping, indices of the images still seen)
    arrays["select_views_img_idx"] = seen_images
    dump("select_views", mv)
    crop_size = (40, 32)
    crop_offsets = torch.tensor([[3, 5], [10, 0], [0, 20], [24, 30], [7, 7]])
    arrays["crop_size"] = np.array(crop_size)
    arrays["crop_offsets"] = crop_offsets
    dump("crop", fresh().crop(crop_size, crop_offsets))
    dump("down4", fresh().downscale_images(4))
    dump("up2", fresh().upscale_images(2))
    dump("up2_nocenter", fresh().upscale_images(2, center=False))
    save("image_ops", **arrays)


def make_interp_golden(ref):
    """sparse_interpolation (image.py:105-170) through get_mapped_features(interpolate=True)
    (image.py:1278-1283): pixels at the mapping resolution, maps at 1/2 and 1/4 of it, every
    border / corner pixel included; output and gradient w.r.t. the maps."""
    import numpy as np
    I = ref.image
    gen = torch.Generator().manual_seed(777)
    arrays = {}
    for tag, (W, H, ds, B, C) in {"half": (64, 32, 2, 3, 8), "quarter": (96, 64, 4, 2, 5)}.items():
        n = 4000
        pix = torch.stack([torch.randint(0, W, (n,), generator=gen), torch.randint(0, H, (n,), generator=gen)], 1)
        edge = torch.tensor([[0, 0], [W - 1, 0], [0, H - 1], [W - 1, H - 1], [W // 2, 0], [0, H // 2],
                             [W - 1, H // 2], [W // 2, H - 1]])
        pix = torch.cat([edge, pix]).long()
        batch = torch.randint(0, B, (pix.shape[0],), generator=gen)
        x = torch.randn(B, C, H // ds, W // ds, generator=gen).requires_grad_(True)
        resolution = torch.Tensor([[W, H]])
        coords = (pix / (resolution - 1))[:, [1, 0]]
        out = I.sparse_interpolation(x, coords, batch)
        w = torch.randn(out.shape, generator=gen)
        (gx,) = torch.autograd.grad((out * w).sum(), [x])
        arrays.update({f"{tag}_pix": pix, f"{tag}_batch": batch, f"{tag}_x": x, f"{tag}_out": out,
                       f"{tag}_w": w, f"{tag}_gx": gx, f"{tag}_size": np.array([W, H, ds])})
    save("sparse_interpolation", **arrays)


def make_neighborhood_golden(ref):
    """NeighborhoodBasedMappingFeatures (core/data_transform/multimodal/image.py:431-612) run by
    the reference on a noisy two-plane cloud; the KeOps argKmin is the dense stand-in of
    oracle/ref_loader.py (exact search, ties by index).  Cases: k list with existing features,
    single k without features, density only / occlusion only."""
    import numpy as np
    T = ref_loader.load_transforms()
    I = ref.image
    gen = torch.Generator().manual_seed(99)
    N, n_img = 2500, 6
    uv = torch.rand(N, 2, generator=gen) * torch.tensor([8.0, 5.0])
    z = torch.where(torch.rand(N, generator=gen) < 0.7, 0.02 * torch.randn(N, generator=gen),
                    1.5 + 0.02 * torch.randn(N, generator=gen))
    pos = torch.cat([uv, z[:, None]], 1)
    pos[10] = pos[11]                                       # exact duplicates: zero distance, index ties
    pos[12] = pos[11]
    (st,) = toy_settings(gen, N, [(64, 48, n_img, 2.5)])
    arrays = dict(pos=pos, pid=st["pid"], iid=st["iid"], pix=st["pix"], feat=st["feat"],
                  size=np.array([64, 48, n_img]))

    def run(tag, with_feat, **kw):
        im = I.SameSettingImageData(path=np.array([f"img_{i}" for i in range(n_img)]), pos=torch.zeros(n_img, 3),
                                    opk=torch.zeros(n_img, 3), ref_size=(64, 48), proj_upscale=1, downscale=1)
        im.mappings = I.ImageMapping.from_dense(st["pid"], st["iid"], st["pix"], st["feat"] if with_feat else None,
                                                num_points=N)
        tr = T.NeighborhoodBasedMappingFeatures(use_cuda=False, use_faiss=False, **kw)
        _, out = tr(ref_loader.load_reference().Data(pos=pos), im)
        arrays[f"{tag}_features"] = out.mappings.features

    run("klist", True, k=[20, 5], voxel=0.05)
    run("k7", False, k=7)
    run("density_only", False, k=[4, 16], voxel=0.1, occlusion=False)
    run("occlusion_only", True, k=10, density=False)
    d = ((pos[:, None, :] - pos[None, :, :]) ** 2).sum(dim=2)
    arrays["neighbors_k20"] = torch.sort(d, dim=1, stable=True).indices[:, :20]
    save("neighborhood_features", **arrays)


def make_integer_golden(ref):
    """Integer side: lex ops, CSR containers, ImageMapping.from_dense / indexing / batching."""
    lex, csr_mod, image = ref.lex, ref.csr, ref.image
    # lex KAT
    a = torch.LongTensor([2, 0, 2, 1, 0])
    b = torch.LongTensor([1, 5, 0, 3, 5])
    u = lex.lexunique(a, b)
    save("kat_lex", a=a, b=b, argsort=lex.lexargsort(a, b), argunique=lex.lexargunique(a, b),
         unique_a=u[0], unique_b=u[1])

    gen = torch.Generator().manual_seed(77)
    n_points, n_items = 500, 4000
    pid = torch.randint(0, n_points, (n_items,), generator=gen)
    iid = torch.randint(0, 5, (n_items,), generator=gen)
    pix = torch.randint(0, 64, (n_items, 2), generator=gen).short()
    feat = torch.rand(n_items, 3, generator=gen)
    # remove duplicates of (point, image, px, py) like MapImages does (image.py transform :328)
    keep = lex.lexargunique(pid, iid, pix[:, 0].long(), pix[:, 1].long())
    pid, iid, pix, feat = pid[keep], iid[keep], pix[keep], feat[keep]
    m = image.ImageMapping.from_dense(pid, iid, pix, feat, num_points=n_points + 7)
    sel = torch.randperm(n_points + 7, generator=gen)[:123]
    ms = m[sel]
    down = m.downscale_images(4)
    save("image_mapping", point_ids=pid, image_ids=iid, pixels=pix, features=feat,
         num_points=np.array(n_points + 7), pointers=m.pointers, images=m.images,
         atomic_pointers=m.values[1].pointers, out_pixels=m.pixels, out_features=m.features,
         sel=sel, sel_pointers=ms.pointers, sel_images=ms.images, sel_atomic_pointers=ms.values[1].pointers,
         sel_pixels=ms.pixels, sel_features=ms.features,
         down_atomic_pointers=down.values[1].pointers, down_pixels=down.pixels,
         fmi_batch=m.feature_map_indexing[0], fmi_h=m.feature_map_indexing[2], fmi_w=m.feature_map_indexing[3])

    # merge re-indexing (image.py:2211-2273) as run after a strided sparse conv (modules.py:232-234)
    merge_idx = torch.randint(0, 180, (n_points + 7,), generator=gen)
    merge_idx[:180] = torch.arange(180)  # every output voxel present (image.py:2220)
    mg = m.select_points(merge_idx, mode="merge")
    save("image_mapping_merge", merge_idx=merge_idx, pointers=mg.pointers, images=mg.images,
         atomic_pointers=mg.values[1].pointers, pixels=mg.pixels, features=mg.features)

    # z-buffer + projection + splat from the numba CPU path (visibility.py)
    vis = ref.visibility
    gen = torch.Generator().manual_seed(5)
    n = 8000
    xyz = (torch.rand(n, 3, generator=gen) - 0.5) * torch.tensor([12., 12., 4.])
    img_xyz = torch.tensor([0.3, -0.2, 0.1])
    img_opk = torch.tensor([0.05, -0.1, 0.7])
    W, H = 512, 256
    for crop_top, crop_bottom, tag in ((0, 0, "nocrop"), (16, 24, "crop")):
        idx, dist, xp, yp = vis.camera_projection_cpu(
            xyz, img_xyz, img_opk=img_opk, img_size=(W, H), crop_top=crop_top, crop_bottom=crop_bottom,
            r_max=8, r_min=0.5, camera="s3dis_equirectangular")
        arrays = dict(xyz=xyz, img_xyz=img_xyz, img_opk=img_opk, size=np.array([W, H]),
                      crop=np.array([crop_top, crop_bottom]), r=np.array([0.5, 8.0]),
                      rotation=vis.pose_to_rotation_matrix_cpu(img_opk.numpy()),
                      proj_idx=idx, dist=dist, x_proj=xp, y_proj=yp)
        splat = vis.equirectangular_splat_cpu(xp.numpy(), yp.numpy(), dist.numpy(), img_size=(W, H),
                                              crop_top=crop_top, crop_bottom=crop_bottom, voxel=0.05,
                                              k_swell=1.0, d_swell=1000)
        arrays["splat"] = splat
        for exact in (False, True):
            i2, xpix, ypix = vis.visibility_from_splatting_cpu(
                xp, yp, dist, xyz[idx], img_size=(W, H), crop_top=crop_top, crop_bottom=crop_bottom,
                voxel=0.05, k_swell=1.0, d_swell=1000, exact=exact, camera="s3dis_equirectangular")
            arrays[f"vis_idx_{int(exact)}"] = i2
            arrays[f"vis_x_{int(exact)}"] = xpix
            arrays[f"vis_y_{int(exact)}"] = ypix
        save(f"zbuffer_{tag}", **arrays)

    # pinhole splat boxes (scannet-like intrinsics), numba pinhole_splat_cpu (visibility.py:761-827)
    intr = np.eye(4, dtype=np.float32)
    intr[0, 0], intr[1, 1], intr[0, 2], intr[1, 2] = 577.87, 577.87, 319.5, 239.5
    m2 = 5000
    xp = (torch.rand(m2, generator=gen).double() * 640).numpy()
    yp = (torch.rand(m2, generator=gen).double() * 480).numpy()
    d = (torch.rand(m2, generator=gen) * 6 + 0.3).numpy()
    sp = vis.pinhole_splat_cpu(xp, yp, d, intr, img_size=(640, 480), crop_top=0, crop_bottom=0,
                               voxel=0.03, k_swell=1.0, d_swell=1000)
    save("splat_pinhole", x_proj=xp, y_proj=yp, dist=d, fx=np.array(577.87, dtype=np.float32),
         fy=np.array(577.87, dtype=np.float32), size=np.array([640, 480]), splat=sp)


if __name__ == "__main__":
    # `--only f1,f2` regenerates single fixture families without touching the others
    if len(sys.argv) > 2 and sys.argv[1] == "--only":
        _ref = ref_loader.load_reference()
        for _name in sys.argv[2].split(","):
            globals()[_name](_ref)
    else:
        main()
