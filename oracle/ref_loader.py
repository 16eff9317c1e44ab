"""Load the reference's hot-path modules BY FILE PATH (oracle; test infrastructure).

Only usable in the build container, where /root/reference exists: used by
oracle/make_golden.py to generate tests/golden/*.npz and by tests marked `needs_reference` to
pin the oracle against the executed reference.  Nothing that runs on the GPU box imports this.

Package-level import of the reference is impossible here (torch_points3d/utils/__init__.py
needs hydra, core/multimodal/data.py torch_geometric, modules/multimodal/modules.py torchsparse,
visibility.py pykeops), so parent packages are stubbed and the files are loaded in dependency
order.  torch_scatter is replaced by oracle/scatter_standin.py and
TorchScript is disabled (PYTORCH_JIT=0) so that the reference's @torch.jit.script helpers call
the stand-in as plain Python.
"""
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("DVA_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "torch_points3d"))


def _stub(name):
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
    return m


def _load(name, relpath):
    path = os.path.join(REFERENCE_ROOT, relpath)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


_loaded = None


def load_reference(with_visibility=True):
    """Returns a namespace with .pooling, .fusion, .base_modules, .lex (utils.multimodal), .csr,
    .image, .visibility of the reference."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    if os.environ.get("PYTORCH_JIT", "1") != "0":
        raise RuntimeError("set PYTORCH_JIT=0 before importing torch to load the reference "
                           "(its @torch.jit.script helpers must call the torch_scatter stand-in)")
    from oracle import scatter_standin
    sys.modules["torch_scatter"] = scatter_standin

    for pkg in ["torch_points3d", "torch_points3d.core", "torch_points3d.core.common_modules",
                "torch_points3d.core.multimodal", "torch_points3d.utils", "torch_points3d.modules",
                "torch_points3d.modules.multimodal"]:
        _stub(pkg)
    # pykeops is only used by the Biasutti visibility model (out of scope)
    pk = _stub("pykeops")
    pkt = _stub("pykeops.torch")
    pkt.LazyTensor = object
    pk.torch = pkt

    ns = types.SimpleNamespace()
    ns.base_modules = _load("torch_points3d.core.common_modules.base_modules",
                            "torch_points3d/core/common_modules/base_modules.py")
    cm = sys.modules["torch_points3d.core.common_modules"]
    cm.MLP = ns.base_modules.MLP
    cm.base_modules = ns.base_modules
    ns.pooling = _load("torch_points3d.modules.multimodal.pooling",
                       "torch_points3d/modules/multimodal/pooling.py")
    ns.fusion = _load("torch_points3d.modules.multimodal.fusion",
                      "torch_points3d/modules/multimodal/fusion.py")
    ns.lex = _load("torch_points3d.utils.multimodal", "torch_points3d/utils/multimodal.py")
    ns.csr = _load("torch_points3d.core.multimodal.csr", "torch_points3d/core/multimodal/csr.py")
    mm = sys.modules["torch_points3d.core.multimodal"]
    mm.csr = ns.csr
    mm.CSRData, mm.CSRBatch = ns.csr.CSRData, ns.csr.CSRBatch
    # image.py imports VisibilityModel, so visibility.py (numba) is always loaded; its
    # @njit(cache=True) needs a writable cache dir because /root/reference is read-only.
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/dva_numba_cache")
    ns.visibility = _load("torch_points3d.core.multimodal.visibility",
                          "torch_points3d/core/multimodal/visibility.py")
    mm.visibility = ns.visibility
    ns.image = _load("torch_points3d.core.multimodal.image", "torch_points3d/core/multimodal/image.py")
    # modules.py: needs MODALITY_NAMES (core/multimodal/data.py:9-10, torch_geometric-free stub),
    # ModalityDropout and a torchsparse stub (hard import at modules.py:10; unused on dense tensors)
    data_stub = _stub("torch_points3d.core.multimodal.data")
    data_stub.MODALITY_NAMES = ["image"]
    tsp = _stub("torchsparse")
    tsp_nn = _stub("torchsparse.nn")
    tsp_f = _stub("torchsparse.nn.functional")
    tsp_f.sphash = tsp_f.sphashquery = None
    tsp.nn, tsp_nn.functional = tsp_nn, tsp_f
    tsp.SparseTensor = type("SparseTensor", (), {})
    ns.dropout = _load("torch_points3d.modules.multimodal.dropout",
                       "torch_points3d/modules/multimodal/dropout.py")
    ns.modules = _load("torch_points3d.modules.multimodal.modules",
                       "torch_points3d/modules/multimodal/modules.py")
    _loaded = ns
    return ns


class DenseLazyTensor:
    """Dense stand-in for the subset of pykeops.torch.LazyTensor that
    NeighborhoodBasedMappingFeatures uses (core/data_transform/multimodal/image.py:504-514):
    broadcasting `-`, `** 2`, `.sum(dim=2)` and `.argKmin(K, dim=1)`.  KeOps documents argKmin as
    the indices of the K smallest values along `dim`; the dense emulation takes them from a stable
    sort, i.e. ties (and only ties) are ordered by index.  Small N only (N x N x 3 floats)."""

    def __init__(self, t):
        self.t = t

    def __sub__(self, other):
        return DenseLazyTensor(self.t - other.t)

    def __pow__(self, p):
        return DenseLazyTensor(self.t ** p)

    def sum(self, dim):
        return DenseLazyTensor(self.t.sum(dim=dim))

    def argKmin(self, K, dim):
        import torch
        return torch.sort(self.t, dim=dim, stable=True).indices[:, :K]


def load_transforms():
    """The reference's image transforms module (core/data_transform/multimodal/image.py), with
    stand-ins for what it imports but the neighbourhood-feature path never touches:
    torch_geometric `Data` (attribute bag with `num_nodes`), the 3D samplers, FAISS finder."""
    ns = load_reference()
    if hasattr(ns, "transforms"):
        return ns.transforms
    tg = _stub("torch_geometric")
    tgd = _stub("torch_geometric.data")

    class Data:
        def __init__(self, **kw):
            self.__dict__.update(kw)

        @property
        def num_nodes(self):
            return self.pos.shape[0]

    tgd.Data = Data
    tg.data = tgd
    dt = _stub("torch_points3d.core.data_transform")
    for name in ("SphereSampling", "CylinderSampling", "GridSampling3D", "SaveOriginalPosId"):
        setattr(dt, name, type(name, (), {}))
    so = _stub("torch_points3d.core.spatial_ops")
    nf = _stub("torch_points3d.core.spatial_ops.neighbour_finder")
    nf.FAISSGPUKNNNeighbourFinder = type("FAISSGPUKNNNeighbourFinder", (), {})
    so.neighbour_finder = nf
    sys.modules["pykeops.torch"].LazyTensor = DenseLazyTensor
    _stub("torch_points3d.core.data_transform.multimodal")
    ns.transforms = _load("torch_points3d.core.data_transform.multimodal.image",
                          "torch_points3d/core/data_transform/multimodal/image.py")
    ns.Data = Data
    return ns.transforms
