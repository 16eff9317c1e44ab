"""UnimodalBranch / MultimodalBlockDown with the reference's interface
(torch_points3d/modules/multimodal/modules.py:23-575), re-plumbed so that one modality branch is

    conv  ->  fused pixel gather + atomic pool  ->  [E_mod]  ->  fused row gather + view attention  ->  fusion

instead of the reference's  conv -> [P,C] NCHW advanced-index copy -> segment_csr ->
cat(...)[idx_sorting] [V,C] copy -> pool  (modules.py:396-416, 481-540).  Behaviour kept: the
mm_data_dict keys ('x_3d', 'x_seen', 'modalities'), list-of-settings ImageData, empty-modality
and empty-setting shortcuts, `checkpointing` flags, `keep_last_view`, `out_channels` inference.
No hard torchsparse import (the reference has one at modules.py:10): sparse-tensor branches are
taken only when MinkowskiEngine / torchsparse are importable.
"""
import torch
import torch.nn as nn
from torch.utils.checkpoint import checkpoint

from ... import ops
from ...core.common_modules import Identity
from .pooling import BimodalCSRPool, GroupBimodalCSRPool, QKVBimodalCSRPool

try:  # optional sparse backends (absent in this environment)
    import MinkowskiEngine as me
except Exception:  # noqa: BLE001
    me = None
try:
    import torchsparse as ts
    from torchsparse.nn.functional import sphash, sphashquery
except Exception:  # noqa: BLE001
    ts = None

# Supported modalities (core/multimodal/data.py:9-10)
MODALITY_NAMES = ["image"]

__all__ = ["MultimodalBlockDown", "UnimodalBranch", "IdentityBranch", "MODALITY_NAMES"]


class MultimodalBlockDown(nn.Module):
    """3D conv -> modality branches -> 3D conv (modules.py:23-236)."""

    def __init__(self, block_1, block_2, **kwargs):
        super().__init__()
        self.block_1 = block_1 if block_1 is not None else Identity()
        self.block_2 = block_2 if block_2 is not None else Identity()
        self._modalities = []
        for m, branch in kwargs.items():
            assert m in MODALITY_NAMES, f"Invalid kwarg modality '{m}', expected one of {MODALITY_NAMES}."
            assert isinstance(branch, (UnimodalBranch, IdentityBranch)), \
                f"Expected a UnimodalBranch module for '{m}' modality but got {type(branch)} instead."
            setattr(self, m, branch)
            self._modalities.append(m)
        self.sampler = [getattr(self.block_1, "sampler", None), getattr(self.block_2, "sampler", None)]

    @property
    def modalities(self):
        return self._modalities

    def forward(self, mm_data_dict):
        mm_data_dict = self.forward_3d_block_down(mm_data_dict, self.block_1)
        for m in self.modalities:
            mm_data_dict = getattr(self, m)(mm_data_dict, m)
        return self.forward_3d_block_down(mm_data_dict, self.block_2)

    @staticmethod
    def forward_3d_block_down(mm_data_dict, block):
        """Run a 3D block and re-index seen flags and mappings after any sampling / strided conv:
        dense tensors -> 'pick' with the sampler's last_idx; sparse tensors -> 'merge' with the
        child->parent voxel index (modules.py:101-236)."""
        if isinstance(block, Identity):
            return mm_data_dict
        x_3d, x_seen = mm_data_dict['x_3d'], mm_data_dict['x_seen']
        idx, mode = None, 'pick'
        if isinstance(x_3d, torch.Tensor):
            block.sampler.last_idx = None
            n_in = x_3d.shape[0]
            x_3d = block(x_3d)
            idx = block.sampler.last_idx
            if idx is not None and idx.numel() == n_in and bool((idx == torch.arange(n_in, device=idx.device)).all()):
                idx = None
        elif me is not None and isinstance(x_3d, me.SparseTensor):
            mode = 'merge'
            stride_in = x_3d.tensor_stride[0]
            x_3d = block(x_3d)
            stride_out = x_3d.tensor_stride[0]
            if stride_in != stride_out:
                src, target = x_3d.coords_man.get_coords_map(stride_in, stride_out)
                idx = target[src.argsort()]
        elif ts is not None and isinstance(x_3d, ts.SparseTensor):
            mode = 'merge'
            stride_in = x_3d.s
            x_3d = block(x_3d)
            stride_out = x_3d.s
            if stride_in != stride_out:
                in_coords = x_3d.coord_maps[stride_in]
                in_coords[:, :3] = ((in_coords[:, :3].float() / stride_out).floor() * stride_out).int()
                idx = sphashquery(sphash(in_coords), sphash(x_3d.coord_maps[stride_out]))
        else:
            raise NotImplementedError(
                f"Unsupported format for x_3d: {type(x_3d)}. If you are trying to use MinkowskiEngine or "
                f"TorchSparse, make sure those are properly installed.")
        if x_seen is not None and idx is not None:
            if mode == 'pick':
                x_seen = x_seen[idx]
            else:  # any child seen -> parent seen (modules.py:225: scatter-sum of a bool mask)
                n_out = int(idx.max().item()) + 1
                x_seen = torch.zeros(n_out, dtype=torch.long, device=idx.device).index_add_(
                    0, idx, x_seen.long()) > 0
        mm_data_dict['x_3d'], mm_data_dict['x_seen'] = x_3d, x_seen
        for m in mm_data_dict['modalities'].keys():
            mm_data_dict['modalities'][m] = mm_data_dict['modalities'][m].select_points(idx, mode=mode)
        return mm_data_dict


class UnimodalBranch(nn.Module):
    """conv -> atomic pool -> view pool -> fusion for one modality (modules.py:249-566)."""

    def __init__(self, conv, atomic_pool, view_pool, fusion, drop_3d=0, drop_mod=0, hard_drop=False,
                 keep_last_view=False, checkpointing='', out_channels=None, interpolate=False):
        super().__init__()
        if hard_drop:
            raise NotImplementedError("hard_drop (ModalityDropout) is dead code in the reference "
                                      "(its constructor rejects `inplace`, modules.py:272-278)")
        self.conv, self.atomic_pool, self.view_pool, self.fusion = conv, atomic_pool, view_pool, fusion
        self.drop_3d = nn.Dropout(p=drop_3d, inplace=False) if drop_3d is not None and drop_3d > 0 else None
        self.drop_mod = nn.Dropout(p=drop_mod, inplace=True) if drop_mod is not None and drop_mod > 0 else None
        self.keep_last_view = keep_last_view
        self._out_channels = out_channels
        self.interpolate = interpolate
        assert not checkpointing or isinstance(checkpointing, str), \
            f'Expected checkpointing to be of type str but received {type(checkpointing)} instead.'
        self.checkpointing = ''.join(sorted(set('cavf').intersection(set(checkpointing or ''))))

    @property
    def out_channels(self):
        if self._out_channels is None:
            raise ValueError(f'{self.__class__.__name__}.out_channels has not been set. Please set it to '
                             f'allow inference even when the modality has no data.')
        return self._out_channels

    # -- forward ------------------------------------------------------------------------------------
    def forward(self, mm_data_dict, modality):
        is_sparse_3d = not isinstance(mm_data_dict['x_3d'], (torch.Tensor, type(None)))
        x_3d = mm_data_dict['x_3d'].F if is_sparse_3d else mm_data_dict['x_3d']
        mod_data = mm_data_dict['modalities'][modality]
        assert isinstance(mod_data.x, list), \
            "modality data must be the list-of-settings ImageData form (modules.py:516-539)"

        # no data at all for this modality: pad x_3d to out_channels (modules.py:317-365)
        if len(mod_data) == 0 or all(e.x.shape[0] == 0 for e in mod_data):
            nc_out, nc_3d = self.out_channels, x_3d.shape[1]
            if nc_out < nc_3d:
                raise ValueError(f'{self.__class__.__name__}.out_channels is smaller than number of features '
                                 f'in x_3d: {nc_out} < {nc_3d}')
            nc_2d = nc_out - nc_3d if nc_out > nc_3d else nc_3d
            if len(mod_data) > 0:
                mod_data.x = [x[:, [0]].repeat_interleave(nc_2d, dim=1) for x in mod_data.x]
            if nc_out > nc_3d:
                x_3d = torch.cat((x_3d, x_3d.new_zeros(x_3d.shape[0], nc_2d)), dim=1)
            self._write_x3d(mm_data_dict, x_3d, is_sparse_3d)
            mm_data_dict['modalities'][modality] = mod_data
            return mm_data_dict

        # some settings without images: run on the others, then restore them (modules.py:372-393)
        if any(e.x.shape[0] == 0 for e in mod_data):
            num = len(mod_data)
            removed = {i: e for i, e in enumerate(mod_data) if e.x.shape[0] == 0}
            kept_idx = [i for i in range(num) if i not in removed]
            mm_data_dict['modalities'][modality] = mod_data[kept_idx]
            mm_data_dict = self.forward(mm_data_dict, modality)
            new = mm_data_dict['modalities'][modality]
            joined = {**{k: e for k, e in zip(kept_idx, new)}, **removed}
            mm_data_dict['modalities'][modality] = new.__class__([joined[i] for i in range(num)])
            return mm_data_dict

        mod_data = self.forward_conv(mod_data)
        x_mod = self.forward_atomic_pool(x_3d, mod_data)
        x_mod, mod_data, csr_idx = self.forward_view_pool(x_3d, x_mod, mod_data)
        x_seen = csr_idx[1:] > csr_idx[:-1]
        x_3d, x_mod, mod_data = self.forward_dropout(x_3d, x_mod, mod_data)
        x_3d = self.forward_fusion(x_3d, x_mod)
        if self._out_channels is None:
            self._out_channels = x_3d.shape[1]
        self._write_x3d(mm_data_dict, x_3d, is_sparse_3d)
        mm_data_dict['modalities'][modality] = mod_data
        prev = mm_data_dict['x_seen']
        mm_data_dict['x_seen'] = x_seen if prev is None else torch.logical_or(x_seen, prev)
        return mm_data_dict

    @staticmethod
    def _write_x3d(mm_data_dict, x_3d, is_sparse_3d):
        if is_sparse_3d:
            mm_data_dict['x_3d'].F = x_3d
        else:
            mm_data_dict['x_3d'] = x_3d

    def forward_conv(self, mod_data, reset=True):
        """2D conv on every setting's feature maps (modules.py:442-479)."""
        if not self.conv:
            return mod_data
        for i in range(len(mod_data)):
            im = mod_data[i]
            if 'c' in self.checkpointing:
                im.x = checkpoint(self.conv, im.x.requires_grad_(), torch.BoolTensor([i == 0]),
                                  use_reentrant=True)
            else:
                im.x = self.conv(im.x, True)
        return mod_data

    def _atomic_one(self, x_3d, im):
        """[V_s, C] view features of one setting."""
        fused = isinstance(self.atomic_pool, BimodalCSRPool) and not self.atomic_pool.save_last \
            and im.x.is_cuda
        if fused:
            # pixel gather (or bilinear interpolation) + pool in one kernel; a channels_last
            # feature map is read coalesced
            bilinear = self.interpolate and im.downscale != 1
            maps = im.scaled_mappings(interpolate=self.interpolate)
            x = im.x
            cl = x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last) and not x.is_contiguous()
            fmap = x.permute(0, 2, 3, 1) if cl else x
            if bilinear:
                return ops.interp_pool(fmap, maps.images, maps.pixels, maps.atomic_csr_indexing,
                                       im.mapping_size, reduce=self.atomic_pool._mode, channels_last=cl)
            return ops.gather_pool(fmap, maps.images, maps.pixels, maps.atomic_csr_indexing,
                                   reduce=self.atomic_pool._mode, channels_last=cl)
        x_pix = im.get_mapped_features(interpolate=self.interpolate)
        csr = im.atomic_csr_indexing if self.interpolate else im.scaled_mappings(False).atomic_csr_indexing
        if 'a' in self.checkpointing:
            return checkpoint(self.atomic_pool, x_3d, x_pix, None, csr, use_reentrant=True)
        return self.atomic_pool(x_3d, x_pix, None, csr)

    def forward_atomic_pool(self, x_3d, mod_data):
        """Atomic (pixel -> view) pooling per setting (modules.py:481-501 + image.py:1262-1287)."""
        return [self._atomic_one(x_3d, im) for im in mod_data]

    def forward_view_pool(self, x_3d, x_mod, mod_data):
        """View (view -> point) pooling over all settings (modules.py:503-540)."""
        idx_sorting = mod_data.view_cat_sorting
        csr_idx = mod_data.view_cat_csr_indexing
        x_cat = torch.cat(x_mod, dim=0)
        x_map = torch.cat(mod_data.mapping_features, dim=0)[idx_sorting]
        fold_gather = isinstance(self.view_pool, (GroupBimodalCSRPool, QKVBimodalCSRPool)) \
            and not self.keep_last_view and not self.view_pool.save_last and x_cat.is_cuda
        if fold_gather:
            args = (x_3d, x_cat, x_map, csr_idx, idx_sorting, True)  # gather folded into the kernel; view_cat_sorting is a permutation
        else:
            x_sorted = x_cat[idx_sorting]
            if self.keep_last_view:
                mod_data.last_view_x_mod = x_sorted
                mod_data.last_view_x_map = x_map
                mod_data.last_view_csr_idx = csr_idx
            args = (x_3d, x_sorted, x_map, csr_idx)
        if 'v' in self.checkpointing:
            x_pool = checkpoint(self.view_pool, *args, use_reentrant=True)
        else:
            x_pool = self.view_pool(*args)
        return x_pool, mod_data, csr_idx

    def forward_fusion(self, x_3d, x_mod):
        if 'f' in self.checkpointing:
            return checkpoint(self.fusion, x_3d, x_mod, use_reentrant=True)
        return self.fusion(x_3d, x_mod)

    def forward_dropout(self, x_3d, x_mod, mod_data):
        if self.drop_3d:
            x_3d = self.drop_3d(x_3d)
        if self.drop_mod:
            x_mod = self.drop_mod(x_mod)
            if self.keep_last_view:
                mod_data.last_view_x_mod = self.drop_mod(mod_data.last_view_x_mod)
        return x_3d, x_mod, mod_data

    def extra_repr(self) -> str:
        return "\n".join(f'{a}={getattr(self, a)}' for a in ['drop_3d', 'drop_mod', 'keep_last_view', 'checkpointing'])


class IdentityBranch(nn.Module):
    def forward(self, mm_data_dict, modality):
        return mm_data_dict
