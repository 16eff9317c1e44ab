"""Point -> image mapping construction (the reference's MapImages transform,
torch_points3d/core/data_transform/multimodal/image.py:162-428) on the GPU kernels.

For every image of a SameSettingImageData: range filter + camera projection + splat z-buffer
(core/multimodal/visibility.py) -> pixel coordinates at the mapping resolution
(`// proj_upscale`, `- crop_offsets`, in-crop filter, `// downscale`, :307-319) -> duplicate
(point, px, py) removed (:328) -> one ImageMapping.from_dense over all images (:415-417), images no
point sees are dropped (:392-394).  The reference's per-image KD-tree sphere sampling (:242-245)
only pre-filters points by distance, which the projection kernel does itself (r_min < d < r_max),
so it has no counterpart here.  Tensors in, tensors out: the torch_geometric `Data` holder of the
reference is not needed on this path.
"""
import torch

from ...utils.multimodal import lexargunique, lexunique
from . import visibility as visibility_module
from .image import ImageMapping, SameSettingImageData


class MapImages:
    def __init__(self, method='SplattingVisibility', proj_upscale=None, ref_size=None, use_cuda=True,
                 verbose=False, cylinder=False, **kwargs):
        if not use_cuda:
            raise RuntimeError("deepviewagg_b200.MapImages runs on CUDA only (no CPU fallback)")
        if method != 'SplattingVisibility':
            raise NotImplementedError(f"visibility method '{method}' is out of scope (see DESIGN.md)")
        self.method, self.proj_upscale, self.ref_size = method, proj_upscale, ref_size
        self.verbose, self.cylinder, self.kwargs = verbose, cylinder, kwargs

    def __call__(self, pos, images: SameSettingImageData, mapping_index=None, linearity=None, planarity=None,
                 scattering=None, normals=None, device='cuda'):
        """pos [N,3]; images: poses in `images.pos` / `images.opk` or `images.extras['extrinsic']`
        (+ 'intrinsic_pinhole' [B,4,4] / 'intrinsic_fisheye' [B,7]); returns a copy of `images`
        restricted to the seen images, with `.mappings` set."""
        assert images.num_views >= 1, "At least one image must be provided."
        if self.ref_size is not None:
            images.ref_size = tuple(self.ref_size)
            images.crop_size = images.ref_size
        if self.proj_upscale is not None:
            images.proj_upscale = self.proj_upscale
        proj_size = tuple(int(v * images.proj_upscale) for v in images.ref_size)
        model = getattr(visibility_module, self.method)(img_size=proj_size, **self.kwargs)
        dev = torch.device(device)
        pos_d = pos.float().to(dev)
        n_points = pos.shape[0]
        ids = torch.arange(n_points, device=dev) if mapping_index is None else mapping_index.to(dev).long()
        to_d = lambda t: t.to(dev) if t is not None else None  # noqa: E731
        lin, pla, sca, nor = to_d(linearity), to_d(planarity), to_d(scattering), to_d(normals)
        ex = images.extras
        crop_off = images.crop_offsets if images.crop_offsets is not None else \
            torch.zeros((images.num_views, 2), dtype=torch.long)
        image_ids, point_ids, features, pixels = [], [], [], []
        for i in range(images.num_views):
            kw = {}
            if images.opk is not None:
                kw['img_opk'] = images.opk[i].float()
            if 'extrinsic' in ex:
                kw['img_extrinsic'] = ex['extrinsic'][i].float()
            if 'intrinsic_pinhole' in ex:
                kw['img_intrinsic_pinhole'] = ex['intrinsic_pinhole'][i].float()
            if 'intrinsic_fisheye' in ex:
                kw['img_intrinsic_fisheye'] = ex['intrinsic_fisheye'][i].float()
            out = model(pos_d, images.pos[i].float(), linearity=lin, planarity=pla, scattering=sca, normals=nor,
                        **kw)
            if out['idx'].shape[0] == 0:
                continue
            pid = ids[out['idx']]
            px = out['x'].long() // images.proj_upscale - int(crop_off[i, 0])
            py = out['y'].long() // images.proj_upscale - int(crop_off[i, 1])
            inside = torch.where((px >= 0) & (py >= 0) & (px < images.crop_size[0]) & (py < images.crop_size[1]))[0]
            px = (px[inside] // images.downscale).long()
            py = (py[inside] // images.downscale).long()
            pid, feat = pid[inside], out['features'].float()[inside]
            keep = lexargunique(pid, px, py)
            image_ids.append(i)
            point_ids.append(pid[keep])
            features.append(feat[keep])
            pixels.append(torch.stack((px[keep], py[keep]), dim=1).to(images.pixel_dtype))
        if len(image_ids) == 0:
            raise ValueError(
                "No mappings were found between the 3D points and any of the provided images. Make sure your "
                "images are located in the vicinity of your point cloud and that the projection parameters "
                "allow for at least one point-image-pixel mapping.")
        seen = torch.tensor(image_ids, dtype=torch.long)
        out_images = images[seen]                      # unseen images dropped, the rest renumbered
        new_ids = torch.arange(len(image_ids), device=dev).repeat_interleave(
            torch.tensor([p.shape[0] for p in point_ids], device=dev))
        n_total = int(ids.max().item()) + 1
        out_images.mappings = ImageMapping.from_dense(torch.cat(point_ids), new_ids, torch.cat(pixels),
                                                      torch.cat(features), num_points=n_total)
        out_images.visibility = model
        return out_images
