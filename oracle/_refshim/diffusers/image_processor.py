"""diffusers.image_processor.VaeImageProcessor, the tensor path only (what EasyAnimateInpaintPipeline feeds it,
pipeline_easyanimate_inpaint.py:322-326,1236,1339): [N, C, H, W] tensors already at the target size.  Restated from the published
0.30/0.31 behaviour: `do_normalize` maps [0, 1] -> [-1, 1] unless the input already has negative values; `do_binarize` thresholds
at 0.5; `do_convert_grayscale` is a PIL-path option (a [N, 1, H, W] mask tensor passes through).  Input preparation, outside the
hot path; third-party and unpinned like the rest of the stand-in."""
import torch

from .configuration_utils import FrozenDict


class VaeImageProcessor:
    def __init__(self, do_resize=True, vae_scale_factor=8, resample="lanczos", do_normalize=True, do_binarize=False,
                 do_convert_rgb=False, do_convert_grayscale=False):
        self.config = FrozenDict(do_resize=do_resize, vae_scale_factor=vae_scale_factor, resample=resample,
                                 do_normalize=do_normalize, do_binarize=do_binarize, do_convert_rgb=do_convert_rgb,
                                 do_convert_grayscale=do_convert_grayscale)

    def preprocess(self, image, height=None, width=None, resize_mode="default", crops_coords=None):
        if not torch.is_tensor(image) or image.ndim != 4:
            raise NotImplementedError("diffusers shim: VaeImageProcessor.preprocess takes [N, C, H, W] tensors only")
        if self.config.do_resize and (height, width) != tuple(image.shape[-2:]):
            image = torch.nn.functional.interpolate(image, size=(height, width))
        do_normalize = self.config.do_normalize
        if do_normalize and image.min() < 0:
            do_normalize = False
        if do_normalize:
            image = 2.0 * image - 1.0
        if self.config.do_binarize:
            image = image.clone()
            image[image < 0.5] = 0
            image[image >= 0.5] = 1
        return image
