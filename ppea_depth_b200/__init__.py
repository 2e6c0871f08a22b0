"""ppea_depth_b200 -- B200-native (sm_100a) view-synthesis loss path of PPEA-Depth.

Public surface (mirrors /root/reference/ppeadepth/layers.py and the loss
methods of /root/reference/ppeadepth/trainer.py):

    from ppea_depth_b200 import (BackprojectDepth, Project3D, SSIM, get_smooth_loss,
                                 disp_to_depth, upsample, transformation_from_parameters,
                                 generate_images_pred, compute_reprojection_loss,
                                 compute_loss_masks, compute_losses, install,
                                 view_synthesis_loss, VslConfig)
"""
from .layers import (BackprojectDepth, Project3D, SSIM, disp_to_depth, get_smooth_loss, get_translation_matrix,
                     rot_from_axisangle, transformation_from_parameters, upsample)
from .functional import VslConfig, VslResult, view_synthesis_loss
from .images import images_to_float
from .decoder import FusedDispHead, disp_head, disp_head_with_depth, install_decoder
from .pyramid import ImagePyramid, pack_rgbx, resize_lanczos_u8
from .glue import DeviceDepthBins, matching_glue, zero_missing_poses
from .matching import cost_volume_tail, install_matching, match_features, match_features_dyn
from .loss import (ViewSynthesisLoss, compute_loss_masks, compute_losses, compute_reprojection_loss,
                   generate_images_pred, install)

__all__ = [
    "BackprojectDepth", "Project3D", "SSIM", "disp_to_depth", "get_smooth_loss", "get_translation_matrix",
    "rot_from_axisangle", "transformation_from_parameters", "upsample", "VslConfig", "VslResult",
    "view_synthesis_loss", "ViewSynthesisLoss", "compute_loss_masks", "compute_losses",
    "compute_reprojection_loss", "generate_images_pred", "install", "images_to_float", "matching_glue", "DeviceDepthBins", "zero_missing_poses", "match_features", "match_features_dyn", "install_matching", "cost_volume_tail",
    "disp_head", "disp_head_with_depth", "FusedDispHead", "install_decoder",
    "ImagePyramid", "pack_rgbx", "resize_lanczos_u8",
]
