"""contrastive_lift_b200 - the B200 (sm_100a) implementation of Contrastive-Lift's render + contrastive hot path.

Public surface = the reference's own class/function names for that path:

    TensorVMSplit, TensoRFRenderer                      model + renderer drop-ins
    slow_fast_loss, contrastive_loss, ema_update, TVLoss    loss drop-ins
    get_rays                                            per-frame ray generation
    FusedAdam                                           torch.optim.Adam drop-in, one launch per optimizer step (8f rank 2)
    nearest_centroid, assign_clusters                   embedding -> centroid labels at inference (8f rank 4)

Everything computes inside ``libclift_b200.so`` (C ABI in include/clift_b200.h); importing the package
does not load it, the first compute call does, and fails loudly if it is missing.
"""
from .field import TensorVMSplit  # noqa: F401
from .renderer import TensoRFRenderer  # noqa: F401
from .loss import TVLoss, contrastive_loss, ema_update, ema_update_slownet, plane_tv, slow_fast_loss  # noqa: F401
from .rays import get_rays, get_rays_checked  # noqa: F401
from .optim import FusedAdam  # noqa: F401
from .cluster import assign_clusters, nearest_centroid  # noqa: F401
