"""nerf_atlas_b200 -- B200-native (sm_100a) volume-rendering path for JulianKnodt/nerf_atlas.

Only the hot path of SURVEY.md section 8 lives here: csrc/ (CUDA kernels + the C ABI of
include/nerf_b200.h) and the host-side mirror of the reference's CommonNeRF.forward surface."""
from .model import (FusedNeRF, FusedPlainNeRF, FusedTinyNeRF, FusedVolSDF, FusedDynamicNeRF, FusedSDF, RenderEngine,  # noqa: F401
                    describe_plain, describe_tiny, describe_volsdf, describe_dyn, volumetric_integrate)
from .shard import shard_rays, ShardedRenderer, GradientAllReducer  # noqa: F401
from . import autograd  # noqa: F401  (differentiable hash_encode / composite: CUDA forward + backward)
