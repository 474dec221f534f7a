from cvpr23_lfdm_b200.dm.video_flow_diffusion import *  # noqa: F401,F403
from cvpr23_lfdm_b200.dm.video_flow_diffusion import Unet3D, GaussianDiffusion, extract, cosine_beta_schedule  # noqa: F401
