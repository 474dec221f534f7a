from cvpr23_lfdm_b200.dm.video_flow_diffusion_model import FlowDiffusion  # noqa: F401
