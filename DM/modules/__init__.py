"""Drop-in import path of the reference layout (`from DM.modules.video_flow_diffusion_model import FlowDiffusion`)."""
