from cvpr23_lfdm_b200.lfae.pixelwise_flow_predictor import PixelwiseFlowPredictor  # noqa: F401
