from cvpr23_lfdm_b200.lfae.region_predictor import RegionPredictor  # noqa: F401
