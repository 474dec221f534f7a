from cvpr23_lfdm_b200.lfae.bg_motion_predictor import BGMotionPredictor  # noqa: F401
