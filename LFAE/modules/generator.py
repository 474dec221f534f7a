from cvpr23_lfdm_b200.lfae.generator import Generator  # noqa: F401
