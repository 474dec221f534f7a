from cvpr23_lfdm_b200.lfae.util import *  # noqa: F401,F403
