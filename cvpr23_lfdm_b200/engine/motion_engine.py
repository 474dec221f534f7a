"""Engines of the full-LFAE branch (Generator.forward's dense-motion network, RegionPredictor, BGMotionPredictor;
SURVEY.md §8 rows a16/a17).  Implemented after the sampling + decode path (see DESIGN.md "scope / next")."""


class _Pending:
    what = ""

    def __init__(self, mod):
        self.mod = mod

    def forward(self, *a, **k):
        raise NotImplementedError(
            f"{self.what} is not yet ported to the sm_100a kernels (SURVEY.md §8f item 2: real-video branch); "
            "the sampling path (compute_fea / sample / forward_with_flow / decode_video) does not use it")


class PixelwiseFlowEngine(_Pending):
    what = "PixelwiseFlowPredictor (Generator.forward)"


class RegionEngine(_Pending):
    what = "RegionPredictor.forward"


class BGEngine(_Pending):
    what = "BGMotionPredictor.forward"
