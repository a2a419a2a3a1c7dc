"""densefusion_b200 -- B200-native (sm_100a) implementation of DenseFusion's per-pixel pose-hypothesis
hot path behind the reference's Python API.

    from densefusion_b200.lib.network import PoseNet, PoseRefineNet
    from densefusion_b200.lib.loss import Loss
    from densefusion_b200.lib.loss_refiner import Loss_refine
    from densefusion_b200.lib.knn import KNearestNeighbor
    from densefusion_b200.pipeline import PoseEstimator          # batched, on-device estimate + refine

The kernels live in libdensefusion_b200.so (C ABI: include/densefusion_b200.h); importing any compute module
fails loudly when that library is absent -- there is no CPU or eager-torch fallback."""
__version__ = "0.1.0"
