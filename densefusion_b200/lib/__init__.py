"""Drop-in counterparts of the reference's `lib` package (same module and class names):
lib.network.PoseNet / PoseRefineNet, lib.loss.Loss, lib.loss_refiner.Loss_refine,
lib.knn.KNearestNeighbor, lib.transformations.quaternion_matrix / quaternion_from_matrix."""
