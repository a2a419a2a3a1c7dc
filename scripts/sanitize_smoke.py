"""One small pass through every hand-written kernel family (for compute-sanitizer memcheck / racecheck runs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from densefusion_b200 import ops, synth
from densefusion_b200.lib.loss import Loss
from densefusion_b200.lib.loss_refiner import Loss_refine
from densefusion_b200.pipeline import PoseEstimator
from densefusion_b200.trainer import DataParallelTrainer
from util import build_nets

torch.backends.cudnn.allow_tf32 = False
n, o, m = 500, 21, 500
est, ref, _, _ = build_nets(n, o, seed=0)
crops = [synth.synth_crop(c, n, m, o, (80, 80), ob) for c, ob in ((1, 12), (2, 3))]
keys = ("img", "points", "choose", "idx", "target", "model_points")
b = {k: torch.cat([c[k] for c in crops], 0).cuda() for k in keys}
for prec in ("fp32", "3xtf32", "hybrid"):
    pipe = PoseEstimator(est, ref, iterations=2, precision=prec, chunk_crops=1)
    pose = pipe.estimate(b["img"], b["points"], b["choose"], b["idx"])
    assert torch.isfinite(pose).all()
with torch.no_grad():
    r, t, c, emb = est.forward_batched(b["img"], b["points"], b["choose"], b["idx"])
loss, dis, npts, ntgt = Loss(m, synth.YCB_SYM)(r, t, c, b["target"], b["model_points"], b["idx"], b["points"], 0.015, False)
ref_pts = b["target"][0].t().contiguous()[None]
ops.knn(ref_pts, ref_pts[:, :, :333].contiguous(), 1)
tr = DataParallelTrainer(est, ref, m, synth.YCB_SYM, phase="estimator")
tr.step([b])
tr.set_phase("refiner")
tr.step([b])
torch.cuda.synchronize()
print("sanitize smoke ok", float(loss.sum()))
