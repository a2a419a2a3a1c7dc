"""GPU debug: per-parameter gradient error of the estimator training path vs the oracle, plus d(loss)/d(feature map)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from conftest import golden
from test_training_cpu import _crops
from util import build_nets, rel
from oracle import df_oracle as O
from densefusion_b200.lib.loss import Loss

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
g = golden("c4_train_ycb")
crops, n, o, m, seed, iters = _crops(g)
sym, w = [int(s) for s in g["sym_list"]], float(g["w"])
est, ref, est_sd, ref_sd = build_nets(n, o, seed)
est.requires_grad_(True)
keys = ("img", "points", "choose", "idx", "target", "model_points")
b = {k: torch.cat([c[k] for c in crops], 0).cuda() for k in keys}
feat = est.cnn(b["img"])
feat.retain_grad()
from densefusion_b200 import training
r, t, c, emb = training.posenet_head_train(est, feat, b["points"], b["choose"], b["idx"])
loss, dis, _, _ = Loss(m, sym)(r, t, c, b["target"], b["model_points"], b["idx"], b["points"], w, False)
loss.sum().backward()
# oracle with feature-map gradient
leaf = {k: v.detach().clone().requires_grad_(True) for k, v in est_sd.items()}
dfe = []
for d in crops:
    out_img = O.psp_encoder(leaf, d["img"]); out_img.retain_grad()
    e = O.gather_embedding(out_img, d["choose"])
    rr, tt, cc = O.posenet_head(leaf, d["points"], e, d["idx"], o)
    tot, _, _, _ = O.loss(rr, tt, cc, d["target"], d["model_points"], d["idx"], d["points"], w, False, m, sym)
    tot.backward()
    dfe.append(out_img.grad)
dfe = torch.cat(dfe, 0)
print("feature map fwd err", rel(feat, torch.cat([O.psp_encoder(est_sd, d["img"]) for d in crops], 0)))
print("dfeat err", rel(feat.grad, dfe), "max", float(dfe.abs().max()))
rows = []
for name, p in est.named_parameters():
    og = leaf[name].grad
    if og is None:
        continue
    rows.append((rel(p.grad, og), name, float(og.abs().max())))
for e, name, mx in sorted(rows, reverse=True)[:25]:
    print(f"{e:.3e}  {name}  max|g|={mx:.3e}")
print("head-only worst", max(e for e, nme, _ in rows if not nme.startswith("cnn.")))
# same encoder backward in float64 on the GPU from OUR dfeat: separates cuDNN backward error from head error
est64 = est.cnn.double()
img64 = b["img"].double()
f64 = est64(img64)
est64.zero_grad()
f64.backward(dfe.cuda().double())
for name, p in est64.named_parameters():
    og = leaf["cnn." + name].grad
    if og is not None and "layer4.1.conv1" in name:
        print("fp64 encoder backward from oracle dfeat vs oracle fp32 grads:", name, rel(p.grad, og))
# (2) fp64 encoder backward from OUR dfeat: isolates the head-side error
ours = feat.grad.detach().double()
est64.zero_grad()
f64 = est64(img64)
f64.backward(ours)
worst = 0
for name, p in est64.named_parameters():
    og = leaf["cnn." + name].grad
    if og is not None:
        e = rel(p.grad, og); worst = max(worst, e)
        if "layer4.1.conv1" in name or "feats.conv1" in name:
            print("fp64 encoder backward from OUR dfeat:", name, e)
print("worst over cnn params (fp64 backward, our dfeat):", worst)
d = (feat.grad.detach().cpu() - dfe).abs()
ref_abs = dfe.abs()
print("dfeat: max err", float(d.max()), "rms err", float(d.pow(2).mean().sqrt()), "rms ref", float(ref_abs.pow(2).mean().sqrt()),
      "nonzero", int((ref_abs > 0).sum()), "frac err>1e-3*max", float((d > 1e-3 * ref_abs.max()).float().sum() / (ref_abs > 0).sum()))
idx = torch.nonzero(d == d.max())[0]
print("worst element", idx.tolist(), float(feat.grad[tuple(idx)]), float(dfe[tuple(idx)]))
# emb-gradient in fp64 head? compare our demb per point: error concentrated on few points?
dpts = d.sum(1).view(d.shape[0], -1)      # (B, HW)
top = torch.topk(dpts.view(-1), 5)
print("top-5 per-pixel abs err sums", top.values.tolist(), "total err sum", float(dpts.sum()))
