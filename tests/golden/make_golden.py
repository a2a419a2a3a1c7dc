"""Generate tests/golden/*.npz by running the REFERENCE ITSELF (imported read-only from
/root/reference, CPU fp32) on the seeded synthetic inputs of densefusion_b200/synth.py.

Run in the build container only:   python tests/golden/make_golden.py
The GPU box has no /root/reference; it consumes the committed .npz files.

What is the reference and what is restated here:
  * PoseNet / PoseRefineNet / Loss / Loss_refine : the reference modules, unmodified.
  * ADD-S branch: the fork's lib/loss.py:44 calls lib/nn.py with the wrong contract and raises
    (SURVEY.md 0.3).  The reference loss files still run unmodified once the module attribute
    `nn_distance` is rebound to a function with the upstream KNearestNeighbor(1) contract
    (lib/knn/__init__.py:15-23); the bound function is oracle/knn_ref.c (bit-exact emulator of
    lib/knn/src/knn_cuda_kernel.cu).  The reference CUDA kernel itself is checked against the
    emulator on the GPU box (tests/test_knn_gpu.py).
  * eval-time refine loop: restated from tools/eval_ycb.py:193-233 around the reference's own
    lib/transformations.py quaternion_matrix / quaternion_from_matrix.
"""
import copy
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
warnings.filterwarnings("ignore")

from lib.network import PoseNet, PoseRefineNet          # noqa: E402  (reference)
from lib.loss import Loss                               # noqa: E402  (reference)
from lib.loss_refiner import Loss_refine                # noqa: E402  (reference)
import lib.loss as ref_loss_mod                         # noqa: E402
import lib.loss_refiner as ref_loss_refiner_mod         # noqa: E402
from lib.transformations import quaternion_matrix, quaternion_from_matrix  # noqa: E402

from densefusion_b200 import synth                      # noqa: E402
from oracle import df_oracle                            # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(8)


def knn1(ref, query):
    return df_oracle.knn(ref, query, 1)


ref_loss_mod.nn_distance = knn1
ref_loss_refiner_mod.nn_distance = knn1


def npy(t):
    return t.detach().cpu().numpy()


def build_nets(num_points, num_obj, seed):
    est = PoseNet(num_points=num_points, num_obj=num_obj)
    ref = PoseRefineNet(num_points=num_points, num_obj=num_obj)
    est.load_state_dict(synth.synth_state_dict(synth.shapes_of(est), seed))
    ref.load_state_dict(synth.synth_state_dict(synth.shapes_of(ref), seed + 1))
    est.eval()
    ref.eval()
    return est, ref


def eval_loop(refiner, cloud, emb, index, my_r, my_t, iteration, num_points):
    """tools/eval_ycb.py:205-229 restated around the reference's transformations functions."""
    for _ in range(iteration):
        T = torch.from_numpy(my_t.astype(np.float32)).view(1, 3).repeat(num_points, 1).contiguous().view(1, num_points, 3)
        my_mat = quaternion_matrix(my_r)
        R = torch.from_numpy(my_mat[:3, :3].astype(np.float32)).view(1, 3, 3)
        my_mat[0:3, 3] = my_t
        new_cloud = torch.bmm((cloud - T), R).contiguous()
        pred_r, pred_t = refiner(new_cloud, emb, index)
        pred_r = pred_r.view(1, 1, -1)
        pred_r = pred_r / (torch.norm(pred_r, dim=2).view(1, 1, 1))
        my_r_2 = pred_r.view(-1).data.numpy()
        my_t_2 = pred_t.view(-1).data.numpy()
        my_mat_2 = quaternion_matrix(my_r_2)
        my_mat_2[0:3, 3] = my_t_2
        my_mat_final = np.dot(my_mat, my_mat_2)
        my_r_final = copy.deepcopy(my_mat_final)
        my_r_final[0:3, 3] = 0
        my_r_final = quaternion_from_matrix(my_r_final, True)
        my_t_final = np.array([my_mat_final[0][3], my_mat_final[1][3], my_mat_final[2][3]])
        my_r, my_t = my_r_final, my_t_final
    return my_r, my_t


def full_case(name, case, num_points, num_obj, num_pt_mesh, hw, obj, sym_list, seed, w=0.015, iters=2):
    est, refiner = build_nets(num_points, num_obj, seed)
    d = synth.synth_crop(case, num_points, num_pt_mesh, num_obj, hw, obj)
    out = {"meta": np.array([case, num_points, num_obj, num_pt_mesh, hw[0], hw[1], obj, seed, iters]),
           "sym_list": np.array(sym_list), "w": np.array(w)}
    with torch.no_grad():
        pred_r, pred_t, pred_c, emb = est(d["img"], d["points"], d["choose"], d["idx"])
    out.update(pred_r=npy(pred_r), pred_t=npy(pred_t), pred_c=npy(pred_c), emb=npy(emb))
    # loss forward + gradients wrt the three prediction tensors
    pr, pt, pc = [t.clone().requires_grad_(True) for t in (pred_r, pred_t, pred_c)]
    crit = Loss(num_pt_mesh, sym_list)
    crit_ref = Loss_refine(num_pt_mesh, sym_list)
    loss, dis, new_points, new_target = crit(pr, pt, pc, d["target"], d["model_points"], d["idx"], d["points"], w, False)
    loss.backward()
    out.update(loss=npy(loss), dis=npy(dis), new_points=npy(new_points), new_target=npy(new_target),
               g_pred_r=npy(pr.grad), g_pred_t=npy(pt.grad), g_pred_c=npy(pc.grad),
               which_max=np.array(int(torch.max(pred_c.view(1, -1), 1)[1][0])))
    c_sorted = torch.sort(pred_c.view(-1), descending=True)[0]
    out["c_top2_margin"] = npy(c_sorted[0] - c_sorted[1])
    # refine=True variant of Loss (never takes the symmetric branch)
    with torch.no_grad():
        l2, d2, np2, nt2 = crit(pred_r, pred_t, pred_c, d["target"], d["model_points"], d["idx"], d["points"], w, True)
    out.update(loss_refineflag=npy(l2), dis_refineflag=npy(d2))
    # training-style refine chain (tools/train.py:156-159)
    pts, tgt = new_points, new_target
    for it in range(iters):
        r_in = pts.clone()
        rr, tt = refiner(r_in, emb, d["idx"])
        rr_l, tt_l = rr.detach().clone().requires_grad_(True), tt.detach().clone().requires_grad_(True)
        dis_r, pts, tgt = crit_ref(rr_l, tt_l, tgt, d["model_points"], d["idx"], pts)
        dis_r.backward()
        out.update({f"train_r{it}": npy(rr), f"train_t{it}": npy(tt), f"train_dis{it}": npy(dis_r),
                    f"train_pts{it}": npy(pts), f"train_tgt{it}": npy(tgt),
                    f"train_g_r{it}": npy(rr_l.grad), f"train_g_t{it}": npy(tt_l.grad)})
    # eval-style pose (tools/eval_ycb.py:193-233)
    with torch.no_grad():
        q = pred_r / torch.norm(pred_r, dim=2).view(1, num_points, 1)
        which = torch.max(pred_c.view(1, num_points), 1)[1]
        my_r = q[0][which[0]].view(-1).data.numpy()
        my_t = (d["points"].view(num_points, 1, 3) + pred_t.view(num_points, 1, 3))[which[0]].view(-1).data.numpy()
        out.update(pose0=np.append(my_r, my_t))
        for n_it in (1, iters, 4):
            r_f, t_f = eval_loop(refiner, d["points"], emb, d["idx"], my_r.copy(), my_t.copy(), n_it, num_points)
            out[f"pose_iter{n_it}"] = np.append(r_f, t_f)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "loss", float(loss), "dis", float(dis), "margin", float(out["c_top2_margin"]),
          "pose", out[f"pose_iter{iters}"], "->", os.path.getsize(path), "bytes")


def loss_only_case(name, case, num_points, num_pt_mesh, obj, sym_list, w=0.015):
    """Loss / Loss_refine on head-independent predictions (ties in pred_c, M != N, big M)."""
    d = synth.synth_crop(case, num_points, num_pt_mesh, 21, (40, 40), obj)
    pred_r, pred_t, pred_c = synth.synth_predictions(case, num_points)
    # exact ties on the maximum: first index must win (torch.max semantics, lib/loss.py:54)
    pred_c[0, 7, 0] = 0.97
    pred_c[0, 300 % num_points, 0] = 0.97
    pr, pt, pc = [t.clone().requires_grad_(True) for t in (pred_r, pred_t, pred_c)]
    crit = Loss(num_pt_mesh, sym_list)
    loss, dis, new_points, new_target = crit(pr, pt, pc, d["target"], d["model_points"], d["idx"], d["points"], w, False)
    loss.backward()
    out = dict(meta=np.array([case, num_points, num_pt_mesh, obj]), sym_list=np.array(sym_list), w=np.array(w),
               pred_c_tied=npy(pred_c), loss=npy(loss), dis=npy(dis), new_points=npy(new_points),
               new_target=npy(new_target), g_pred_r=npy(pr.grad), g_pred_t=npy(pt.grad), g_pred_c=npy(pc.grad))
    # Loss_refine with a single hypothesis taken from the same predictions
    r1 = pred_r[0, 5].view(1, 4).clone().requires_grad_(True)
    t1 = (pred_t[0, 5] + torch.tensor([0.0, 0.0, 0.8])).view(1, 3).clone().requires_grad_(True)
    crit_ref = Loss_refine(num_pt_mesh, sym_list)
    dis_r, np_r, nt_r = crit_ref(r1, t1, d["target"], d["model_points"], d["idx"], d["points"])
    dis_r.backward()
    out.update(ref_dis=npy(dis_r), ref_new_points=npy(np_r), ref_new_target=npy(nt_r),
               ref_g_r=npy(r1.grad), ref_g_t=npy(t1.grad))
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "loss", float(loss), "dis", float(dis), "ref_dis", float(dis_r), "->", os.path.getsize(path), "bytes")


def grad_summary(named_grads):
    """Per-parameter summary small enough to commit: L2 norm, sum, and 32 strided samples."""
    out = {}
    for name, g in named_grads:
        if g is None:
            out["gnone." + name] = np.array(1)
            continue
        f = g.detach().reshape(-1).double()
        stride = max(1, f.numel() // 32)
        out["gnorm." + name] = np.array(float(f.norm()))
        out["gsum." + name] = np.array(float(f.sum()))
        out["gsamp." + name] = f[::stride][:32].float().numpy()
    return out


def train_case(name, cases, objs, num_points, num_obj, num_pt_mesh, hw, sym_list, seed, w=0.015, iters=2):
    """tools/train.py:143-169 on the reference modules: gradient ACCUMULATION over bs=1 samples (no division),
    estimator phase (loss.backward()) and refiner phase (dis.backward() inside the iteration loop), followed by the
    reference's optimiser (torch.optim.Adam, lr 1e-4, tools/train.py:97).  eval() mode: Dropout2d off."""
    est, refiner = build_nets(num_points, num_obj, seed)
    crops = [synth.synth_crop(c, num_points, num_pt_mesh, num_obj, hw, o) for c, o in zip(cases, objs)]
    crit, crit_ref = Loss(num_pt_mesh, sym_list), Loss_refine(num_pt_mesh, sym_list)
    out = {"meta": np.array([num_points, num_obj, num_pt_mesh, hw[0], hw[1], seed, iters]), "cases": np.array(cases),
           "objs": np.array(objs), "sym_list": np.array(sym_list), "w": np.array(w)}
    # ---- estimator phase ----
    opt = torch.optim.Adam(est.parameters(), lr=1e-4)
    opt.zero_grad()
    losses, dists = [], []
    for d in crops:
        pred_r, pred_t, pred_c, emb = est(d["img"], d["points"], d["choose"], d["idx"])
        loss, dis, _, _ = crit(pred_r, pred_t, pred_c, d["target"], d["model_points"], d["idx"], d["points"], w, False)
        loss.backward()
        losses.append(float(loss)); dists.append(float(dis))
    out.update(est_losses=np.array(losses), est_dis=np.array(dists))
    out.update({"est." + k: v for k, v in grad_summary([(n, p.grad) for n, p in est.named_parameters()]).items()})
    before = {n: p.detach().clone() for n, p in est.named_parameters()}
    opt.step()
    for n, p in est.named_parameters():
        f = (p.detach() - before[n]).reshape(-1)
        out["est.delta." + n] = f[::max(1, f.numel() // 32)][:32].numpy()
    # ---- refiner phase (fresh estimator weights, as loaded) ----
    est, _ = build_nets(num_points, num_obj, seed)
    opt = torch.optim.Adam(refiner.parameters(), lr=1e-4)
    opt.zero_grad()
    dists = []
    for d in crops:
        pred_r, pred_t, pred_c, emb = est(d["img"], d["points"], d["choose"], d["idx"])
        loss, dis, pts, tgt = crit(pred_r, pred_t, pred_c, d["target"], d["model_points"], d["idx"], d["points"], w, True)
        for _ in range(iters):
            rr, tt = refiner(pts, emb, d["idx"])
            dis, pts, tgt = crit_ref(rr, tt, tgt, d["model_points"], d["idx"], pts)
            dis.backward()
        dists.append(float(dis))
    out.update(ref_dis=np.array(dists))
    out.update({"ref." + k: v for k, v in grad_summary([(n, p.grad) for n, p in refiner.named_parameters()]).items()})
    before = {n: p.detach().clone() for n, p in refiner.named_parameters()}
    opt.step()
    for n, p in refiner.named_parameters():
        f = (p.detach() - before[n]).reshape(-1)
        out["ref.delta." + n] = f[::max(1, f.numel() // 32)][:32].numpy()
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "losses", losses, "ref dis", dists, "->", os.path.getsize(path), "bytes")


def ply_points(path):
    """binary_little_endian PLY with `element vertex N` and three double properties."""
    raw = open(path, "rb").read()
    end = raw.index(b"end_header\n") + len(b"end_header\n")
    header = raw[:end].decode("ascii")
    n = int([l for l in header.split("\n") if l.startswith("element vertex")][0].split()[-1])
    return np.frombuffer(raw[end:end + n * 24], dtype="<f8").reshape(n, 3).copy()


def ply_case():
    """Real-geometry fixture shipped in the reference tree (tools/eval_cad.py:130-136 output)."""
    pred = ply_points("/root/reference/pred_pcld_output.ply")
    tgt = ply_points("/root/reference/target_pcld_output.ply")
    add = np.mean(np.linalg.norm(pred - tgt, axis=1))
    p32, t32 = torch.from_numpy(pred.astype(np.float32)), torch.from_numpy(tgt.astype(np.float32))
    inds = knn1(t32.t().contiguous().unsqueeze(0), p32.t().contiguous().unsqueeze(0)).view(-1) - 1
    adds = torch.mean(torch.norm(p32 - t32[inds], dim=1)).item()
    np.savez_compressed(os.path.join(OUT, "ply_pair.npz"), pred=pred, target=tgt, add=add, adds=adds,
                        inds=inds.numpy().astype(np.int16))
    print("ply_pair ADD", add, "ADD-S", adds)


def customcad_case():
    """Real geometry shipped in the reference tree: datasets/customCAD/{depth_projected,model,target}.ply
    = an observed cloud (1000 points), the CAD model (500) and the model under the ground-truth pose (500).  The reference's
    OWN Loss / Loss_refine (ADD and, with the kNN rebound as above, ADD-S) and its transformations functions run on them with
    seeded per-point hypotheses scattered around the true pose."""
    base = "/root/reference/datasets/customCAD/"
    cloud = torch.from_numpy(ply_points(base + "depth_projected.ply").astype(np.float32)).view(1, -1, 3)
    model = torch.from_numpy(ply_points(base + "model.ply").astype(np.float32)).view(1, -1, 3)
    target = torch.from_numpy(ply_points(base + "target.ply").astype(np.float32)).view(1, -1, 3)
    n, m = cloud.shape[1], model.shape[1]
    g = torch.Generator().manual_seed(2024)
    # the rigid motion model -> target (Kabsch on the corresponding points), then noisy per-point hypotheses around it
    mc, tc = model[0].double().mean(0), target[0].double().mean(0)
    H = (model[0].double() - mc).t() @ (target[0].double() - tc)
    U, _, Vt = np.linalg.svd(H.numpy())
    R = Vt.T @ np.diag([1, 1, np.sign(np.linalg.det(Vt.T @ U.T))]) @ U.T
    M4 = np.eye(4); M4[:3, :3] = R
    q_gt = quaternion_from_matrix(M4, True)
    t_gt = tc.numpy() - R @ mc.numpy()
    pred_r = (torch.from_numpy(q_gt).float().view(1, 1, 4) * (1.0 + 0.3 * torch.rand(1, n, 1, generator=g))
              + 0.05 * torch.randn(1, n, 4, generator=g)).contiguous()          # un-normalised, as the head emits them
    pred_t = (torch.from_numpy(t_gt).float().view(1, 1, 3) - cloud + 0.004 * torch.randn(1, n, 3, generator=g)).contiguous()
    pred_c = (torch.rand(1, n, 1, generator=g) * 0.9 + 0.05).contiguous()
    out = dict(cloud=npy(cloud), model=npy(model), target=npy(target), pred_r=npy(pred_r), pred_t=npy(pred_t), pred_c=npy(pred_c),
               q_gt=q_gt, t_gt=t_gt, sym_list=np.array([5]), w=np.array(0.015))
    for tag, obj in (("add", 3), ("adds", 5)):
        idx = torch.tensor([[obj]])
        pr, pt, pc = [t.clone().requires_grad_(True) for t in (pred_r, pred_t, pred_c)]
        loss, dis, new_points, new_target = Loss(m, [5])(pr, pt, pc, target, model, idx, cloud, 0.015, False)
        loss.backward()
        out.update({f"{tag}_loss": npy(loss), f"{tag}_dis": npy(dis), f"{tag}_new_points": npy(new_points),
                    f"{tag}_new_target": npy(new_target), f"{tag}_g_r": npy(pr.grad), f"{tag}_g_t": npy(pt.grad),
                    f"{tag}_g_c": npy(pc.grad)})
        r1 = (torch.tensor([1.0, 0.01, -0.02, 0.015]) * 1.3).view(1, 4).requires_grad_(True)
        t1 = torch.tensor([[0.002, -0.001, 0.003]], requires_grad=True)
        dis_r, np_r, nt_r = Loss_refine(m, [5])(r1, t1, new_target, model, idx, new_points)
        dis_r.backward()
        out.update({f"{tag}_ref_dis": npy(dis_r), f"{tag}_ref_new_points": npy(np_r), f"{tag}_ref_new_target": npy(nt_r),
                    f"{tag}_ref_g_r": npy(r1.grad), f"{tag}_ref_g_t": npy(t1.grad)})
        print("customcad", tag, "loss", float(loss), "dis", float(dis), "ref_dis", float(dis_r))
    # eval-style selection and one pose composition with the reference's transformations (tools/eval_ycb.py:193-229)
    q = pred_r / torch.norm(pred_r, dim=2).view(1, n, 1)
    which = int(torch.max(pred_c.view(1, n), 1)[1][0])
    my_r = q[0][which].numpy()
    my_t = (cloud.view(n, 1, 3) + pred_t.view(n, 1, 3))[which].view(-1).numpy()
    my_mat = quaternion_matrix(my_r)
    Rm = torch.from_numpy(my_mat[:3, :3].astype(np.float32)).view(1, 3, 3)
    T = torch.from_numpy(my_t.astype(np.float32)).view(1, 1, 3)
    new_cloud = torch.bmm(cloud - T, Rm)
    my_mat[0:3, 3] = my_t
    r2 = np.array([0.99, 0.02, -0.03, 0.01]); r2n = r2 / np.linalg.norm(r2)
    t2 = np.array([0.003, -0.002, 0.001])
    m2 = quaternion_matrix(r2n); m2[0:3, 3] = t2
    final = np.dot(my_mat, m2)
    rot = copy.deepcopy(final); rot[0:3, 3] = 0
    out.update(which=np.array(which), pose0=np.append(my_r, my_t), new_cloud=npy(new_cloud), r2=r2.astype(np.float32), t2=t2.astype(np.float32),
               pose1=np.append(quaternion_from_matrix(rot, True), final[0:3, 3]))
    path = os.path.join(OUT, "customcad_triple.npz")
    np.savez_compressed(path, **out)
    print("customcad_triple ->", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    if "--ply-only" in sys.argv:
        ply_case()
        sys.exit(0)
    if "--customcad-only" in sys.argv:
        customcad_case()
        sys.exit(0)
    # C4 shape: training step (gradient accumulation over 3 samples: two objects share id 12 (symmetric), one is not)
    train_case("c4_train_ycb", cases=[40, 41, 42], objs=[12, 3, 12], num_points=500, num_obj=21, num_pt_mesh=500,
               hw=(80, 80), sym_list=synth.YCB_SYM, seed=4)
    if "--train-only" in sys.argv:
        sys.exit(0)
    # C0: LineMOD PoseNet(500,13), non-symmetric object, 80x80
    full_case("c0_linemod_add", case=0, num_points=500, num_obj=13, num_pt_mesh=500, hw=(80, 80), obj=3,
              sym_list=synth.LINEMOD_SYM, seed=0)
    # C1/C2 shape: YCB PoseNet(500,21), symmetric object -> ADD-S through the kNN (R=500, Q=250000)
    full_case("c1_ycb_adds", case=1, num_points=500, num_obj=21, num_pt_mesh=500, hw=(120, 120), obj=12,
              sym_list=synth.YCB_SYM, seed=2)
    # loss-only: confidence ties, N != M, symmetric and not
    loss_only_case("loss_add_n500_m500", case=10, num_points=500, num_pt_mesh=500, obj=3, sym_list=synth.YCB_SYM)
    loss_only_case("loss_adds_n500_m500", case=11, num_points=500, num_pt_mesh=500, obj=15, sym_list=synth.YCB_SYM)
    loss_only_case("loss_adds_n100_m2600", case=12, num_points=100, num_pt_mesh=2600, obj=19, sym_list=synth.YCB_SYM)
    loss_only_case("loss_add_n1000_m500", case=13, num_points=1000, num_pt_mesh=500, obj=0, sym_list=synth.YCB_SYM)
    ply_case()
    customcad_case()
