"""GPU parity at the BASELINE size (100 k points per frame, BASELINE.json configs[1]/[2]) against the CPU oracle --
the small-case suites stop at 30 k points.  Same bars as there: indices bit-exact, flow <= 1e-4 abs."""
import numpy as np
import pytest
import torch

from himo_b200 import chamfer3d_ext, deflowpp, fastnsf, frames, weights
from oracle import deflowpp_ref, fastnsf_ref, leaf

pytestmark = pytest.mark.gpu
N = 100_000
FLOW_TOL = 1e-4


def _batch(tr):
    b = {k: torch.from_numpy(np.ascontiguousarray(tr[k]))[None].cuda() for k in ("pc0", "pc1", "pch1")}
    b.update({k: [torch.from_numpy(np.asarray(tr[k]))] for k in ("pose0", "pose1", "poseh1")})
    return b


@pytest.fixture(scope="module")
def net100k():
    return deflowpp.DeFlowPP(max_points=N)


@pytest.mark.parametrize("kind,seed", [("lidar", 41), ("uniform", 42)])
def test_seflowpp_100k_matches_oracle(kind, seed, net100k):
    """DeFlowPP.forward (OSF/src/models/deflow.py:115-158) on a 100 k-point triple: valid indices bit-exact, flow <= 1e-4."""
    sd = weights.synth_deflowpp_state_dict(seed)
    net100k.load_state_dict(sd)
    tr = frames.lidar_triple(N, seed) if kind == "lidar" else frames.uniform_triple(N, seed)
    ref = deflowpp_ref.deflowpp_forward(sd, tr["pch1"], tr["pc0"], tr["pc1"], tr["poseh1"], tr["pose0"], tr["pose1"])
    out = net100k(_batch(tr))
    assert (out["pc0_valid_point_idxes"][0].cpu() == ref["pc0_valid_point_idxes"]).all()
    err = (out["flow"][0].cpu() - ref["flow"]).abs().max().item()
    assert err <= FLOW_TOL, err


@pytest.mark.parametrize("kind,seed", [("lidar", 43), ("uniform", 44)])
def test_chamfer_100k_sampled_queries_bit_exact(kind, seed):
    """chamfer3D.forward (chamfer3D.cu:33-105) at 100 k x 100 k: 2 000 sampled queries per direction against the
    brute-force oracle over the WHOLE other cloud -- squared distance and index bit-equal."""
    if kind == "lidar":
        tr = frames.lidar_triple(N, seed)
        a, b = np.ascontiguousarray(tr["pc0"][:, :3]), np.ascontiguousarray(tr["pc1"][:, :3])
    else:
        a, b = frames.uniform_frame(N, seed)[:, :3].copy(), frames.uniform_frame(N - 11, seed + 1)[:, :3].copy()
    pa, pb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    d0 = torch.zeros(len(a), device="cuda"); d1 = torch.zeros(len(b), device="cuda")
    i0 = torch.zeros(len(a), dtype=torch.int32, device="cuda"); i1 = torch.zeros(len(b), dtype=torch.int32, device="cuda")
    assert chamfer3d_ext.forward(pa, pb, d0, d1, i0, i1) == 1
    rng = np.random.default_rng(seed)
    qa, qb = rng.choice(len(a), 2000, replace=False), rng.choice(len(b), 2000, replace=False)
    r0 = leaf.chamfer_forward(np.ascontiguousarray(a[qa]), b)          # (dist0, dist1, idx0, idx1); only direction 0 is used
    r1 = leaf.chamfer_forward(np.ascontiguousarray(b[qb]), a)
    assert (d0.cpu().numpy()[qa] == r0[0]).all() and (i0.cpu().numpy()[qa] == r0[2]).all()
    assert (d1.cpu().numpy()[qb] == r1[0]).all() and (i1.cpu().numpy()[qb] == r1[2]).all()


def test_fastnsf_100k_one_iteration_matches_autograd():
    """FastNSF.optimize (OSF/src/models/fastnsf.py:137-164), one iteration at 100 k points on an identical distance
    volume: loss <= 1e-5 relative, full parameter gradient (read back through Adam's first moment) cos > 1 - 1e-5,
    first-iteration flow <= 1e-4."""
    tr = frames.lidar_triple(N, 45)
    pc0, pc1 = torch.from_numpy(tr["pc0"]), torch.from_numpy(tr["pc1"])
    T = deflowpp_ref.pose0to1(torch.from_numpy(tr["pose0"]), torch.from_numpy(tr["pose1"]))
    sel0 = pc0[fastnsf_ref.range_mask(pc0)]
    pc0 = (sel0 @ T[:3, :3].T + T[:3, 3]).contiguous()
    pc1 = pc1[fastnsf_ref.range_mask(pc1)].contiguous()
    sd = weights.synth_neural_prior_state_dict(45)
    lo_ref, hi_ref = fastnsf_ref.dt_bounds(pc0, pc1, 10.0)
    D_ref = fastnsf_ref.dt_build(pc1, lo_ref, hi_ref, 10.0)
    params = [p.requires_grad_(True) for p in fastnsf_ref.params_from_state_dict(sd)]
    flow = fastnsf_ref.mlp_forward(params, pc0[None])[0]
    loss = fastnsf_ref.dt_lookup(D_ref, lo_ref, 10.0, pc0 + flow).mean()
    loss.backward()
    g_ref = torch.cat([p.grad.reshape(-1) for p in params])
    one = fastnsf.FastNSF(itr_num=2, early_patience=1, min_delta=1e9)    # 2 loss evaluations, 1 update
    o1 = one.optimize(pc0.cuda(), pc1.cuda(), init_state_dict=sd, D=D_ref.cuda(), lo=lo_ref.numpy(),
                      dims=tuple(D_ref.shape), return_params=True)
    g = o1["exp_avg"].cpu() / 0.1
    assert torch.nn.functional.cosine_similarity(g, g_ref, dim=0).item() > 1 - 1e-5
    assert (g - g_ref).abs().max().item() <= 1e-3 * g_ref.abs().max().item()
    f1 = fastnsf.FastNSF(itr_num=1).optimize(pc0.cuda(), pc1.cuda(), init_state_dict=sd, D=D_ref.cuda(),
                                             lo=lo_ref.numpy(), dims=tuple(D_ref.shape))
    assert abs(f1["loss"] - float(loss)) <= 1e-5 * max(1.0, abs(float(loss)))
    assert (f1["flow"].cpu() - flow.detach()).abs().max().item() <= FLOW_TOL
    # and the CUDA distance volume at this size is the oracle's, bit for bit
    lo, dims = fastnsf.volume_geometry(pc0.cuda(), pc1.cuda(), 10.0)
    assert dims == tuple(D_ref.shape)
    assert torch.equal(fastnsf.dt_build(pc1.cuda(), lo, dims, 10.0).cpu(), D_ref)
