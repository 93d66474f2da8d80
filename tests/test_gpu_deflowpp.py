"""GPU parity of the fused SeFlow++ path (embedder -> tcgen05 backbone -> ConvGRU decoder) against the
CPU oracle restatement of DeFlowPP.forward and against the goldens produced by the reference's own
class.  Bar: valid-point indices bit-exact; flow <= 1e-4 abs (north_star) in the fp32 (split-bf16) mode."""
import ctypes
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from himo_b200 import _lib, deflowpp, frames, weights
from oracle import deflowpp_ref

pytestmark = pytest.mark.gpu
FLOW_TOL = 1e-4


def _batch(tr):
    b = {k: torch.from_numpy(np.ascontiguousarray(tr[k]))[None].cuda() for k in ("pc0", "pc1", "pch1")}
    b.update({k: [torch.from_numpy(np.asarray(tr[k]))] for k in ("pose0", "pose1", "poseh1")})
    return b


@pytest.fixture(scope="module")
def net0():
    net = deflowpp.DeFlowPP(max_points=32768)
    return net


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "deflowpp_*.npz"))))
def test_matches_reference_golden(path, net0):
    z = np.load(path)
    net0.load_state_dict(weights.synth_deflowpp_state_dict(int(z["weight_seed"])))
    out = net0(_batch(z))
    assert (out["pc0_valid_point_idxes"][0].cpu().numpy() == z["pc0_valid_point_idxes"]).all()
    np.testing.assert_array_equal(out["pose_flow"][0].cpu().numpy(), z["pose_flow"])   # bit-exact rigid warp (NaN rows equal)
    err = np.abs(out["flow"][0].cpu().numpy() - z["flow"]).max()
    assert err <= FLOW_TOL, err


@pytest.mark.parametrize("kind,n,seed", [("lidar", 20000, 31), ("uniform", 30000, 32)])
def test_stagewise_against_oracle(kind, n, seed, net0):
    sd = weights.synth_deflowpp_state_dict(seed)
    net0.load_state_dict(sd)
    tr = frames.lidar_triple(n, seed) if kind == "lidar" else frames.uniform_triple(n, seed)
    ref = deflowpp_ref.deflowpp_forward(sd, tr["pch1"], tr["pc0"], tr["pc1"], tr["poseh1"], tr["pose0"],
                                        tr["pose1"], accum="exact", keep=True)
    out = net0(_batch(tr))
    torch.cuda.synchronize()
    # --- embedder: voxel order / counts / inverse map bit-exact, features <= 1e-5
    _WS["ws"] = net0._ws
    v = net0.views()
    ev = _embed_view(net0)
    names = ("info_h1", "info_0", "info_1")
    for f, nm in enumerate(names):
        info = ref[nm]
        m = int(ev["num_voxels"][f])
        assert m == info["voxel_coors"].shape[0]
        key_ref = (info["voxel_coors"][:, 1] * 512 + info["voxel_coors"][:, 2]).numpy()
        assert (ev["voxel_key"][f][:m] == key_ref).all()
        assert (ev["voxel_count"][f][:m] == info["voxel_count"].numpy()).all()
        rank = ev["rank"][f][: tr[("pch1", "pc0", "pc1")[f]].shape[0]]
        valid = rank >= 0
        assert (np.nonzero(valid)[0] == info["point_idxes"].numpy()).all()
        assert (rank[valid] == info["point2voxel"].numpy()).all()
        np.testing.assert_allclose(ev["voxel_mean"][f][:m], info["voxel_mean"].numpy(), rtol=0, atol=2e-6)
        np.testing.assert_allclose(ev["voxel_feats"][f][:m], info["voxel_feats"].numpy(), rtol=1e-5, atol=1e-5)
    # --- backbone output and final flow
    V = _read_f32(v["V"], 512 * 512 * 96).reshape(512, 512, 96)
    after = ref["after"].permute(1, 2, 0).numpy()
    assert np.abs(V - after).max() <= 1e-4 * max(1.0, np.abs(after).max())
    assert (out["pc0_valid_point_idxes"][0].cpu() == ref["pc0_valid_point_idxes"]).all()
    err = (out["flow"][0].cpu() - ref["flow"]).abs().max().item()
    assert err <= FLOW_TOL, err


def test_bf16_mode_runs_and_is_close():
    sd = weights.synth_deflowpp_state_dict(5)
    tr = frames.lidar_triple(8000, 5)
    ref = deflowpp_ref.deflowpp_forward(sd, tr["pch1"], tr["pc0"], tr["pc1"], tr["poseh1"], tr["pose0"], tr["pose1"])
    net = deflowpp.DeFlowPP(precision="bf16", max_points=8192).load_state_dict(sd)
    out = net(_batch(tr))
    assert (out["pc0_valid_point_idxes"][0].cpu() == ref["pc0_valid_point_idxes"]).all()
    err = (out["flow"][0].cpu() - ref["flow"]).abs().max().item()
    scale = ref["flow"].abs().max().item()
    assert err < 0.1 * scale, (err, scale)       # plain bf16: a few % of the flow scale, far above 1e-4


def test_empty_and_tiny_frames(net0):
    net0.load_state_dict(weights.synth_deflowpp_state_dict(0))
    tr = frames.uniform_triple(64, 1)
    tr["pc0"] = np.zeros((0, 3), np.float32)
    out = net0(_batch(tr))
    assert out["flow"][0].shape == (0, 3) and out["pc0_valid_point_idxes"][0].shape == (0,)
    tr = frames.uniform_triple(5, 2)
    tr["pc0"][:] = 1000.0            # every point out of range
    out = net0(_batch(tr))
    assert out["flow"][0].shape == (0, 3)


# ---------------------------------------------------------------------------------------------- helpers
_WS = {}


def _slice(ptr, nbytes):
    ws = _WS["ws"]
    off = ptr - ws.data_ptr()
    assert 0 <= off and off + nbytes <= ws.numel()
    torch.cuda.synchronize()
    return ws[off:off + nbytes]


def _read_f32(ptr, count):
    return _slice(ptr, count * 4).view(torch.float32).cpu().numpy()


def _read_i32(ptr, count):
    return _slice(ptr, count * 4).view(torch.int32).cpu().numpy()


class _EmbedView(ctypes.Structure):
    _fields_ = [(k, ctypes.c_void_p) for k in ("pt4", "bitmap", "word_prefix", "num_voxels", "rank",
                                               "voxel_count", "seg_start", "sorted_idx", "voxel_feats",
                                               "voxel_mean", "voxel_key")] + [("n_words", ctypes.c_int)]


def _embed_view(net):
    L = _lib.lib()
    L.himo_embed_views.restype = ctypes.c_int
    L.himo_embed_views.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                   ctypes.POINTER(_EmbedView)]
    vs = (ctypes.c_float * 3)(0.2, 0.2, 6.0)
    cr = (ctypes.c_float * 6)(-51.2, -51.2, -3.0, 51.2, 51.2, 3.0)
    ev = _EmbedView()
    st = L.himo_embed_views(3, net._n_max, vs, cr, ctypes.c_void_p(net.views()["embed_ws"]), ctypes.byref(ev))
    assert st == 0
    n = net._n_max
    res = {"num_voxels": _read_i32(ev.num_voxels, 3)}
    res["rank"] = _read_i32(ev.rank, 3 * n).reshape(3, n)
    res["voxel_count"] = _read_i32(ev.voxel_count, 3 * (n + 1)).reshape(3, n + 1)
    res["voxel_key"] = _read_i32(ev.voxel_key, 3 * n).reshape(3, n)
    res["voxel_mean"] = _read_f32(ev.voxel_mean, 3 * n * 3).reshape(3, n, 3)
    res["voxel_feats"] = _read_f32(ev.voxel_feats, 3 * n * 32).reshape(3, n, 32)
    return res


def test_fused_decoder_matches_unfused_path():
    """The fused persistent ConvGRU decoder (csrc/decfused.cu) against the separate GEMM + element-wise launches
    it replaces (same weights, same frame): the two differ only in accumulation order / the 22-bit hidden state."""
    from himo_b200 import _lib
    L = _lib.lib()
    sd = weights.synth_deflowpp_state_dict(3)
    tr = frames.lidar_triple(20000, 41)
    net = deflowpp.DeFlowPP(max_points=20480)
    net.load_state_dict(sd)
    batch = _batch(tr)
    try:
        L.himo_deflowpp_set_fused_decoder(0)
        ref = net(batch)["flow"][0].clone()
        L.himo_deflowpp_set_fused_decoder(1)
        got = net(batch)["flow"][0].clone()
    finally:
        L.himo_deflowpp_set_fused_decoder(1)
    assert ref.shape == got.shape and ref.abs().max() > 0
    assert (ref - got).abs().max().item() < 2e-5


@pytest.mark.parametrize("n0", [1, 127, 256, 257, 1000])
def test_ragged_point_counts_against_oracle(n0, net0):
    """Tile-boundary cases of the fused decoder (256-point CTA-pair tiles, padded rows) and of the embedder:
    pc0 with 1 / 127 / 256 / 257 / 1000 points against dense neighbour frames, flow vs the CPU oracle."""
    sd = weights.synth_deflowpp_state_dict(7)
    net0.load_state_dict(sd)
    tr = frames.lidar_triple(6000, 50 + n0)
    tr = dict(tr)
    tr["pc0"] = np.ascontiguousarray(tr["pc0"][:n0])
    ref = deflowpp_ref.deflowpp_forward(sd, tr["pch1"], tr["pc0"], tr["pc1"], tr["poseh1"], tr["pose0"], tr["pose1"])
    out = net0(_batch(tr))
    assert (out["pc0_valid_point_idxes"][0].cpu() == ref["pc0_valid_point_idxes"]).all()
    got = out["flow"][0].cpu()
    assert got.shape == ref["flow"].shape
    if got.numel():
        assert (got - ref["flow"]).abs().max().item() <= FLOW_TOL
