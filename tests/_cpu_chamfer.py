"""Test helper: an `nnChamferDis` look-alike on CPU tensors backed by the brute-force oracle (oracle/leaf_ops.c), so
that the host logic built on top of `himo_b200.chamfer3d` (losses, NSFP loop) can be exercised without a GPU.  Lives
under tests/ because only tests may touch oracle/."""
import numpy as np
import torch

from oracle import leaf


class _CpuChamfer(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pc0, pc1):
        a, b = pc0.detach().contiguous(), pc1.detach().contiguous()
        d0, d1, i0, i1 = leaf.chamfer_forward(a.numpy(), b.numpy())
        i0, i1 = torch.from_numpy(i0), torch.from_numpy(i1)
        ctx.save_for_backward(a, b, i0, i1)
        ctx.mark_non_differentiable(i0, i1)
        return torch.from_numpy(d0), torch.from_numpy(d1), i0, i1

    @staticmethod
    def backward(ctx, g0, g1, _a, _b):
        a, b, i0, i1 = ctx.saved_tensors
        ga, gb = leaf.chamfer_backward(a.numpy(), b.numpy(), i0.numpy(), i1.numpy(),
                                       g0.contiguous().numpy(), g1.contiguous().numpy())
        return torch.from_numpy(np.asarray(ga)), torch.from_numpy(np.asarray(gb))


class CpuChamferDis:
    def __call__(self, a, b, truncate_dist=-1):
        d0, d1, _, _ = _CpuChamfer.apply(a, b)
        if truncate_dist <= 0:
            return d0.mean() + d1.mean()
        return torch.nanmean(d0[d0 <= truncate_dist]) + torch.nanmean(d1[d1 <= truncate_dist])

    def dis_res(self, a, b):
        return _CpuChamfer.apply(a, b)[:2]

    def disid_res(self, a, b):
        return _CpuChamfer.apply(a, b)

    def truncated_dis(self, a, b, truncate_dist=2):
        d0, d1, _, _ = _CpuChamfer.apply(a, b)
        z = torch.zeros(())
        return torch.where(d0 >= truncate_dist, z, d0).mean() + torch.where(d1 >= truncate_dist, z, d1).mean()
