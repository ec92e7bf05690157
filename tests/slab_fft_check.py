"""Slab-decomposed FFT Poisson solve (ipplb_poisson_create_slab + ipplb_loop_poisson_solve) on in-process ranks of ONE GPU
against (a) the single-rank cuFFT solver on the whole domain and (b) numpy's half-spectrum solve with the oracle's k-space
multipliers.  Run as a script by tests/test_zz_slab_fft_gpu.py (own process: the code below has not run on a GPU yet, and a
CUDA fault must not take the rest of the suite with it).  Prints SLAB_FFT_OK <json> on success."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import oracle  # noqa: E402
from test_gpu_loop import orb_like_boxes  # noqa: E402


def main():
    import torch
    import ippl_b200 as ib
    worst = {}
    for ng, world, kind in (((24, 16, 16), 2, "default"), ((24, 16, 16), 4, "orb"), ((32, 24, 20), 8, "default"),
                            ((24, 16, 16), 8, "orb"), ((18, 10, 6), 4, "default")):
        origin, h = (0.0, 0.5, -1.0), (0.3, 0.25, 0.4)
        ctxs = [ib.Context(0) for _ in range(world)]
        dev = ctxs[0].device
        loop = ib.Loop(ctxs)
        layout = ib.Layout(ng, world)
        if kind == "orb":
            layout.set_boxes(orb_like_boxes(ng, world))
        boxes = layout.boxes()
        rng = np.random.default_rng(11)
        rho_g = rng.normal(size=(ng[2], ng[1], ng[0]))
        rho_g -= rho_g.mean()
        N, nxh = ng[0] * ng[1] * ng[2], ng[0] // 2 + 1
        rhat = np.fft.rfftn(rho_g) / N
        want = np.stack([np.fft.irfftn(rhat * np.broadcast_to(M, rho_g.shape)[:, :, :nxh], s=rho_g.shape, axes=(0, 1, 2)) * N
                         for M in oracle.poisson_kspace_multipliers(ng, origin, h)], axis=-1)
        g = 1
        rhos, efs, solvers, meshes = [], [], [], []
        for r in range(world):
            ctxs[r].set_layout(layout, origin, h)
            mesh = layout.mesh(r, origin, h)
            b = boxes[r]
            nl = [int(b[3 + d] - b[d] + 1) for d in range(3)]
            a = np.full((nl[2] + 2 * g, nl[1] + 2 * g, nl[0] + 2 * g), np.nan)
            a[g:-g, g:-g, g:-g] = rho_g[b[2]:b[5] + 1, b[1]:b[4] + 1, b[0]:b[3] + 1]
            rhos.append(torch.from_numpy(a.ravel().copy()).to(dev))
            efs.append(torch.full((a.size * 3,), float("nan"), dtype=torch.float64, device=dev))
            solvers.append(ib.Poisson(ctxs[r], None, layout=layout, origin=origin, h=h, slab=True))
            meshes.append(mesh)
        for rep in range(2):   # twice: the second solve must not depend on what the first left in the work buffers
            for r in range(world):
                b = boxes[r]
                nl = [int(b[3 + d] - b[d] + 1) for d in range(3)]
                a = np.full((nl[2] + 2 * g, nl[1] + 2 * g, nl[0] + 2 * g), np.nan)
                a[g:-g, g:-g, g:-g] = rho_g[b[2]:b[5] + 1, b[1]:b[4] + 1, b[0]:b[3] + 1]
                rhos[r].copy_(torch.from_numpy(a.ravel().copy()))
            loop.poisson_solve(solvers, rhos, efs)
            torch.cuda.synchronize()
        scale = np.abs(want).max()
        err = 0.0
        for r in range(world):
            b = boxes[r]
            nl = [int(b[3 + d] - b[d] + 1) for d in range(3)]
            ef = efs[r].cpu().numpy().reshape(nl[2] + 2 * g, nl[1] + 2 * g, nl[0] + 2 * g, 3)
            got = ef[g:-g, g:-g, g:-g]
            ref = want[b[2]:b[5] + 1, b[1]:b[4] + 1, b[0]:b[3] + 1]
            assert np.isfinite(got).all(), f"{ng} {world} {kind} rank {r}: E interior not fully written"
            err = max(err, float(np.max(np.abs(got - ref)) / scale))
            halo = ef.copy()
            halo[g:-g, g:-g, g:-g] = np.nan
            assert np.isnan(halo).all(), f"rank {r}: the solve wrote into E's ghost layers"
            rho = rhos[r].cpu().numpy().reshape(nl[2] + 2 * g, nl[1] + 2 * g, nl[0] + 2 * g)
            assert np.array_equal(rho[g:-g, g:-g, g:-g], got[..., 2]), f"rank {r}: rho interior != last gradient component"
        assert err <= 1e-12, (ng, world, kind, err)
        worst[f"{ng} x{world} {kind}"] = err
        # (a) the single-rank cuFFT solver on the whole domain (even sizes: its 3-D plan and the 2-D + 1-D plans agree to rounding)
        one = ib.Context(0)
        whole = ib.Layout(ng, 1)
        m1 = whole.mesh(0, origin, h)
        a = np.zeros((ng[2] + 2, ng[1] + 2, ng[0] + 2))
        a[1:-1, 1:-1, 1:-1] = rho_g
        rho1 = torch.from_numpy(a.ravel().copy()).to(dev)
        ef1 = torch.zeros(a.size * 3, dtype=torch.float64, device=dev)
        s1 = ib.Poisson(one, m1)
        s1.solve(rho1, ef1)
        torch.cuda.synchronize()
        e1 = ef1.cpu().numpy().reshape(ng[2] + 2, ng[1] + 2, ng[0] + 2, 3)[1:-1, 1:-1, 1:-1]
        assert np.max(np.abs(e1 - want)) / scale <= 1e-12
        s1.close(); whole.close(); one.close()
        for s in solvers:
            s.close()
        loop.close()
        layout.close()
        for c in ctxs:
            c.close()
    print("SLAB_FFT_OK", json.dumps(worst))


if __name__ == "__main__":
    main()
