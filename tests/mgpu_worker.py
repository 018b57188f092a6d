"""torchrun worker of tests/test_gpu_multi.py: block-decomposed run over NCCL vs the single-GPU run of the same
global problem.  Prints one RESULT line on rank 0."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cholla_b200 as cb  # noqa: E402
from cholla_b200 import initial_conditions as ic  # noqa: E402
from cholla_b200.grid import Grid3D  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, steps, gamma = (64, 48, 40), 6, 1.66666667
    flags = [int(v) for v in os.environ.get("MGPU_FLAGS", "1,1,1,1,1,1").split(",")]
    strict = os.environ.get("MGPU_STRICT", "1") == "1"
    mhd = os.environ.get("MGPU_MHD", "0") == "1"      # Orszag-Tang, PPMC + HLLD + CT (BASELINE config 4, reduced size)
    overlap = os.environ.get("MGPU_OVERLAP", "1") == "1"
    transport = os.environ.get("MGPU_TRANSPORT", "p2p")
    # a -DDE -DSCALAR -DSTATIC_GRAV build: 7 fields through the exchange, Select_Internal_Energy reading neighbours across
    # rank boundaries, and an analytic field evaluated from each rank's global offsets (Milky-Way model, custom_grav = 1)
    passive = os.environ.get("MGPU_PASSIVE", "0") == "1"
    nf = 8 if mhd else (7 if passive else 5)
    solver = cb.HLLD if mhd else cb.HLLC
    if mhd:
        n, steps, gamma = (64, 64, 64), 5, 1.666666666666667
    kw = dict(recon=cb.PPMC, solver=solver, gamma=gamma, bc=flags, strict=strict, mhd=mhd)
    if passive:
        kw.update(n_scalars=1, dual_energy=True, bounds_min=(-0.53, -0.49, -0.51))
    G = Grid3D(n, (1.0, 1.0, 1.0), device="cuda:%d" % local, dist_group=dist.group.WORLD, rank=rank, nproc=world,
               halo_transport=transport, overlap=overlap, **kw)
    if passive:
        G.Set_Static_Gravity(1)
    if mhd:
        from cholla_b200.grid import n_ghost_for
        with ic.block(G.decomp.offset, G.decomp.n_local):
            G.upload(ic.orszag_tang(n, n_ghost_for(cb.PPMC, True), gamma=gamma).reshape(-1))
    else:
        fields = ic.spherical_overpressure(n, gamma=gamma)
        if passive:
            fields = ic.with_passive(fields, 1, True, gamma)
        G.Set_Initial_Conditions(fields)
    G.Set_Boundary_Conditions_Grid()
    dts = G.advance(steps)
    mine = torch.from_numpy(G.real_cells()).cuda()
    halo_kind = G.halo_kind()
    # gather every rank's block on rank 0
    out = None
    if rank == 0:
        out = np.zeros((nf, n[2], n[1], n[0]))
    for r in range(world):
        D = type(G.decomp).build(world, r, list(n), flags)
        shape = (nf, D.n_local[2], D.n_local[1], D.n_local[0])
        buf = mine if r == rank else torch.empty(shape, dtype=torch.float64, device="cuda")
        if r != 0:
            if rank == r:
                dist.send(mine, 0)
            elif rank == 0:
                dist.recv(buf, r)
        if rank == 0:
            o = D.offset
            out[:, o[2]:o[2] + shape[1], o[1]:o[1] + shape[2], o[0]:o[0] + shape[3]] = (mine if r == 0 else buf).cpu().numpy()
    G.close()
    if rank == 0:
        S = Grid3D(n, (1.0, 1.0, 1.0), device="cuda:0", **kw)
        if passive:
            S.Set_Static_Gravity(1)
        if mhd:
            S.upload(ic.orszag_tang(n, S.H.n_ghost, gamma=gamma).reshape(-1))
        else:
            S.Set_Initial_Conditions(fields)
        S.Set_Boundary_Conditions_Grid()
        dts1 = S.advance(steps)
        one = S.real_cells()
        S.close()
        err = float(np.abs(out - one).sum() / np.abs(one).sum())
        print("RESULT " + json.dumps(dict(world=world, equal=bool(np.array_equal(out, one)), dt_equal=bool(dts == dts1),
                                          rel_l1=err, decomposition=G.decomp.np, mhd=mhd, halo=halo_kind,
                                          overlap=overlap)))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
