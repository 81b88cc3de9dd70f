"""Worker for the world_size-2 gloo tests (CPU): exercises the host side of the
multi-rank path -- rendezvous, export exchange, Decomposition neighbour rings
(the same direction table the CUDA kernel uses), block scatter/gather -- with a
numpy emulation of one block's ghost-frame step standing in for the GPU block.
Writes result .npy files that the parent test compares with the single-rank oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch.distributed as dist                                  # noqa: E402

from latticeboltzmann_b200 import distributed as D               # noqa: E402
from latticeboltzmann_b200._lib import DIRS, LbExport             # noqa: E402
from latticeboltzmann_b200.decomposition import Decomposition    # noqa: E402
from oracle import oracle as orc                                  # noqa: E402

CX, CY = orc.C_IC[:, 0], orc.C_IC[:, 1]
OPP = orc.OPPOSITE
# populations that leave through each direction slot (same table as csrc/step_kernel.cuh push_halo)
LEAVING = {0: (3, 6, 7), 1: (1, 5, 8), 2: (4, 7, 8), 3: (2, 5, 6), 4: (7,), 5: (6,), 6: (8,), 7: (5,)}


def ghost_slices(d, lnx, lny, ghost):
    """Index of the strip exchanged with direction d: the rim strip of the sender
    (ghost=False) or the ghost strip of the receiver seen from ITS side (ghost=True)."""
    dx, dy = DIRS[d]
    if ghost:
        dx, dy = -dx, -dy     # receiver: data from direction -d lands on that side's ghost
        sx = slice(1, lnx + 1) if dx == 0 else (slice(0, 1) if dx < 0 else slice(lnx + 1, lnx + 2))
        sy = slice(1, lny + 1) if dy == 0 else (slice(0, 1) if dy < 0 else slice(lny + 1, lny + 2))
    else:
        sx = slice(1, lnx + 1) if dx == 0 else (slice(1, 2) if dx < 0 else slice(lnx, lnx + 1))
        sy = slice(1, lny + 1) if dy == 0 else (slice(1, 2) if dy < 0 else slice(lny, lny + 1))
    return sx, sy


def opposite(d):
    return d ^ 1 if d < 4 else 11 - d          # 0<->1, 2<->3, 4<->7, 5<->6 (csrc/lattice.cuh dir_opp)


def exchange(G, decomp, rank, lnx, lny):
    """Ghost-frame exchange over torch.distributed p2p (stands in for the kernel's peer stores):
    for every direction slot d, send the rim strip that leaves through d to neighbour(d) and
    receive, from neighbour(opposite(d)), the strip that lands on my opposite-side ghost."""
    import torch
    sends, recvs = [], []
    for d in range(8):
        pops = list(LEAVING[d])
        sx, sy = ghost_slices(d, lnx, lny, ghost=False)
        gx, gy = ghost_slices(d, lnx, lny, ghost=True)
        out = np.ascontiguousarray(G[pops][:, sx, sy])
        dst, src = decomp.neighbour(rank, d), decomp.neighbour(rank, opposite(d))
        target = np.ix_(pops, range(gx.start, gx.stop), range(gy.start, gy.stop))
        if dst == rank and src == rank:            # ring closed on this rank
            G[target] = out
            continue
        buf = torch.empty(out.shape, dtype=torch.float64)
        sends.append(dist.isend(torch.from_numpy(out), dst, tag=d))
        recvs.append((dist.irecv(buf, src, tag=d), buf, target))
    for r in sends:
        r.wait()
    for r, buf, target in recvs:
        r.wait()
        G[target] = buf.numpy()


def block_step(G, blk, gnx, gny, omega, u0, boundary):
    """One fused step of a ghosted block (numpy): pull, global-coordinate walls, lid, collide."""
    lnx, lny = blk.lnx, blk.lny
    R = np.stack([G[i, 1 - CX[i]:1 - CX[i] + lnx, 1 - CY[i]:1 - CY[i] + lny] for i in range(9)])
    own = G[:, 1:lnx + 1, 1:lny + 1]
    post = R.copy()
    if boundary != "periodic":
        gk = (blk.x0 + np.arange(lnx))[:, None] * np.ones((1, lny), int)
        gl = (blk.y0 + np.arange(lny))[None, :] * np.ones((lnx, 1), int)
        walls = boundary == "cavity"
        bottom, top = gl == 0, gl == gny - 1
        left, right = walls & (gk == 0), walls & (gk == gnx - 1)
        for i in range(1, 9):
            outside = ((CY[i] == 1) & bottom) | ((CY[i] == -1) & top) | ((CX[i] == 1) & left) | ((CX[i] == -1) & right)
            post[i] = np.where(outside, own[OPP[i]], post[i])
        rho = own[6] + own[2] + own[5] + R[6] + R[2] + R[5] + R[3] + R[0] + R[1]
        lid = 6 * np.float64(1 / 36) * rho * u0
        post[8] = np.where(top & ~left, own[6] + lid, post[8])
        post[7] = np.where(top & ~right, own[5] - lid, post[7])
    post = np.ascontiguousarray(post)
    orc.collide(post.reshape(9, -1), omega)
    G[:, 1:lnx + 1, 1:lny + 1] = post


def main():
    outdir, boundary, ndx, ndy, nx, ny, nsteps = sys.argv[1], sys.argv[2], *map(int, sys.argv[3:8])
    rank, world, _ = D.init_process_group("gloo")
    decomp = Decomposition(nx, ny, ndx, ndy)
    blk = decomp.block(rank)
    # export exchange round trip (struct <-> bytes), as DistributedLattice does
    e = LbExport(pid=os.getpid(), lnx=blk.lnx, lny=blk.lny, device=rank)
    exports = [LbExport.from_buffer_copy(b) for b in D.exchange_blobs(bytes(e))]
    assert [x.device for x in exports] == list(range(world))
    assert exports[rank].pid == os.getpid()
    for d, nb in enumerate(decomp.neighbours(rank)):
        # faces must match, exactly what lb_connect validates
        if DIRS[d][1] == 0:
            assert exports[nb].lny == blk.lny
        if DIRS[d][0] == 0:
            assert exports[nb].lnx == blk.lnx
    f0 = orc.perturbed_state(nx, ny, seed=21)
    G = np.zeros((9, blk.lnx + 2, blk.lny + 2))
    G[:, 1:-1, 1:-1] = decomp.scatter(f0, rank)
    exchange(G, decomp, rank, blk.lnx, blk.lny)
    for _ in range(nsteps):
        block_step(G, blk, nx, ny, 1.7, 0.1, boundary)
        exchange(G, decomp, rank, blk.lnx, blk.lny)
    g = D.gather_blocks(np.ascontiguousarray(G[:, 1:-1, 1:-1]), decomp, dst=0)
    t = D.max_over_ranks(float(rank + 1))
    assert t == float(world)
    if rank == 0:
        np.save(os.path.join(outdir, "gathered.npy"), g)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
