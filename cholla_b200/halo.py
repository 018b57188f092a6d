"""Ghost-cell exchange schedule of the block decomposition -- host logic only, transport = torch.distributed.

Mirrors Grid3D::Set_Boundaries_MPI_BLOCK (src/grid/mpi_boundaries.cpp:29-111): for each axis in x, y, z order
    Load_and_Send_MPI_Comm_Buffers(axis)   pack both slabs, post receives and sends        (:300-777)
    Set_Boundaries(2*axis), (2*axis+1)     physical faces of this axis, if any             (boundary_conditions.cpp:165)
    Wait_and_Unload_MPI_Comm_Buffers(axis) wait, unpack into the ghost slabs               (:779-830)
so the y slabs carry the x ghosts and the z slabs the x and y ghosts (edges and corners need no diagonal messages).

The class knows nothing about devices: the caller supplies `pack(face, slab, buffer)`, `unpack(face, slab, buffer)`
and `alloc(n)`.  Grid3D plugs in the CUDA kernels (chl_pack_buffer / chl_unpack_buffer) and NCCL; the CPU tests plug
in the oracle's pack/unpack and gloo (tests/test_halo_gloo.py)."""
from dataclasses import dataclass

BC_MPI = 5


@dataclass
class Slab:
    """Arguments of one PackBuffers3D / UnpackBuffers3D call (cuda_boundaries.cu:11-69)."""
    idxoffset: int
    isize: int
    jsize: int
    ksize: int

    @property
    def ncells(self):
        return self.isize * self.jsize * self.ksize


def halo_slab(nx, ny, nz, g, face, recv):
    """Slab sent from (recv=False) or received into (recv=True) face 0..5.

    Geometry of Load_/Unload_Hydro_DeviceBuffer_{X,Y,Z}{0,1} (src/grid/mpi_boundaries.cpp:113-298): g layers
    along the face's axis; the full extent (ghosts included) of the axes before it; the real cells of the
    axes after it -- so edges and corners travel with the later phases.
    """
    n = [nx, ny, nz]
    a, side = face // 2, face % 2
    lo, size = [0, 0, 0], [1, 1, 1]
    for d in range(3):
        if n[d] == 1:
            continue
        if d < a:
            lo[d], size[d] = 0, n[d]
        elif d > a:
            lo[d], size[d] = g, n[d] - 2 * g
        else:
            size[d] = g
            if recv:
                lo[d] = 0 if side == 0 else n[d] - g
            else:
                lo[d] = g if side == 0 else n[d] - 2 * g
    return Slab(lo[0] + nx * (lo[1] + ny * lo[2]), size[0], size[1], size[2])


class HaloExchange:
    def __init__(self, dist, group, dims, n_ghost, n_fields, flags, dest, alloc, pack, unpack):
        self.dist, self.group = dist, group
        self.nx, self.ny, self.nz = dims
        self.g, self.nf = n_ghost, n_fields
        self.flags, self.dest = list(flags), list(dest)
        self.alloc, self.pack, self.unpack = alloc, pack, unpack
        self.ndim = 1 + (self.ny > 1) + (self.nz > 1)
        self._bufs = {}
        self.messages = 0  # sends posted (bookkeeping for tests / bench)

    def _buf(self, key, n):
        b = self._bufs.get(key)
        if b is None or b.numel() != n:
            b = self.alloc(n)
            self._bufs[key] = b
        return b

    def post(self, axis):
        """Load_and_Send_MPI_Comm_Buffers(axis): returns the outstanding work handles."""
        dist = self.dist
        sends, recvs = [], []
        for face in (2 * axis, 2 * axis + 1):
            if self.flags[face] != BC_MPI:
                continue
            s = halo_slab(self.nx, self.ny, self.nz, self.g, face, recv=False)
            sb = self._buf(("s", face), s.ncells * self.nf)
            self.pack(face, s, sb)
            rb = self._buf(("r", face), s.ncells * self.nf)
            # my low slab lands in my low neighbour's HIGH ghosts (tag = receiving face), and vice versa
            sends.append(dist.P2POp(dist.isend, sb, self.dest[face], self.group, tag=face ^ 1))
            recvs.append((face, dist.P2POp(dist.irecv, rb, self.dest[face], self.group, tag=face)))
            self.messages += 1
        # a peer that is both neighbours (np = 2, periodic) sends low-then-high; what it sends first fills my
        # HIGH ghosts, so receives are posted high-then-low to match in-order pairing (NCCL has no tags)
        recvs.sort(key=lambda fr: -fr[0])
        ops = sends + [r for _, r in recvs]
        return dist.batch_isend_irecv(ops) if ops else []

    def finish(self, axis, works):
        """Wait_and_Unload_MPI_Comm_Buffers(axis)."""
        for w in works:
            w.wait()
        for face in (2 * axis, 2 * axis + 1):
            if self.flags[face] != BC_MPI:
                continue
            r = halo_slab(self.nx, self.ny, self.nz, self.g, face, recv=True)
            self.unpack(face, r, self._buf(("r", face), r.ncells * self.nf))

    def exchange(self, fill_physical_face):
        """The whole x -> y -> z schedule; `fill_physical_face(face)` is Grid3D::Set_Boundaries for one face."""
        for axis in range(self.ndim):
            works = self.post(axis)
            fill_physical_face(2 * axis)
            fill_physical_face(2 * axis + 1)
            self.finish(axis, works)
