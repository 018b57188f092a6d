"""Host-side mirror of the reference's `Grid3D` for the Van Leer hot path.

Same method names, argument meaning and call order as the reference's host code, so that a driver written
against `Grid3D` reads like `src/main.cpp:247-379`:

    dti = G.Calc_Inverse_Timestep()
    while G.H.t < tout:
        G.set_dt(dti)                       # grid3D.cpp:373-409 (all-reduce MAX across ranks)
        dti = G.Update_Hydro_Grid()         # grid3D.cpp:474-585 (VL step, temperature floor, Calc_dt)
        G.Update_Time()
        G.Set_Boundary_Conditions_Grid()    # boundary_conditions.cpp:19- / mpi_boundaries.cpp:29-111

All arithmetic happens in libcholla_b200.so (CUDA).  PyTorch supplies device memory, the stream and
torch.distributed (NCCL on GPUs, gloo for CPU tests of the exchange schedule); numpy is only used to stage
initial conditions.  Multi-GPU = the reference's block decomposition (src/mpi/mpi_routines.cpp:239-532),
one process per GPU.
"""
import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import _lib
from ._lib import BC_MPI, BC_PERIODIC, Config, check

N_FIELDS = 5
C_CFL = 0.3  # grid3D.cpp:157


def n_ghost_for(recon, mhd=False):
    """Grid3D constructor, grid3D.cpp:52-76: PCM 2, PLMP 3, PLMC 3, PPMC 4, one more in an MHD build."""
    return {_lib.PCM: 2, _lib.PLMP: 3, _lib.PLMC: 3, _lib.PPMP: 4, _lib.PPMC: 4}[recon] + (1 if mhd else 0)


@dataclass
class Decomposition:
    """DomainDecompositionBLOCK (src/mpi/mpi_routines.cpp:270-532) for one rank."""
    nproc: int
    rank: int
    np: List[int]              # process grid (x, y, z)
    coords: List[int]          # this rank's (ix, iy, iz); ranks are laid out x-fastest
    n_local: List[int]         # real cells per axis on this rank
    offset: List[int]          # global index of the first real cell per axis
    dest: List[int]            # neighbour ranks for faces 0..5 (periodic wrap, as in the reference)
    flags: List[int]           # local boundary flags for faces 0..5 (5 = exchange with a neighbour)
    global_flags: List[int]

    @staticmethod
    def build(nproc, rank, n_global, global_flags):
        nxg, nyg, nzg = n_global
        np_ = _lib.tile_decomposition(nproc, nxg, nyg, nzg)
        ix = rank % np_[0]
        iy = (rank // np_[0]) % np_[1]
        iz = rank // (np_[0] * np_[1])
        coords = [ix, iy, iz]
        n_local, offset = [], []
        for a in range(3):
            n, o = _lib.local_extent(n_global[a], np_[a], coords[a])
            n_local.append(n)
            offset.append(o)

        def rank_of(c):
            return c[0] + np_[0] * (c[1] + np_[1] * c[2])

        dest = []
        for a in range(3):
            for s in (-1, +1):
                c = list(coords)
                c[a] = (c[a] + s) % np_[a]
                dest.append(rank_of(c))
        flags = list(global_flags)
        for a in range(3):  # mpi_routines.cpp:453-532
            if np_[a] == 1:
                continue
            lo, hi = 2 * a, 2 * a + 1
            if coords[a] == 0:
                flags[hi] = BC_MPI
                if flags[lo] == BC_PERIODIC:
                    flags[lo] = BC_MPI
            elif coords[a] == np_[a] - 1:
                flags[lo] = BC_MPI
                if flags[hi] == BC_PERIODIC:
                    flags[hi] = BC_MPI
            else:
                flags[lo] = flags[hi] = BC_MPI
        return Decomposition(nproc, rank, np_, coords, n_local, offset, dest, flags, list(global_flags))


from .halo import HaloExchange, Slab, halo_slab  # noqa: E402,F401  (re-exported: tests import them from here)


@dataclass
class Header:
    """The subset of Grid3D::Header (grid3D.h:112-279) the hot path reads."""
    nx: int = 1
    ny: int = 1
    nz: int = 1
    nx_real: int = 1
    ny_real: int = 1
    nz_real: int = 1
    n_ghost: int = 0
    n_fields: int = N_FIELDS
    n_cells: int = 1
    dx: float = 1.0
    dy: float = 1.0
    dz: float = 1.0
    xbound: float = 0.0
    ybound: float = 0.0
    zbound: float = 0.0
    t: float = 0.0
    dt: float = 0.0
    n_step: int = 0
    max_dti: float = 0.0


class Grid3D:
    """One rank's grid: conserved state in HBM + the chl context that advances it."""

    def __init__(self, n_global, domain_len, bounds_min=(0.0, 0.0, 0.0), recon=_lib.PPMC, solver=_lib.HLLC,
                 gamma=5.0 / 3.0, bc=(1, 1, 1, 1, 1, 1), precision=2, strict=True, density_floor=0.0,
                 temperature_floor_U=0.0, device=None, dist_group=None, rank=0, nproc=1, fused=True, mhd=False,
                 integrator="vl", halo_transport="p2p", overlap=True, n_scalars=0, dual_energy=False, cooling=None):
        import torch
        self.torch = torch
        if integrator not in ("vl", "simple"):  # -DVL / -DSIMPLE, builds/make.type.hydro:8-12
            raise ValueError("integrator must be 'vl' or 'simple'")
        self.integrator = integrator
        self.recon, self.solver, self.gamma = recon, solver, float(gamma)
        self.precision, self.strict, self.fused = precision, bool(strict), bool(fused)
        self.mhd = bool(mhd) or solver == _lib.HLLD
        # -DSCALAR / -DDE builds: fields 5 .. are the scalars, GasEnergy is the last field (grid_enum.h:33-78)
        self.n_scalars, self.dual_energy = int(n_scalars), bool(dual_energy)
        if cooling not in (None, "cie"):  # -DCOOLING_GPU (CIE curve); the Cloudy tables of -DCLOUDY_COOL are not built
            raise ValueError("cooling must be None or 'cie'")
        self.cooling = cooling
        if self.mhd:
            self.solver = solver = _lib.HLLD  # error_handling.cpp:93-96: MHD only supports HLLD
        self.dist_group, self.nproc, self.rank = dist_group, nproc, rank
        n_global = [int(v) for v in n_global]
        self.n_global = n_global
        self.ndim = 1 + (n_global[1] > 1) + (n_global[2] > 1)
        self.decomp = Decomposition.build(nproc, rank, n_global, list(bc))
        g = n_ghost_for(recon, self.mhd)
        H = Header()
        H.n_ghost = g
        H.n_fields = 8 if self.mhd else N_FIELDS + self.n_scalars + int(self.dual_energy)
        H.nx_real, H.ny_real, H.nz_real = self.decomp.n_local
        H.nx = H.nx_real + 2 * g
        H.ny = H.ny_real + 2 * g if n_global[1] > 1 else 1
        H.nz = H.nz_real + 2 * g if n_global[2] > 1 else 1
        H.n_cells = H.nx * H.ny * H.nz
        H.dx = domain_len[0] / n_global[0]
        H.dy = domain_len[1] / n_global[1]
        H.dz = domain_len[2] / n_global[2]
        H.xbound, H.ybound, H.zbound = bounds_min
        self.H = H
        self.flags = self.decomp.flags
        self.U_floor = float(temperature_floor_U)
        self.dtype = torch.float64 if precision == 2 else torch.float32
        self.device = torch.device(device if device is not None else "cuda:0")
        if self.device.type != "cuda":
            raise _lib.ChollaB200Error("Grid3D needs a CUDA device: there is no CPU path")
        torch.cuda.set_device(self.device)
        self.C_device = torch.zeros(H.n_fields * H.n_cells, dtype=self.dtype, device=self.device)
        cfg = Config(H.nx, H.ny, H.nz, g, H.n_fields, recon, solver, precision, int(self.strict), int(self.mhd), self.gamma,
                     H.dx, H.dy, H.dz, float(density_floor), self.U_floor, self.n_scalars, int(self.dual_energy))
        self._cfg = cfg
        self._ctx = C.c_void_p()
        check(_lib.lib().chl_create(C.byref(cfg), C.byref(self._ctx)), "chl_create")
        self.gpu_launches = 0
        self._halo_x = None
        # multi-rank: "p2p" = the library's own exchange (peer-mapped stores over NVLink, chl_halo_*), optionally
        # overlapped with the interior of the last corrector pass; "nccl" = torch.distributed send / recv (halo.py)
        if halo_transport not in ("p2p", "nccl"):
            raise ValueError("halo_transport must be 'p2p' or 'nccl'")
        self.halo_transport = halo_transport if nproc > 1 else "none"
        self.overlap = (bool(overlap) and self.halo_transport == "p2p" and self.integrator == "vl" and self.fused
                        and not self.cooling)
        self._halo_h = None          # chl_halo*
        self._exchanged = False      # the step just taken already exchanged the ghost cells
        self._dti_is_global = False  # the dti just returned is already the maximum over ranks
        self._key_t = None
        if self.halo_transport == "p2p":
            self._halo_p2p_setup()

    # ---- plumbing ----
    def close(self):
        if getattr(self, "_halo_h", None) is not None and self._halo_h.value:
            _lib.lib().chl_halo_destroy(self._halo_h)
            self._halo_h = None
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            _lib.lib().chl_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def _ptr(self, t=None):
        return C.c_void_p((self.C_device if t is None else t).data_ptr())

    def _cflags(self):
        return (C.c_int * 6)(*self.flags)

    def set_tuning(self, dir_chunk=0, fused_predictor=False):
        """chl_set_tuning: marching length per thread (0 = automatic) and the single-launch predictor."""
        check(_lib.lib().chl_set_tuning(self._ctx, int(dir_chunk), int(bool(fused_predictor))), "chl_set_tuning")

    def set_profiling(self, on):
        check(_lib.lib().chl_set_profiling(self._ctx, int(bool(on))), "chl_set_profiling")

    def stage_times(self):
        """(predictor_ms, corrector_ms, n_steps) summed over the steps since the last call (CUDA events)."""
        ms, n = (C.c_double * 2)(), C.c_int()
        check(_lib.lib().chl_stage_times(self._ctx, ms, C.byref(n)), "chl_stage_times")
        return ms[0], ms[1], n.value

    # ---- state in / out ----
    def local_slices(self):
        """numpy slices of this rank's real cells inside a global (nz, ny, nx) array."""
        o, n = self.decomp.offset, self.decomp.n_local
        return (slice(o[2], o[2] + n[2]), slice(o[1], o[1] + n[1]), slice(o[0], o[0] + n[0]))

    def Set_Initial_Conditions(self, fields_global):
        """fields_global: float64 array (n_fields, nz_g, ny_g, nx_g) of conserved real cells (host).  The local
        block is copied into the real cells (initial_conditions.cpp:102-104 does the same H2D of C.host)."""
        H, g, torch = self.H, self.H.n_ghost, self.torch
        fields_global = np.asarray(fields_global)
        host = np.zeros((H.n_fields, H.nz, H.ny, H.nx), dtype=np.float64)
        sz, sy, sx = self.local_slices()
        gz = g if H.nz > 1 else 0
        gy = g if H.ny > 1 else 0
        host[:, gz:gz + H.nz_real, gy:gy + H.ny_real, g:g + H.nx_real] = fields_global[:, sz, sy, sx]
        self.C_device.copy_(torch.from_numpy(host.reshape(-1)).to(self.dtype))
        self.H.t, self.H.n_step = 0.0, 0

    def Set_Initial_Conditions_Local(self, fields_local):
        """Same, from this rank's block only: (n_fields, nz_real, ny_real, nx_real)."""
        H, g, torch = self.H, self.H.n_ghost, self.torch
        host = np.zeros((H.n_fields, H.nz, H.ny, H.nx), dtype=np.float64)
        gz = g if H.nz > 1 else 0
        gy = g if H.ny > 1 else 0
        host[:, gz:gz + H.nz_real, gy:gy + H.ny_real, g:g + H.nx_real] = fields_local
        self.C_device.copy_(torch.from_numpy(host.reshape(-1)).to(self.dtype))
        self.H.t, self.H.n_step = 0.0, 0

    def upload(self, flat_host):
        """Copy a full local array (ghosts included, flat SoA) to the device."""
        self.C_device.copy_(self.torch.from_numpy(np.ascontiguousarray(flat_host).reshape(-1)).to(self.dtype))

    def download(self):
        return self.C_device.detach().to("cpu", self.torch.float64).numpy().copy()

    def real_cells(self):
        H, g = self.H, self.H.n_ghost
        a = self.download().reshape(H.n_fields, H.nz, H.ny, H.nx)
        gz = g if H.nz > 1 else 0
        gy = g if H.ny > 1 else 0
        return a[:, gz:gz + H.nz_real, gy:gy + H.ny_real, g:g + H.nx_real].copy()

    def snapshot_field(self, field, mhd_direction=-1, out_precision=None, extra=0):
        """Write_Grid_HDF5_Field_GPU / Write_HDF5_Field_3D (io.cpp:1318-1324, io_gpu.cu:97-140): the real cells of one
        field transposed on the device into the snapshot's [i][j][k] order, then copied to the host.  Face-centred
        magnetic fields (mhd_direction 0/1/2) carry one extra face along their own axis."""
        H, torch = self.H, self.torch
        out_precision = out_precision or self.precision
        ext = [H.nx_real, H.ny_real, H.nz_real]
        if mhd_direction >= 0:
            ext[mhd_direction] += 1
        if extra:  # Output_Float32's magnetic datasets (io.cpp:347-362): one more cell on the low side of every axis
            ext = [e + extra for e in ext]
        ndim = 1 + (H.ny > 1) + (H.nz > 1)
        buf = torch.empty(ext[0] * ext[1] * ext[2], dtype=torch.float64 if out_precision == 2 else torch.float32,
                          device=self.device)
        esz = 8 if self.precision == 2 else 4
        src = C.c_void_p(self.C_device.data_ptr() + field * H.n_cells * esz)
        check(_lib.lib().chl_pack_snapshot_field(self._ptr(buf), src, H.nx, H.ny, H.nz, H.n_ghost - extra, ext[0], ext[1],
                                                 ext[2], mhd_direction, self.precision, out_precision, self._stream()),
              "chl_pack_snapshot_field")
        self.gpu_launches += 1
        return buf.cpu().numpy().reshape(ext[:ndim])

    # ---- the hot path, reference names ----
    def Calc_Inverse_Timestep(self):
        """Grid3D::Calc_Inverse_Timestep (grid3D.cpp:115-120) -> Calc_dt_GPU."""
        out = C.c_double()
        check(_lib.lib().chl_calc_dti(self._ctx, self._ptr(), C.byref(out), self._stream()), "chl_calc_dti")
        self.gpu_launches += 1
        return out.value

    def set_dt(self, dti):
        """Grid3D::set_dt (grid3D.cpp:373-409): global max of the inverse time step, dt = C_cfl / max_dti."""
        max_dti = dti
        if self.nproc > 1 and self._dti_is_global:
            self._dti_is_global = False  # Update_Hydro_Grid reduced the CFL keys across ranks on the device
        elif self.nproc > 1:
            dist, torch = self.torch.distributed, self.torch
            dev = self.device if dist.get_backend(self.dist_group) == "nccl" else "cpu"
            t = torch.tensor([dti], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.dist_group)
            max_dti = float(t.item())
        self.H.max_dti = max_dti
        self.H.dt = C_CFL / max_dti
        return self.H.dt

    def Execute_Hydro_Integrator(self):
        """Grid3D::Execute_Hydro_Integrator (grid3D.cpp:413-470) -> VL_Algorithm_{1,2,3}D_CUDA, or
        Simple_Algorithm_{1,2,3}D_CUDA in a -DSIMPLE build."""
        if self.integrator == "simple":
            check(_lib.lib().chl_simple_step(self._ctx, self._ptr(), self.H.dt, self._stream()), "chl_simple_step")
        else:
            check(_lib.lib().chl_vl_step(self._ctx, self._ptr(), self.H.dt, self._stream()), "chl_vl_step")
        self.gpu_launches += self._sweep_launches()

    def _sweep_launches(self):
        """Kernels one step launches: VL a predictor and a corrector sweep per direction (direction-split default);
        SIMPLE the primitive pass and one sweep per direction; a dual-energy context adds the two p div v passes and
        Select_Internal_Energy / Sync_Energies."""
        return 1 + self.ndim if self.integrator == "simple" else 2 * self.ndim + (4 if self.dual_energy else 0)

    def Set_Static_Gravity(self, custom_grav):
        """A -DSTATIC_GRAV build with Parameters::custom_grav = custom_grav (static_grav.h; < 0 switches gravity off): the
        analytic acceleration is evaluated once on the device from this rank's offsets and the domain corner, every later
        step adds the time-centred source terms of Update_Conserved_Variables_{1,2,3}D."""
        o, H = self.decomp.offset, self.H
        check(_lib.lib().chl_set_static_gravity(self._ctx, int(custom_grav), o[0], o[1], o[2], H.xbound, H.ybound, H.zbound,
                                                self._stream()), "chl_set_static_gravity")
        self.static_gravity = custom_grav >= 0
        self.gpu_launches += 1
        if self.static_gravity:
            self.overlap = False  # the source terms follow the whole corrector: the exchange cannot start before them

    def Set_Gravity_Potential(self, potential, five_point=True):
        """A -DGRAVITY build: `potential` is this rank's potential with ghost cells (numpy (nz, ny, nx) or a device tensor),
        as the reference's Poisson solver leaves it in d_grav_potential; None switches the source terms off.  The array is
        kept on the device and read by every later step (chl_set_gravity_potential)."""
        if potential is None:
            self._potential = None
            check(_lib.lib().chl_set_gravity_potential(self._ctx, None, 0), "chl_set_gravity_potential")
            return
        torch = self.torch
        if not isinstance(potential, torch.Tensor):
            potential = torch.from_numpy(np.ascontiguousarray(potential, dtype=np.float64).reshape(-1))
        self._potential = potential.to(self.device, self.dtype).contiguous()
        assert self._potential.numel() == self.H.n_cells
        check(_lib.lib().chl_set_gravity_potential(self._ctx, self._ptr(self._potential), int(bool(five_point))),
              "chl_set_gravity_potential")
        self.overlap = False

    def gravity_field(self):
        """The device's [ndim][n_cells] acceleration arrays as a host array (None when gravity is off)."""
        p = C.c_void_p()
        check(_lib.lib().chl_gravity_field_ptr(self._ctx, C.byref(p)), "chl_gravity_field_ptr")
        if not p.value:
            return None
        out = self.torch.empty(self.ndim * self.H.n_cells, dtype=self.dtype, device=self.device)
        check(_lib.lib().chl_copy_gravity_field(self._ctx, self._ptr(out), self._stream()), "chl_copy_gravity_field")
        return out.cpu().numpy().astype(np.float64)

    def Apply_Scalar_Floor(self, field_num, scalar_floor):
        """Apply_Scalar_Floor (hydro_cuda.cu:1267-1307; grid3D.cpp:505-510 calls it for the dust density)."""
        check(_lib.lib().chl_apply_scalar_floor(self._ctx, self._ptr(), int(field_num), float(scalar_floor), self._stream()),
              "chl_apply_scalar_floor")
        self.gpu_launches += 1

    def Update_Hydro_Grid(self):
        """Grid3D::Update_Hydro_Grid (grid3D.cpp:474-585): integrator, temperature floor, new inverse dt."""
        L = _lib.lib()
        if self.fused and not self.cooling:
            if self.integrator == "simple":
                check(L.chl_simple_step_fused(self._ctx, self._ptr(), self.H.dt, self._stream()), "chl_simple_step_fused")
            elif self.overlap:
                # step + ghost-cell exchange in one call: boundary shell first, exchange overlapped with the interior
                check(L.chl_vl_step_exchange(self._ctx, self._halo_h, self._ptr(), self.H.dt, self._stream()),
                      "chl_vl_step_exchange")
                self._exchanged = True
                self.gpu_launches += self._exchange_launches() + (1 if self.ndim == 3 and not self.mhd else 0)
            else:
                check(L.chl_vl_step_fused(self._ctx, self._ptr(), self.H.dt, self._stream()), "chl_vl_step_fused")
            self.gpu_launches += self._sweep_launches()
            if self.nproc > 1 and self.torch.distributed.get_backend(self.dist_group) == "nccl":
                # MAX over ranks of the ordered CFL keys, on the device; one host read per step.  (The key of a
                # non-negative value has its top bit set, so as int64 the keys keep their order.)
                torch, dist = self.torch, self.torch.distributed
                if self._key_t is None:
                    self._key_t = torch.zeros(1, dtype=torch.int64, device=self.device)
                check(L.chl_dti_copy_key(self._ctx, self._ptr(self._key_t), self._stream()), "chl_dti_copy_key")
                dist.all_reduce(self._key_t, op=dist.ReduceOp.MAX, group=self.dist_group)
                self._dti_is_global = True
                return L.chl_dti_decode(C.c_ulonglong(int(self._key_t.item()) & 0xFFFFFFFFFFFFFFFF))
            out = C.c_double()
            check(L.chl_dti_result(self._ctx, C.byref(out), self._stream()), "chl_dti_result")
            return out.value
        # the reference's own order (grid3D.cpp:491-575): integrator, temperature floor, cooling, Calc_dt
        self.Execute_Hydro_Integrator()
        if self.U_floor >= 0.0:
            check(L.chl_apply_temperature_floor(self._ctx, self._ptr(), self.U_floor, self._stream()),
                  "chl_apply_temperature_floor")
            self.gpu_launches += 1
        if self.cooling:
            check(L.chl_cooling_update(self._ctx, self._ptr(), self.H.dt, self._stream()), "chl_cooling_update")
            self.gpu_launches += 1
        return self.Calc_Inverse_Timestep()

    def bad_cells(self):
        """Cells whose density came out <= 0 (or NaN) in the corrector since the last call (chl_bad_cells): the host-side
        stand-in for the reference's device printf + neighbour averaging (hydro_cuda.cu:292-297) -- abort or repair here."""
        n = C.c_ulonglong()
        check(_lib.lib().chl_bad_cells(self._ctx, C.byref(n), self._stream()), "chl_bad_cells")
        return int(n.value)

    def max_magnetic_divergence(self):
        """mhd::checkMagneticDivergence (magnetic_divergence.cu:79-112): max |div B| reduced on the device and, across
        ranks, with a MAX all-reduce (ReduceRealMax).  The caller applies the reference's 1e-9 limit."""
        out = C.c_double()
        check(_lib.lib().chl_mhd_max_divergence(self._ctx, self._ptr(), C.byref(out), self._stream()),
              "chl_mhd_max_divergence")
        self.gpu_launches += 1
        v = out.value
        if self.nproc > 1:
            dist, torch = self.torch.distributed, self.torch
            dev = self.device if dist.get_backend(self.dist_group) == "nccl" else "cpu"
            t = torch.tensor([v], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.dist_group)
            v = float(t.item())
        return v

    def Update_Time(self):
        self.H.t += self.H.dt
        self.H.n_step += 1

    def Set_Boundaries(self, direction):
        """Grid3D::Set_Boundaries(dir, flags) (boundary_conditions.cpp:165-417) -> SetGhostCells."""
        if direction // 2 >= self.ndim or self.flags[direction] in (_lib.BC_CUSTOM, _lib.BC_MPI):
            return
        check(_lib.lib().chl_set_ghost_cells(self._ctx, self._ptr(), self._cflags(), direction, self._stream()),
              "chl_set_ghost_cells")
        self.gpu_launches += 1

    def Set_Boundary_Conditions_Grid(self):
        """Grid3D::Set_Boundary_Conditions (boundary_conditions.cpp:54-): the six physical faces in order, or,
        when any face is an exchange face, the x -> y -> z phased exchange of Set_Boundaries_MPI_BLOCK
        (mpi_boundaries.cpp:29-111)."""
        if self.nproc == 1:
            for d in range(2 * self.ndim):
                self.Set_Boundaries(d)
            return
        if self._exchanged:  # chl_vl_step_exchange did it, overlapped with the interior update
            self._exchanged = False
            return
        if self.halo_transport == "p2p":
            check(_lib.lib().chl_halo_exchange(self._halo_h, self._ptr(), self._stream()), "chl_halo_exchange")
            self.gpu_launches += self._exchange_launches()
            return
        self._halo().exchange(self.Set_Boundaries)

    # ---- the library's own exchange: set-up (the only host-side communication it needs) ----
    def _halo_p2p_setup(self):
        L, dist = _lib.lib(), self.torch.distributed
        h = C.c_void_p()
        check(L.chl_halo_create(self._ctx, self._cflags(), self.rank, C.byref(h)), "chl_halo_create")
        self._halo_h = h
        blob = C.create_string_buffer(L.chl_halo_blob_bytes())
        check(L.chl_halo_export(h, blob), "chl_halo_export")
        blobs = [None] * self.nproc
        dist.all_gather_object(blobs, bytes(blob.raw), group=self.dist_group)  # MPI_Allgather in a Cholla build
        for face in range(2 * self.ndim):
            if self.flags[face] == BC_MPI:
                check(L.chl_halo_connect(h, face, blobs[self.decomp.dest[face]]), "chl_halo_connect")
        dist.barrier(group=self.dist_group)  # every rank has mapped its neighbours before anyone sends

    def _exchange_launches(self):
        n_mpi = sum(1 for f in range(2 * self.ndim) if self.flags[f] == BC_MPI)
        return 2 * n_mpi + (2 * self.ndim - n_mpi)

    def halo_kind(self):
        return self.halo_transport

    def halo_check(self):
        """Raises if a receive gave up waiting for its neighbour (chl_halo_error); synchronises the stream."""
        if self._halo_h is not None:
            check(_lib.lib().chl_halo_error(self._halo_h, self._stream()), "chl_halo_error")

    # ---- halo exchange over torch.distributed (NCCL send/recv over NVLink); schedule in halo.py ----
    def _halo(self):
        if self._halo_x is None:
            H, L, torch = self.H, _lib.lib(), self.torch

            def pack(face, s, buf):
                check(L.chl_pack_buffer(self._ptr(buf), self._ptr(), H.nx, H.ny, H.n_fields, H.n_cells, s.idxoffset,
                                        s.isize, s.jsize, s.ksize, self.precision, self._stream()), "chl_pack_buffer")
                self.gpu_launches += 1

            def unpack(face, s, buf):
                check(L.chl_unpack_buffer(self._ptr(buf), self._ptr(), H.nx, H.ny, H.n_fields, H.n_cells, s.idxoffset,
                                          s.isize, s.jsize, s.ksize, self.precision, self._stream()),
                      "chl_unpack_buffer")
                self.gpu_launches += 1

            self._halo_x = HaloExchange(torch.distributed, self.dist_group, (H.nx, H.ny, H.nz), H.n_ghost, H.n_fields,
                                        self.flags, self.decomp.dest,
                                        lambda n: torch.empty(n, dtype=self.dtype, device=self.device), pack, unpack)
        self._halo_x.flags = list(self.flags)
        return self._halo_x

    # ---- convenience driver: the body of main()'s loop ----
    def advance(self, n_steps):
        """n_steps iterations of main.cpp:247-379 (no output).  Returns the dt history."""
        dts = []
        dti = self.Calc_Inverse_Timestep()
        for _ in range(n_steps):
            self.set_dt(dti)
            dti = self.Update_Hydro_Grid()
            self.Update_Time()
            self.Set_Boundary_Conditions_Grid()
            dts.append(self.H.dt)
        self._last_dti = dti
        self.halo_check()
        return dts
