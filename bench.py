#!/usr/bin/env python3
"""bench.py -- cell-updates/s of the Van Leer hydro hot path (3-D, PPMC + HLLC, fp64) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--size S] [--mode strict|fast]
                    [--recon pcm|plmc|ppmc|plmp|ppmp] [--solver hllc|roe|hll|exact] [--integrator vl|simple]
                    [--workload blast|kh|noise] [--precision 1|2]        (defaults = the headline: PPMC + HLLC, VL, fp64)

A "step" is one pass of the hot path over the local grid: the reference's main-loop body (main.cpp:247-379)
    set_dt (all-reduce MAX) -> VL step (+ temperature floor, CFL max) -> ghost fill / halo exchange.
Workload (BASELINE.json metric, config 5): spherical blast wave, S^3 cells PER GPU (default 512^3), periodic,
PPMC + HLLC, fp64; weak scaling over the reference's block decomposition (2x1x1, 2x2x1, 2x2x2).

One JSON line on rank 0.  `value` = cell-updates/s with the state resident in HBM (CUDA events around K steps,
barrier + synchronize either side, max over ranks).  `e2e` = the same through host buffers: pinned-host ->
device copy of the state, K steps, device -> pinned-host copy of the result, all inside the timed region.
`roofline` = the step's algorithmic bytes (200 B per cell-update) over the CUDA-event time of the six
direction sweeps that make up a step, against the measured HBM copy bandwidth in MEASURED_PEAKS.json; `traffic` =
the DRAM bytes of those launches from the committed ncu capture (profiles/roofline_traffic.json).  `other_mode` =
the same workload in the other arithmetic mode.  `cpu_baseline` = the CPU
oracle (a C restatement of the reference algorithm; the reference has no CPU hydro path) on a bounded sample.

--impl reference: the reference's OWN CUDA kernels, unmodified, compiled for sm_100 (oracle/_ref, built by
oracle/ref_build/Makefile), same workload, same loop, on rank 0's GPU.  The reference has no CPU hydro
implementation and needs MPI (absent from this image) for multi-GPU, so this arm is single-GPU by construction.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

B_ALG_STEP = 200.0       # algorithmic bytes per cell-update, fp64 hydro: 5 * NF * sizeof(Real)
GAMMA = 1.66666667


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.proc = index, [], threading.Event(), None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag.is_set():
                    break
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def finish(self):
        self.stop_flag.set()
        if self.proc is not None:
            try:
                self.proc.kill()
            except Exception:
                pass
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """DRAM bytes of one step's six sweeps per arithmetic mode, from the committed ncu capture (scratch/ncu_traffic.py)."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as fh:
            return json.load(fh)
    except Exception:
        return None


def cpu_baseline(sample_n=192, steps=1):
    """CPU oracle (oracle/vl_oracle.c, a port) on a bounded sample of the same workload; scalar, 1 thread."""
    import numpy as np

    import oracle
    from cholla_b200 import initial_conditions as ic
    n = (sample_n,) * 3
    g = 4
    f = ic.spherical_overpressure(n, gamma=GAMMA)
    a = np.zeros((5, n[2] + 2 * g, n[1] + 2 * g, n[0] + 2 * g))
    a[:, g:-g, g:-g, g:-g] = f
    flat = np.ascontiguousarray(a.reshape(-1))
    dims = (n[0] + 2 * g,) * 3
    oracle.set_boundaries(flat, *dims, 5, g, [1] * 6)
    t0 = time.perf_counter()
    oracle.advance(flat, *dims, g, 1.0 / sample_n, 1.0 / sample_n, 1.0 / sample_n, GAMMA, oracle.PPMC, oracle.HLLC,
                   [1] * 6, steps)
    dt = time.perf_counter() - t0
    return {"value": sample_n ** 3 * steps / dt, "unit": "cell-updates/s", "cores": 1, "kind": "port",
            "sample": "%d^3 blast, PPMC+HLLC fp64, %d step(s), oracle/vl_oracle.c single thread, %.1f s"
                      % (sample_n, steps, dt)}


def run_reference(args):
    """The reference's own kernels (oracle/_ref) on rank 0's GPU.  Other ranks exit without work."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    import numpy as np
    import torch

    from cholla_b200 import initial_conditions as ic
    from oracle import reference
    recon, solver = {"pcm": 0, "plmc": 1, "ppmc": 2, "plmp": 3, "ppmp": 4}[args.recon], {"hllc": 0, "roe": 1, "hll": 3, "exact": 4}[args.solver]
    out = {"impl": "reference", "metric": "cell-updates/s", "unit": "cell-updates/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic"}
    nosync = args.ref_variant == "nosync"
    simple = args.integrator == "simple"
    if not reference.available(recon, solver, nosync, simple=simple) or not torch.cuda.is_available():
        print(json.dumps({"impl": "reference", "unavailable":
                          "oracle/_ref/libcholla_ref_%s_%s%s.so not built or no GPU"
                          % (args.recon, args.solver, "_nosync" if nosync else "")}))
        return 0
    S, g = args.size, {"pcm": 2, "plmc": 3, "plmp": 3, "ppmc": 4, "ppmp": 4}[args.recon]
    n = (S, S, S)
    dims = (S + 2 * g,) * 3
    torch.cuda.set_device(0)
    host = np.zeros((5,) + dims[::-1])
    host[:, g:-g, g:-g, g:-g] = ic.spherical_overpressure(n, gamma=GAMMA)
    U = torch.from_numpy(host.reshape(-1)).cuda()
    del host
    d = 1.0 / S
    flags = [1] * 6
    # ghost fill + warm-up (also sizes the reference's global scratch: 55 doubles per cell)
    L = reference.lib(recon, solver, nosync, simple=simple)
    L.ref_set_gamma(GAMMA)
    import ctypes as C
    L.ref_set_boundaries(C.c_void_p(U.data_ptr()), *dims, 5, g, (C.c_int * 6)(*flags))
    reference.advance(U, dims, g, d, d, d, GAMMA, recon, solver, flags, max(args.warmup, 1), nosync=nosync, simple=simple)
    sampler = ClockSampler(0)
    sampler.start()
    torch.cuda.synchronize()
    _, _, ms = reference.advance(U, dims, g, d, d, d, GAMMA, recon, solver, flags, args.steps, nosync=nosync, timed=True, simple=simple)
    torch.cuda.synchronize()
    clocks = sampler.finish()
    value = S ** 3 * args.steps / (ms * 1e-3)
    peak, which = measured_hbm_peak()
    out.update({"value": value, "ms_per_step": ms / args.steps, "clocks": clocks,
                "config": {"workload": "blast %d^3 %s+%s fp64 periodic, %s integrator, reference kernels%s" %
                           (S, args.recon.upper(), args.solver.upper(), args.integrator.upper(), " (-DDISABLE_GPU_ERROR_CHECKING)" if nosync else " (as shipped: device sync after every stage)"),
                           "l2": "inputs larger than L2", "build": L.ref_build_flags().decode()},
                "cpu_baseline": {"value": value, "unit": "cell-updates/s", "cores": 1, "kind": "reference",
                                 "sample": "the reference has no CPU hydro path: this is its own unmodified CUDA "
                                           "build (oracle/_ref) on one B200, full workload, 1 host thread"},
                "e2e": {"value": value, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "roofline": {"bound": "hbm", "achieved": value * B_ALG_STEP / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": value * B_ALG_STEP / 1e9 / peak, "traffic": None,
                             "note": "whole step at the 200 B/cell algorithmic figure; peak " + which},
                "gpu_launches": 0})
    print(json.dumps(out))
    return 0


def run_b200(args, mode, world, rank, local, group, with_e2e):
    """One arm of the bench: K timed steps of the hot path in `mode` arithmetic.  Returns the result dict (all ranks)."""
    import torch
    import torch.distributed as dist

    import cholla_b200 as cb
    from cholla_b200 import initial_conditions as ic
    from cholla_b200.grid import Grid3D

    S = args.size
    np_grid = cb.tile_decomposition(world, S, S, S)
    n_global = [S * np_grid[0], S * np_grid[1], S * np_grid[2]]
    lengths = [float(np_grid[0]), float(np_grid[1]), float(np_grid[2])]  # unit cells, unit box per GPU
    strict = mode == "strict" and args.precision == 2
    recon = {"pcm": cb.PCM, "plmc": cb.PLMC, "ppmc": cb.PPMC, "plmp": cb.PLMP, "ppmp": cb.PPMP}[args.recon]
    solver = {"hllc": cb.HLLC, "roe": cb.ROE, "hll": cb.HLL, "exact": cb.EXACT}[args.solver]
    G = Grid3D(n_global, lengths, recon=recon, solver=solver, gamma=GAMMA, bc=[1] * 6, precision=args.precision,
               strict=strict, device="cuda:%d" % local, dist_group=group, rank=rank, nproc=world, fused=True,
               integrator=args.integrator)
    with ic.block(G.decomp.offset, G.decomp.n_local):
        if args.workload == "kh":      # BASELINE config 3: every cell moves, two shear layers
            fields = ic.kh_res_ind(n_global, lengths=lengths, lo=(0.0, -0.5 * lengths[1], -0.5 * lengths[2]), gamma=GAMMA)
        elif args.workload == "noise":  # worst case: cell-to-cell noise, every limiter branch taken everywhere
            fields = ic.random_smooth(G.decomp.n_local, seed=1 + rank, gamma=GAMMA)
        else:                           # BASELINE config 5 (the headline): blast wave into a medium at rest
            fields = ic.spherical_overpressure(n_global, lengths=lengths, gamma=GAMMA)
    G.Set_Initial_Conditions_Local(fields)
    del fields
    G.Set_Boundary_Conditions_Grid()
    cells_local = G.H.nx_real * G.H.ny_real * G.H.nz_real
    cells_total = cells_local * world
    state_bytes = G.C_device.numel() * G.C_device.element_size()

    def barrier():
        if world > 1:
            dist.barrier(group=group)
        torch.cuda.synchronize()

    def step_loop(k, dti):
        for _ in range(k):
            G.set_dt(dti)
            dti = G.Update_Hydro_Grid()
            G.Update_Time()
            G.Set_Boundary_Conditions_Grid()
        return dti

    # ---- warm-up ----
    dti = G.Calc_Inverse_Timestep()
    dti = step_loop(args.warmup, dti)
    # ---- timed region: K steps, state resident in HBM ----
    G.set_profiling(True)
    G.gpu_launches = 0
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    dti = step_loop(args.steps, dti)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.finish() if sampler else None
    launches = G.gpu_launches
    pred_ms, corr_ms, nprof = G.stage_times()
    G.set_profiling(False)
    t = torch.tensor([ms, corr_ms / max(nprof, 1), pred_ms / max(nprof, 1)], dtype=torch.float64, device=G.device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    ms, corr_ms_step, pred_ms_step = [float(v) for v in t.tolist()]
    value = cells_total * args.steps / (ms * 1e-3)

    # ---- e2e: host buffers, copies inside the timed region ----
    e2e = None
    if with_e2e:
        host_in = torch.empty(G.C_device.numel(), dtype=G.C_device.dtype, pin_memory=True)
        host_out = torch.empty_like(host_in, pin_memory=True)
        host_in.copy_(G.C_device)
        barrier()
        e0.record()
        G.C_device.copy_(host_in, non_blocking=True)
        d2 = G.Calc_Inverse_Timestep()
        d2 = step_loop(args.steps, d2)
        host_out.copy_(G.C_device, non_blocking=True)
        e1.record()
        barrier()
        t2 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=G.device)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX, group=group)
        ms2 = float(t2.item())
        e2e = {"value": cells_total * args.steps / (ms2 * 1e-3), "unit": "cell-updates/s",
               "h2d_bytes_per_step": state_bytes / args.steps, "d2h_bytes_per_step": state_bytes / args.steps + 8,
               "note": "pinned host -> HBM copy of the state, K steps (each returning its CFL scalar to the host), "
                       "HBM -> pinned host copy of the result, all timed; ms=%.2f" % ms2}
        del host_in, host_out
    G.close()
    del G
    torch.cuda.empty_cache()
    return dict(value=value, ms=ms, corr_ms_step=corr_ms_step, pred_ms_step=pred_ms_step, e2e=e2e, launches=launches,
                clocks=clocks, cells_local=cells_local, n_global=n_global, np_grid=np_grid, state_bytes=state_bytes)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=512, help="cells per axis PER GPU")
    ap.add_argument("--mode", default="strict", choices=["strict", "fast"],
                    help="strict = bit-reproduces the reference (headline); fast = product arithmetic (FMA, shared "
                         "reciprocals), reported beside it")
    ap.add_argument("--workload", default="blast", choices=["blast", "kh", "noise"],
                    help="blast = BASELINE config 5 (headline); kh = Kelvin-Helmholtz (config 3 initial conditions); "
                         "noise = random cell-to-cell state (worst case for the limiters)")
    ap.add_argument("--precision", type=int, default=2, choices=[1, 2])
    ap.add_argument("--recon", default="ppmc", choices=["pcm", "plmc", "ppmc", "plmp", "ppmp"], help="headline: ppmc")
    ap.add_argument("--solver", default="hllc", choices=["hllc", "roe", "hll", "exact"], help="headline: hllc")
    ap.add_argument("--integrator", default="vl", choices=["vl", "simple"],
                    help="headline: vl; simple = the reference's -DSIMPLE integrator (PCM / PLMC only)")
    ap.add_argument("--ref-variant", default="shipped", choices=["shipped", "nosync"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--single-mode", action="store_true", help="do not also time the other arithmetic mode")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    world, rank, local = env_int("WORLD_SIZE", 1), env_int("RANK", 0), env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # NCCL prints its version banner on stdout: keep the one JSON line clean
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        group = dist.group.WORLD
    S = args.size
    mode = args.mode if args.precision == 2 else "fast"
    r = run_b200(args, mode, world, rank, local, group, not args.no_e2e)
    other = None
    if not args.single_mode and args.precision == 2:
        o = run_b200(args, "fast" if mode == "strict" else "strict", world, rank, local, group, False)
        other = {"mode": "fast" if mode == "strict" else "strict", "value": o["value"], "unit": "cell-updates/s",
                 "ms_per_step": o["ms"] / args.steps, "predictor_ms_per_step": o["pred_ms_step"],
                 "corrector_ms_per_step": o["corr_ms_step"],
                 "note": "same workload and timing in the other arithmetic mode (strict = the reference's bits; fast = "
                         "FMA + shared reciprocals, within 1e-12 relative L1 on smooth data)"}
    if rank != 0:
        dist.destroy_process_group()
        return 0
    peak, which = measured_hbm_peak()
    sweeps_ms = r["pred_ms_step"] + r["corr_ms_step"]
    b_alg = B_ALG_STEP if args.precision == 2 else B_ALG_STEP / 2
    achieved = r["cells_local"] * b_alg / (sweeps_ms * 1e-3) / 1e9 if sweeps_ms > 0 else None
    tr = (ncu_traffic() or {}).get(mode if args.precision == 2 else "f32")
    out = {
        "metric": "cell-updates/s", "value": r["value"], "unit": "cell-updates/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r["ms"] / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64" if args.precision == 2 else "f32", "data": "synthetic",
        "config": {"workload": "%s %d^3 per GPU, %s+%s, periodic, %s integrator, %s arithmetic"
                               % (args.workload, S, args.recon.upper(), args.solver.upper(), args.integrator.upper(), mode),
                   "global_cells": r["n_global"], "decomposition": r["np_grid"],
                   "l2": "inputs larger than L2 (%.1f GB state per GPU)" % (r["state_bytes"] / 1e9),
                   "step": "set_dt + %s step (fused T-floor/CFL) + ghost fill/halo exchange" % args.integrator.upper()},
        "e2e": r["e2e"],
        "gpu_launches": r["launches"],
        "clocks": r["clocks"],
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak if achieved else None,
                     "traffic": tr.get("dram_bytes_per_step") if tr else None,
                     "kernel": "vl_dir_kernel<PPMC,HLLC>: the six direction sweeps of one step (predictor x,y,z + "
                               "corrector x,y,z), timed together by CUDA events on the launching stream",
                     "ms_per_launch": sweeps_ms, "launches_per_step": 6,
                     "algorithmic_bytes_per_launch": r["cells_local"] * b_alg,
                     "predictor_ms_per_step": r["pred_ms_step"], "corrector_ms_per_step": r["corr_ms_step"],
                     "step_frac_of_hbm_roofline": r["value"] / world * b_alg / 1e9 / peak,
                     "fp64_pipe_busy": tr.get("fp64_pipe_busy_pct") if tr else None,
                     "note": "achieved = %d B per cell-update (SURVEY 8d) x cells per step / time of the six sweeps; "
                             "traffic = DRAM bytes of the same six launches from ncu (profiles/); peak " % b_alg + which +
                             "; the sweeps are fp64-issue bound, not HBM bound, see DESIGN.md 3.1"},
    }
    if other:
        out["other_mode"] = other
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline()
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
