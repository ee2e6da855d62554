#!/usr/bin/env python
"""Benchmark of the hot path: GDoF/s of the hyperFS (finite-strain Neo-Hookean) Jacobian
MatMult at p = 4 on a synthetic box mesh (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # /gpu/b200 arm
    python bench.py --impl reference --gpus N ...            # CPU arm (restated /cpu/self)

One "step" = one MatShell MatMult of the fine-level Jacobian: ApplyJacobian_Ceed
(/root/reference/src/matops.c:98-112) = zero Dirichlet inputs, DMGlobalToLocal (+ halo),
zero Yloc, CeedOperatorApply (ONE fused kernel), DMLocalToGlobal (+ halo).
Workload at N = 1: BASELINE.json configs[2], box 64^3, degree 4 (262 144 elements,
50.9 M DoFs).  N > 1: weak scaling, one 64^3 brick per GPU (2x1x1, 2x2x1, 2x2x2 bricks),
NCCL halo exchange; `--scaling strong` shards the ~100 M-DoF 80^3 box of configs[3] instead.
Inputs (4.5 GB of per-quadrature-point data per GPU) are far larger than L2, so no explicit
L2 flush is needed between timed iterations.
"""
import argparse
import json
import os
import sys
import threading
import time

# the image exports NCCL_DEBUG=VERSION, which makes NCCL print a banner on STDOUT: keep stdout = one JSON line
if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
    os.environ["NCCL_DEBUG"] = "WARN"

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

def alg_bytes(problem, p, qextra=0):
    """8 Q^3 (10 + g) + 4 P^3 + 48 p^3 bytes per element (22 572 for hyperFS p = 4)."""
    P, Q = p + 1, p + 1 + qextra
    g = 0 if problem == "linElas" else 9
    return 8 * Q ** 3 * (10 + g) + 4 * P ** 3 + 48 * p ** 3


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def result(self):
        self.stop_flag = True
        self.join(timeout=1)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cpu_leg(problem, p, n_cpu, steps, warmup):
    """Times the CPU oracle (restated /cpu/self operator; the reference's own QFunctions when
    oracle/_ref is built) on a bounded sample of the workload, all host threads."""
    from helpers import OracleProblem
    from oracle import oracle
    pr = OracleProblem(problem, n_cpu, p, which=oracle.default_which())
    x = np.random.default_rng(1).standard_normal(pr.lsize)
    y = np.zeros_like(x)
    for _ in range(max(1, warmup)):
        pr_y = oracle.operator_apply(problem, True, (0.3, 1.0), pr.nelem, pr.P, pr.Q, pr.B, pr.D, pr.offsets,
                                     pr.qdata, pr.gradu, x, oracle.default_which(), y)
    t0 = time.perf_counter()
    for _ in range(steps):
        pr_y = oracle.operator_apply(problem, True, (0.3, 1.0), pr.nelem, pr.P, pr.Q, pr.B, pr.D, pr.offsets,
                                     pr.qdata, pr.gradu, x, oracle.default_which(), y)
    dt = (time.perf_counter() - t0) / steps
    del pr_y
    dofs = pr.lsize
    return {"value": dofs / dt / 1e9, "unit": "GDoF/s", "cores": oracle.num_threads(),
            "kind": "port",
            "sample": f"{problem} p={p} Jacobian apply on a {n_cpu}^3 box ({pr.nelem} elements, {dofs} DoFs), "
                      f"{steps} applies, restated /cpu/self operator (oracle/ceed_oracle.c, OpenMP over elements) calling "
                      f"{'the reference QFunctions compiled from /root/reference (oracle/_ref)' if oracle.default_which() == 'ref' else 'the ported QFunctions (oracle/qf_port.c)'}",
            "ms_per_apply": dt * 1e3}


def solve_bench(args, rank, world, local_rank, dist, config):
    """BASELINE configs[4]: full Newton-Krylov-p-MG solve, hyperFS, 10 load steps, Chebyshev/Jacobi smoothing
    with GPU diagonal assembly; wall time of the load-increment loop only, max over ranks
    (/root/reference/elasticity.c:632-676,755-764)."""
    import torch
    from ceedpetscsolid_b200 import ceed as libceed
    from ceedpetscsolid_b200.elasticity import AppCtx, Elasticity
    from ceedpetscsolid_b200.mesh import BoxMesh, grid_for
    grid = grid_for(world)
    n = (args.n * grid[0], args.n * grid[1], args.n * grid[2]) if args.scaling == "weak" else (args.n,) * 3
    lengths = tuple(float(g) for g in grid) if args.scaling == "weak" else (1.0, 1.0, 1.0)
    umesh = None
    if args.mesh:
        from ceedpetscsolid_b200.exodus import HexMesh, tube_mesh
        umesh = tube_mesh(*[int(v) for v in args.mesh[5:].split(",")]) if args.mesh.startswith("tube:") else HexMesh.from_file(args.mesh)
    app = AppCtx(problem=args.problem, degree=args.degree, n=n, num_steps=args.load_steps, perturb=0.05,
                 clamp={(2, 0): [0, 0, 0, 0, 0, 1, 0], (2, 1): [0, 0, -0.1 * lengths[2], 0, 0, 1, 0]})
    gmesh = BoxMesh(n=n, perturb=app.perturb, seed=0, lengths=lengths)
    if umesh is not None:
        ids = [int(v) for v in args.clamp_sets.split(",")]
        ext = float(np.ptp(umesh.vertices, axis=0).max())
        app.mesh, gmesh = umesh, None
        app.clamp = {ids[0]: [0, 0, 0, 0, 0, 1, 0], ids[1]: [0, -0.05 * ext, 0.1 * ext, 0, 0, 1, 0]}
    el = Elasticity(app, dist=dist if world > 1 else None, rank=rank, world=world, device_id=local_rank, gmesh=gmesh,
                    coarse_rtol=args.coarse_rtol, coarse=args.coarse, assemble=args.assemble, masked=args.dm == "masked",
                    halo=args.halo)
    el.pc.coarse_maxit = args.coarse_maxit
    libceed.launch_count_reset()
    sampler = ClockSampler(local_rank)
    sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    out = el.solve(log=(lambda m: print(m, file=sys.stderr)) if rank == 0 else None)
    t = torch.tensor([out["time_s"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    clocks = sampler.result()
    launches = int(libceed.launch_count())
    energy = el.strain_energy()   # elasticity.c:820-830: printed by the reference after the solve (all ranks: reduction)
    if rank == 0:
        config.update({"workload": f"{args.problem} degree {args.degree} Newton-Krylov-pMG solve, box {args.n}^3 per GPU, "
                                   f"{args.load_steps} load steps, levels {el.degrees}" + (f", mesh {args.mesh}" if args.mesh else ""),
                       "elements_per_gpu": el.mesh.nelem,
                       "bricks": "x".join(map(str, grid))})
        print(json.dumps({"metric": "SNES solve time", "value": float(t.item()), "unit": "s", "n_gpus": world, "steps": 1,
                          "warmup": 0, "ms_per_step": float(t.item()) * 1e3, "higher_is_better": False,
                          "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                          "snes_its": out["snes_its"], "ksp_its": out["ksp_its"], "converged": out["converged"],
                          "coarse_pcg_its": out["coarse_its"], "coarse_rtol": args.coarse_rtol, "coarse": args.coarse, "assemble": args.assemble, "dofs_unconstrained": out["dofs_global_unconstrained"],
                          "mdofs_per_sec_in_snes": out["mdofs_per_sec_in_snes"], "strain_energy": energy, "gpu_launches": launches,
                          "clocks": clocks}))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--problem", default="hyperFS")
    ap.add_argument("--degree", type=int, default=4)
    ap.add_argument("--box", "--n", dest="n", type=int, default=64, help="elements per direction per GPU (weak) / of the whole box (strong)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--cpu-box", "--n-cpu", dest="n_cpu", type=int, default=20, help="box size of the bounded CPU sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--solve", action="store_true",
                    help="time the full Newton-Krylov-p-MG solve (BASELINE configs[4]) instead of the MatMult")
    ap.add_argument("--load-steps", type=int, default=10)
    ap.add_argument("--mesh", default=None, help="--solve on an unstructured hex8 mesh: an Exodus II file (-mesh, "
                    "setupdm.c:40-68) or 'tube:NR,NT,NZ' (synthetic); side sets --clamp-sets are clamped")
    ap.add_argument("--clamp-sets", default="1,2", help="side-set ids: first fixed, second translated by (0,-0.05,0.1) x length scale")
    ap.add_argument("--coarse-rtol", type=float, default=1e-2)
    ap.add_argument("--coarse-maxit", type=int, default=500)
    ap.add_argument("--dm", default="masked", choices=["masked", "compressed"],
                    help="global-vector layout of the DM stand-in (matops.LevelDM)")
    ap.add_argument("--overlap", action="store_true",
                    help="N > 1: interface elements first, halo exchange on a side stream overlapped with the interior "
                         "elements (measured slower than the plain sequence at 8 GPUs without high-priority NCCL streams)")
    ap.add_argument("--halo", default="nccl", choices=["nccl", "p2p"],
                    help="N > 1: interface exchange through NCCL send/recv, or stored straight into the neighbours' "
                         "windows over NVLink peer memory (csrc/b200_halo.cu)")
    ap.add_argument("--assemble", default="coo", choices=["coo", "color"], help="p=1 matrix: CeedOperatorLinearAssemble element matrices, or 81 coloured applies (misc.c:151-183)")
    ap.add_argument("--coarse", default="hmg", choices=["hmg", "pcg"], help="coarse solve on the assembled p=1 level: h-multigrid (GAMG stand-in) or Jacobi-PCG")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    problem, p = args.problem, args.degree
    workload = (f"{problem} degree {p} Jacobian MatMult, box {args.n}^3 per GPU" if args.scaling == "weak"
                else f"{problem} degree {p} Jacobian MatMult, box {args.n}^3 sharded")
    config = {"workload": workload, "problem": problem, "degree": p, "levels": "fine level of {1,2,4}",
              "elements_per_gpu": None, "scatter": "fp64 atomics", "l2": "inputs larger than L2 (no flush needed)",
              "bricks": None, "halo": "none (1 GPU)" if world == 1 else "NCCL p2p, one sum-and-share exchange per MatMult" + (
                  ", overlapped with the interior elements" if (args.overlap and args.dm == "masked") else "")}

    # ------------------------------------------------------------------ CPU reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        config["elements_per_gpu"] = args.n ** 3
        config["bricks"] = "1x1x1" if world == 1 else "x".join(map(str, __import__("ceedpetscsolid_b200.mesh", fromlist=["grid_for"]).grid_for(world)))
        steps = max(1, min(args.steps, 20))
        cb = cpu_leg(problem, p, args.n_cpu, steps, min(args.warmup, 2))
        line = {"impl": "reference", "metric": "GDoF/s of hyperFS Jacobian MatMult at p=4", "value": cb["value"],
                "unit": "GDoF/s", "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 2),
                "ms_per_step": cb["ms_per_apply"], "higher_is_better": True, "scaling": args.scaling,
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config, "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "GDoF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ /gpu/b200 arm
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1 and args.overlap:
        os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")  # exchange kernels must not queue behind the interior CTAs
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if args.solve:
        return solve_bench(args, rank, world, local_rank, dist, config)
    from ceedpetscsolid_b200 import ceed as libceed
    from ceedpetscsolid_b200 import matops, setuplibceed
    from ceedpetscsolid_b200.mesh import BoxMesh, grid_for, smooth_displacement

    grid = grid_for(world)
    if args.scaling == "weak":
        gmesh = BoxMesh(n=(args.n * grid[0], args.n * grid[1], args.n * grid[2]), perturb=0.08, seed=0,
                        lengths=tuple(float(g) for g in grid))  # the domain grows with the mesh: cubic elements
    else:
        gmesh = BoxMesh(n=(args.n,) * 3, perturb=0.08, seed=0)
    mesh = gmesh.brick(grid, rank, interface_first=args.dm == "masked" and args.overlap) if world > 1 else gmesh
    config["elements_per_gpu"] = mesh.nelem
    config["bricks"] = "x".join(map(str, grid))

    ceed = libceed.Ceed(f"/gpu/b200:device_id={local_rank}")
    degrees, data, phys = setuplibceed.setup_all(ceed, mesh, problem, p)
    fine = len(degrees) - 1
    halo = None
    if world > 1:
        from ceedpetscsolid_b200.halo import Halo
        halo = Halo(gmesh, grid, rank, p, dist)
        if args.halo == "p2p":
            halo.enable_p2p()
            config["halo"] = "NVLink peer-memory windows (CUDA IPC), one sum-and-share per MatMult, no NCCL on the data path"
    # N > 1: shared-dof global vectors -> one symmetric sum-and-share halo exchange per MatMult
    masked = args.dm == "masked"
    config["dm"] = ("masked constrained dofs: Krylov vectors have the L-vector layout, no G2L/L2G copies" if masked
                    else "compressed global vectors: G2L gather + L2G scatter around every apply (PETSc DM style)")
    dm = matops.LevelDM(mesh, p, bc_faces="all", halo=halo, shared=True, masked=masked)
    user = matops.setup_jacobian_ctx(dm, ceed, data[fine], phys)

    # state: smooth admissible displacement -> residual fills gradu
    u = torch.from_numpy(smooth_displacement(mesh.node_coords(p)).reshape(-1)).cuda()
    uc, rc = ceed.Vector(u.numel()), ceed.Vector(u.numel())
    r = torch.zeros_like(u)
    uc.set_array(u); rc.set_array(r)
    data[fine].opApply.apply(uc, rc)
    uc.take_array(); rc.take_array()
    del u, r
    X, Y = dm.create_global_vector(), dm.create_global_vector()
    X.copy_(torch.from_numpy(np.random.default_rng(1 + rank).standard_normal(dm.nglobal)))
    dm.zero_constrained(X)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        matops.ApplyJacobian_Ceed(user, X, Y)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    libceed.launch_count_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        matops.ApplyJacobian_Ceed(user, X, Y)
    e1.record()
    barrier()
    launches = libceed.launch_count()
    clocks = sampler.result()
    if halo is not None:
        halo.check_p2p()   # a timed-out peer-memory exchange invalidates the run: fail loudly
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    global_nodes = gmesh.num_nodes(p)
    dofs = 3 * global_nodes
    dofs_unconstrained = 3 * int(np.prod([gmesh.n[d] * p - 1 for d in range(3)]))
    value = dofs * args.steps / (ms * 1e-3) / 1e9

    # ---- dominant kernel alone (fused Jacobian apply), CUDA events on the launching stream
    xl, yl = dm.create_local_vector(), dm.create_local_vector()
    xl.copy_(torch.from_numpy(np.random.default_rng(7).standard_normal(dm.lsize)))
    xc, yc = data[fine].xceed, data[fine].yceed
    xc.set_array(xl); yc.set_array(yl)
    for _ in range(3):
        data[fine].opJacob.apply_add(xc, yc)
    torch.cuda.synchronize()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nk = max(args.steps, 10)
    k0.record()
    for _ in range(nk):
        data[fine].opJacob.apply_add(xc, yc)
    k1.record()
    torch.cuda.synchronize()
    kms = k0.elapsed_time(k1) / nk
    xc.take_array(); yc.take_array()
    peak, peak_src = measured_peaks()
    ab = alg_bytes(problem, p) * mesh.nelem
    achieved = ab / (kms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "k_fused_apply<P=5,Q=5,hyperFS,Jacobian>", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "kernel_ms": kms, "algorithmic_bytes_per_launch": ab,
                "algorithmic_bytes_per_element": alg_bytes(problem, p),
                "kernel_gdofs": 3 * mesh.num_nodes(p) / (kms * 1e-3) / 1e9}
    # FP64 cross-check: DFMA-class instructions the kernel executes per element (12 line stages x
    # 3 Q^4 + Q^3 x ~105 for the cached hyperFS Jacobian) against a DFMA microbenchmark run now
    if problem == "hyperFS" and p == 4:
        try:
            import ctypes as C
            rate = C.c_double(0.0)
            libceed.b2(libceed.lib.b200_fp64_probe(C.byref(rate)))
            Qn = p + 1
            dfma_elem = 12 * 3 * Qn ** 4 + Qn ** 3 * 105
            roofline["fp64"] = {"dfma_per_element": dfma_elem, "kernel_tdfma_per_s": dfma_elem * mesh.nelem / (kms * 1e-3) / 1e12,
                                "probe_tdfma_per_s": rate.value / 1e12,
                                "frac": dfma_elem * mesh.nelem / (kms * 1e-3) / rate.value}
        except Exception as exc:  # the probe is informational: never fail the bench on it
            roofline["fp64"] = {"error": str(exc)}
    traffic_path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_path):
        with open(traffic_path) as f:
            tj = json.load(f)
        # one `ncu --set full` capture on a 32^3 box; traffic is per element, scaled to this launch
        roofline["traffic"] = tj.get("dram_bytes_per_element", 0) * mesh.nelem or None
        roofline["traffic_source"] = tj.get("source")

    # ---- e2e: the libCEED boundary with HOST buffers (-memtype host): SetArray(HOST), Apply, TakeArray(HOST)
    e2e = None
    if not args.no_e2e:
        xh = torch.empty(dm.lsize, dtype=torch.float64).pin_memory()
        yh = torch.zeros(dm.lsize, dtype=torch.float64).pin_memory()
        xh.copy_(xl.cpu())
        nrep = max(3, min(args.steps, 10))
        for it in range(2 + nrep):
            if it == 2:
                barrier()
                t0 = time.perf_counter()
            xc.set_array(xh, libceed.MEM_HOST)
            yc.set_array(yh, libceed.MEM_HOST)
            data[fine].opJacob.apply(xc, yc)
            xc.take_array(libceed.MEM_HOST)
            yc.take_array(libceed.MEM_HOST)
        barrier()
        dt = (time.perf_counter() - t0) / nrep
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": dofs / dt / 1e9, "unit": "GDoF/s", "h2d_bytes_per_step": int(dm.lsize * 8),
               "d2h_bytes_per_step": int(dm.lsize * 8), "ms_per_step": dt * 1e3,
               "what": "CeedVectorSetArray(HOST,USE_POINTER) + CeedOperatorApply + CeedVectorTakeArray(HOST) on pinned host L-vectors"}

    if rank == 0:
        cb = None
        if world == 1 and not args.no_cpu:
            cb = cpu_leg(problem, p, args.n_cpu, 5, 1)
        line = {"metric": "GDoF/s of hyperFS Jacobian MatMult at p=4", "value": value, "unit": "GDoF/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": config, "dofs": dofs, "dofs_unconstrained": dofs_unconstrained,
                "value_unconstrained_dofs": dofs_unconstrained * args.steps / (ms * 1e-3) / 1e9,
                "roofline": roofline, "cpu_baseline": cb, "e2e": e2e, "gpu_launches": int(launches),
                "clocks": clocks}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
