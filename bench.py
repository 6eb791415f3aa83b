#!/usr/bin/env python
"""Benchmark of the fractional-step time loop (contract: see the task statement / DESIGN.md section 7).

  python bench.py --gpus N --steps K --warmup W            # our CUDA path, one JSON line on rank 0
  python bench.py --impl reference --steps K --warmup W    # the CPU oracle (reference restatement), host cores

A "step" is one pass of the time-loop body (BCH.cpp:2859-3118 / GQ.cpp:3308-3416) over the whole mesh.
metric = time steps/s.  `value`: state resident in HBM, CUDA-event time on the launch stream.
`e2e`: same steps driven through cfd_timeStep_host with pinned HOST buffers (H2D + step + D2H per step).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "cfd-with-cuda_b200"))

WORKLOADS = {
    # BASELINE.json configs[4] (the config the metric is quoted on), single-GPU form
    "cavity256_q1_implicit": dict(n=256, NENv=8, scheme="implicit", cluster=0.0, viscosity=0.01, dt=1e-3),
    # configs[3]
    "cavity128_q1_implicit": dict(n=128, NENv=8, scheme="implicit", cluster=0.0, viscosity=0.01, dt=1e-3),
    "cavity128_q1_explicit": dict(n=128, NENv=8, scheme="explicit", cluster=0.0, viscosity=0.01, dt=1e-3),
    "cavity64_q2_explicit": dict(n=64, NENv=27, scheme="explicit", cluster=0.0, viscosity=0.01, dt=1e-3),
    # (explicit viscous term: dt below h^2 / (6 nu) with h = the Q2 node spacing -- 1e-3 blows up on these two meshes)
    "cavity96_q2_explicit": dict(n=96, NENv=27, scheme="explicit", cluster=0.0, viscosity=0.01, dt=2.5e-4),
    "cavity128_q2_explicit": dict(n=128, NENv=27, scheme="explicit", cluster=0.0, viscosity=0.01, dt=2e-4),
    "cavity128_q2_implicit": dict(n=128, NENv=27, scheme="implicit", cluster=0.0, viscosity=0.01, dt=1e-3),
    "cavity64_q2_implicit": dict(n=64, NENv=27, scheme="implicit", cluster=0.0, viscosity=0.01, dt=1e-3),
    # configs[0] / configs[1]: the shipped lidDrivenCavity_NE27000.inp (30^3 Q2/Q1, sinh clustering 2.0)
    "ne27000_explicit": dict(n=30, NENv=27, scheme="explicit", cluster=2.0, viscosity=0.01, dt=1e-3, inp=True),
    "ne27000_implicit": dict(n=30, NENv=27, scheme="implicit", cluster=2.0, viscosity=0.001, dt=1e-3, inp=True),
    # configs[2]: stand-in for the missing bendingSquareDuct_49x33x33.inp (straight 48x32x32 channel, DESIGN.md / BASELINE.md C3)
    "channel_48x32x32_implicit": dict(n=48, ny=32, nz=32, NENv=27, scheme="implicit", cluster=0.0, viscosity=0.01, dt=1e-3, inp=True, channel=True),
    # the same duct with the plug profile prescribed at the outlet as well: the stand-in on which the reference's implicit scheme converges
    # (with a free outlet -- the workload above -- the reference algorithm itself diverges: every Krylov solve runs into the 1000 cap)
    "duct_48x32x32_implicit": dict(n=48, ny=32, nz=32, NENv=27, scheme="implicit", cluster=0.0, viscosity=0.01, dt=1e-3, inp=True, channel=True, outlet="plug"),
}
DEFAULT_WORKLOAD = "cavity256_q1_implicit"
CPU_SAMPLE_N = 128   # mesh edge of the bounded CPU sample for Q1/Q1 workloads larger than this (the oracle steps 128^3 in seconds)


def sizes(w):
    nx, ny, nz, q2 = w["n"], w.get("ny", w["n"]), w.get("nz", w["n"]), w["NENv"] == 27
    NN = (2 * nx + 1) * (2 * ny + 1) * (2 * nz + 1) if q2 else (nx + 1) * (ny + 1) * (nz + 1)
    NNp = (nx + 1) * (ny + 1) * (nz + 1)
    return dict(NE=nx * ny * nz, NN=NN, NNp=NNp, DOF=3 * NN + NNp)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return None
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, reasons, mx = [], set(), None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return None
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(workload, kernel_class):
    """dram bytes per launch of the dominant kernel from a committed ncu --set full capture, if any."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(workload, {}).get(kernel_class)
        except Exception:
            return None
    return None


def make_inp(w, path, n=None, **extra):
    import meshgen
    if w.get("channel"):
        f = 1.0 if n is None else n / float(w["n"])
        meshgen.channel(path, nx=max(2, int(w["n"] * f)), ny=max(2, int(w["ny"] * f)), nz=max(2, int(w["nz"] * f)), NENv=w["NENv"],
                        viscosity=w["viscosity"], dt=w["dt"], outlet=w.get("outlet", "free"), **extra)
        return
    meshgen.cube(path, n or w["n"], cluster=w["cluster"], NENv=w["NENv"], viscosity=w["viscosity"], dt=w["dt"],
                 full_precision=not w.get("inp", False), **extra)


REF_BIN = {"explicit": os.path.join(ROOT, "oracle", "_ref", "bch_cpu"), "implicit": os.path.join(ROOT, "oracle", "_ref", "gq_cpu")}
REF_SAMPLE_N = 30   # the reference's own preprocessing and the serial symmetric SpMV of its CG are slow beyond the shipped mesh size


def run_reference_binary(w, steps, warmup, n):
    """The UNMODIFIED reference program (oracle/_ref: its source compiled against the MKL / CSparse / PARALUTION stand-ins of
    oracle/shim, see oracle/Makefile) run the way its users run it: ProblemName.txt + <name>.inp in the working directory.
    Times come from the `TimeSpend` column of its own monitor table.  It sets its thread count itself
    (N_OPENMP_THREADS = 8, BCH.cpp:86).  Returns seconds per step of the last `steps` steps."""
    with tempfile.TemporaryDirectory() as d:
        total = steps + warmup
        make_inp(w, os.path.join(d, "case.inp"), n, t_final=(total - 0.5) * w["dt"], converge=-1e30)
        with open(os.path.join(d, "ProblemName.txt"), "w") as f:
            f.write("case\n")
        env = dict(os.environ)
        env.pop("OMP_NUM_THREADS", None)   # the program sets its own 8 threads (BCH.cpp:85-86); the stand-ins' loops follow the same runtime
        out = subprocess.run([REF_BIN[w["scheme"]]], cwd=d, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, stdin=subprocess.DEVNULL,
                             text=True, check=True, env=env).stdout
    ncol, tcol = (9, 7) if w["scheme"] == "explicit" else (8, 6)
    rows = [ln.split() for ln in out.splitlines()]
    rows = [r for r in rows if len(r) == ncol and r[0].isdigit()]
    if len(rows) != total:
        raise RuntimeError("reference binary printed %d monitor rows, expected %d" % (len(rows), total))
    return sum(float(r[tcol]) for r in rows[warmup:]) / steps


def host_threads():
    """Cores this process may use.  Never inherited from OMP_NUM_THREADS: torch.distributed.run exports OMP_NUM_THREADS=1 to its
    workers, which in round 1 made the reference arm of the N > 1 runs a one-core measurement."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def run_oracle(w, steps, warmup, n, threads=0):
    """The CPU oracle (line-faithful restatement of the reference's CPU path) on an n^3 mesh of the workload's kind, on `threads`
    host threads (0 = every core of the box, set explicitly).  Returns seconds per step."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import orc
    threads = threads or host_threads()
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "sample.inp")
        make_inp(w, path, n)
        t0 = time.time()
        o = orc.Oracle(path, w["scheme"], threads)
        setup = time.time() - t0
        if warmup:
            o.run(warmup)
        t0 = time.time()
        o.run(steps)
        dt = (time.time() - t0) / steps
        cg = o.arr("cgIters")
        info = dict(setup_s=round(setup, 2), cg_iterations_per_solve=float(cg.mean()) if len(cg) else 0.0,
                    cg_gdof_iter_per_s=(o.i("NNp") * float(cg.sum()) / o.d("cgSeconds") / 1e9) if o.d("cgSeconds") > 0 else None,
                    threads=orc.threads())
        o.close()
    return dt, info


def cpu_sample(w, steps, warmup):
    if w["NENv"] == 27 and all(os.path.exists(b) for b in REF_BIN.values()):
        # Q2/Q1 workloads: the reference itself (it has no Q1/Q1 element: its NENv == 8 branch is commented out, BCH.cpp:2303-2305)
        n_s = min(w["n"], REF_SAMPLE_N)
        sec = run_reference_binary(w, steps, warmup, n_s)
        scale = (n_s / float(w["n"])) ** 3
        sample = ("the reference program itself (%s compiled against oracle/shim stand-ins for MKL/CSparse/PARALUTION), %s Q2/Q1 %s mesh, "
                  "%d timed steps after %d warm-up, per-step times from its own monitor table"
                  % ("blascoCodinaHuerta.cpp" if w["scheme"] == "explicit" else "guermondQuartapelle.cpp",
                     ("%d^3" % n_s) if not w.get("channel") else "%d/%d-scale channel" % (n_s, w["n"]), w["scheme"], steps, warmup))
        if n_s != w["n"]:
            sample += "; steps/s scaled by element count %d^3/%d^3 to the workload mesh" % (n_s, w["n"])
        return dict(value=scale / sec, unit="steps/s", cores=min(8, host_threads()), kind="reference", sample=sample,
                    sample_steps_per_s=1.0 / sec, extrapolated=n_s != w["n"], sample_mesh=n_s, workload_mesh=w["n"],
                    extrapolation=None if n_s == w["n"] else "steps/s x (sample NE / workload NE); Krylov iteration growth with n ignored, which favours the CPU")
    n_s = min(w["n"], CPU_SAMPLE_N)
    sec, info = run_oracle(w, steps, warmup, n_s)
    scale = (n_s / float(w["n"])) ** 3   # per-step work ~ elements (Krylov iteration growth with n ignored: favours the CPU)
    sample = ("oracle (C++/OpenMP restatement of the reference CPU path), %d^3 %s %s mesh, %d timed steps after %d warm-up"
              % (n_s, "Q2/Q1" if w["NENv"] == 27 else "Q1/Q1", w["scheme"], steps, warmup))
    if n_s != w["n"]:
        sample += "; steps/s scaled by element count %d^3/%d^3 to the workload mesh" % (n_s, w["n"])
    note = ("the reference program itself (oracle/_ref) is timed on the Q2/Q1 workloads; it has no Q1/Q1 element (its NENv == 8 shape-function "
            "branch is commented out, blascoCodinaHuerta.cpp:2303-2305), so this workload uses the oracle port, which reproduces the reference "
            "program's output on Q2/Q1 inputs to 1e-12 (tests/test_reference_pin.py)") if w["NENv"] == 8 else "oracle/_ref not built: oracle port"
    return dict(value=scale / sec, unit="steps/s", cores=info["threads"], kind="port", sample=sample,
                sample_steps_per_s=1.0 / sec, cg_gdof_iter_per_s=info["cg_gdof_iter_per_s"], note=note,
                extrapolated=n_s != w["n"], sample_mesh=n_s, workload_mesh=w["n"], sample_cg_iterations_per_solve=info["cg_iterations_per_solve"],
                extrapolation=None if n_s == w["n"] else "steps/s x (sample NE / workload NE); Krylov iteration growth with n ignored, which favours the CPU")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("CFD_BENCH_WORKLOAD", DEFAULT_WORKLOAD), choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cg-variant", type=int, default=0, help="cfd_options.cg_variant: 0 library default, 1 classical PCG, 2 single-reduction PCG")
    ap.add_argument("--deterministic", type=int, default=0, help="cfd_options.deterministic")
    ap.add_argument("--path", default="csr", choices=["csr", "ebe"],
                    help="operator path of the explicit scheme: csr = assembled matrices (Sell SpMV), ebe = matrix-free element kernels (cfd_options.operator_path)")
    ap.add_argument("--step0", default="device", choices=["host", "device"],
                    help="set-up: device (default on one GPU) = step0 (element integrals, Z, layout conversion) on the GPU (cfd_options.step0_device); "
                         "host = on the host cores, as the library itself defaults to (partitioned runs always)")
    a = ap.parse_args()
    w = WORKLOADS[a.workload]
    sz = sizes(w)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = dict(workload=a.workload, scheme=w["scheme"], element="Q2/Q1" if w["NENv"] == 27 else "Q1/Q1",
                  mesh=("%dx%dx%d hex channel" % (w["n"], w["ny"], w["nz"])) if w.get("channel") else "%d^3 hex cavity" % w["n"],
                  NE=sz["NE"], DOF=sz["DOF"], dt=w["dt"], viscosity=w["viscosity"],
                  krylov="Jacobi-PCG 1e-6/1000 (pressure)" + (", Jacobi-BiCGStab 1e-6/1000 x3 (momentum)" if w["scheme"] == "implicit" else ""))

    if a.impl == "reference":
        if rank != 0:
            return 0
        cb = cpu_sample(w, max(1, a.steps), a.warmup)
        line = dict(metric="time steps/s", value=cb["value"], unit="steps/s", n_gpus=a.gpus, steps=a.steps, warmup=a.warmup,
                    ms_per_step=1e3 / cb["value"], higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f64",
                    data="synthetic", impl="reference", config=config, cpu_baseline=cb,
                    e2e=dict(value=cb["value"], unit="steps/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
        if cb.get("extrapolated"):
            # machine-readable: `value` / `ms_per_step` are for the workload mesh, the process only ran the sample
            line["consistency"] = dict(fits_in_driver_run=False, same_config=False, measured_ms_per_step_on_sample=1e3 / cb["sample_steps_per_s"],
                                       why="the CPU path cannot step the %d^3 workload within minutes: timed on a %d^3 sample of the same mesh kind and scaled by "
                                           "element count (CPU-favouring)" % (cb["workload_mesh"], cb["sample_mesh"]))
        print(json.dumps(line))
        return 0

    import numpy as np
    import torch
    import fracstep
    if os.environ.get("BENCH_WATCHDOG_S"):   # diagnostics: dump every thread's Python stack and exit if the run takes longer than this
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["BENCH_WATCHDOG_S"]), exit=True)
    t_start = time.time()

    def stage(msg):
        if os.environ.get("BENCH_VERBOSE"):
            print("[bench rank %d %7.1f s] %s" % (rank, time.time() - t_start, msg), file=sys.stderr, flush=True)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if rank == 0:
        fracstep.build()
    if dist is not None:
        dist.barrier()
    opts = fracstep.options(device=local_rank, cg_variant=a.cg_variant, deterministic=a.deterministic, operator_path=1 if a.path == "ebe" else 0,
                            step0_device=1 if a.step0 == "device" and world == 1 else 0)
    t0 = time.time()
    tmp = None
    if w.get("inp"):
        tmp = tempfile.TemporaryDirectory()
        path = os.path.join(tmp.name, a.workload + ".inp")
        make_inp(w, path)
        s = fracstep.Solver(path, w["scheme"], opts=opts, world=world, rank=rank)
    else:
        s = fracstep.Solver(scheme=w["scheme"], opts=opts, world=world, rank=rank,
                            cavity=dict(nx=w["n"], cluster=w["cluster"], NENv=w["NENv"], dt=w["dt"], viscosity=w["viscosity"]))
    if world > 1:
        # element-block partition: rank r holds block r; NCCL communicator for halo swaps and allreduces
        # (+ CUDA-IPC exchange windows: the pressure solve then runs as one persistent kernel per GPU that talks to
        #  its neighbours through peer memory; CFD_NO_P2P=1 keeps the host-driven NCCL version)
        fracstep.setup_distributed(s, dist, torch, peer_memory=True)
    setup_s = time.time() - t0
    stage('set-up done')
    c = s.ctx

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up, then the timed region (state resident in HBM)
    if a.warmup:
        c.time_loop(a.warmup)
    # the state the timed region starts from, kept on the host so that the end-to-end leg repeats the SAME K steps
    state_names = ("Un", "Pn", "Unp1_prev") if w["scheme"] == "explicit" else ("Uk_prev", "Pk_prev", "Pk_prevprev")
    state0 = {nm: c.download(nm) for nm in state_names}
    stage('warm-up done')
    clocks = ClockSampler(local_rank)
    barrier()
    clocks.start()
    c.profile(True)
    l0 = c.kernel_launches()
    reps, ms = c.time_loop_timed(a.steps)
    launches = c.kernel_launches() - l0
    barrier()
    clk = clocks.stop()
    classes = ["ebe", "step1", "gt_rhs", "pcg", "update", "step3", "spmv_k", "assemble", "gq_rhs", "bicgstab", "halo"]
    prof = {k: c.profile_get(k) for k in classes}
    c.profile(False)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())

    def global_norms():
        """||U||_2, ||P||_2 over the whole mesh (owned entries of every rank): identical numbers at N = 1, 2, 4, 8 mean identical fields."""
        _, Uo, _, Po = s.owned_fields()
        v = torch.tensor([float((Uo * Uo).sum()), float((Po * Po).sum())], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(v)
        return float(v[0].sqrt().item()), float(v[1].sqrt().item())

    normU, normP = global_norms()
    stage('timed region done')

    # ---- end to end: the same K steps again, from the same state, host buffers in / host buffers out every step
    for nm in state_names:
        c.upload(nm, state0[nm])
    Uh = torch.from_numpy(state0[state_names[0]].copy()).pin_memory().numpy()
    Ph = torch.from_numpy(state0[state_names[1]].copy()).pin_memory().numpy()
    barrier()
    e2e_ms = 0.0
    e2e_reps = []
    for _ in range(a.steps):
        r1, m1 = c.time_step_host(Uh, Ph)
        e2e_reps.append(r1)
        e2e_ms += m1
    barrier()
    t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    normU2, normP2 = global_norms()
    stage('end-to-end leg done')

    def finish():
        if dist is not None:
            fracstep.teardown_distributed(s, dist)
            dist.destroy_process_group()
        else:
            s.close()

    if rank != 0:
        finish()
        return 0
    n_rep = 1       # one partitioned job over all ranks (strong scaling: the mesh is fixed, each rank holds an element block)
    peak, peak_src = measured_peak()
    total_ms = sum(p["ms"] for k, p in prof.items() if k != "halo")   # halo swaps inside Krylov solves are already in pcg / bicgstab
    dom = max((k for k in prof if k != "halo"), key=lambda k: prof[k]["ms"])
    d = prof[dom]
    sec = d["ms"] * 1e-3
    alg = d["bytes"] / sec / 1e9 if sec > 0 else 0.0
    lay = d["layout_bytes"] / sec / 1e9 if sec > 0 else 0.0
    cg_its = sum(r["cg_iterations"] for r in reps)
    traffic = ncu_traffic(a.workload, dom)
    cg_ms = prof["pcg"]["ms"]
    kernels = {k: dict(ms=round(v["ms"], 3), launches=v["launches"],
                       gbs=round(v["layout_bytes"] / (v["ms"] * 1e-3) / 1e9, 1) if v["ms"] > 0 else 0.0,
                       gbs_algorithmic=round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1) if v["ms"] > 0 else 0.0,
                       share=round(v["ms"] / total_ms, 4) if total_ms else 0.0) for k, v in prof.items() if v["launches"]}
    # what is resident where: stored bytes of the matrices on this rank against the 126 MB L2
    mb = {}
    for nm, nval in (("K", 1 if w["scheme"] == "explicit" else 3), ("G", 3), ("GT", 3), ("Z", 1)):
        try:
            st = c.matrix_stats(nm)
            mb[nm] = (8.0 * st["stored_values"] * nval + 4.0 * st["stored_indices"]) / 1e6
        except Exception:
            mb[nm] = 0.0
    per_step_gb = sum(v["layout_bytes"] for v in prof.values()) / a.steps / 1e9
    if sum(mb.values()) > 4 * 126:
        l2_note = "matrices %.0f MB on this rank, %.2f GB streamed per step >> 126 MB L2 (no flush needed)" % (sum(mb.values()), per_step_gb)
    else:
        l2_note = ("small workload, timed as it runs in production (a time loop re-reads the same arrays every step; no flush): pressure matrix Z %.0f MB "
                   "stays in shared memory / L2 across the CG iterations of a solve, K %.0f MB + G, G^T %.0f MB (> L2 together) stream from HBM once per kernel"
                   % (mb["Z"], mb["K"], mb["G"] + mb["GT"]))
    config.update(parallelism="1 GPU" if world == 1 else "%d element blocks (z-slabs), one per GPU; peer-memory Krylov kernels + NCCL halo swap / allreduce" % world,
                  l2=l2_note,
                  setup_s=round(setup_s, 1), cg_iterations_per_step=cg_its / float(a.steps), cg_variant=a.cg_variant, operator_path=a.path, step0=a.step0 if world == 1 else "host",
                  bicgstab_iterations_last_step=reps[-1]["bicg_iterations"])
    krylov_ok = all(r["cg_converged"] for r in reps) and all(all(r["bicg_converged"]) for r in reps)
    line = dict(metric="time steps/s", value=n_rep * a.steps / (ms_max * 1e-3), unit="steps/s", n_gpus=a.gpus, steps=a.steps, warmup=a.warmup,
                ms_per_step=ms_max / a.steps, higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f64", data="synthetic",
                config=config, clocks=clk,
                e2e=dict(value=n_rep * a.steps / (e2e_ms * 1e-3), unit="steps/s", h2d_bytes_per_step=int(8 * (Uh.size + Ph.size)),
                         d2h_bytes_per_step=int(8 * (Uh.size + Ph.size)),
                         note="per rank: its block's U,P (owned + ghost) each step; the same K steps as `value`, restarted from the same saved state",
                         fields_vs_device_run=dict(normU_rel_diff=abs(normU2 / normU - 1.0) if normU else 0.0,
                                                   normP_rel_diff=abs(normP2 / normP - 1.0) if normP else 0.0)),
                gpu_launches=int(launches),
                # fields after the timed steps: the same numbers at every N = the partitioned runs compute the same thing
                fields=dict(normU=normU, normP=normP, monitor=reps[-1]["monitor"], maxAcc=reps[-1]["maxAcc"],
                            cg_iterations=[r["cg_iterations"] for r in reps], krylov_converged=krylov_ok,
                            bicg_true_residual_max=max([v for r in reps for v in r["bicg_true_residual"]] + [-1.0])),
                roofline=dict(bound="hbm", kernel=dom, achieved=round(lay, 1), peak=peak, unit="GB/s", frac=round(lay / peak, 4),
                              frac_of_nominal_8000=round(lay / 8000.0, 4),   # both the measured-copy and the nominal peak
                              achieved_algorithmic=round(alg, 1), frac_algorithmic=round(alg / peak, 4),
                              traffic=traffic, peak_source=peak_src, share_of_step=round(d["ms"] / total_ms, 4) if total_ms else None,
                              note="achieved = bytes the kernel has to move for its launches in the layout it stores (Sell: zero padding in, per-slice "
                                   "compressed column indices counted once; vectors as in the algorithmic model) / CUDA-event time of those launches: "
                                   "cannot exceed the bandwidth.  achieved_algorithmic = the CSR model of SURVEY.md 8(d) (12 B per non-zero + vectors) over "
                                   "the same time: what a CSR kernel would have had to sustain, > 1 of peak where the compression pays.  traffic = ncu "
                                   "dram__bytes per launch from the committed capture (profiles/traffic.json)"),
                cg_gdof_iter_per_s=(sz["NNp"] * cg_its / (cg_ms * 1e-3) / 1e9) if cg_ms > 0 else None,
                kernels=kernels)
    if dom == "pcg" and world == 1 and 0 < mb["Z"] <= 33.0:
        line["roofline"]["resident"] = ("pressure matrix (%.0f MB) resident in the SMs' shared memory for the whole solve (k_pcg_res): DRAM traffic per solve = one read "
                                        "of the matrix (`traffic`), so `achieved` is what a streaming kernel would have had to sustain, not HBM traffic" % mb["Z"])
    if world == 1 and not a.no_cpu_baseline:
        line["cpu_baseline"] = cpu_sample(w, 3, 1)
    print(json.dumps(line))
    finish()
    return 0


if __name__ == "__main__":
    sys.exit(main())
