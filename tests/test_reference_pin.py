"""Pins the oracle -- and the CUDA path -- against outputs of the REFERENCE ITSELF.

tests/golden/ref_*.npz were written by the unmodified reference programs (blascoCodinaHuerta.cpp,
guermondQuartapelle.cpp) compiled here against stand-ins for MKL / CSparse / PARALUTION (oracle/shim, `make -C oracle ref`)
and run the way a user runs them; tests/golden/make_ref_golden.py is the generating script.

  as-is fixtures  the reference's own solver settings.  Explicit (Jacobi-CG, ||r||^2 <= 1e-12 ||r0||^2, BCH.cpp:4011-4018)
                  equals the oracle's / product's default, so fields must agree to print precision and CG counts exactly.
                  Implicit (plain CG to 1e-3, BiCGStab to 1e-6 on an under-integrated momentum matrix) is compared with the
                  same settings but a loose bar: where in the iteration the residual is tested differs between BiCGStab
                  implementations and moves the fields by ~1e-5.
  tight fixtures  Krylov tolerances forced to 1e-12 on both sides: isolates everything that is the reference's own code.
"""
import ast
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
ROOT = os.path.dirname(HERE)

SMALL = ["ref_bch_cav4", "ref_bch_cav6c", "ref_bch_chan644", "ref_gq_cav4", "ref_gq_cav4_tight", "ref_gq_cav6c_tight",
         "ref_gq_chan644_tight", "ref_bch_box345_nopin", "ref_bch_cav4_pin1_it1", "ref_gq_box345_nopin_tight"]
FULL = ["ref_bch_ne27000", "ref_gq_ne27000"]
# added after the round's GPU budget was spent: checked against the oracle here, against the CUDA path from round 2 on
CPU_ONLY = ["ref_bch_cav8c_20steps", "ref_bch_chan483232",
            # duct stand-in with a prescribed outflow (meshgen.duct): written by the reference programs at the end of round 2, after the
            # GPU budget was spent -- the oracle is pinned against them here; the CUDA path meets the oracle on the same case in
            # tests/test_gpu_parity.py::test_duct_standin_implicit
            "ref_bch_duct644", "ref_gq_duct644_tight", "ref_gq_duct1288_tight"]


def load(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    meta = dict(scheme=str(g["scheme"]), kind=str(g["mesh_kind"]), kw=ast.literal_eval(str(g["mesh_kw"])),
                env=ast.literal_eval(str(g["shim_env"])), steps=int(g["steps"]), stride=int(g["stride"]))
    return g, meta


def settings(meta):
    """Solver settings that mirror the run that produced the fixture: (oracle options, product options)."""
    env = meta["env"]
    if env:   # tight: REF_SHIM_* told the stand-ins what to do
        ct, bt = float(env["REF_SHIM_CG_TOL"]), float(env["REF_SHIM_BICG_TOL"])
        # GPU side: coloured element scatter + static slice assignment = a bitwise repeatable run.  BiCGStab sits on the rounding
        # floor of the under-integrated Q2 matrix for hundreds of iterations at 1e-12 and the path it takes there moves with every
        # rounding change; what makes the comparison robust is that the product accepts a momentum solve only on its TRUE residual
        # (restart otherwise) -- test_tight_fixture_is_robust_to_rounding_paths repeats the run with the atomic scatter.
        return dict(cg_tol=ct, bicg_tol=bt, cg_precond=1), dict(cg_tol=ct, bicg_tol=bt, ebe_scatter=1, deterministic=1)
    if meta["scheme"] == "implicit":
        # GQ.cpp:4510-4515: no Jacobi, dpar[0] = 1e-6 on SQUARED norms; BiCGStab: one residual test per iteration, which is where the
        # PARALUTION-style solver of the reference's CPU build stops (GQ.cpp:4190-4229)
        return dict(cg_precond=0, cg_tol=1e-3, bicg_halfstep=0), dict(cg_no_precond=1, cg_tol=1e-3, bicg_no_halfstep=1)
    return {}, {}


def errors(U, P, g, meta, NN):
    rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
    if meta["stride"] == 1:
        if meta["kw"].get("pin", 2) <= 1:   # no pin (negative node, or node 1 which is never pinned): P is defined up to a constant
            return rel(U, g["U"]), rel(P - P.mean(), g["P"] - g["P"].mean())
        return rel(U, g["U"]), rel(P, g["P"])
    idx = g["node_idx"]
    Us = np.stack([U[idx], U[NN + idx], U[2 * NN + idx]])
    # the stored subset, plus the norms of the complete fields
    eu = max(rel(Us, g["U_sub"]), abs(np.linalg.norm(U) / float(g["normU"]) - 1))
    ep = max(rel(P[g["pres_idx"]], g["P_sub"]), abs(np.linalg.norm(P) / float(g["normP"]) - 1))
    return eu, ep


def bars(name, meta):
    """(rel-L2 bar for U, for P, allowed CG-count difference per solve or None = not compared)."""
    if meta["kw"].get("pin", 2) <= 1:   # singular Z (no pressure pin): rounding is amplified, CG counts may move by one
        return 1e-7, 1e-7, 2
    if meta["scheme"] == "explicit":
        return 1e-10, 1e-10, 0
    if meta["env"]:   # forced to 1e-12: the last iterations sit on the rounding floor, the count there is noisy (the +-2 bar
        return 1e-7, 1e-7, 5   # of BASELINE.json is for the reference tolerance and is applied to the as-is fixtures)
    # Implicit at the reference's own tolerances (BiCGStab 1e-6, plain CG 1e-3): NOT a parity bar but the measured floor of the
    # comparison.  Both sides solve A u = b to |b - A u| <= 1e-6 |b| with cond(A) ~ 1e4 (under-integrated Q2 mass matrix,
    # GQ.cpp:3490-3500), so two correct runs may differ by ~1e-6 * cond in the fields.  Measured (scripts/asis_spread.py,
    # profiles/r02_asis_spread.txt): the ORACLE against itself with 1 vs 8 threads (= another summation order of its dot
    # products) moves by 2.4e-5 / 2.4e-4 (U / P) on the 4^3 fixture and 3.8e-5 / 2.8e-5 on the shipped mesh; oracle vs reference
    # 1.6e-5..8.6e-5 / 1.4e-4..5.8e-4 and 1.6e-5..3.8e-5 / 4.7e-7..2.8e-5.  Bars = 3x the largest measured value; the arithmetic
    # inside PARALUTION's BiCGStab is a stand-in (oracle/shim), so this comparison is "parity unpinned" by construction and the
    # 1e-6 parity claim rests on the tight fixtures above.
    return (1.2e-4, 1e-4, 6) if "ne27000" in name else (3e-4, 2e-3, 2)   # (plain CG stopped at 1e-3: its count moves by up to 5 of ~150)


def inp_for(mesh, name, meta):
    return mesh(name, kind=meta["kind"], **meta["kw"])


# ---------------------------------------------------------------------------------------------------- CPU: the oracle
@pytest.mark.parametrize("name", SMALL + FULL + CPU_ONLY)
def test_oracle_reproduces_reference_run(mesh, name):
    import orc
    g, meta = load(name)
    # tight fixtures: one thread, so that the summation order -- and with it the path BiCGStab takes along the rounding
    # floor -- is the same in every run of this test
    one_thread = meta["scheme"] == "implicit" and meta["stride"] == 1      # small implicit fixtures (as-is ones too): deterministic
    o = orc.Oracle(inp_for(mesh, name, meta), meta["scheme"], 1 if one_thread else 0, **settings(meta)[0])
    o.run(meta["steps"])
    eu, ep = errors(o.U, o.P, g, meta, o.i("NN"))
    bu, bp, dcg = bars(name, meta)
    assert eu <= bu and ep <= bp, (name, eu, ep)
    cg = list(o.arr("cgIters"))
    assert len(cg) == len(g["cg"])
    if dcg is not None:
        assert all(abs(a - b) <= dcg for a, b in zip(cg, g["cg"][:, 0])), (cg, list(g["cg"][:, 0]))
    if meta["scheme"] == "explicit":
        # second column of the reference's monitor table = inner iterations it printed (maxIter+1 when not converged)
        assert list(o.arr("innerIters")) == [int(v) for v in g["monitor"][:, 1]]
    # monitor-point values as printed by the reference (%12.5f): u, v, w, p of the last step
    mon = o.i("monPoint")
    NN = o.i("NN")
    cols = g["monitor"][-1, 3:7] if meta["scheme"] == "explicit" else g["monitor"][-1, 2:6]
    ours = [o.U[mon], o.U[NN + mon], o.U[2 * NN + mon], o.P[mon]]
    tol = 6e-6 + (0.0 if meta["scheme"] == "explicit" or meta["env"] else 1e-2 * max(1.0, abs(cols[3])))
    assert np.allclose(ours, cols, rtol=0, atol=tol), (ours, cols)


@pytest.mark.parametrize("name", ["ref_bch_cav4", "ref_gq_cav4_tight"])
def test_reference_binary_still_writes_the_fixture(name, tmp_path):
    """Re-runs the compiled reference (oracle/_ref, when present) and compares its .dat with the committed fixture."""
    g, meta = load(name)
    exe = os.path.join(ROOT, "oracle", "_ref", "bch_cpu" if meta["scheme"] == "explicit" else "gq_cpu")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref not built (needs the reference tree: make -C oracle ref)")
    import meshgen
    kw = dict(meta["kw"])
    getattr(meshgen, meta["kind"])(str(tmp_path / "case.inp"), t_final=(meta["steps"] - 0.5) * kw.get("dt", 0.001), **kw)
    (tmp_path / "ProblemName.txt").write_text("case\n")
    subprocess.run([exe], cwd=str(tmp_path), env=dict(os.environ, **meta["env"]), stdout=subprocess.DEVNULL,
                   stdin=subprocess.DEVNULL, check=True, timeout=300)
    t = np.loadtxt(str(tmp_path / "case.dat"), skiprows=3, max_rows=int(g["NN"]))
    U = np.concatenate([t[:, 3], t[:, 4], t[:, 5]])
    # threads may sum dot products in another order (and the implicit momentum matrix amplifies that): rounding-level bar
    eu, ep = errors(U, t[:int(g["NNp"]), 6], g, meta, int(g["NN"]))
    assert eu <= 1e-8 and ep <= 1e-8, (eu, ep)


@pytest.mark.parametrize("name", ["ref_bch_cav6c", "ref_gq_chan644_tight"])
def test_tecplot_writer_against_reference_file(mesh, name, tmp_path):
    """createTecplot (BCH.cpp:4249-4482): the product's writer and the oracle's, fed the reference's fields, against the
    file the reference wrote -- header, node rows (coordinates, u, v, w, pressure interpolated to the generated nodes)
    and the 8-sub-brick connectivity."""
    import fracstep
    import orc
    g, meta = load(name)
    path = inp_for(mesh, name, meta)
    s = fracstep.Solver(path, meta["scheme"], init_gpu=False)
    o = orc.Oracle(path, meta["scheme"])
    NN = int(g["NN"])
    for who, write in (("product", lambda f: s.write_tecplot_fields(f, "case", g["U"], g["P"])),):
        f = str(tmp_path / (who + ".dat"))
        write(f)
        lines = open(f).read().split("\n")
        assert "\n".join(lines[:3]) + "\n" == str(g["dat_header"])
        t = np.array([[float(v) for v in ln.split()] for ln in lines[3:3 + NN]])
        assert np.array_equal(t[:, :3], g["coord"])
        assert np.array_equal(t[:, 3:6].T.ravel(), g["U"])
        # generated nodes: average of the END-node pressures, formed here from the 12-digit values the reference printed
        assert np.allclose(t[:, 6], g["P_all_nodes"], rtol=1e-10, atol=1e-13)
        conn = np.array([[int(v) for v in ln.split()] for ln in lines[3 + NN:] if ln.strip()], dtype=np.int32)
        assert np.array_equal(conn, g["conn"])
    # the oracle's own writer after its own run: same header / connectivity
    o.run(1)
    f = str(tmp_path / "oracle.dat")
    o.write_tecplot(f, "case")
    lines = open(f).read().split("\n")
    assert "\n".join(lines[:3]) + "\n" == str(g["dat_header"])
    conn = np.array([[int(v) for v in ln.split()] for ln in lines[3 + NN:] if ln.strip()], dtype=np.int32)
    assert np.array_equal(conn, g["conn"])


RESTART = ["ref_bch_cav4_restart", "ref_gq_cav4_restart_tight"]


def restart_dir(tmp_path, g, meta):
    """case.inp with isRestart 1 and case_restart.dat rebuilt from the rows the reference's FIRST run wrote
    (`scientific, precision(11)`, BCH.cpp:4288-4292 == "%.11e")."""
    import meshgen
    getattr(meshgen, meta["kind"])(str(tmp_path / "case.inp"), is_restart=1, **meta["kw"])
    with open(str(tmp_path / "case_restart.dat"), "w") as f:
        f.write(str(g["dat_header"]))
        for row in g["restart_table"]:
            f.write(" ".join("%.11e" % v for v in row) + "\n")
    return str(tmp_path / "case.inp"), str(tmp_path / "case_restart.dat")


@pytest.mark.parametrize("name", RESTART)
def test_oracle_reproduces_reference_restart(name, tmp_path):
    """readRestartFile (BCH.cpp:4214-4242, GQ.cpp likewise): the reference continued from its own .dat file."""
    import orc
    g, meta = load(name)
    inp, dat = restart_dir(tmp_path, g, meta)
    o = orc.Oracle(inp, meta["scheme"], 1 if meta["env"] else 0, **settings(meta)[0])
    o.read_restart(dat)
    o.run(meta["steps"])
    eu, ep = errors(o.U, o.P, g, meta, o.i("NN"))
    assert eu <= 1e-7 and ep <= 1e-7, (eu, ep)
    assert all(abs(a - b) <= (5 if meta["env"] else 0) for a, b in zip(o.arr("cgIters"), g["cg"][:, 0]))


@pytest.mark.parametrize("name", RESTART)
def test_product_reads_the_reference_restart_file(name, tmp_path):
    """The product's (parallel) restart reader on the .dat the reference wrote: u, v, w of every node, p of the corner nodes."""
    import fracstep
    g, meta = load(name)
    inp, dat = restart_dir(tmp_path, g, meta)
    s = fracstep.Solver(inp, meta["scheme"], init_gpu=False)
    U, P = s.read_restart_fields(dat)
    t = g["restart_table"]
    assert np.array_equal(U, np.concatenate([t[:, 3], t[:, 4], t[:, 5]])) and np.array_equal(P, t[:int(g["NNp"]), 6])


# ---------------------------------------------------------------------------------------------------- GPU: the product
@pytest.mark.gpu
@pytest.mark.parametrize("name", RESTART)
def test_cuda_path_reproduces_reference_restart(name, tmp_path):
    import fracstep
    g, meta = load(name)
    inp, _ = restart_dir(tmp_path, g, meta)
    popt = settings(meta)[1]
    s = fracstep.Solver(inp, meta["scheme"], opts=fracstep.options(**popt) if popt else None)   # reads case_restart.dat itself
    reps = s.ctx.time_loop(meta["steps"])
    U, P = s.ctx.fields()
    eu, ep = errors(U, P, g, meta, s.i("NN"))
    assert eu <= 1e-6 and ep <= 1e-6, (eu, ep)
    assert abs(sum(r["cg_iterations"] for r in reps) - int(g["cg"][:, 0].sum())) <= (5 if meta["env"] else 2) * len(g["cg"])
    s.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", SMALL + FULL)
def test_cuda_path_reproduces_reference_run(mesh, name):
    """Whole product pipeline (own reader, node generation, step0, CUDA time loop through the C ABI) against what the
    reference program wrote for the same input file.  BASELINE.json bar: rel-L2 <= 1e-6, CG counts +-2."""
    import fracstep
    g, meta = load(name)
    popt = settings(meta)[1]
    s = fracstep.Solver(inp_for(mesh, name, meta), meta["scheme"], opts=fracstep.options(**popt) if popt else None)
    reps = s.ctx.time_loop(meta["steps"])
    U, P = s.ctx.fields()
    eu, ep = errors(U, P, g, meta, s.i("NN"))
    bu, bp, dcg = bars(name, meta)
    bu, bp = max(bu, 1e-6), max(bp, 1e-6)
    assert eu <= bu and ep <= bp, (name, eu, ep)
    assert all(r["cg_converged"] for r in reps)
    if meta["scheme"] == "explicit":
        assert [r["iter"] for r in reps] == [int(v) for v in g["monitor"][:, 1]]
        assert abs(sum(r["cg_iterations"] for r in reps) - int(g["cg"][:, 0].sum())) <= 2 * len(g["cg"])
    else:
        assert all(abs(r["cg_iterations"] - k) <= max(dcg, 2) for r, k in zip(reps, g["cg"][:, 0]))
        # every momentum solve is accepted on its true residual |b - A x| / |b|: at the reference tolerance that is the stopping
        # rule itself; at 1e-12 the attainable accuracy of this matrix is about 1e-12, a decade is allowed
        tol = popt.get("bicg_tol", 1e-6)
        maxit = popt.get("bicg_maxit", 1000)
        for r in reps:
            for d in range(3):
                if r["bicg_converged"][d]:
                    assert r["bicg_true_residual"][d] <= tol, (d, r)
                else:
                    # the reference's cap of 1000 iterations is reached now and then on the shipped mesh at Re 1000 (its own BiCGStab
                    # needs 640-680 there and rounding moves that by hundreds): reported as unconverged, never silently accepted; at
                    # 1e-12 a solve may also stop on the attainable accuracy of the matrix, a decade above the request
                    assert r["bicg_iterations"][d] >= maxit or r["bicg_true_residual"][d] <= 10 * tol, (d, r)
    s.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["ref_gq_cav6c_tight", "ref_gq_chan644_tight"])
def test_tight_fixture_is_robust_to_rounding_paths(mesh, name):
    """Round 1's failure mode: the same fixture gave 3e-8 on one box and 6.5e-5 on another.  Ten runs in a row with the ATOMIC
    element scatter and the dynamically scheduled sweeps (every run rounds differently, so BiCGStab takes another path along the
    rounding floor each time) must all meet the 1e-6 bar, each solve with a verified true residual."""
    import fracstep
    g, meta = load(name)
    popt = dict(settings(meta)[1], ebe_scatter=0, deterministic=0)
    worst = (0.0, 0.0, 0.0)
    for _ in range(10):
        s = fracstep.Solver(inp_for(mesh, name, meta), meta["scheme"], opts=fracstep.options(**popt))
        reps = s.ctx.time_loop(meta["steps"])
        U, P = s.ctx.fields()
        eu, ep = errors(U, P, g, meta, s.i("NN"))
        tr = max(v for r in reps for v in r["bicg_true_residual"])
        worst = (max(worst[0], eu), max(worst[1], ep), max(worst[2], tr))
        s.close()
    assert worst[0] <= 1e-6 and worst[1] <= 1e-6 and worst[2] <= 1e-11, worst


@pytest.mark.gpu
@pytest.mark.parametrize("name,env", [("ref_bch_cav4", {}), ("ref_bch_chan644", {}),
                                      ("ref_gq_cav4_tight", {"CFD_CG_TOL": "1e-12", "CFD_BICG_TOL": "1e-12"})])
def test_solverGPU_program_against_reference_files(name, env, tmp_path):
    """Process-level drop-in: bin/solverGPU_<scheme> run like the reference's executable -- no arguments,
    ProblemName.txt + case.inp in the working directory -- must write the case.dat the reference wrote and print the same
    monitor table (all columns but the wall-clock one)."""
    import meshgen
    g, meta = load(name)
    exe = os.path.join(ROOT, "cfd-with-cuda_b200", "bin", "solverGPU_" + meta["scheme"])
    kw = dict(meta["kw"])
    getattr(meshgen, meta["kind"])(str(tmp_path / "case.inp"), t_final=(meta["steps"] - 0.5) * kw.get("dt", 0.001), **kw)
    (tmp_path / "ProblemName.txt").write_text("case\n")
    out = subprocess.run([exe], cwd=str(tmp_path), env=dict(os.environ, **env), capture_output=True, text=True, stdin=subprocess.DEVNULL,
                         timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    NN = int(g["NN"])
    lines = open(str(tmp_path / "case.dat")).read().split("\n")
    assert "\n".join(lines[:3]) + "\n" == str(g["dat_header"])
    t = np.array([[float(v) for v in ln.split()] for ln in lines[3:3 + NN]])
    assert np.array_equal(t[:, :3], g["coord"])
    eu, ep = errors(t[:, 3:6].T.ravel(), t[:int(g["NNp"]), 6], g, meta, NN)
    assert eu <= 1e-6 and ep <= 1e-6, (eu, ep)
    assert np.allclose(t[:, 6], g["P_all_nodes"], rtol=0, atol=1e-6 * np.abs(g["P_all_nodes"]).max())
    conn = np.array([[int(v) for v in ln.split()] for ln in lines[3 + NN:] if ln.strip()], dtype=np.int32)
    assert np.array_equal(conn, g["conn"])
    # monitor table: explicit "timeN iter time u v w p TimeSpend maxAcc", implicit "timeN time u v w p TimeSpend maxAcc"
    ncol, skip = (9, 7) if meta["scheme"] == "explicit" else (8, 6)
    rows = [ln.split() for ln in out.stdout.split("\n")]
    rows = np.array([[float(v) for v in r] for r in rows if len(r) == ncol and r[0].isdigit()])
    ref = g["monitor"]
    assert rows.shape == ref.shape
    keep = [c for c in range(ncol) if c != skip]
    if meta["scheme"] == "implicit":
        # the implicit CPU loop's maxAcc is not a maximum: `if (abs(a) > maxAcc) maxAcc = a` keeps SIGNED values, so later
        # entries are compared against a possibly negative running value (GQ.cpp:3336-3341); the GPU builds take max|a|
        keep.remove(ncol - 1)
    assert np.allclose(rows[:, keep], ref[:, keep], rtol=1e-4, atol=2e-5), (rows, ref)
