"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Tolerances: FP64 fields rel-L2 <= 1e-6 after N steps (BASELINE.json north_star); single phases
agree to summation-order rounding (<= 1e-11 relative); CG iteration counts within +-2."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = {
    "q2_c4": dict(n=4, cluster=2.0),
    "q2_u5": dict(n=5, cluster=0.0, full_precision=True),
    "q1_u8": dict(n=8, NENv=8, full_precision=True),
}


def rel(a, b):
    d = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (d if d > 0 else 1.0)


def rel_free(U, Uref, src):
    """rel-L2 over the non-Dirichlet velocity DOFs only (the lid values would otherwise dominate the norm)."""
    NN = src.i("NN")
    free = np.ones(NN, bool)
    free[src.arr("BCvelNodes")[0::2]] = False
    m = np.tile(free, 3)
    return rel(U[m], Uref[m])


def make_pair(mesh, name, scheme, **opt):
    import fracstep
    import orc
    path = mesh(name, **CASES[name])
    o = orc.Oracle(path, scheme)
    ctx = fracstep.Context(fracstep.Problem(o, scheme), fracstep.options(**opt) if opt else None)
    return o, ctx


@pytest.mark.parametrize("name", list(CASES))
def test_explicit_phases(mesh, name):
    """Every seam function of the explicit scheme against the oracle's phase, two inner iterations."""
    o, c = make_pair(mesh, name, "explicit")
    o.run(2)   # a non-trivial state
    c.call("cfd_timeLoop", 2, 0, None, None)
    U, P = c.fields()
    assert rel(U, o.U) < 1e-7 and rel(P, o.P) < 1e-5   # both sides stop CG at 1e-6; Q1/Q1 explicit Z is close to singular
    # bring both to the same state and compare phase by phase
    for nm in ("Un", "Pn", "Unp1_prev"):
        c.upload(nm, o.arr(nm))
    # begin of step: UnpHalf_prev <- Un etc. (BCH.cpp:2866-2869)
    c.call("bch_beginTimeStep")
    for it in (1, 2):
        if it == 1:
            # oracle: emulate the loop-top copies
            for nm, src in (("UnpHalf_prev", "Un"), ("Pnp1_prev", "Pn")):
                np.copyto(o.arr(nm, copy=False), o.arr(src))
            o.arr("Acc_prev", copy=False)[:] = 0
        o.phase("bch_calculateMatrixA")
        c.call("bch_calculateMatrixAGPU")
        # the convective term alone is ~|u|^2 (tiny this early in the run): measured against the
        # full step-1 right-hand side below; test_ebe_random_fields checks it in isolation
        r1_ebe_gpu, r1_ebe_ref = c.download("R1"), o.arr("R1")
        o.phase("bch_step1")
        c.call("bch_step1GPUpart", it)
        assert rel(c.download("R1"), o.arr("R1")) < 1e-11
        # (inputs of the second pass already differ by rounding, amplified by 1/h in the gradient)
        assert np.linalg.norm(r1_ebe_gpu - r1_ebe_ref) < 1e-8 * np.linalg.norm(o.arr("R1"))
        assert rel(c.download("UnpHalf"), o.arr("UnpHalf")) < 1e-12
        o.phase("bch_step2", it)
        c.call("bch_step2GPU", it)
        assert rel(c.download("R2"), o.arr("R2")) < 1e-10
        assert abs(c.last_cg() - o.arr("cgIters")[-1]) <= 2
        assert rel(c.download("Pdot"), o.arr("Pdot")) < 1e-5   # solver tolerance 1e-6 on both sides
        # continue from identical Pdot so that later phases compare to rounding
        c.upload("Pdot", o.arr("Pdot"))
        c.upload("Pnp1", o.arr("Pnp1"))
        o.phase("bch_step3", it)
        c.call("bch_step3GPU", it)
        for nm in ("R3", "Acc", "Unp1"):
            assert rel(c.download(nm), o.arr(nm)) < 1e-11, nm
        # rotation + K*Acc_prev
        for nm, src in (("UnpHalf_prev", "UnpHalf"), ("Unp1_prev", "Unp1"), ("Acc_prev", "Acc"), ("Pnp1_prev", "Pnp1")):
            np.copyto(o.arr(nm, copy=False), o.arr(src))
        c.call("bch_rotateIterates")
        c.call("bch_calculate_KtimesAcc_prevGPU")
        import scipy.sparse as sp
        NN = o.i("NN")
        K = sp.csr_matrix((o.arr("sparseKvalue"), o.arr("sparseMcol"), o.arr("sparseMrowStarts")), shape=(NN, NN))
        ka = np.concatenate([K @ o.arr("Acc_prev")[d * NN:(d + 1) * NN] for d in range(3)])
        np.copyto(o.arr("KtimesAcc_prev", copy=False), ka)
        assert rel(c.download("KtimesAcc_prev"), ka) < 1e-12


# Krylov settings per run.  "reference": what the reference hard-codes (1e-6, 1000 iterations).
# "tight": 1e-12 on both sides (reached reliably on these small Re = 100 cases; on harder ones -- Re = 1000, the channel --
# BiCGStab sits on the rounding floor at 1e-12 and can stagnate, and 1e-10 already moves the fields by ~1e-6, measured).
# The implicit Q2/Q1 momentum matrices are under-integrated
# (27-node mass matrix with 8 Gauss points, GQ.cpp:3490-3500) and badly conditioned, so two correct
# BiCGStab runs stopped at 1e-6 differ by ~1e-4 in the fields (even the iteration counts move with
# rounding); the 1e-6 field bar is therefore asserted at the tight setting, and at the reference
# setting the two are required to agree to the solver tolerance.
TIGHT = dict(cg_tol=1e-12, bicg_tol=1e-12, cg_maxit=5000, bicg_maxit=5000)


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("scheme,setting", [("explicit", "reference"), ("implicit", "reference"), ("implicit", "tight")])
def test_time_loop_parity(mesh, name, scheme, setting):
    """N whole time steps: fields rel-L2 <= 1e-6, CG iteration counts within +-2 of the oracle."""
    import fracstep
    import orc
    kw = TIGHT if setting == "tight" else {}
    path = mesh(name, **CASES[name])
    # tight: one oracle thread and a bitwise reproducible product run (coloured scatter, static slice assignment)
    q2_implicit_ref = scheme == "implicit" and setting == "reference" and CASES[name].get("NENv", 27) == 27
    repeatable = bool(kw) or q2_implicit_ref
    o = orc.Oracle(path, scheme, 1 if repeatable else 0, **kw)
    c = fracstep.Context(fracstep.Problem(o, scheme), fracstep.options(ebe_scatter=1, deterministic=1, **kw) if repeatable else None)
    n = 5
    o.run(n)
    reps = c.time_loop(n)
    U, P = c.fields()
    # Q2/Q1 implicit at the reference tolerance (1e-6 on a matrix with cond ~ 1e4): the measured floor of the comparison, not a
    # parity bar -- see tests/test_reference_pin.py::bars and profiles/r02_asis_spread.txt; the 1e-6 bar is asserted at "tight".
    # Five steps here (the fixtures: three): the spread grows with the step count, measured up to 4.2e-4 (U) on the B200.
    bar, barP = (1e-3, 5e-3) if q2_implicit_ref else (1e-6, 1e-6)
    assert rel(U, o.U) <= bar, rel(U, o.U)
    # interior field alone (~1e-3 of the lid speed this early).  At the reference tolerance the stopping rule |b - A u| <= 1e-6 |b| is
    # dominated by the lid rows of b, so the interior is only determined to ~1e-6 x 1e3 x cond: measured 1.4e-2 on the 4^3 mesh (B200)
    assert rel_free(U, o.U, o) <= (0.1 if q2_implicit_ref else 2e-5), rel_free(U, o.U, o)
    assert rel(P, o.P) <= barP, rel(P, o.P)
    assert all(r["cg_converged"] for r in reps)
    if scheme == "implicit":
        tol = kw.get("bicg_tol", 1e-6)
        assert all(v <= (tol if tol > 1e-9 else 10 * tol) for r in reps for v in r["bicg_true_residual"]), [r["bicg_true_residual"] for r in reps]
        if not kw:
            assert all(all(r["bicg_converged"]) for r in reps)
    ocg = o.arr("cgIters")
    if scheme == "explicit":
        assert [r["iter"] for r in reps] == list(o.arr("innerIters"))
        assert abs(sum(r["cg_iterations"] for r in reps) - ocg.sum()) <= 2 * len(ocg)
    else:
        slack = 8 if q2_implicit_ref else 2   # (as-is floor: the velocity fields differ by ~1e-4 there; measured 5 on the 4^3 mesh)
        for r, k in zip(reps, ocg):
            assert abs(r["cg_iterations"] - k) <= slack
        ob = o.arr("bicgIters").reshape(-1, 3)
        for r, k in zip(reps, ob):
            # BiCGStab iteration counts on these systems move with rounding (the oracle's own counts move by +-10 % with its
            # thread count, profiles/r02_asis_spread.txt): a sanity bound, not a parity claim
            assert all(abs(a - b) <= max(12, 0.35 * b) and a < 1000 for a, b in zip(r["bicg_iterations"], k)), (r["bicg_iterations"], k)
    mon = np.array(reps[-1]["monitor"])
    mp, NN = o.i("monPoint"), o.i("NN")
    ref = np.array([o.U[mp], o.U[NN + mp], o.U[2 * NN + mp], o.P[mp]])
    scale = np.array([np.abs(o.U).max()] * 3 + [np.abs(o.P).max()])
    assert np.all(np.abs(mon - ref) <= 10 * bar * scale)
    # maxAcc = max |dU| / dt is a pointwise quantity of the interior field: as-is floor measured 7.9e-3 (5^3 mesh, B200)
    assert abs(reps[-1]["maxAcc"] - o.d("maxAcc")) <= (3e-2 if q2_implicit_ref else 1e-5) * abs(o.d("maxAcc"))


@pytest.mark.parametrize("name", list(CASES))
def test_ebe_random_fields(mesh, name):
    """Element-by-element convective operator on O(1) random fields (apply form, explicit seam) and the
    assembled form (implicit seam, checked through the Jacobi diagonal and a BiCGStab solve)."""
    o, c = make_pair(mesh, name, "explicit")
    rng = np.random.default_rng(11)
    NN = o.i("NN")
    un, up = rng.standard_normal(3 * NN), rng.standard_normal(3 * NN)
    np.copyto(o.arr("Un", copy=False), un)
    np.copyto(o.arr("UnpHalf_prev", copy=False), up)
    c.upload("Un", un)
    c.upload("UnpHalf_prev", up)
    o.phase("bch_calculateMatrixA")
    c.call("bch_calculateMatrixAGPU")
    assert rel(c.download("R1"), o.arr("R1")) < 1e-11   # atomic scatter order varies run to run


def test_spmv_and_pcg_kernels(mesh):
    """Stand-alone kernels on the oracle's matrices: multi-RHS SpMV (1, 3, 6 vectors) and Jacobi-PCG."""
    import fracstep
    import orc
    import scipy.sparse as sp
    o = orc.Oracle(mesh("q2_c4", **CASES["q2_c4"]), "explicit")
    NN, NNp = o.i("NN"), o.i("NNp")
    rs, cl, v = o.arr("sparseMrowStarts"), o.arr("sparseMcol"), o.arr("sparseKvalue")
    K = sp.csr_matrix((v, cl, rs), shape=(NN, NN))
    rng = np.random.default_rng(7)
    for nvec in (1, 3, 6):
        x = rng.standard_normal((nvec, NN))
        y0 = rng.standard_normal((nvec, NN))
        y, _ = fracstep.csr_spmv(rs, cl, v, x, y0, alpha=-1.5, beta=0.5)
        ref = -1.5 * (K @ x.T).T + 0.5 * y0
        assert rel(y, ref) < 1e-13
    zr, zc, zv = o.arr("sparseZrowStarts"), o.arr("sparseZcol"), o.arr("sparseZvalue")
    Z = sp.csr_matrix((zv, zc, zr), shape=(NNp, NNp))
    b = rng.standard_normal(NNp)
    x, its, _ = fracstep.csr_pcg(zr, zc, zv, b, tol=1e-10, maxit=2000)
    assert np.linalg.norm(Z @ x - b) <= 1.01e-10 * np.linalg.norm(b)
    assert 0 < its < 2000
    # the single-reduction (Chronopoulos-Gear) recurrences: same iterates in exact arithmetic -> same count (+-2), same solution
    x2, its2, _ = fracstep.csr_pcg(zr, zc, zv, b, tol=1e-10, maxit=2000, variant=2)
    assert np.linalg.norm(Z @ x2 - b) <= 1.5e-10 * np.linalg.norm(b) and abs(its2 - its) <= 2 and rel(x2, x) < 1e-8
    x0 = rng.standard_normal(NNp)   # warm start
    xa, ia, _ = fracstep.csr_pcg(zr, zc, zv, b, x0=x0, tol=1e-8, maxit=2000)
    xb, ib, _ = fracstep.csr_pcg(zr, zc, zv, b, x0=x0, tol=1e-8, maxit=2000, variant=2)
    assert abs(ia - ib) <= 2 and rel(xb, xa) < 1e-6
    # empty right-hand side: zero iterations, x untouched (CUSP monitor semantics)
    for variant in (1, 2):
        x, its, _ = fracstep.csr_pcg(zr, zc, zv, np.zeros(NNp), variant=variant)
        assert its == 0 and not x.any()


@pytest.mark.parametrize("name,scheme", [("q2_c4", "explicit"), ("q1_u8", "implicit"), ("q2_u5", "explicit")])
def test_both_cg_recurrences_against_oracle(mesh, name, scheme):
    """cfd_options.cg_variant: the classical PCG (what the reference runs) and the single-reduction form the library picks for
    small / partitioned systems must both meet the parity bar: fields 1e-6, CG counts within +-2 of the oracle per solve."""
    import fracstep
    import orc
    path = mesh(name, **CASES[name])
    o = orc.Oracle(path, scheme)
    o.run(4)
    ocg = o.arr("cgIters")
    for variant in (1, 2):
        c = fracstep.Context(fracstep.Problem(o, scheme), fracstep.options(cg_variant=variant))
        reps = c.time_loop(4)
        U, P = c.fields()
        assert rel(U, o.U) <= 1e-6 and rel(P, o.P) <= (1e-6 if scheme == "implicit" else 1e-5), (variant, rel(U, o.U), rel(P, o.P))
        assert abs(sum(r["cg_iterations"] for r in reps) - int(ocg.sum())) <= 2 * len(ocg), (variant, [r["cg_iterations"] for r in reps], list(ocg))
        assert all(r["cg_converged"] for r in reps)
        c.close()


GOLDEN = ["q2_c4_explicit", "q2_c6_explicit", "q1_u8_explicit", "q1_u8_implicit", "q2_c4_implicit_tight"]


@pytest.mark.parametrize("name", GOLDEN)
def test_against_golden_fixtures(mesh, name):
    """Whole product pipeline (own .inp reader, node generation, step0, CUDA time loop) against the
    committed oracle fixtures (tests/golden/make_golden.py)."""
    import ast
    import os
    import fracstep
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    mk, opt = ast.literal_eval(str(g["mesh"])), ast.literal_eval(str(g["options"]))
    s = fracstep.Solver(mesh("gold_" + name, **mk), str(g["scheme"]), opts=fracstep.options(**opt) if opt else None)
    reps = s.ctx.time_loop(int(g["steps"]))
    U, P = s.ctx.fields()
    assert rel(U, g["U"]) <= 1e-6 and rel_free(U, g["U"], s) <= 2e-5
    assert rel(P, g["P"]) <= (1e-6 if "implicit" in name else 1e-5)   # explicit P: CG tolerance x cond(Z), see test_oracle
    if str(g["scheme"]) == "explicit":
        assert [r["iter"] for r in reps] == list(g["innerIters"])
        assert abs(sum(r["cg_iterations"] for r in reps) - int(g["cgIters"].sum())) <= 2 * len(g["cgIters"])
    else:
        assert all(abs(r["cg_iterations"] - k) <= 2 for r, k in zip(reps, g["cgIters"]))


@pytest.mark.parametrize("scheme,steps", [("explicit", 3), ("implicit", 1)])
def test_ne27000_full_size(mesh, scheme, steps):
    """BASELINE configs[0]/[1] at full size: the shipped 30^3 Taylor-Hood cavity (regenerated byte-identically),
    product pipeline on the GPU against the oracle."""
    import fracstep
    import orc
    visc = 0.01 if scheme == "explicit" else 0.001
    path = mesh("ne27000_%s" % scheme, n=30, cluster=2.0, viscosity=visc)
    o = orc.Oracle(path, scheme)
    s = fracstep.Solver(path, scheme)
    o.run(steps)
    reps = s.ctx.time_loop(steps)
    U, P = s.ctx.fields()
    if scheme == "explicit":
        assert rel(U, o.U) <= 1e-6 and rel(P, o.P) <= 1e-6
        assert [r["iter"] for r in reps] == list(o.arr("innerIters"))
        assert abs(sum(r["cg_iterations"] for r in reps) - int(o.arr("cgIters").sum())) <= 2 * len(o.arr("cgIters"))
    else:
        # first implicit step: u converges in a few BiCGStab iterations, v = w = 0; P to the CG tolerance
        assert rel(U, o.U) <= 1e-6 and rel(P, o.P) <= 1e-4
        assert abs(reps[0]["cg_iterations"] - int(o.arr("cgIters")[0])) <= 2


def test_momentum_solves_meet_their_true_residual(mesh):
    """|b - A x| <= tol |b| for the three momentum systems, recomputed on the host with the ORACLE's assembled A (same state, so the
    same matrix to rounding) from the right-hand sides and solutions the product holds -- and the status the product reports."""
    import scipy.sparse as sp
    import fracstep
    o, c = make_pair(mesh, "q2_c4", "implicit")
    o.run(2)
    c.time_loop(2)
    for nm in ("Uk_prev", "Pk_prev", "Pk_prevprev"):
        c.upload(nm, o.arr(nm))
    o.phase("gq_step1")
    c.call("gq_step1GPU")
    NN = o.i("NN")
    A = sp.csr_matrix((o.arr("sparseAvalue"), o.arr("sparseMcol"), o.arr("sparseMrowStarts")), shape=(NN, NN))
    b, x = c.download("R1"), c.download("Uk")
    st = c.krylov_status()
    for d in range(3):
        bd, xd = b[d * NN:(d + 1) * NN], x[d * NN:(d + 1) * NN]
        true = np.linalg.norm(bd - A @ xd) / np.linalg.norm(bd)
        assert true <= 1.0001e-6, (d, true)
        assert st["bicg_converged"][d] and abs(st["bicg_true_residual"][d] - true) <= 1e-9 and st["bicg_restarts"][d] == 0
    # an unattainable tolerance is REPORTED, not silently accepted (1e-15 IS reached on this small system: measured 7e-16;
    # 1e-18 is below the rounding of b - A x itself)
    c2 = fracstep.Context(fracstep.Problem(o, "implicit"), fracstep.options(bicg_tol=1e-18, bicg_maxit=3000))
    for nm in ("Uk_prev", "Pk_prev", "Pk_prevprev"):
        c2.upload(nm, o.arr(nm))
    c2.call("gq_step1GPU")
    st = c2.krylov_status()
    assert not any(st["bicg_converged"]) and all(0 <= v < 1e-10 for v in st["bicg_true_residual"]), st
    x2 = c2.download("Uk")
    assert np.isfinite(x2).all() and rel(x2, x) < 1e-2   # (the 1e-6 solve itself is only good to 1e-6 x cond ~ 1e-3)


HEADLINE = {
    # BASELINE.json configs[3]/[4] element and scheme (Q1/Q1, implicit) at 128^3 -- the largest mesh the CPU oracle steps in seconds
    "q1_128_implicit": dict(n=128, NENv=8, scheme="implicit", steps=2),
    # configs[3] "EBE" element (Q2/Q1), explicit, 64^3 = 2.1 M velocity nodes
    "q2_64_explicit": dict(n=64, NENv=27, scheme="explicit", steps=1),
}


@pytest.mark.parametrize("name", list(HEADLINE))
def test_headline_config_fields_against_oracle(mesh, name):
    """Fields of the benchmark configurations themselves, not only of small meshes: rel-L2 <= 1e-6, CG counts +-2.
    (Q1/Q1 is oracle-only parity by construction: the reference's NENv == 8 shape-function branch is commented out,
    blascoCodinaHuerta.cpp:2303-2305, so the reference program cannot run these meshes.)"""
    import fracstep
    import orc
    h = HEADLINE[name]
    path = mesh("headline_" + name, n=h["n"], NENv=h["NENv"], full_precision=True)
    o = orc.Oracle(path, h["scheme"])
    s = fracstep.Solver(path, h["scheme"])
    o.run(h["steps"])
    reps = s.ctx.time_loop(h["steps"])
    U, P = s.ctx.fields()
    assert rel(U, o.U) <= 1e-6, rel(U, o.U)
    # pressure: both sides stop CG at 1e-6 |b|; the field error is that times cond(Z) (oracle-vs-oracle: test_oracle.py)
    assert rel(P, o.P) <= (1e-6 if h["scheme"] == "implicit" else 1e-5), rel(P, o.P)
    ocg = list(o.arr("cgIters"))
    if h["scheme"] == "explicit":
        assert [r["iter"] for r in reps] == list(o.arr("innerIters"))
        assert abs(sum(r["cg_iterations"] for r in reps) - sum(ocg)) <= 2 * len(ocg)
    else:
        assert all(abs(r["cg_iterations"] - k) <= 2 for r, k in zip(reps, ocg)), ([r["cg_iterations"] for r in reps], ocg)
        assert all(all(r["bicg_converged"]) and max(r["bicg_true_residual"]) <= 1e-6 for r in reps)
    s.close()
    o.close()


def test_large_mesh_properties():
    """Size-independent properties on a mesh the oracle would take minutes for (64^3 Q1/Q1, implicit):
    Dirichlet values are reproduced to the solver tolerance, the pressure solve meets its stopping rule, the velocity
    systems are solved to their tolerance (checked through a host SpMV with the product's own CSR arrays)."""
    import scipy.sparse as sp
    import fracstep
    s = fracstep.Solver(scheme="implicit", cavity=dict(nx=64, NENv=8))
    c = s.ctx
    reps = c.time_loop(2)
    assert all(0 < r["cg_iterations"] < 1000 for r in reps)
    U, P = c.fields()
    NN, NNp = s.i("NN"), s.i("NNp")
    bc = s.arr("BCvelNodes").reshape(-1, 2)
    vals = s.arr("BCstr").reshape(-1, 3)
    for d in range(3):
        # identity rows go through BiCGStab like every other row (GQ.cpp:4614-4643): exact to the solver tolerance
        assert np.allclose(U[d * NN + bc[:, 0]], vals[bc[:, 1], d], rtol=0, atol=1e-5)
    Z = sp.csr_matrix((s.arr("sparseZvalue"), s.arr("sparseZcol"), s.arr("sparseZrowStarts")), shape=(NNp, NNp))
    R2, Pdiff = c.download("R2"), c.download("Pdiff")
    assert np.linalg.norm(Z @ Pdiff - R2) <= 1.05e-6 * np.linalg.norm(R2)
    assert R2[s.i("zeroPressureNode")] == 0.0


@pytest.mark.parametrize("peer_memory", [True, False])
def test_two_gpu_partitioned_run(peer_memory, tmp_path):
    """Multi-GPU: the partitioned run (element blocks, halo swap, all-reduced dot products) reproduces the single-GPU
    fields -- with the peer-memory persistent PCG and with the host-driven NCCL PCG.  Needs 2 GPUs on the box."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ)
    if not peer_memory:
        env["NO_P2P"] = "1"
    else:
        env["CG_VARIANTS"] = "1,2"   # both peer-memory PCG kernels: classical and single-reduction
    port = 29600 + os.getpid() % 300 + (7 if peer_memory else 0)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", str(port), os.path.join(root, "tests", "mgpu_check.py")], cwd=root, env=env,
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    lines = [l for l in out.stdout.split("\n") if "relU" in l]
    assert len(lines) == (8 if peer_memory else 4), out.stdout[-3000:]
    for l in lines:
        relU, relP = float(l.split("relU")[1].split()[0]), float(l.split("relP")[1].split()[0])
        assert relU <= 1e-6 and relP <= 1e-5, l


def test_channel_standin_implicit(mesh):
    """BASELINE configs[2] stand-in (straight channel, uniform inlet, free outlet, pin at the outlet): implicit scheme
    against the oracle.  The anisotropic channel is the hardest case for Jacobi-BiCGStab on the reference's under-integrated
    Q2 momentum matrix: 1e-12 is not attainable (measured: the solver stagnates for thousands of iterations, the count varies
    by +-50 % from run to run with the summation order -- in the OpenMP oracle too -- and can end in the rho = 0 breakdown),
    so the momentum tolerance here is 1e-9, which both sides reach."""
    import fracstep
    import orc
    path = mesh("chan_gpu", kind="channel", nx=12, ny=8, nz=8, full_precision=True)
    opt = dict(cg_tol=1e-12, bicg_tol=1e-9, cg_maxit=5000, bicg_maxit=5000)
    o = orc.Oracle(path, "implicit", 1, **opt)
    s = fracstep.Solver(path, "implicit", opts=fracstep.options(**opt))
    o.run(3)
    reps = s.ctx.time_loop(3)
    U, P = s.ctx.fields()
    assert np.isfinite(U).all() and np.isfinite(P).all() and np.isfinite(o.U).all()
    assert all(max(r["bicg_iterations"]) < 5000 for r in reps) and o.arr("bicgIters").max() < 5000
    assert all(all(r["bicg_converged"]) and max(r["bicg_true_residual"]) <= 1e-9 for r in reps), reps
    assert rel(U, o.U) <= 1e-6 and rel_free(U, o.U, o) <= 2e-5 and rel(P, o.P) <= 1e-5


def test_solver_main_drop_in(mesh, tmp_path):
    """The whole-program drop-in: ProblemName.txt -> <name>.inp -> monitor table -> <name>.dat, like the reference's
    executable (BCH.cpp:383-520).  The .dat written at t_final equals the oracle's createTecplot output to 1e-6."""
    import shutil
    import subprocess
    import sys
    import orc
    src = mesh("main_q2", n=4, cluster=2.0, t_final=0.003)     # 3 steps of dt = 1e-3
    shutil.copy(src, tmp_path / "cavity.inp")
    (tmp_path / "ProblemName.txt").write_text("cavity\n")
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = "import sys; sys.path.insert(0, %r); import fracstep; fracstep.solver_main(%r, 'explicit')" % (
        os.path.join(root, "cfd-with-cuda_b200"), str(tmp_path))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    rows = [l for l in out.stdout.split("\n") if l.strip() and l.split()[0].isdigit()]
    assert len(rows) == 3 and [int(r.split()[0]) for r in rows] == [1, 2, 3]          # one monitor line per time step
    assert "Time step  Iter     Time       u_monitor" in out.stdout
    o = orc.Oracle(src, "explicit")
    o.run(3)
    o.write_tecplot(str(tmp_path / "oracle.dat"), "cavity")
    a = open(tmp_path / "cavity.dat").read().split("\n")
    b = open(tmp_path / "oracle.dat").read().split("\n")
    assert a[:3] == b[:3] and len(a) == len(b)
    NN = o.i("NN")
    A = np.array([l.split() for l in a[3:3 + NN]], float)
    B = np.array([l.split() for l in b[3:3 + NN]], float)
    assert np.array_equal(A[:, :3], B[:, :3])                                           # coordinates: same bytes
    assert rel(A[:, 3:6], B[:, 3:6]) <= 1e-6 and rel(A[:, 6], B[:, 6]) <= 1e-5
    assert a[3 + NN:] == b[3 + NN:]                                                     # connectivity: same bytes


def test_reference_gpu_build_switches(mesh):
    """The reference's own GPU build differs from its CPU build: CG warm start from the previous
    Pdot (BCH.cu:418-451) and maxAcc without the 1/dt factor (BCH.cu:107-117).  Both are switches on both sides."""
    import fracstep
    import orc
    path = mesh("q2_c4", **CASES["q2_c4"])
    o = orc.Oracle(path, "explicit", cg_warm=1, maxacc_rule=1)
    c = fracstep.Context(fracstep.Problem(o, "explicit"), fracstep.options(cg_warm_start=1, maxacc_rule=1))
    o.run(4)
    reps = c.time_loop(4)
    U, P = c.fields()
    assert rel(U, o.U) <= 1e-6 and rel(P, o.P) <= 1e-6
    assert abs(sum(r["cg_iterations"] for r in reps) - int(o.arr("cgIters").sum())) <= 2 * len(o.arr("cgIters"))
    assert abs(reps[-1]["maxAcc"] - o.d("maxAcc")) <= 1e-5 * o.d("maxAcc")
    # warm start must actually save iterations compared with the zero start
    o0 = orc.Oracle(path, "explicit")
    o0.run(4)
    assert o.arr("cgIters").sum() < o0.arr("cgIters").sum()


def test_no_pressure_pin_and_restart(mesh, tmp_path):
    """Edge cases of the input grammar: pin node 0 (1-based; 0-based -1 = "no pin", and node 1 = 0-based 0 is never
    pinned either, BCH.cpp:4172) and a restart from <name>_restart.dat (isRestart 1, BCH.cpp:4214-4242)."""
    import shutil
    import fracstep
    import orc
    src = open(mesh("q1_u8", **CASES["q1_u8"])).read()
    lines = src.split("\n")
    k = lines.index("Node number where pressure is taken to be zero")
    for pin in ("0", "1"):
        lines[k + 1] = pin
        p = tmp_path / ("nopin%s.inp" % pin)
        p.write_text("\n".join(lines))
        o = orc.Oracle(str(p), "implicit")
        s = fracstep.Solver(str(p), "implicit")
        assert s.i("pinLocal") == -1
        o.run(2)
        reps = s.ctx.time_loop(2)
        U, P = s.ctx.fields()
        # without a pin the pressure is defined up to a constant: compare after removing the mean
        assert rel(U, o.U) <= 1e-6 and rel(P - P.mean(), o.P - o.P.mean()) <= 1e-4
        assert all(abs(r["cg_iterations"] - k) <= 2 for r, k in zip(reps, o.arr("cgIters")))
    # restart: run 2 steps, write the Tecplot file, restart a new solver from it and continue 2 steps
    base = mesh("q2_c4", **CASES["q2_c4"])
    a = fracstep.Solver(base, "explicit")
    a.ctx.time_loop(2)
    rdir = tmp_path / "restart"
    rdir.mkdir()
    a.write_tecplot(str(rdir / "cav_restart.dat"), "cav")
    (rdir / "cav.inp").write_text(open(base).read().replace("isRestart: 0", "isRestart: 1"))
    b = fracstep.Solver(str(rdir / "cav.inp"), "explicit")
    Ua, Pa = a.ctx.fields()
    Ub, Pb = b.ctx.fields()
    assert np.allclose(Ub, Ua, rtol=2e-11, atol=1e-300) and np.allclose(Pb, Pa, rtol=2e-11, atol=1e-300)   # 11 digits in the file
    o = orc.Oracle(str(rdir / "cav.inp"), "explicit")
    o.read_restart(str(rdir / "cav_restart.dat"))
    o.run(2)
    b.ctx.time_loop(2)
    Ub, Pb = b.ctx.fields()
    assert rel(Ub, o.U) <= 1e-6 and rel(Pb, o.P) <= 1e-5


@pytest.mark.parametrize("name", ["q2_u5", "q1_u8"])
def test_coloured_scatter_is_reproducible(mesh, name):
    """ebe_scatter = 1: element colouring (valid even on the odd meshes where the reference's colouring races), one launch
    per colour, plain stores.  Same result as the atomic scatter to rounding, and bitwise identical from run to run."""
    import fracstep
    import orc
    path = mesh(name, **CASES[name])
    o = orc.Oracle(path, "explicit")
    rng = np.random.default_rng(5)
    NN = o.i("NN")
    un, up = rng.standard_normal(3 * NN), rng.standard_normal(3 * NN)
    out = []
    for mode in (1, 1, 0):
        c = fracstep.Context(fracstep.Problem(o, "explicit"), fracstep.options(ebe_scatter=mode))
        c.upload("Un", un)
        c.upload("UnpHalf_prev", up)
        c.call("bch_calculateMatrixAGPU")
        out.append(c.download("R1"))
    assert np.array_equal(out[0], out[1])
    assert rel(out[0], out[2]) < 1e-13
    for scheme in ("explicit", "implicit"):
        oo = orc.Oracle(path, scheme)
        cc = fracstep.Context(fracstep.Problem(oo, scheme), fracstep.options(ebe_scatter=1))
        oo.run(3)
        cc.time_loop(3)
        U, P = cc.fields()
        bar = 1e-6 if (scheme == "explicit" or name.startswith("q1")) else 5e-3
        assert rel(U, oo.U) <= bar and rel(P, oo.P) <= bar


def test_pcg_against_library_solver(mesh):
    """Independent cross-check of the persistent PCG kernel: the same system solved by a plain cuSPARSE + cuBLAS CG
    (baseline/lib_pcg.cu, the reference's GPU call sequence on current APIs) -- same iteration count, same solution."""
    import ctypes as C
    import os
    import subprocess
    import fracstep
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    try:
        subprocess.check_call(["make", "-C", os.path.join(root, "baseline")], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        LB = C.CDLL(os.path.join(root, "baseline", "liblibpcg.so"))
    except Exception as e:   # cuSPARSE / cuBLAS not available
        pytest.skip("library baseline not buildable: %s" % e)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    LB.lib_pcg.argtypes = [C.c_int, C.c_int, ip, ip, dp, dp, dp, dp, C.c_double, C.c_int, ip, dp]
    s = fracstep.Solver(mesh("q2_c6_lib", n=6, cluster=2.0), "explicit")
    s.ctx.time_loop(2)
    R2, x_ours, its_ours = np.ascontiguousarray(s.ctx.download("R2")), s.ctx.download("Pdot"), s.ctx.last_cg()
    rs = np.ascontiguousarray(s.arr("sparseZrowStarts"), np.int32)
    cl = np.ascontiguousarray(s.arr("sparseZcol"), np.int32)
    v = np.ascontiguousarray(s.arr("sparseZvalue"), np.float64)
    n = len(rs) - 1
    rows = np.repeat(np.arange(n), np.diff(rs))
    dinv = np.ones(n)
    dinv[rows[rows == cl]] = 1.0 / v[rows == cl]
    x = np.zeros(n)
    its, ms = C.c_int(), C.c_double()
    assert LB.lib_pcg(n, len(cl), rs.ctypes.data_as(ip), cl.ctypes.data_as(ip), v.ctypes.data_as(dp), dinv.ctypes.data_as(dp),
                      R2.ctypes.data_as(dp), x.ctypes.data_as(dp), 1e-6, 1000, C.byref(its), C.byref(ms)) == 0
    assert abs(its.value - its_ours) <= 1
    assert rel(x_ours, x) <= 1e-9


@pytest.mark.parametrize("name", list(CASES))
def test_matrix_free_operators_equal_the_assembled_ones(mesh, name):
    """cfd_options.operator_path / cfd_operator_apply: K, G, G^T and the FE Laplacian applied element by element from the integrals of
    step0 (BCH.cpp:3127-3315, GQ.cpp:3677-3772) equal the assembled (Sell SpMV) products to summation-order rounding -- and both equal
    scipy on the oracle's CSR arrays."""
    import scipy.sparse as sp
    import fracstep
    import orc
    rng = np.random.default_rng(3)
    for scheme in ("explicit", "implicit"):
        o = orc.Oracle(mesh(name, **CASES[name]), scheme)
        c = fracstep.Context(fracstep.Problem(o, scheme), None)
        NN, NNp = o.i("NN"), o.i("NNp")
        K = sp.csr_matrix((o.arr("sparseKvalue"), o.arr("sparseMcol"), o.arr("sparseMrowStarts")), shape=(NN, NN))
        G = [sp.csr_matrix((o.arr("sparseG%dvalue" % d), o.arr("sparseGcol"), o.arr("sparseGrowStarts")), shape=(NN, NNp)) for d in (1, 2, 3)]
        xv, xp = rng.standard_normal(3 * NN), rng.standard_normal(NNp)
        ref = {"K": np.concatenate([K @ xv[d * NN:(d + 1) * NN] for d in range(3)]),
               "G": np.concatenate([G[d] @ xp for d in range(3)]),
               "GT": sum(G[d].T @ xv[d * NN:(d + 1) * NN] for d in range(3))}
        ops = ["K", "G", "GT"]
        if scheme == "implicit":
            Z = sp.csr_matrix((o.arr("sparseZvalue"), o.arr("sparseZcol"), o.arr("sparseZrowStarts")), shape=(NNp, NNp))
            ref["Z"] = Z @ xp
            ops.append("Z")
        for op in ops:
            x = xp if op in ("G", "Z") else xv
            y1, _, _ = c.operator_apply(op, x, path=1)
            assert rel(y1, ref[op]) < 1e-12, (scheme, op, rel(y1, ref[op]))
            if not (op == "K" and scheme == "implicit"):   # (implicit contexts keep M + K and M, not K alone)
                y0, _, _ = c.operator_apply(op, x, path=0)
                assert rel(y0, ref[op]) < 1e-13, (scheme, op, rel(y0, ref[op]))
        c.close()


@pytest.mark.parametrize("name", list(CASES))
def test_matrix_free_time_loop(mesh, name):
    """The explicit time loop with operator_path = 1 (step 1 = ONE element pass for -C(U)U - K U - G P; G^T, G Pdot and K Acc element by
    element): same fields as the oracle (1e-6), same inner iterations and CG counts, and equal to the assembled path to rounding."""
    import fracstep
    import orc
    o = orc.Oracle(mesh(name, **CASES[name]), "explicit")
    o.run(4)
    out = {}
    for path in (0, 1):
        c = fracstep.Context(fracstep.Problem(o, "explicit"), fracstep.options(operator_path=path))
        reps = c.time_loop(4)
        out[path] = c.fields()
        U, P = out[path]
        assert rel(U, o.U) <= 1e-6 and rel(P, o.P) <= 1e-5, (path, rel(U, o.U), rel(P, o.P))
        assert [r["iter"] for r in reps] == list(o.arr("innerIters"))
        assert abs(sum(r["cg_iterations"] for r in reps) - int(o.arr("cgIters").sum())) <= 2 * len(o.arr("cgIters"))
        c.close()
    assert rel(out[1][0], out[0][0]) <= 1e-9 and rel(out[1][1], out[0][1]) <= 1e-6   # (P: CG stopped at 1e-6 on both)


def test_shared_memory_resident_pcg_equals_the_streaming_kernel(mesh, monkeypatch):
    """Small pressure systems run a RESIDENT kernel (one CTA per SM for the whole solve): k_pcg_res (row blocks per CTA: values in
    shared memory, the vector entries a block touches staged per sweep, vectors in registers) or, if a block does not fit, k_pcg_smem
    (values in shared memory only); larger ones the streaming kernel k_pcg_t.  Same Jacobi-PCG: on the same system all give the same
    iteration counts and the same solution to rounding -- checked through the explicit time loop of a mesh small enough for the resident
    kernels, with each of the three, cold and warm started."""
    import fracstep
    import orc
    path = mesh("q2_c6_smem", n=6, cluster=2.0)
    o = orc.Oracle(path, "explicit")
    o.run(3)
    out = {}
    for kernel in ("res", "smem", "stream"):
        monkeypatch.delenv("CFD_NO_PCG_SMEM", raising=False)
        monkeypatch.delenv("CFD_NO_PCG_RES", raising=False)
        if kernel == "smem":
            monkeypatch.setenv("CFD_NO_PCG_RES", "1")
        if kernel == "stream":
            monkeypatch.setenv("CFD_NO_PCG_SMEM", "1")
        c = fracstep.Context(fracstep.Problem(o, "explicit"), None)
        reps = c.time_loop(3)
        out[kernel] = (c.fields(), [r["cg_iterations"] for r in reps])
        U, P = out[kernel][0]
        assert rel(U, o.U) <= 1e-6 and rel(P, o.P) <= 1e-5, kernel
        assert all(r["cg_converged"] for r in reps)
        c.close()
        # warm start (the reference's GPU build keeps the previous Pdot, BCH.cu:418-451): r = b - Z x through the same sweep
        c = fracstep.Context(fracstep.Problem(o, "explicit"), fracstep.options(cg_warm_start=1))
        reps = c.time_loop(3)
        out[kernel + "_warm"] = (c.fields(), [r["cg_iterations"] for r in reps])
        assert all(r["cg_converged"] for r in reps)
        c.close()
    for kernel in ("res", "smem"):
        for sfx in ("", "_warm"):
            a, b = out[kernel + sfx], out["stream" + sfx]
            assert max(abs(x - y) for x, y in zip(a[1], b[1])) <= 1, (kernel, sfx, a[1], b[1])
            assert rel(a[0][0], b[0][0]) <= 1e-9 and rel(a[0][1], b[0][1]) <= 2e-6, (kernel, sfx)   # (P: CG stopped at 1e-6 on both)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("scheme", ["explicit", "implicit"])
def test_step0_on_the_device(mesh, name, scheme):
    """cfd_options.step0_device = 1: the host builds the patterns only; element integrals, lumped mass,
    Z (explicit: G^T Md^-1 G as a device sparse product, implicit: FE Laplacian), pin and the conversion into the device layout
    run on the GPU.  Checked three ways: (1) the assembled operators K, G, G^T (and the implicit Z) applied to random vectors equal the
    oracle's (= the reference's) CSR arrays to rounding; (2) the time loop gives the fields of the host-step0 context to rounding
    (this covers M, the lumped masses and the explicit Z); (3) and the oracle's fields to the parity bar."""
    import scipy.sparse as sp
    import fracstep
    import orc
    kw = TIGHT if scheme == "implicit" else {}
    path = mesh(name, **CASES[name])
    o = orc.Oracle(path, scheme, 1, **kw)
    rng = np.random.default_rng(7)
    NN, NNp = o.i("NN"), o.i("NNp")
    xv, xp = rng.standard_normal(3 * NN), rng.standard_normal(NNp)
    G = [sp.csr_matrix((o.arr("sparseG%dvalue" % d), o.arr("sparseGcol"), o.arr("sparseGrowStarts")), shape=(NN, NNp)) for d in (1, 2, 3)]
    ref = {"G": np.concatenate([G[d] @ xp for d in range(3)]), "GT": sum(G[d].T @ xv[d * NN:(d + 1) * NN] for d in range(3))}
    if scheme == "explicit":
        K = sp.csr_matrix((o.arr("sparseKvalue"), o.arr("sparseMcol"), o.arr("sparseMrowStarts")), shape=(NN, NN))
        ref["K"] = np.concatenate([K @ xv[d * NN:(d + 1) * NN] for d in range(3)])
    else:
        Z = sp.csr_matrix((o.arr("sparseZvalue"), o.arr("sparseZcol"), o.arr("sparseZrowStarts")), shape=(NNp, NNp))
        ref["Z"] = Z @ xp
    n = 3
    o.run(n)
    out = {}
    for dev in (1, 0):
        s = fracstep.Solver(path, scheme, opts=fracstep.options(step0_device=dev, ebe_scatter=1, deterministic=1, **kw))
        if dev:
            assert len(s.arr("sparseKvalue")) == 0 and len(s.arr("sparseZvalue")) == 0 and len(s.arr("sparseG1value")) == 0   # nothing assembled on the host
            for op in ref:
                y, _, _ = s.ctx.operator_apply(op, xp if op in ("G", "Z") else xv, path=0)
                assert rel(y, ref[op]) < 1e-12, (scheme, op, rel(y, ref[op]))
        reps = s.ctx.time_loop(n)
        out[dev] = s.ctx.fields()
        assert all(r["cg_converged"] for r in reps)
        U, P = out[dev]
        assert rel(U, o.U) <= 1e-6 and rel(P, o.P) <= 1e-6, (dev, rel(U, o.U), rel(P, o.P))
        s.close()
    assert rel(out[1][0], out[0][0]) <= 1e-7 and rel(out[1][1], out[0][1]) <= 1e-6, (rel(out[1][0], out[0][0]), rel(out[1][1], out[0][1]))


@pytest.mark.gpu
def test_duct_standin_implicit(mesh):
    """configs[2] stand-in with a prescribed outflow (meshgen.duct): the implicit scheme converges on it, so the CUDA path is compared
    field by field with the oracle -- both sides iterated to 1e-12, the setting at which the 1e-6 bar is meaningful on the
    reference's under-integrated Q2/Q1 matrices (at the reference's own 1e-6 the oracle differs from ITSELF by 1.1e-3 in U on this
    case when its thread count changes, and by 1.1e-2 from the converged fields: measured, /tests/README.md)."""
    import fracstep
    import orc
    path = mesh("duct_gpu", kind="duct", nx=12, ny=8, nz=8)
    o = orc.Oracle(path, "implicit", 1, **TIGHT)
    s = fracstep.Solver(path, "implicit", opts=fracstep.options(ebe_scatter=1, deterministic=1, **TIGHT))
    n = 4
    o.run(n)
    reps = s.ctx.time_loop(n)
    U, P = s.ctx.fields()
    assert rel(U, o.U) <= 1e-6 and rel(P, o.P) <= 1e-6, (rel(U, o.U), rel(P, o.P))
    assert all(r["cg_converged"] for r in reps)
    assert all(v <= 1e-11 for r in reps for v in r["bicg_true_residual"]), [r["bicg_true_residual"] for r in reps]
    for r, k in zip(reps, o.arr("cgIters")):
        assert abs(r["cg_iterations"] - k) <= 5
    assert max(max(r["bicg_iterations"]) for r in reps) < 5000
    s.close()
    o.close()
