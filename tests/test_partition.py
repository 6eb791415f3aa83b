"""CPU: element-block partition for multi-GPU runs.

test_partition_consistency: every rank's owned rows equal the global operators' rows, every node is owned
exactly once, send lists match the receivers' ghost ranges.
test_halo_exchange_gloo: two processes (torch.distributed, gloo, world_size 2) run the exchange protocol the
CUDA library runs over NCCL -- owners send their interface values in plan order, receivers store them into
their ghost ranges -- and check a distributed K*x, G^T*d and Z*p against the global products."""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def csr(s, v, which, shape):
    return sp.csr_matrix((s.arr(v), s.arr("sparse%scol" % which), s.arr("sparse%srowStarts" % which)), shape=shape)


@pytest.mark.parametrize("cav,scheme,W", [(dict(nx=6, NENv=27), "explicit", 2), (dict(nx=6, NENv=27), "implicit", 3),
                                           (dict(nx=8, NENv=8), "explicit", 4), (dict(nx=5, ny=4, nz=7, NENv=27), "explicit", 3),
                                           (dict(nx=8, NENv=8), "implicit", 8),
                                           # blocks that are not whole element layers; one element per rank
                                           (dict(nx=3, ny=3, nz=2, NENv=27), "explicit", 5), (dict(nx=2, ny=2, nz=2, NENv=8), "implicit", 8)])
def test_partition_consistency(cav, scheme, W):
    import fracstep
    g = fracstep.Solver(scheme=scheme, cavity=cav, init_gpu=False)
    NN, NNp = g.i("NN"), g.i("NNp")
    Kg, Zg, Gg = csr(g, "sparseKvalue", "M", (NN, NN)), csr(g, "sparseZvalue", "Z", (NNp, NNp)), csr(g, "sparseG2value", "G", (NN, NNp))
    locs = [fracstep.Solver(scheme=scheme, cavity=cav, init_gpu=False, world=W, rank=r) for r in range(W)]
    ownedV, ownedP = np.zeros(NN, int), np.zeros(NNp, int)
    for r, s in enumerate(locs):
        ng, pg = s.arr("part_nodeGlobal"), s.arr("part_presGlobal")
        NNl, NPl, NNo, NPo = s.i("NN"), s.i("NNp"), s.i("NN_owned"), s.i("NNp_owned")
        ownedV[ng[:NNo]] += 1
        ownedP[pg[:NPo]] += 1
        Pn = sp.csr_matrix((np.ones(NNl), (np.arange(NNl), ng)), shape=(NNl, NN))
        Pp = sp.csr_matrix((np.ones(NPl), (np.arange(NPl), pg)), shape=(NPl, NNp))
        assert abs(csr(s, "sparseKvalue", "M", (NNl, NNl))[:NNo] @ Pn - Kg[ng[:NNo]]).max() < 1e-13 * abs(Kg).max()
        assert abs(csr(s, "sparseZvalue", "Z", (NPl, NPl))[:NPo] @ Pp - Zg[pg[:NPo]]).max() < 1e-12 * abs(Zg).max()
        assert abs(csr(s, "sparseG2value", "G", (NNl, NPl))[:NNo] @ Pp - Gg[ng[:NNo]]).max() < 1e-13 * abs(Gg).max()
        if scheme == "explicit":
            assert np.allclose(s.arr("MdInv")[:NNo], g.arr("MdInv")[ng[:NNo]], rtol=1e-13)
        gb = dict(g.arr("BCvelNodes").reshape(-1, 2))
        bl = s.arr("BCvelNodes").reshape(-1, 2)
        assert all(gb.get(ng[n], -1) == w for n, w in bl[bl[:, 0] < NNo])
        peers, svp, svi = s.arr("peerRank"), s.arr("sendVptr"), s.arr("sendVidx")
        spp, spi = s.arr("sendPptr"), s.arr("sendPidx")
        for k, q in enumerate(peers):
            t = locs[q]
            kk = list(t.arr("peerRank")).index(r)
            off, cnt = t.arr("recvVoff")[kk], t.arr("recvVcnt")[kk]
            assert np.array_equal(ng[svi[svp[k]:svp[k + 1]]], t.arr("part_nodeGlobal")[off:off + cnt])
            off, cnt = t.arr("recvPoff")[kk], t.arr("recvPcnt")[kk]
            assert np.array_equal(pg[spi[spp[k]:spp[k + 1]]], t.arr("part_presGlobal")[off:off + cnt])
            if svp[k + 1] > svp[k]:   # peer-memory pushes: my block inside the peer's ghost range
                assert s.arr("remoteVoff")[k] == t.arr("recvVoff")[kk]
            if spp[k + 1] > spp[k]:
                assert s.arr("remotePoff")[k] == off
    assert (ownedV == 1).all() and (ownedP == 1).all()
    pins = [s.i("pinLocal") for s in locs]
    owners = [r for r, s in enumerate(locs) if 0 <= pins[r] < s.i("NNp_owned")]
    assert len(owners) == 1 and locs[owners[0]].arr("part_presGlobal")[pins[owners[0]]] == g.i("zeroPressureNode")


def _worker(rank, world, port, q, scheme="explicit", cav=None):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, os.path.join(ROOT, "cfd-with-cuda_b200"))
    import torch
    import torch.distributed as dist
    import fracstep
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cav = cav or dict(nx=6, NENv=27)
        s = fracstep.Solver(scheme=scheme, cavity=cav, init_gpu=False, world=world, rank=rank)
        g = fracstep.Solver(scheme=scheme, cavity=cav, init_gpu=False)
        NN, NNp = g.i("NN"), g.i("NNp")
        NNl, NPl, NNo, NPo = s.i("NN"), s.i("NNp"), s.i("NN_owned"), s.i("NNp_owned")
        ng, pg = s.arr("part_nodeGlobal"), s.arr("part_presGlobal")
        peers = s.arr("peerRank")

        def halo(vec, kind):
            """ghost entries <- owners (the protocol of device.cu: halo())."""
            ptr, idx = s.arr("send%sptr" % kind), s.arr("send%sidx" % kind)
            off, cnt = s.arr("recv%soff" % kind), s.arr("recv%scnt" % kind)
            reqs, bufs = [], []
            for j, p in enumerate(peers):
                snd = torch.from_numpy(np.ascontiguousarray(vec[idx[ptr[j]:ptr[j + 1]]]))
                if snd.numel():
                    reqs.append(dist.isend(snd, int(p)))
                if cnt[j]:
                    rb = torch.empty(int(cnt[j]), dtype=torch.float64)
                    bufs.append((rb, int(off[j])))
                    reqs.append(dist.irecv(rb, int(p)))
            for r in reqs:
                r.wait()
            for rb, o in bufs:
                vec[o:o + rb.numel()] = rb.numpy()

        rng = np.random.default_rng(3)
        xg, pgl = rng.standard_normal(NN), rng.standard_normal(NNp)
        x = np.full(NNl, np.nan)
        x[:NNo] = xg[ng[:NNo]]          # owned values only; ghosts must come from the exchange
        halo(x, "V")
        assert np.array_equal(x, xg[ng])
        pv = np.full(NPl, np.nan)
        pv[:NPo] = pgl[pg[:NPo]]
        halo(pv, "P")
        assert np.array_equal(pv, pgl[pg])
        K = csr(s, "sparseKvalue", "M", (NNl, NNl))[:NNo]
        Kg = csr(g, "sparseKvalue", "M", (NN, NN))
        assert np.allclose(K @ x, (Kg @ xg)[ng[:NNo]], rtol=1e-12, atol=1e-13)
        Z = csr(s, "sparseZvalue", "Z", (NPl, NPl))[:NPo]
        Zg = csr(g, "sparseZvalue", "Z", (NNp, NNp))
        assert np.allclose(Z @ pv, (Zg @ pgl)[pg[:NPo]], rtol=1e-11, atol=1e-12)
        # G^T d on the owned pressure rows needs the full local G (ghost velocity rows included)
        G = csr(s, "sparseG1value", "G", (NNl, NPl))
        Gg = csr(g, "sparseG1value", "G", (NN, NNp))
        assert np.allclose((G.T @ x)[:NPo], (Gg.T @ xg)[pg[:NPo]], rtol=1e-11, atol=1e-12)
        # dot product over owned entries + allreduce == global dot
        t = torch.tensor([float(x[:NNo] @ x[:NNo])], dtype=torch.float64)
        dist.all_reduce(t)
        assert abs(t.item() - xg @ xg) < 1e-10 * (xg @ xg)
        q.put((rank, "ok"))
    except Exception as e:   # surface the failure to the parent
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,scheme,cav", [(2, "explicit", dict(nx=6, NENv=27)),
                                               # three ranks (the middle one has two peers), one halo layer, Q1/Q1, non-cubic
                                               (3, "implicit", dict(nx=5, ny=4, nz=7, NENv=8))])
def test_halo_exchange_gloo(world, scheme, cav):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + 7 * world) % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, scheme, cav)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(60)
    assert all(m == "ok" for _, m in res), res


@pytest.mark.parametrize("scheme", ["explicit", "implicit"])
def test_restart_file_on_a_partition(mesh, tmp_path, scheme):
    """isRestart 1 on a partitioned run (readRestartFile, BCH.cpp:4214-4242): every rank reads the restart file of the whole mesh and
    keeps the rows of its own nodes -- owned and ghost -- in its local numbering; pressure from the corner-node rows."""
    import shutil
    import fracstep
    import meshgen
    base = str(tmp_path / "case")
    meshgen.cube(base + ".inp", 4, cluster=2.0, is_restart=1)
    plain = str(tmp_path / "plain.inp")
    meshgen.cube(plain, 4, cluster=2.0)
    g = fracstep.Solver(plain, scheme, init_gpu=False)
    NN, NNp = g.i("NN"), g.i("NNp")
    rng = np.random.default_rng(3)
    U, P = rng.standard_normal(3 * NN), rng.standard_normal(NNp)
    g.write_tecplot_fields(base + "_restart.dat", "restart", U, P)
    Ur, Pr = g.read_restart_fields(base + "_restart.dat")          # what the file holds (11 significant digits)
    g.close()
    for world in (2, 3):
        seen = np.zeros(NN, dtype=bool)
        for rank in range(world):
            s = fracstep.Solver(base + ".inp", scheme, world=world, rank=rank, init_gpu=False)
            ng, pg = s.arr("part_nodeGlobal"), s.arr("part_presGlobal")
            U0, P0 = s.arr("U0"), s.arr("P0")
            assert len(U0) == 3 * len(ng) and len(P0) == len(pg)
            assert np.array_equal(U0.reshape(3, -1), Ur.reshape(3, NN)[:, ng])
            assert np.array_equal(P0, Pr[pg])
            seen[ng] = True
            s.close()
        assert seen.all()
    # and a whole-mesh solver still reads it the old way
    w = fracstep.Solver(base + ".inp", scheme, init_gpu=False)
    assert np.array_equal(w.arr("U0"), Ur) and np.array_equal(w.arr("P0"), Pr)
    w.close()
