"""torchrun --nproc-per-node N tests/mgpu_check.py : partitioned run vs single-GPU run of the same problem.
Environment: NO_P2P=1 -> host-driven NCCL Krylov path; CG_VARIANTS=0,1,2 -> pressure-CG recurrences to run (cfd_options.cg_variant)."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'cfd-with-cuda_b200'))
import fracstep
rank, world, lr = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(lr)
dist.init_process_group('nccl', device_id=torch.device('cuda', lr))
def bcast_id():
    t = torch.zeros(128, dtype=torch.uint8, device='cuda')
    if rank == 0: t.copy_(torch.frombuffer(bytearray(fracstep.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(t, 0); return bytes(t.cpu().numpy().tobytes())
cases = [(dict(nx=8, NENv=27), 'explicit', 3, {}), (dict(nx=12, NENv=8), 'implicit', 3, {}),
         (dict(nx=8, NENv=27, cluster=2.0), 'implicit', 2, dict(cg_tol=1e-12, bicg_tol=1e-12, cg_maxit=5000, bicg_maxit=5000)),
         (dict(nx=12, NENv=8), 'explicit', 2, {})]
variants = [int(v) for v in os.environ.get('CG_VARIANTS', '0').split(',')]
if os.environ.get('CASES'):
    cases = [cases[int(k)] for k in os.environ['CASES'].split(',')]
for (cav, scheme, steps, kw), variant in [(cs, v) for cs in cases for v in variants]:
    opts = fracstep.options(device=lr, cg_variant=variant, **kw)
    s = fracstep.Solver(scheme=scheme, cavity=cav, opts=opts, world=world, rank=rank)
    fracstep.setup_distributed(s, dist, torch, peer_memory=os.environ.get('NO_P2P') is None)
    reps = s.ctx.time_loop(steps)
    ng, U, pg, P = s.owned_fields()
    if rank == 0:
        g = fracstep.Solver(scheme=scheme, cavity=cav, opts=opts)
        gre = g.ctx.time_loop(steps)
        Ug, Pg = g.ctx.fields(); NN = g.i('NN'); Ug = Ug.reshape(3, NN)
    # gather errors
    errs = torch.zeros(4, dtype=torch.float64, device='cuda')
    Uref = torch.zeros(3 * s.i('globalNN'), dtype=torch.float64, device='cuda'); Pref = torch.zeros(s.i('globalNNp'), dtype=torch.float64, device='cuda')
    if rank == 0: Uref.copy_(torch.from_numpy(Ug.reshape(-1))); Pref.copy_(torch.from_numpy(Pg))
    dist.broadcast(Uref, 0); dist.broadcast(Pref, 0)
    Ur = Uref.cpu().numpy().reshape(3, -1)[:, ng]; Pr = Pref.cpu().numpy()[pg]
    errs[0] = float(((U - Ur) ** 2).sum()); errs[1] = float((Ur ** 2).sum()); errs[2] = float(((P - Pr) ** 2).sum()); errs[3] = float((Pr ** 2).sum())
    dist.all_reduce(errs)
    if rank == 0:
        e = errs.cpu().numpy()
        print(cav, scheme, 'cg_variant', variant, 'world', world, 'relU %.2e relP %.2e' % (np.sqrt(e[0] / e[1]), np.sqrt(e[2] / e[3])))
        print('   part: iter', [r['iter'] for r in reps], 'cg', [r['cg_iterations'] for r in reps], 'bicg', [r['bicg_iterations'] for r in reps], 'mon', reps[-1]['monitor'], 'maxAcc', reps[-1]['maxAcc'])
        print('   single: iter', [r['iter'] for r in gre], 'cg', [r['cg_iterations'] for r in gre], 'bicg', [r['bicg_iterations'] for r in gre], 'mon', gre[-1]['monitor'], 'maxAcc', gre[-1]['maxAcc'])
        g.close()
    fracstep.teardown_distributed(s, dist)
dist.barrier()
dist.destroy_process_group()
