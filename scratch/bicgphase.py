import sys, os
sys.path.insert(0, 'cfd-with-cuda_b200')
import fracstep, meshgen
wl = sys.argv[1] if len(sys.argv) > 1 else "ne27000"
if wl == "ne27000":
    meshgen.cube('/tmp/c.inp', 30, cluster=2.0, viscosity=0.001)
elif wl == "q2_64":
    meshgen.cube('/tmp/c.inp', 64, cluster=0.0, viscosity=0.01, full_precision=True)
s = fracstep.Solver('/tmp/c.inp', 'implicit', opts=fracstep.options(device=0))
reps, ms = s.ctx.time_loop_timed(4)
print("ms/step", ms / 4, [r["bicg_iterations"] for r in reps], [r["cg_iterations"] for r in reps], file=sys.stderr)
