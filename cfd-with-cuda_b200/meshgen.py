"""Structured hexahedral `.inp` writers in the reference's input grammar.

Reproduces the node/element numbering, BC-face order and text layout of the reference's
MATLAB generators
  oldFiles/meshGenerators&Converters/HexaMeshGeneratorInACube_GeneratesCornerNodes.m
  oldFiles/meshGenerators&Converters/HexaMeshGeneratorInAChannel_GeneratesCornerNodes.m
plus the `converge :` header line that the C++ reader expects
(fractionalStep/explicit/Cpp/blascoCodinaHuerta.cpp:545-562).  With
cube(30, cluster=2.0) the output equals the shipped
inputFiles/lidDrivenCavity/lidDrivenCavity_NE27000.inp byte for byte
(tests/test_meshgen.py checks this whenever /root/reference is mounted).

Only corner nodes are written: the solver generates mid-edge / mid-face / centre nodes.
"""
import math


def _axis(n_elems, length, cluster):
    """Corner-node offsets along one axis (generator lines 57-99)."""
    nc = n_elems + 1
    if cluster == 0:
        return [i * length / (nc - 1) for i in range(nc)]
    if n_elems % 2:
        raise ValueError("clustering needs an even element count per direction")
    half = (nc + 1) // 2
    mx = math.sinh(cluster)
    d = [0.0] * nc
    for i in range(1, half + 1):
        xx = (1.0 * (i - 1)) / (half - 1)
        d[i - 1] = length / 2.0 / mx * math.sinh(cluster * xx)
    for i in range(half, nc + 1):
        d[i - 1] = length - d[nc + 1 - i - 1]
    return d


def _header(title, NE, NCN, NENv, dt, t_final, max_iter, tolerance, converge, viscosity, density=1.0,
            is_restart=0, t_ini=0.0, alpha=1.0):
    return (
        f"{title} \n"
        "================================================\n"
        "eType    : 1 \n"
        f"NE       : {NE} \n"
        f"NCN      : {NCN} \n"
        f"NENv     : {NENv} \n"
        "NENp     : 8 \n"
        "NGP      : 8 \n"
        f"alpha    : {alpha:f}\n"
        f"dt       : {dt:f}\n"
        f"t_ini    : {t_ini:f} \n"
        f"t_final  : {t_final:f} \n"
        f"maxIter  : {max_iter} \n"
        f"tolerance: {tolerance:f}\n"
        f"converge : {converge:f} \n"
        f"isRestart: {is_restart}\n"
        f"density  : {density:f} \n"
        f"viscosity: {viscosity:f} \n"
        "fx       : 0.0 \n"
        "fy       : 0.0 \n"
        "================================================\n"
    )


def _fmt_coord(v, full_precision):
    return repr(float(v)) if full_precision else None


def _write(path, title, nx, ny, nz, xs, ys, zs, vel_faces, pin_node, monitor, NENv, dt, t_final, max_iter,
           tolerance, converge, viscosity, bc_lines, full_precision, is_restart=0):
    NE = nx * ny * nz
    NCN = (nx + 1) * (ny + 1) * (nz + 1)
    out = [_header(title, NE, NCN, NENv, dt, t_final, max_iter, tolerance, converge, viscosity,
                   is_restart=is_restart)]
    out.append("Corner Node No         x                y                z\n")
    n = 0
    rows = []
    for k in range(nz + 1):
        for j in range(ny + 1):
            for i in range(nx + 1):
                n += 1
                if full_precision:
                    # `operator>>` parses any precision (blascoCodinaHuerta.cpp:579); %.17g round-trips doubles
                    rows.append("%9d %s %s %s\n" % (n, "%.17g" % xs[i], "%.17g" % ys[j], "%.17g" % zs[k]))
                else:
                    rows.append("%9d %18.7f %16.7f %16.7f\n" % (n, xs[i], ys[j], zs[k]))
    out.append("".join(rows))
    out.append("================================================\n")
    out.append("Elem No   corner1  corner2  corner3  corner4  corner5  corner6  corner7  corner8\n")
    rows = []
    e = 0
    pl = (nx + 1) * (ny + 1)
    for k in range(1, nz + 1):
        for j in range(1, ny + 1):
            for i in range(1, nx + 1):
                n1 = (k - 1) * pl + (j - 1) * (nx + 1) + i
                n2 = n1 + 1
                n3 = n1 + nx + 2
                n4 = n3 - 1
                e += 1
                rows.append("%6d  %8d %8d %8d %8d %8d %8d %8d %8d \n" % (e, n1, n2, n3, n4, n1 + pl, n2 + pl, n3 + pl, n4 + pl))
    out.append("".join(rows))
    out.append("================================================\n")
    out.append("BCs (Number of specified BCs, their types and strings) \n")
    out.append("nBC       : %d \n" % len(bc_lines))
    for q, (u, v, w) in enumerate(bc_lines):
        out.append("BC %d      : 1  %s : %s : %s\n" % (q + 1, u, v, w))
    out.append("================================================\n")
    out.append("nVelFaces : %d \n" % len(vel_faces))
    out.append("nOutFaces : %d \n" % 0)
    out.append("================================================\n")
    out.append("Velocity BC (Elem# Face# BC#)\n")
    out.append("".join("%5d    %d    %d\n" % t for t in vel_faces))
    out.append("================================================\n")
    out.append("Outflow BC (Elem# Face# BC#)\n")
    out.append("================================================\n")
    out.append("Node number where pressure is taken to be zero\n")
    out.append("%d\n" % pin_node)
    out.append("================================================\n")
    out.append("Monitor point coordinates\n")
    out.append("%s  %s  %s\n" % monitor)
    with open(path, "w") as f:
        f.write("".join(out))
    return dict(NE=NE, NCN=NCN, nx=nx, ny=ny, nz=nz)


def cube(path, n, cluster=0.0, NENv=27, viscosity=0.01, dt=0.001, t_final=1.0, max_iter=4, tolerance=0.001,
         converge=0.000001, full_precision=False, title="3D Lid-driven cavity problem", ny=None, nz=None,
         is_restart=0, pin=None, lid=("1.0", "0.0", "0.0")):
    """Unit-cube lid-driven cavity: no-slip BC 1 on 5 walls, lid BC 2 u=(1,0,0) on the top (face 6).
    `pin` overrides the pressure-pin node (1-based; negative = none), `lid` the lid velocity strings."""
    nx = n
    ny = ny or n
    nz = nz or n
    xs, ys, zs = _axis(nx, 1.0, cluster), _axis(ny, 1.0, cluster), _axis(nz, 1.0, cluster)
    el = lambda i, j, k: (k - 1) * nx * ny + (j - 1) * nx + i
    faces = []
    faces += [(el(i, j, 1), 1, 1) for j in range(1, ny + 1) for i in range(1, nx + 1)]       # bottom
    faces += [(el(1, j, k), 5, 1) for k in range(1, nz + 1) for j in range(1, ny + 1)]       # left
    faces += [(el(nx, j, k), 3, 1) for k in range(1, nz + 1) for j in range(1, ny + 1)]      # right
    faces += [(el(i, 1, k), 2, 1) for k in range(1, nz + 1) for i in range(1, nx + 1)]       # front
    faces += [(el(i, ny, k), 4, 1) for k in range(1, nz + 1) for i in range(1, nx + 1)]      # back
    faces += [(el(i, j, nz), 6, 2) for j in range(1, ny + 1) for i in range(1, nx + 1)]      # top (lid)
    nxc = nx + 1
    if pin is None:
        pin = (ny // 2) * nxc + (nxc + 1) // 2
    return _write(path, title, nx, ny, nz, xs, ys, zs, faces, pin, ("0.5", "0.5", "0.5"), NENv, dt, t_final,
                  max_iter, tolerance, converge, viscosity, [("0.0", "0.0", "0.0"), tuple(lid)],
                  full_precision, is_restart=is_restart)


def channel(path, nx=48, ny=32, nz=32, length=10.0, NENv=27, viscosity=0.01, dt=0.001, t_final=1.0, max_iter=4,
            tolerance=0.001, converge=0.000001, full_precision=False, title="3D channel flow problem", outlet="free"):
    """Straight channel [0,L]x[0,1]^2 (stand-in for the missing bendingSquareDuct_49x33x33.inp):
    no-slip walls (BC 1), uniform inlet u=(1,0,0) on x=0 (BC 2, written last so it wins on the
    inlet edges), pressure pinned at the outlet mid node.  outlet = "free": no condition on x=L -- the reference reads outflow
    faces but never uses them, and its scheme then DIVERGES within a few steps on any mesh size (oracle and reference
    program alike, profiles/r02_channel_oracle_diverges.txt); outlet = "plug": the inlet profile is prescribed on x=L as well
    (mass-conserving, all-Dirichlet like the cavity): the well-posed duct stand-in, see duct()."""
    xs, ys, zs = _axis(nx, length, 0.0), _axis(ny, 1.0, 0.0), _axis(nz, 1.0, 0.0)
    el = lambda i, j, k: (k - 1) * nx * ny + (j - 1) * nx + i
    faces = []
    faces += [(el(i, j, 1), 1, 1) for j in range(1, ny + 1) for i in range(1, nx + 1)]       # bottom
    faces += [(el(i, 1, k), 2, 1) for k in range(1, nz + 1) for i in range(1, nx + 1)]       # front
    faces += [(el(i, ny, k), 4, 1) for k in range(1, nz + 1) for i in range(1, nx + 1)]      # back
    faces += [(el(i, j, nz), 6, 1) for j in range(1, ny + 1) for i in range(1, nx + 1)]      # top
    faces += [(el(1, j, k), 5, 2) for k in range(1, nz + 1) for j in range(1, ny + 1)]       # inlet
    if outlet == "plug":
        faces += [(el(nx, j, k), 3, 2) for k in range(1, nz + 1) for j in range(1, ny + 1)]  # outlet: the same plug profile
    elif outlet != "free":
        raise ValueError("outlet must be 'free' or 'plug'")
    pin = nz // 2 * (nx + 1) * (ny + 1) + (ny // 2 + 1) * (nx + 1)
    return _write(path, title, nx, ny, nz, xs, ys, zs, faces, pin, ("%.1f" % length, "0.5", "0.5"), NENv, dt,
                  t_final, max_iter, tolerance, converge, viscosity,
                  [("0.0", "0.0", "0.0"), ("1.0", "0.0", "0.0")], full_precision)


def duct(path, **kw):
    """Laminar square-duct flow with prescribed in- and outflow: channel(outlet="plug").  The configs[2] stand-in the implicit scheme
    actually converges on (BiCGStab 100-300 iterations per system and step at 48x32x32, fields settle within a few steps)."""
    kw.setdefault("title", "3D duct flow problem")
    return channel(path, outlet="plug", **kw)


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("kind", choices=["cube", "channel"])
    ap.add_argument("path")
    ap.add_argument("-n", type=int, default=8)
    ap.add_argument("--cluster", type=float, default=0.0)
    ap.add_argument("--nenv", type=int, default=27, choices=[8, 27])
    ap.add_argument("--viscosity", type=float, default=0.01)
    a = ap.parse_args()
    if a.kind == "cube":
        print(cube(a.path, a.n, a.cluster, a.nenv, a.viscosity))
    else:
        print(channel(a.path, NENv=a.nenv, viscosity=a.viscosity))
