"""Randomised configurations through the three CPU statements of the dense stage: the unmodified reference
(oracle/_ref/libtess_ref.so), the plain-C port (oracle/dense_oracle.c) and the kernels' __host__ __device__ logic
(tests/emul/emul.cpp).  Test infrastructure only.

    python tests/fuzz_logic.py [seed] [seconds]

Random particle sets (uniform / clustered / jittered lattice, domains of extent 0.3 .. 100 at offsets up to 1e4),
1-16 regular or kd-tree blocks from the repo's host tess(), grids of 4..90 points per axis, both algorithms, 3-D
and projected, eps 1e-6..1e-2, 0-3 given bounds wider or narrower than the data.  Every case runs in a forked
child: the reference has no bounds checks, and where a deposit falls outside its block's sub-grid (counted by
the port: `out_of_range`) the reference's result is undefined and only port == device logic is asked for.
Round 1: 74 000 cases over 11 seeds (5 x 300 s, 6 x 1200 s), no difference; it found the projection-with-narrow-z case (test_emul.py).
Round 2: the emulation runs DENSE_CIC in the gather form of k_cic_prepare / k_cic_gather (3-D): seeds 777 (240 s), 4242 (1200 s) and, with the final merge loop, 99 (300 s): 18 905 cases, no difference."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def same(o1, o2):
    if o1["block_min_idx"] != o2["block_min_idx"] or o1["block_num_idx"] != o2["block_num_idx"]:
        return "geometry"
    return sum(int((a.view(np.uint32) != b.view(np.uint32)).sum()) for a, b in zip(o1["block_density"], o2["block_density"]))


def random_case(rng):
    from tess2_b200 import host_tess
    from tess2_b200.harness import particles
    n = int(rng.integers(30, 3000))
    off = float(rng.choice([0.0, 0.0, -50.0, 1000.0, 1e4]))
    ext = np.array(rng.choice([1.0, 16.0, 100.0]) * rng.uniform(0.3, 1.0, 3), np.float32)
    if rng.random() < 0.5:
        ext[:] = ext[0]
    dmin = np.full(3, off, np.float32)
    dmax = (dmin + ext).astype(np.float32)
    kind = str(rng.choice(["uniform", "clustered", "lattice"]))
    if kind == "uniform":
        p = particles.uniform_particles(n, dmin, dmax, seed=int(rng.integers(1 << 30)))
    elif kind == "clustered":
        p = particles.clustered_particles(n, dmin, dmax, seed=int(rng.integers(1 << 30)), n_clumps=int(rng.integers(1, 8)))
    else:
        k = max(3, int(round(n ** (1 / 3))))
        g = np.stack(np.meshgrid(*[np.linspace(0.05, 0.95, k)] * 3, indexing="ij"), -1).reshape(-1, 3)
        g = g + rng.normal(0, rng.choice([1e-3, 1e-2]), g.shape)
        p = (dmin + np.clip(g, 0.001, 0.999) * ext).astype(np.float32)
    p = np.unique(p, axis=0)
    nb = int(rng.choice([1, 2, 4, 8, 16]))
    if rng.random() < 0.5:
        bounds, owner = host_tess.regular_blocks(dmin, dmax, nb), None
    else:
        bounds, owner = host_tess.kdtree_blocks(p, dmin, dmax, nb)
    blocks = host_tess.tess(p, owner, bounds, dmin, dmax)
    gs = tuple(int(x) for x in rng.integers(4, int(rng.choice([12, 40, 90])), 3))
    if rng.random() < 0.4:
        gs = (gs[0],) * 3
    gb = None
    if rng.random() < 0.5:
        k = int(rng.integers(1, 4))
        pad = ext * rng.uniform(-0.3, 0.3, 3)
        gb = ([float(x) for x in (dmin - pad)[:k]], [float(x) for x in (dmax + pad)[:k]])
    args = dict(alg=int(rng.integers(0, 2)), project=bool(rng.random() < 0.4), eps=float(rng.choice([1e-4, 1e-4, 1e-6, 1e-2])),
                mass=float(rng.choice([1.0, 1.0, 0.37, 1e3])), given_bounds=gb)
    desc = dict(n=len(p), kind=kind, off=off, ext=ext.tolist(), nb=nb, kd=owner is not None, gs=gs, **args)
    return blocks, gs, args, desc


def run(seed, seconds=None, cases=None, checkers=None, log=print):
    """Returns (cases run, list of failure descriptions)."""
    from oracle import ref
    if checkers is None:
        emul_so = os.path.join(ROOT, "tests", "emul", "_build", "libtess_emul.so")
        checkers = (ref.Checker("reference"), ref.Checker("port"), ref.Checker("emul", emul_so, "emu_"))
    reference, port, emul = checkers
    rng = np.random.default_rng(seed)
    failures = []
    t0, it = time.time(), 0
    while (seconds is None or time.time() - t0 < seconds) and (cases is None or it < cases):
        it += 1
        try:
            blocks, gs, args, desc = random_case(rng)
        except RuntimeError as e:
            log("tess failed", e)
            continue
        r, w = os.pipe()
        pid = os.fork()
        if pid == 0:
            os.close(r)
            msg = ""
            try:
                o1 = port.dense(blocks, gs, **args)
                o2 = emul.dense(blocks, gs, **args)
                d = same(o1, o2)
                if d:
                    msg = f"port vs device logic: {d}"
                elif o1["out_of_range"] == 0:
                    o0 = reference.dense(blocks, gs, **args)
                    d = same(o0, o1)
                    if d:
                        msg = f"reference vs port: {d}"
            except RuntimeError as e:
                # -3: a block without grid points; -1: a block's index box exceeds the grid (narrow given bounds):
                # rejected, not a failure
                if not str(e).endswith(("-3", "-1")):
                    msg = f"exception: {e}"
            os.write(w, msg.encode())
            os._exit(0)
        os.close(w)
        msg = os.read(r, 4096).decode()
        os.close(r)
        _, status = os.waitpid(pid, 0)
        if status != 0:
            msg = f"crash (wait status {status})"
        if msg:
            failures.append((msg, desc))
            log("FAIL", msg, desc)
    return it, failures


if __name__ == "__main__":
    n, f = run(int(sys.argv[1]) if len(sys.argv) > 1 else 0, seconds=float(sys.argv[2]) if len(sys.argv) > 2 else 60.0)
    print(f"{n} cases, {len(f)} failures")
    sys.exit(1 if f else 0)
