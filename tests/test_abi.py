"""The C-ABI library: loads, exports every symbol include/tess_b200.h declares, and fails loudly
(no CPU fallback) when there is no CUDA device.  No compute calls here."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols(name="tess_b200.h"):
    txt = open(os.path.join(ROOT, "include", name)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(tessb200_[a-z0-9_]+)\s*\(", txt)))


def test_header_declares_the_expected_entry_points():
    from tess2_b200 import lib
    assert header_symbols() == sorted(lib.EXPORTS)


def test_library_exports_every_declared_symbol():
    from tess2_b200 import lib
    l = lib.load()
    for name in header_symbols():
        assert hasattr(l, name), f"libtess_b200.so does not export {name}"
    assert l.tessb200_version() == 1


def test_host_library_exports_every_declared_symbol(tmp_path):
    # the CPU side of the tessellation stage (include/tess_b200_host.h, libtess_b200_host.so)
    import subprocess
    from tess2_b200 import host_tess
    syms = [s for s in header_symbols("tess_b200_host.h")]
    assert sorted(syms) == sorted(host_tess.EXPORTS)
    l = host_tess.load()
    for name in syms:
        assert hasattr(l, name), f"libtess_b200_host.so does not export {name}"
    src = tmp_path / "szh.c"
    src.write_text('''#include <stdio.h>
#include <stddef.h>
#include "tess_b200_host.h"
int main(void) { printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(tessb200_host_block), offsetof(tessb200_host_block, particles), offsetof(tessb200_host_block, seconds),
                        sizeof(tessb200_host_dblock), offsetof(tessb200_host_dblock, particles), offsetof(tessb200_host_dblock, vert_to_tet)); return 0; }''')
    exe = tmp_path / "szh"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert got == [C.sizeof(host_tess.HostBlock), host_tess.HostBlock.particles.offset, host_tess.HostBlock.seconds.offset,
                   C.sizeof(host_tess.HostDBlock), host_tess.HostDBlock.particles.offset, host_tess.HostDBlock.vert_to_tet.offset]


def test_struct_layouts_match_the_header(tmp_path):
    # sizeof/offsetof as the C compiler sees the header vs the ctypes mirror: drift would corrupt memory
    import subprocess
    from tess2_b200 import lib
    src = tmp_path / "sz.c"
    src.write_text('''#include <stdio.h>
#include <stddef.h>
#include "tess_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(tessb200_block), sizeof(tessb200_dense_params), sizeof(tessb200_dense_stats),
         offsetof(tessb200_block, density), offsetof(tessb200_block, num_grid_pts),
         offsetof(tessb200_dense_params, data_mins), offsetof(tessb200_dense_stats, ms_upload));
  return 0; }''')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    want = [C.sizeof(lib.Block), C.sizeof(lib.DenseParams), C.sizeof(lib.DenseStats), lib.Block.density.offset,
            lib.Block.num_grid_pts.offset, lib.DenseParams.data_mins.offset, lib.DenseStats.ms_upload.offset]
    assert got == want


def test_no_cpu_fallback_without_a_device():
    import torch
    from tess2_b200 import lib
    l = lib.load()
    h = C.c_void_p()
    rc = l.tessb200_create(C.byref(h), 0)
    if torch.cuda.is_available():
        assert rc == 0
        l.tessb200_destroy(h)
    else:
        assert rc == -2 and not h.value            # TESSB200_ECUDA
        assert b"no CPU fallback" in l.tessb200_last_error()
        from tess2_b200 import Context, TessB200Error
        with pytest.raises(TessB200Error):
            Context(0)


def test_product_does_not_import_the_oracle():
    # the oracle is test infrastructure: nothing under tess2_b200/ may reference it
    bad = []
    for dp, _, fs in os.walk(os.path.join(ROOT, "tess2_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                s = open(os.path.join(dp, f), errors="replace").read()
                if re.search(r"(from|import)\s+oracle|oracle/|libtess_oracle|libtess_ref|dense_oracle", s):
                    bad.append(os.path.join(dp, f))
    assert not bad, bad


@pytest.mark.parametrize("name,gs", [("u16x8", (32, 32, 32)), ("clump8", (48, 48, 48)), ("aniso", (40, 28, 17))])
def test_write_grid_equals_the_reference_writer(reference, tmp_path, name, gs):
    # WriteGrid / ProjectGrid (src/dense.cpp:751-1023) are host code on both sides: tessb200_write_grid over the block
    # densities the unmodified reference left in its DBlocks must write the bytes the reference's own MPI-IO writer wrote
    from conftest import dataset, write_grid_from, assert_same_bits
    import numpy as np
    blocks = dataset(name)
    for alg in (0, 1):
        for proj in (False, True):
            f1, f2 = tmp_path / "ref.raw", tmp_path / "lib.raw"
            o = reference.dense(blocks, gs, alg=alg, project=proj, outfile=str(f1))
            write_grid_from(o, blocks, gs, proj, f2)
            a, b = np.fromfile(f1, np.float32), np.fromfile(f2, np.float32)
            assert a.size == (gs[0] * gs[1] if proj else gs[0] * gs[1] * gs[2])
            assert_same_bits(a, b, f"{name} alg{alg} proj{proj} dense.raw")


def test_check_block_finds_malformed_input():
    # tessb200_check_block is host code: the kernels trust their input (as the reference does), this is the opt-in guard
    import numpy as np
    import tess2_b200
    from conftest import dataset
    blocks = dataset("u16x8")
    tess2_b200.check_blocks(blocks, deep=True)           # SciPy-Qhull blocks of the harness
    from tess2_b200 import host_tess
    p = np.random.default_rng(3).random((500, 3), dtype=np.float32)
    native = host_tess.tess(p, None, host_tess.regular_blocks([0, 0, 0], [1, 1, 1], 2), [0, 0, 0], [1, 1, 1])
    tess2_b200.check_blocks(native, deep=True)           # blocks of the repo's own tess()

    def broken(**kw):
        b = dict(blocks[0])
        for k, f in kw.items():
            a = np.array(b[k], copy=True)
            f(a)
            b[k] = a
        return [b]

    def set_(idx, v):
        return lambda a: a.__setitem__(idx, v)

    n, t = len(blocks[0]["particles"]), len(blocks[0]["tets"])
    cases = {
        "vertex": (broken(tets=set_((5, 2), n)), False),
        "neighbour": (broken(tets=set_((5, 6), t)), False),
        "neighbour ": (broken(tets=set_((5, 6), -2)), False),
        "repeats": (broken(tets=set_((7, 1), int(blocks[0]["tets"][7, 0]))), False),
        "vert_to_tet": (broken(vert_to_tet=set_(3, t + 4)), False),
        "does not hold": (broken(vert_to_tet=set_(3, int(np.argmax((blocks[0]["tets"][:, :4] != 3).all(axis=1))))), False),
        "non-finite": (broken(particles=set_((2, 1), np.nan)), False),
        "not the tet across": (broken(tets=set_((5, 6), (int(blocks[0]["tets"][5, 6]) + 1) % t)), True),
    }
    for what, (blk, deep) in cases.items():
        with pytest.raises(tess2_b200.TessB200Error, match=what.strip()):
            tess2_b200.check_blocks(blk, deep=deep)
    b = dict(blocks[0], num_orig=n + 1)
    with pytest.raises(tess2_b200.TessB200Error, match="inconsistent counts"):
        tess2_b200.check_blocks([b])
