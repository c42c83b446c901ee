"""The drop-in boundary in the reference's own language: include/tess_b200_diy.hpp gives
tessb200::dense() the signature of the reference's dense() over a diy::Master.  The test library
oracle/_ref/libtess_dropin.so is the reference's driver code with exactly that one call swapped
(oracle/ref_driver.cpp, -DTESSB200_DROPIN); the outputs left in the DBlocks -- and the bytes the
reference's own WriteGrid then writes from them -- must equal what the unmodified reference produces."""
import os
import tempfile

import numpy as np
import pytest

from conftest import dataset, assert_same_bits, ROOT

DROPIN = os.path.join(ROOT, "oracle", "_ref", "libtess_dropin.so")


def test_dropin_header_compiles_against_the_reference_headers():
    # compile-only (no GPU needed): the header must stay source-compatible with include/tess/dense.hpp
    if not os.path.isdir("/root/reference/include"):
        pytest.skip("/root/reference absent")
    import subprocess
    src = '#include "tess_b200_diy.hpp"\nint main() { return 0; }\n'
    with tempfile.TemporaryDirectory() as td:
        f = os.path.join(td, "t.cpp")
        open(f, "w").write(src)
        subprocess.run(["g++", "-std=c++14", "-fsyntax-only", "-w", "-DTESS_NO_OPENMP", "-DDIY_NO_THREADS", "-I", os.path.join(ROOT, "oracle", "stub"),
                        "-I", "/root/reference/include", "-I", os.path.join(ROOT, "include"), f], check=True)


@pytest.mark.gpu
@pytest.mark.parametrize("name,gs", [("u16x8", (32, 32, 32)), ("clump8", (48, 48, 48))])
def test_dropin_equals_reference(reference, name, gs):
    if not os.path.exists(DROPIN):
        pytest.skip("oracle/_ref/libtess_dropin.so not built (needs /root/reference at build time)")
    from oracle import ref
    drp = ref.Checker("dropin")
    blocks = dataset(name)
    for alg in (0, 1):
        for proj in (False, True):
            with tempfile.TemporaryDirectory() as td:
                f1, f2 = os.path.join(td, "ref.raw"), os.path.join(td, "gpu.raw")
                o1 = reference.dense(blocks, gs, alg=alg, project=proj, outfile=f1)
                o2 = drp.dense(blocks, gs, alg=alg, project=proj, outfile=f2)   # WriteGrid is the reference's own
                raw1, raw2 = np.fromfile(f1, np.float32), np.fromfile(f2, np.float32)
            assert o1["block_min_idx"] == o2["block_min_idx"] and o1["block_num_idx"] == o2["block_num_idx"]
            for i, (d1, d2) in enumerate(zip(o1["block_density"], o2["block_density"])):
                assert_same_bits(d1, d2, f"{name} alg{alg} proj{proj} block {i}")
            assert_same_bits(raw1, raw2, f"{name} alg{alg} proj{proj} dense.raw")
            assert_same_bits(o1["step"], o2["step"], "grid_step_size")


@pytest.mark.gpu
def test_dropin_joins_the_masters_communicator(reference, monkeypatch):
    # The multi-rank set-up of tessb200::dense (layout of every block gathered over master.communicator(), NCCL unique id
    # from rank 0, tessb200_comm_init, tessb200_dense_set_layout) run in one process: the MPI stand-in has one rank, so
    # the gathers are copies and the communicator has one member; the result must not change.
    if not os.path.exists(DROPIN):
        pytest.skip("oracle/_ref/libtess_dropin.so not built (needs /root/reference at build time)")
    import torch  # noqa: F401  (maps torch's own libnccl.so.2 first: one NCCL per process, tess2_b200/multi.py)
    from oracle import ref
    monkeypatch.setenv("TESSB200_DIY_FORCE_JOIN", "1")
    drp = ref.Checker("dropin")
    blocks = dataset("u16x8")
    o1 = reference.dense(blocks, (32, 32, 32), alg=0)
    o2 = drp.dense(blocks, (32, 32, 32), alg=0)
    for i, (d1, d2) in enumerate(zip(o1["block_density"], o2["block_density"])):
        assert_same_bits(d1, d2, f"joined, block {i}")
