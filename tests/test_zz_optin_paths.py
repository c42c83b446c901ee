"""The opt-in code paths stay bit-identical to the oracle: TESSB200_FUSED=1 (one kernel per cell, fused.cuh), TESSB200_SEGMENTS=1
(shared grid points through per-point segments), TESSB200_DIRECT=0 (no k_cell_direct: round 1's routing).  The switches are read
when a context is created, so every variant runs in its own process."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

SCRIPT = r'''
import json, os, sys
import numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
from conftest import dataset
from oracle import ref
import tess2_b200
port = ref.Checker("port")
ctx = tess2_b200.Context(0)
bad = 0; n = 0
for name, gs in (("u16x8", (32, 32, 32)), ("clump8", (48, 48, 48)), ("c1", (64, 64, 64))):
    blocks = dataset(name)
    for alg in (0, 1):
        o = port.dense(blocks, gs, alg=alg, assemble=False)
        r = ctx.dense(alg, 0, None, None, False, (0.0, 0.0, 1.0), 1.0, 1e-4, gs, blocks, want_grid=False)
        for d1, d2 in zip(r.block_density, o["block_density"]):
            bad += int((np.ascontiguousarray(d1).view(np.uint32) != np.ascontiguousarray(d2).view(np.uint32)).sum()); n += d1.size
ctx.close()
print(json.dumps({"values": n, "differing": bad}))
'''


@pytest.mark.parametrize("env", [{"TESSB200_FUSED": "1"}, {"TESSB200_SEGMENTS": "1"}, {"TESSB200_DIRECT": "0"}, {"TESSB200_FUSED": "1", "TESSB200_SEGMENTS": "1"}, {"TESSB200_CIC_GATHER": "0"}],
                         ids=["fused", "segments", "no_direct", "fused+segments", "cic_records"])
def test_optin_path_matches_oracle(env):
    r = subprocess.run([sys.executable, "-c", SCRIPT, ROOT], capture_output=True, text=True, env=dict(os.environ, **env), timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["values"] > 500000 and out["differing"] == 0, out
