"""Expected dense.raw files for profiles/quick/run.sh: the `tess` driver's blocks through the CPU oracle (port) and the
library's host-side WriteGrid.  CPU only; run here, the files travel to the GPU box with the snapshot."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import write_grid_from  # noqa: E402
from oracle import ref  # noqa: E402
from tess2_b200 import host_tess  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
BIN = os.path.join(HERE, "_bin")
os.makedirs(BIN, exist_ok=True)
subprocess.run(["make", "--no-print-directory", "-C", os.path.join(ROOT, "examples", "drivers"), f"OUT={BIN}"], check=True)
subprocess.run([os.path.join(BIN, "tess"), "8", "-1", "16", "16", "16", "0", "-1", "-1", "0", "0", os.path.join(BIN, "del.out")], check=True)
blocks, _, _, _ = host_tess.read_blocks(os.path.join(BIN, "del.out"))
port = ref.Checker("port")
CASES = {   # name: (alg, gsize, project, given bounds)
    "a_3d_tess": (0, 32, False, None),
    "b_3d_cic": (1, 32, False, None),
    "c_proj_two_bounds": (0, 24, True, ([-1.5, -1.5], [16.5, 16.5])),
    "d_proj_narrow_z": (0, 24, True, ([-1.5, -1.5, 2.5], [16.5, 16.5, 12.5])),
    "e_proj_narrow_z_cic": (1, 24, True, ([-1.5, -1.5, 2.5], [16.5, 16.5, 12.5])),
    "f_3d_narrow_xy": (0, 24, False, ([2.5, 2.5], [12.5, 12.5])),
}
with open(os.path.join(BIN, "cases.txt"), "w") as f:
    for name, (alg, g, proj, gb) in CASES.items():
        o = port.dense(blocks, (g, g, g), alg=alg, project=proj, given_bounds=gb)
        write_grid_from(o, blocks, (g, g, g), proj, os.path.join(BIN, name + ".exp"))
        tail = [str(g)] * 3 + (["0", "0", "1"] if proj else ["!"]) + ["1", str(0 if gb is None else len(gb[0]))]
        if gb:
            tail += [str(x) for x in gb[0]] + [str(x) for x in gb[1]]
        f.write(f"{name} {alg} {' '.join(tail)}\n")
        print(name, "out_of_range", o["out_of_range"], "sum", float(np.fromfile(os.path.join(BIN, name + '.exp'), np.float32).sum()))
