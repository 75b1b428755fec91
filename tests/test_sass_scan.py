"""The built library's SASS holds what the design says it holds (cuobjdump; no GPU): the hot kernels reach shared
memory with LDS / STS only -- a pointer that loses its address space turns every access into a generic LD / ST on the
long scoreboard, which cost W = 65536 five percent until an ncu capture showed it -- keep their
tables in Tensor Memory (LDTM / STTM), and the CTA pair exchanges its half buffers with one bulk copy (UBLKCP)."""
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def table():
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "sass_scan.py")], capture_output=True, text=True, check=True).stdout
    rows = {}
    for line in out.splitlines():
        cells = [c.strip() for c in line.strip("|").split("|")]
        if len(cells) == 12 and cells[1].isdigit():
            rows[cells[0].replace("void ", "")] = dict(zip(("LDS", "STS", "LDG", "STG", "LDTM", "STTM", "UBLKCP", "LD", "ST", "LDL", "STL"), map(int, cells[1:])))
    assert rows, out[:400]
    return rows


def test_hot_kernels_use_shared_memory_instructions(table):
    hot = [k for k in table if re.match(r"stretch_kernel<(9|1[0-4]), false, [1-5]>|stretch_pair_kernel|analysis_pair_kernel", k)]
    assert len(hot) >= 20, sorted(table)
    for k in hot:
        c = table[k]
        # no generic accesses -- but for the pair's posted stores into the peer CTA's buffer (st.shared::cluster is a
        # generic ST in SASS)
        assert c["LD"] == 0 and c["ST"] <= (16 if k == "stretch_pair_kernel" else 0), (k, c)
        assert c["LDS"] > 40 and c["STS"] > 40, (k, c)
        assert c["LDL"] + c["STL"] <= 12, (k, c)           # at most a few spilled words


def test_tensor_memory_and_bulk_copy_are_in_the_sass(table):
    for w in (9, 10, 11, 12, 13, 14):
        c = table["stretch_kernel<%d, false, 2>" % w]
        assert c["LDTM"] > 0 and c["STTM"] > 0, (w, c)
    assert table["stretch_kernel<8, false, 2>"]["LDTM"] == 0      # W = 512 and below: tables in L1 / L2
    pair = table["stretch_pair_kernel"]
    assert pair["LDTM"] > 0 and pair["UBLKCP"] == 1, pair
