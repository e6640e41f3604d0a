"""The blend work models (tests/tools/) still build and run: a small frame, sanity relations between their counters."""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def _run(tool, *args):
    out = subprocess.run([sys.executable, os.path.join(HERE, "tools", tool), *args], check=True, capture_output=True, text=True,
                         timeout=600).stdout
    return out


def test_blend_model_counters_are_consistent():
    out = _run("blend_model.py", "C1", "20000")
    c = {m.group(1): int(m.group(2)) for m in re.finditer(r"^(\w+)\s+(\d+)$", out, re.M)}
    stats = eval(re.search(r"(\{'E'.*\})", out).group(1))
    # the replay blends exactly the pairs the oracle's blend counts as passing the alpha test (up to exp rounding at the
    # saturation test), never more than the reference examines
    assert abs(c["lane_ok"] - stats["E_alpha"]) <= max(16, stats["E_alpha"] // 10000)
    assert c["lane_ok"] <= stats["E"]
    # fewer, larger patches need fewer hit evaluations and fewer segment walks; dense packing never walks more
    assert c["hits_exact"] <= c["hits"] and c["hits_8x8"] <= c["hits"] and c["hits_8x8_exact"] <= c["hits_8x8"]
    assert c["seg_walks_8x8"] <= c["seg_walks"] and c["seg_walks_dense"] <= c["seg_walks"]
    assert max(c["hits_half_lr"], c["hits_half_tb"], c["hits_quarter"], c["hits_rows"]) <= c["hits"]


def test_blend_rounds_walks_do_not_depend_on_the_round_size():
    out = _run("blend_rounds.py", "C1", "20000")
    rows = [[int(x) for x in ln.split()] for ln in out.splitlines() if re.match(r"^\s*\d+\s+\d+", ln)]
    assert [r[0] for r in rows] == [32, 64, 128, 256]
    assert len({r[3] for r in rows}) == 1                      # segment walks
    assert all(a[4] <= b[4] for a, b in zip(rows, rows[1:]))   # hit evaluations grow with the round size (staler box)
    assert all(a[5] <= b[5] for a, b in zip(rows, rows[1:]))   # so does the number of staged candidates
