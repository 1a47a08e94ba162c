"""CPU-only checks of the boundary: the C-ABI library loads and exports every
symbol include/pacman_b200.h declares (no compute call is made without a GPU),
argument validation answers with PM_E* codes, the host-side layout loader and
maze generator reproduce the reference's fixtures, and the env class refuses to
run without CUDA instead of falling back."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

import helpers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "pacman_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pm_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from pacman_marl_2025_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 15 and "pm_step" in names and "pm_distances" in names
    lib = C.CDLL(_lib.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(_lib.EXPORTS) == names, "pacman_marl_2025_b200/_lib.py EXPORTS out of sync with the header"
    assert _lib.load().pm_abi_version() == int(re.search(r"#define PM_ABI_VERSION (\d+)", open(
        os.path.join(ROOT, "include", "pacman_b200.h")).read()).group(1))


def test_struct_layouts_match_the_header():
    """ctypes mirrors of pm_env_desc / pm_step_out have the C compiler's size (checked by compiling a probe)."""
    import subprocess, tempfile
    from pacman_marl_2025_b200 import _lib
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "p.c")
        open(src, "w").write('#include <stdio.h>\n#include <stddef.h>\n#include "pacman_b200.h"\nint main(){printf("%zu %zu %zu %zu\\n",'
                             'sizeof(pm_env_desc),sizeof(pm_step_out),offsetof(pm_env_desc,env_index_base),offsetof(pm_env_desc,shaping_scale));return 0;}\n')
        exe = os.path.join(td, "p")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", exe, src])
        a, b, c, d = map(int, subprocess.check_output([exe]).split())
    assert C.sizeof(_lib.PmEnvDesc) == a and C.sizeof(_lib.PmStepOut) == b
    assert _lib.PmEnvDesc.env_index_base.offset == c and _lib.PmEnvDesc.shaping_scale.offset == d


def _header_struct_fields(name):
    """Field names of `typedef struct <name> {...} <name>;` in the header, in declaration order."""
    text = open(os.path.join(ROOT, "include", "pacman_b200.h")).read()
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), text, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        # "const uint32_t *geom" / "int32_t W, H" / "struct pm_host_record *packed"
        first, *rest = decl.split(",")
        fields.append(re.search(r"(\w+)\s*$", first.replace("*", " ")).group(1))
        fields += [r.replace("*", " ").strip() for r in rest]
    return fields


def test_structs_match_header_ctypes_and_integration_doc_field_by_field():
    """pm_env_desc / pm_step_out: the header's fields, the ctypes classes of _lib.py and the binding printed in
    INTEGRATION.md section 4 list the same names in the same order, and every ctypes offset equals the C
    compiler's offsetof (a maintainer who copies the doc gets the real struct, not a truncated one)."""
    import subprocess, tempfile
    from pacman_marl_2025_b200 import _lib
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    for cname, cls in (("pm_env_desc", _lib.PmEnvDesc), ("pm_step_out", _lib.PmStepOut)):
        hdr = _header_struct_fields(cname)
        assert [n for n, _ in cls._fields_] == hdr, cname
        blk = re.search(r"class %s\(C\.Structure\):.*?_fields_ = \[(.*?)\]\n" % cname, doc, re.S).group(1)
        assert re.findall(r'\("(\w+)",', blk) == hdr, f"INTEGRATION.md lists other fields for {cname}"
        doc_types = dict(re.findall(r'\("(\w+)", C\.(\w+)\)', blk))
        for n, t in cls._fields_:
            assert getattr(C, doc_types[n]) is t, (cname, n)           # (c_int32 is an alias of c_int: compare the types)
        with tempfile.TemporaryDirectory() as td:
            src = os.path.join(td, "p.c")
            prints = "".join('printf("%%zu\\n", offsetof(%s, %s));' % (cname, f) for f in hdr)
            open(src, "w").write('#include <stdio.h>\n#include <stddef.h>\n#include "pacman_b200.h"\nint main(){%s return 0;}\n' % prints)
            exe = os.path.join(td, "p")
            subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", exe, src])
            offs = list(map(int, subprocess.check_output([exe]).split()))
        assert [getattr(cls, f).offset for f in hdr] == offs, cname


def test_argument_errors_without_touching_the_gpu():
    from pacman_marl_2025_b200 import _lib
    L = _lib.load()
    assert L.pm_geom_words(300, 4) == -3            # PM_ETOOBIG
    assert L.pm_geom_words(2, 2) == -1              # PM_EINVAL
    buf = np.zeros(L.pm_layout_words(5, 5), np.uint32)
    hc = C.c_int(0)
    bad = ("%%%%%" "%1 2%" "%   %" "%3  %" "%%%%%").encode()        # only three agents
    assert L.pm_pack_layout(bad, 5, 5, buf.ctypes.data, C.byref(hc)) == -2      # PM_ELAYOUT
    open_border = ("%% %%" "%1 2%" "%   %" "%3 4%" "%%%%%").encode()
    assert L.pm_pack_layout(open_border, 5, 5, buf.ctypes.data, C.byref(hc)) == -2
    ok = ("%%%%%" "%1o2%" "% . %" "%3 4%" "%%%%%").encode()
    assert L.pm_pack_layout(ok, 5, 5, buf.ctypes.data, C.byref(hc)) == 0 and hc.value == 1
    # null descriptor / null pointers are rejected before any CUDA call
    assert L.pm_step(None, None, 1, None, None) == -1
    d = _lib.PmEnvDesc()
    assert L.pm_reset(C.byref(d), None, None, 0, None, None) == -1
    assert L.pm_distances(5, 5, None, 1, None, None, None, 8, None) == -1
    assert L.pm_gae(None, None, None, None, 0.99, 0.95, None, None, 4, 4, None) == -1


def test_env_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from pacman_marl_2025_b200.gymPacMan import gymPacMan_parallel_env
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        gymPacMan_parallel_env(self_play=True)
    from pacman_marl_2025_b200.mappo_ops import compute_gae
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        compute_gae(torch.zeros(4, 2), torch.zeros(4, 2), torch.zeros(4, 2), 0.0)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pacman_marl_2025_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "pm_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f


@pytest.mark.parametrize("seed", list(json.load(open(os.path.join(helpers.GOLDEN, "mazes.json")))))
def test_maze_generator_text_matches_reference(seed):
    """mazeGenerator.generateMaze(seed) (mazeGenerator.py:255-264) and capture.randomLayout() (capture.py:905-911,
    seed drawn from the global RNG): same text as the reference."""
    import random
    from pacman_marl_2025_b200 import mazeGenerator
    want = json.load(open(os.path.join(helpers.GOLDEN, "mazes.json")))[seed]
    if seed.startswith("global"):
        random.seed(int(seed[6:]))
        assert mazeGenerator.randomLayout() == want
    else:
        assert mazeGenerator.generateMaze(int(seed)) == want


@pytest.mark.parametrize("name", helpers.traj_names())
def test_layout_loader_matches_reference_fixture(name):
    """layout.Layout (layout.py:95-130): walls / food / capsules / starts of every fixture board, checked
    against the first recorded state of the reference."""
    from pacman_marl_2025_b200.layout import Layout
    fx = helpers.load_traj(name)
    lay = Layout(fx["layout"])
    W, H = fx["W"], fx["H"]
    assert (lay.width, lay.height) == (W, H)
    rows = fx["layout"]
    for x in range(W):
        for y in range(H):
            ch = rows[H - 1 - y][x]
            assert bool(lay.walls[x][y]) == (ch == "%")
            assert bool(lay.food[x][y]) == (ch == ".")
    assert sorted(lay.capsules) == sorted((x, y) for x in range(W) for y in range(H) if rows[H - 1 - y][x] == "o")
    for i in range(4):
        x, y = lay.starts[i]
        assert rows[H - 1 - y][x] == str(i + 1)


def test_batched_layout_packer_equals_one_by_one_and_reports_the_bad_board():
    """pm_pack_layouts (L boards per call, + the largest open-cell count) == pm_pack_layout per board; a malformed board is
    reported with its index and the single-board error code."""
    import ctypes as C
    import numpy as np
    from pacman_marl_2025_b200 import _lib, mazeGenerator
    from pacman_marl_2025_b200.gymPacMan import pack_tables
    from pacman_marl_2025_b200.layout import Layout
    L = _lib.load()
    lays = [Layout(mazeGenerator.generateMaze(300 + s).split("\n")) for s in range(9)]
    W, H = lays[0].width, lays[0].height
    geom, table, has_caps, max_open = pack_tables(lays, with_max_open=True)
    LS = L.pm_layout_words(W, H)
    for i, lay in enumerate(lays):
        one = np.zeros(LS, np.uint32)
        hc = C.c_int(0)
        assert L.pm_pack_layout("".join(lay.layoutText).encode(), W, H, one.ctypes.data, C.byref(hc)) == 0
        np.testing.assert_array_equal(table[i], one)
    assert max_open == max(int((~l.walls.array).sum()) for l in lays) and has_caps == 1
    texts = ["".join(l.layoutText) for l in lays]
    texts[4] = texts[4].replace("1", " ", 1)                      # only three agents on board 4
    out = np.zeros((9, LS), np.uint32)
    bad = C.c_int(-1)
    rc = L.pm_pack_layouts("".join(texts).encode(), W, H, 9, out.ctypes.data, None, None, C.byref(bad))
    assert rc == -2 and bad.value == 4


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the arm the driver runs first, on the box's host cores; no GPU needed): one JSON line
    with the same metric / unit / config keys as the GPU arm plus impl, cpu_baseline and a zero-copy e2e block."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--envs", "256", "--steps", "2",
                        "--warmup", "1"], capture_output=True, text=True, timeout=300, cwd=root)
    assert p.returncode == 0, p.stderr[-500:]
    line = json.loads([ln for ln in p.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["impl"] == "reference" and line["metric"] == "agent-steps/sec (env step+obs)" and line["unit"] == "agent-steps/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["gpu_launches"] == 0
    assert line["config"]["layout"] == "tinyCapture" and line["config"]["envs_per_gpu"] == 256
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "sample" in cb
    assert line["e2e"] == {"value": line["value"], "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # a second rank of a torchrun launch prints nothing and exits 0
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p2 = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--envs", "256",
                         "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=120, cwd=root, env=env)
    assert p2.returncode == 0 and not p2.stdout.strip()
