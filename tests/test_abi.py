"""CPU-side checks of the drop-in boundary: libxnart.so loads, exports every symbol that
include/xnart.h declares, agrees on struct sizes, reports errors the way the header says, and its
native octree builder produces exactly the oracle's (= the reference's) trees.  No compute calls:
those need a GPU and live in the `-m gpu` tests."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from xna_ray_trace_b200 import capi, records, scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    capi.build()
    return capi.lib()


def test_exports_match_header(built_lib):
    hdr = open(os.path.join(ROOT, "include", "xnart.h")).read()
    declared = sorted(set(re.findall(r"\b(xnart_[a-z0-9_]+)\s*\(", hdr)))
    assert declared == sorted(capi.EXPORTS)
    for name in declared:
        assert hasattr(built_lib, name), name


def test_struct_sizes_match_header():
    # compile a tiny C program against the header and compare sizeof with the ctypes mirrors
    import subprocess
    import tempfile

    src = r'''
#include <stdio.h>
#include "xnart.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu\n", sizeof(XnartTriangle), sizeof(XnartMaterial), sizeof(XnartLight), sizeof(XnartFrame), sizeof(XnartCounters));
  return 0; }
'''
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "s.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(td, "s"), os.path.join(td, "s.c")], check=True)
        out = subprocess.run([os.path.join(td, "s")], capture_output=True, text=True, check=True).stdout.split()
    want = [records.TRIANGLE_DTYPE.itemsize, C.sizeof(records.XnartMaterial), C.sizeof(records.XnartLight),
            C.sizeof(records.XnartFrame), C.sizeof(records.XnartCounters)]
    assert [int(x) for x in out] == want


def _compare_trees(ot, xt):
    """Walk the oracle's (depth-first numbered) and libxnart's (breadth-first numbered) trees together."""
    stack = [(0, 0)]
    leaves = 0
    while stack:
        a, b = stack.pop()
        assert np.array_equal(ot["boxes"][a, :3], xt["lo"][b]) and np.array_equal(ot["boxes"][a, 3:], xt["hi"][b])
        if ot["first_child"][a] >= 0:
            assert xt["w1"][b] < 0, "internal in oracle, leaf in libxnart"
            ci = -1 - xt["w1"][b]
            assert np.array_equal(ot["cons"][a, 3:], xt["cons_max"][ci])
            assert np.array_equal(ot["cons"][a, :3], xt["lo"][b]), "a child's Min never undershoots its parent's"
            for c in range(8):
                stack.append((ot["first_child"][a] + c, xt["w0"][b] + c))
        else:
            assert xt["w1"][b] >= 0
            n = ot["ref_count"][a]
            assert xt["w1"][b] == n
            # same items, same order, and the same position in the depth-first leaf layout
            assert xt["w0"][b] == ot["ref_start"][a]
            assert np.array_equal(ot["refs"][ot["ref_start"][a]: ot["ref_start"][a] + n], xt["refs"][xt["w0"][b]: xt["w0"][b] + n])
            leaves += 1
    return leaves


@pytest.mark.parametrize("name", ["crate", "spheres", "grid", "soup", "many"])
def test_native_builder_equals_oracle_trees(built_lib, oracle_mod, name):
    s = {"crate": scenes.crate_scene, "spheres": scenes.spheres_scene, "grid": lambda: scenes.grid_scene(90),
         "soup": lambda: scenes.soup_scene(60000, half=12.0), "many": scenes.many_objects_scene}[name]()
    o = oracle_mod.Oracle()
    scenes.upload(o, s)
    x = capi.Xnart(devices=())  # host-only context: builds and exports trees, cannot render
    scenes.upload(x, s)
    total = 0
    for mid in range(len(s.meshes)):
        ot, xt = o.tree(mid), x.tree(mid)
        assert len(ot["boxes"]) == len(xt["lo"]) and ot["depth"] == xt["depth"]
        total += _compare_trees(ot, xt)
    ot, xt = o.tree(None), x.tree(None)
    assert len(ot["boxes"]) == len(xt["lo"])
    _compare_trees(ot, xt)
    if name == "many":
        assert len(xt["lo"]) > 1, "more than 20 objects must split the scene octree"
    assert total >= len(s.meshes)


def test_conservative_box_overshoot_exists(built_lib):
    """The 1-ulp child-Max overshoot the traversal guards against really occurs."""
    s = scenes.grid_scene(90)
    x = capi.Xnart(devices=())
    scenes.upload(x, s)
    t = x.tree(0)
    internal = np.where(t["w1"] < 0)[0]
    over = sum(bool((t["cons_max"][-1 - t["w1"][n]] > t["hi"][n]).any()) for n in internal)
    assert over > 0


def test_error_behaviour(built_lib):
    x = capi.Xnart(devices=())
    s = scenes.crate_scene()
    f = s.frame(32, 24)
    # render before build / without a device
    with pytest.raises(capi.XnartError) as e:
        x.render_frame(f)
    assert e.value.code == records.XNART_ERR_NO_DEVICE
    # unbounded recursion guard
    c = np.zeros((60, 3), np.float32)
    tris = scenes.make_triangles(c, c, c)
    with pytest.raises(capi.XnartError) as e:
        x.add_mesh(tris, box=np.zeros(6, np.float32))
    assert e.value.code == records.XNART_ERR_OCTREE_DEPTH
    # bad mesh id, texture flag without texels
    with pytest.raises(capi.XnartError) as e:
        x.add_object([5], np.eye(4), np.eye(4), np.zeros(6))
    assert e.value.code == records.XNART_ERR_INVALID_ARGUMENT
    with pytest.raises(capi.XnartError) as e:
        x.add_mesh(scenes.crate_triangles(), use_texture=True, argb=None, box=np.zeros(6))
    assert e.value.code == records.XNART_ERR_INVALID_ARGUMENT
    assert capi.version().startswith("libxnart")


def test_no_device_is_loud(built_lib):
    """On a box without a GPU, asking for a device context must fail — never fall back to the CPU."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(capi.XnartError) as e:
        capi.Xnart(devices=(0,))
    assert e.value.code == records.XNART_ERR_NO_DEVICE


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "xna-ray-trace_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")) or fn == "Makefile":
                text = open(os.path.join(dirpath, fn), errors="ignore").read()
                assert "import oracle" not in text and "from oracle" not in text, fn
                assert "libxnart_oracle" not in text and "xnart_oracle" not in text, fn
                assert "../oracle" not in text and "oracle/" not in text, fn


@pytest.mark.parametrize("name", ["grid", "soup", "spheres"])
def test_expansion_records_describe_the_trees(xnart_lib, name):
    """The 128-byte node records the persistent kernel expands (device_types.h: 12 child planes, consMax, 8 child
    descriptors) against the breadth-order trees that are verified node for node against the oracle: every
    child box is the product of its per-axis halves, bit for bit, and every descriptor names the right leaf
    run / record.  Host-only context: no GPU needed."""
    from xna_ray_trace_b200 import scenes

    s = {"grid": lambda: scenes.grid_scene(60), "soup": lambda: scenes.soup_scene(20000, half=14.0), "spheres": scenes.spheres_scene}[name]()
    x = xnart_lib.Xnart(devices=())
    scenes.upload(x, s)
    checked = 0
    for mid in range(len(s.meshes)):
        t = x.tree(mid)
        rec, ib, rb = x.inner_records(mid)
        inner = np.nonzero(t["w1"] < 0)[0]
        assert len(rec) == len(inner)
        for n in inner:
            k, f = -1 - int(t["w1"][n]), int(t["w0"][n])
            r = rec[k]
            desc = r[16:32].view(np.int32).reshape(8, 2)
            assert np.array_equal(r[12:15], t["cons_max"][k])
            assert (r[0], r[4], r[8]) == tuple(t["lo"][n])          # the node's own Min is child 0's Min
            for c in range(8):
                i, j, kk = c >> 2, (c >> 1) & 1, c & 1
                lo = (r[0 + 2 * i], r[4 + 2 * j], r[8 + 2 * kk])
                hi = (r[1 + 2 * i], r[5 + 2 * j], r[9 + 2 * kk])
                assert lo == tuple(t["lo"][f + c]) and hi == tuple(t["hi"][f + c])
                if t["w1"][f + c] >= 0:
                    assert tuple(desc[c]) == (rb + int(t["w0"][f + c]), int(t["w1"][f + c]))
                else:
                    assert tuple(desc[c]) == (ib + (-1 - int(t["w1"][f + c])), -1)
            checked += 1
    assert checked > (0 if name == "spheres" else 50) or name == "spheres"
    x.close()


def test_reference_arm_prints_the_bench_contract():
    """`bench.py --impl reference` (the oracle on the host cores) prints ONE JSON line with the contract's keys and the
    same `config` dictionary as the own arm (bench.bench_config), so that the driver compares like with like."""
    import json
    import subprocess
    import sys

    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "Mrays/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    sys.path.insert(0, ROOT)
    import bench
    from xna_ray_trace_b200 import scenes

    assert d["config"] == bench.bench_config(scenes.crate_scene(), 128)
    assert d["config"]["workload"] == "free_crate_1080p_depth8"
