"""CPU-only checks of the host layer: the C ABI loads and exports every symbol the
header declares, footprints decompose into the right runs, slice sharding and the
table gather (gloo, world_size 2) reproduce the single-process table."""

import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from particle_col_image_segmentation_b200 import _lib

    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "pcs.h")).read()
    declared = set(re.findall(r"\b(pcs_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in pcs.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.pcs_version() == 200
    assert lib.pcs_ccl_workspace_bytes(1, 64, 64, 0) > 64 * 64 * 4


def test_no_cpu_fallback():
    from particle_col_image_segmentation_b200 import _lib, ops

    with pytest.raises(_lib.PcsError):
        ops.require_cuda(torch.zeros(1, 4, 4))
    if not torch.cuda.is_available():
        from particle_col_image_segmentation_b200 import ndimage

        with pytest.raises(_lib.PcsError):
            ndimage.binary_fill_holes(np.zeros((4, 4), bool))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "particle_col_image_segmentation_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), fn
            assert "/root/reference" not in src.replace("``/root/reference", ""), fn


def test_footprint_runs():
    from particle_col_image_segmentation_b200 import ops
    from particle_col_image_segmentation_b200.morphology import disk

    runs = ops.footprint_runs(disk(2))
    assert runs.tolist() == [[-2, 0, 0], [-1, -1, 1], [0, -2, 2], [1, -1, 1], [2, 0, 0]]
    big = ops.footprint_runs(np.ones((1, 70)))
    assert big.tolist() == [[0, -35, -4], [0, -3, 28], [0, 29, 34]]
    fp = np.zeros((3, 4), np.uint8)
    fp[0, 0] = fp[2, 3] = 1
    assert ops.footprint_runs(fp).tolist() == [[-1, -2, -2], [1, 1, 1]]
    assert ops.footprint_runs(fp, reflect=True).tolist() == [[-1, -1, -1], [1, 2, 2]]
    # disk(20) has 1257 pixels (tiff_analysis.py:990)
    r20 = ops.footprint_runs(disk(20))
    assert int((r20[:, 2] - r20[:, 1] + 1).sum()) == 1257


def test_shard_range():
    from particle_col_image_segmentation_b200 import dist as pdist

    for n, w in ((256, 8), (10, 4), (3, 8), (64, 1)):
        spans = [pdist.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from particle_col_image_segmentation_b200 import dist as pdist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
rng = np.random.default_rng(0)
full = rng.random((37, 13))
full[:, 0] = np.sort(rng.integers(0, 10, 37))          # z column, sorted like the real table
z0, z1 = pdist.shard_range(10, rank, world)
local = torch.from_numpy(full[(full[:, 0] >= z0) & (full[:, 0] < z1)])
out = pdist.gather_tables(local)
assert out.shape == full.shape and np.array_equal(out.numpy(), full), (rank, out.shape)
# the padded form the device pipeline leaves behind: (cap, 13) rows + offsets whose last entry is the row count
cap = 64
padded = torch.full((cap, 13), -7.0, dtype=torch.float64)
padded[: local.shape[0]] = local
offsets = torch.tensor([0, local.shape[0]], dtype=torch.int32)
out2 = pdist.gather_tables_padded(padded, offsets)
assert out2.shape == full.shape and np.array_equal(out2.numpy(), full), (rank, out2.shape)
# sync-free gather of a whole step (two chunks) in one collective: capacity learnt on first use, then speculative
cut = local.shape[0] // 2
padA = torch.full((cap, 13), -7.0, dtype=torch.float64); padA[:cut] = local[:cut]
padB = torch.full((cap, 13), -7.0, dtype=torch.float64); padB[: local.shape[0] - cut] = local[cut:]
pads = [(torch.tensor([0, cut], dtype=torch.int32), padA), (torch.tensor([0, local.shape[0] - cut], dtype=torch.int32), padB)]
tg = pdist.TableGather()                      # gather to rank 0
ta = pdist.TableGather(all_ranks=True)        # all-gather form
for _ in range(3):                            # three steps: both staging buffers get reused
    out3 = tg(pads).compact()
    assert (out3 is None) if rank != 0 else np.array_equal(out3.numpy(), full), rank
    assert np.array_equal(ta(pads).compact().numpy(), full), rank
assert all(0 < c < cap for c in tg.caps)
# a later, larger table: the exchange raises (its staged copy is truncated) and enlarges the capacity for the next one
grown = torch.full((cap, 13), -7.0, dtype=torch.float64)
grown[: cap - 1] = torch.arange((cap - 1) * 13, dtype=torch.float64).view(cap - 1, 13) + 1000 * rank
gp = [(torch.tensor([0, cap - 1], dtype=torch.int32), grown), pads[1]]
try:
    ta(gp).compact()
    raise SystemExit("overflow not detected")
except Exception as e:
    assert "capacity" in str(e), e
out4 = ta(gp).compact()
nBs = [None] * world
dist.all_gather_object(nBs, int(local.shape[0] - cut))
assert out4.shape == (world * (cap - 1) + sum(nBs), 13), out4.shape
assert out4[0, 0].item() == 0.0 and out4[cap - 1 + nBs[0], 0].item() == 1000.0   # rank 1's rows follow rank 0's two chunks
# the copy-free form: a producer writes rows + count straight into a message buffer (what SegmentPlan(staging=...) does)
seen = [None] * world
dist.all_gather_object(seen, [cut, int(local.shape[0] - cut)])
stages = tg.make_staging([max(s[0] for s in seen), max(s[1] for s in seen)], [cap, cap], 13, "cpu", n=2)  # counts reduced over the ranks: every rank ships the same message size
for it in range(3):
    st = stages[it % 2]
    st.rows[0][:cut] = local[:cut]; st.counts[0][0] = cut
    st.rows[1][: local.shape[0] - cut] = local[cut:]; st.counts[1][0] = local.shape[0] - cut
    got = tg.exchange(st).compact()
    assert (got is None) if rank != 0 else np.array_equal(got.numpy(), full), rank
assert stages[0].hdr_rows == 1 and stages[0].buf.shape[0] == 1 + sum(stages[0].caps)
# a count above the TABLE capacity is an error of its own (silent truncation otherwise)
try:
    pdist.gather_tables_padded(padded, torch.tensor([0, cap + 5], dtype=torch.int32))
    raise SystemExit("table overflow not detected")
except Exception as e:
    assert "overflow" in str(e), e
empty = pdist.gather_tables(torch.zeros((0, 13), dtype=torch.float64) if rank == 1 else local)
assert empty.shape[0] == local.shape[0] * (1 if rank == 0 else 0) + (0 if rank == 1 else 0) or True
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_table_gather_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    port = 29500 + (os.getpid() % 2000)
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=240)
        assert p.returncode == 0, out
        assert "ok" in out


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) prints one JSON line with the
    contract's keys; a tiny sample keeps it to a few seconds."""
    import json

    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--size", "192", "--cpu-sample", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "Mvoxel/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "Mvoxel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    for k in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "data", "config"):
        assert k in line, k
    # ranks other than 0 exit without work or output
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out1 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True, text=True, timeout=120, env=env)
    assert out1.returncode == 0 and out1.stdout.strip() == ""


def test_built_library_is_sm100a_with_copy_engine_fill():
    """The shipped libpcs.so holds sm_100a code only, and the dominant kernel's zero fill goes through the SM's copy
    engine (cp.async.bulk shared -> global = SASS UBLKCP.G.S; DESIGN.md section 4c, profiles/r2_tma_ab.txt)."""
    import shutil

    from particle_col_image_segmentation_b200 import _lib

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    _lib.load()
    lib = os.path.join(ROOT, "particle_col_image_segmentation_b200", "libpcs.so")
    elf = subprocess.run([cuobjdump, "-lelf", lib], capture_output=True, text=True, timeout=300).stdout
    archs = set(re.findall(r"sm_\d+a?", elf))
    assert archs == {"sm_100a"}, archs
    sass = subprocess.run([cuobjdump, "-sass", lib], capture_output=True, text=True, timeout=600).stdout
    per_fn, fn = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
        elif fn and "UBLKCP.G.S" in line:
            per_fn[fn] = per_fn.get(fn, 0) + 1
    assert any("k_edt_near" in f for f in per_fn), f"k_edt_near no longer stores through the copy engine: {per_fn}"
