"""The C-ABI library: loads, exports every symbol include/spnerf_b200.h declares, and its structs
match the ctypes mirrors.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest

import spnerf_b200
from spnerf_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "spnerf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(spnerf_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _cabi.lib()
    names = declared_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/spnerf_b200.h but not exported"


def test_abi_version_and_struct_sizes():
    lib = _cabi.lib()          # _declare() raises on a struct size mismatch
    assert lib.spnerf_abi_version() == 1
    sizes = (ctypes.c_int32 * len(_cabi.STRUCTS))()
    lib.spnerf_struct_sizes(sizes)
    assert [ctypes.sizeof(s) for s in _cabi.STRUCTS] == list(sizes)


def test_plan_sizes_host_only():
    lib = _cabi.lib()
    cfg = _cabi.NetConfig(feat=512, layers=8, skip_layer=4, mapping=1, sem=1, num_sem_classes=3, emb_dim=3, beta=0,
                          t_dim=4)
    s = _cabi.NetSizes()
    assert lib.spnerf_net_sizes(ctypes.byref(cfg), ctypes.byref(s)) == 0
    assert (s.n_out, s.in_dim, s.tile_points) == (11, 63, 128)
    assert s.fwd_steps > 100 and s.bwd_steps > 100 and s.fwd_blob_bytes > 5_000_000
    # the weight stream covers every MAC of the network at least once: 2.69 M weights in fp16
    assert s.fwd_blob_bytes >= 2 * 2_690_000
    assert lib.spnerf_mlp_wgrad_workspace_bytes(ctypes.byref(cfg)) > 0


@pytest.mark.parametrize("kw", [dict(feat=128), dict(feat=384), dict(layers=6), dict(skip_layer=3), dict(num_sem_classes=9, emb_dim=9),
                                dict(mapping=1, num_sem_classes=8, emb_dim=8),
                                dict(mapping=1, num_sem_classes=6, emb_dim=6, beta=1, t_dim=8)])
def test_unsupported_configurations_are_refused(kw):
    base = dict(feat=512, layers=8, skip_layer=4, mapping=0, sem=1, num_sem_classes=3, emb_dim=3, beta=0, t_dim=4)
    base.update(kw)
    cfg = _cabi.NetConfig(**base)
    s = _cabi.NetSizes()
    assert _cabi.lib().spnerf_net_sizes(ctypes.byref(cfg), ctypes.byref(s)) == 2     # SPNERF_ERR_UNSUPPORTED


def test_encoded_input_wider_than_the_slab_is_planned():
    """--mapping with the CLI-default 5 semantic classes: 65 encoded columns (one beyond the 64-column input slab)."""
    lib = _cabi.lib()
    for c, beta in ((5, 0), (5, 1), (7, 0)):
        cfg = _cabi.NetConfig(feat=512, layers=8, skip_layer=4, mapping=1, sem=1, num_sem_classes=c, emb_dim=c, beta=beta,
                              t_dim=4)
        s = _cabi.NetSizes()
        assert lib.spnerf_net_sizes(ctypes.byref(cfg), ctypes.byref(s)) == 0
        assert s.in_dim == 60 + c and s.n_out == 8 + beta + c


def test_class_default_width_is_planned():
    """feat = 256 (the SPNeRF class default, models/spnerf.py:163): half of every tile."""
    lib = _cabi.lib()
    small, big = _cabi.NetSizes(), _cabi.NetSizes()
    for feat, s in ((256, small), (512, big)):
        cfg = _cabi.NetConfig(feat=feat, layers=8, skip_layer=4, mapping=1, sem=1, num_sem_classes=3, emb_dim=3, beta=0, t_dim=4)
        assert lib.spnerf_net_sizes(ctypes.byref(cfg), ctypes.byref(s)) == 0
    assert small.n_out == big.n_out == 11
    assert small.save_slabs_per_tile < big.save_slabs_per_tile and small.grad_slabs_per_tile < big.grad_slabs_per_tile
    assert 0.2 < small.fwd_blob_bytes / big.fwd_blob_bytes < 0.35          # ~4x fewer weights in the trunk


def test_bad_arguments_are_refused_without_a_device():
    lib = _cabi.lib()
    assert lib.spnerf_net_sizes(None, None) == 1
    assert lib.spnerf_sample_coarse(None, None, None, 4, 64, None, None) == 1
    assert lib.spnerf_composite_fwd(None, None) == 1
    assert lib.spnerf_losses(None, None) == 1
    assert lib.spnerf_mlp_fwd(None, None) == 1


def _step_table(cfg, backward):
    lib = _cabi.lib()
    lib.spnerf_debug_step_table.restype = ctypes.c_int
    lib.spnerf_debug_step_table.argtypes = [ctypes.POINTER(_cabi.NetConfig), ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
    tab = (ctypes.c_int32 * (8 * 384))()
    n = lib.spnerf_debug_step_table(ctypes.byref(cfg), backward, tab, 384)
    assert 0 < n <= 384
    keys = ("n", "col", "a_slab", "ksteps", "first", "last", "lane", "half")
    return [dict(zip(keys, tab[8 * i:8 * i + 8])) for i in range(n)]


@pytest.mark.parametrize("kw", [dict(mapping=0, sem=1), dict(mapping=1, sem=1), dict(mapping=0, sem=0),
                                dict(mapping=1, sem=1, beta=1), dict(feat=256, mapping=1, sem=1),
                                dict(feat=256, mapping=0, sem=1, beta=1), dict(mapping=1, sem=1, num_sem_classes=5, emb_dim=5)])
@pytest.mark.parametrize("backward", [0, 1])
def test_step_table_keeps_accumulation_order_deterministic(kw, backward):
    """The two MMA issuer warps must never accumulate into the same accumulator columns inside a phase: the
    tensor pipe executes MMAs in arrival order, so sharing columns across issuers makes the fp32 summation
    order (and with it the result bits) depend on timing.  Every chunk (column range) of a phase belongs to
    one issuer, its overwriting item comes first, and every phase ends with exactly one `last` item."""
    base = dict(feat=512, layers=8, skip_layer=4, mapping=0, sem=1, num_sem_classes=3, emb_dim=3, beta=0, t_dim=4)
    base.update(kw)
    if not base["sem"]:
        base.update(num_sem_classes=0, emb_dim=0)
    steps = _step_table(_cabi.NetConfig(**base), backward)
    assert steps[-1]["last"] == 1
    phase = []
    for s in steps:
        assert s["lane"] in (0, 1) and s["n"] % 16 == 0 and s["col"] + s["n"] <= 512
        phase.append(s)
        if s["last"]:
            owner, started = {}, set()
            for t in phase:
                rng = (t["col"], t["n"])
                for (c0, n0), lane in owner.items():
                    if t["col"] < c0 + n0 and c0 < t["col"] + t["n"]:
                        assert (c0, n0) == rng and lane == t["lane"], (t, c0, n0, lane)
                owner[rng] = t["lane"]
                assert bool(t["first"]) == (rng not in started) or not t["first"], t
                started.add(rng)
            assert {t["lane"] for t in phase} <= {0, 1}
            phase = []
    assert not phase


@pytest.mark.parametrize("kw", [dict(mapping=0, sem=1), dict(mapping=1, sem=1, beta=1), dict(feat=256, mapping=1, sem=1)])
def test_split_phases_of_the_backward_step_table(kw):
    """Split phases (csrc/net_plan.h, SPNERF_SPLIT_BWD): a full-width phase is issued as quarter-width chunks, the
    first accumulator half first.  Exactly one step of such a phase carries the `half` mark (both issuers commit to the
    half-done barrier when they pass it, so a second mark would over-arrive); everything up to the mark accumulates
    into columns below the half, everything after it at or above, and each issuer owns one chunk on either side."""
    base = dict(feat=512, layers=8, skip_layer=4, mapping=0, sem=1, num_sem_classes=3, emb_dim=3, beta=0, t_dim=4)
    base.update(kw)
    steps = _step_table(_cabi.NetConfig(**base), 1)
    half_cols = base["feat"] // 2
    phase, n_split, n_free = [], 0, 0
    for s in steps:
        phase.append(s)
        if not s["last"]:
            continue
        marks = [i for i, t in enumerate(phase) if t["half"] & 1]
        frees = [i for i, t in enumerate(phase) if t["half"] & 2]
        n_free += len(frees)
        assert len(marks) <= 1 and len(frees) <= 1, phase
        if frees:
            # nothing after the mark reads the A slabs under the first half's columns, and the mark sits in the second half
            assert marks and frees[0] > marks[0]
            assert all(t["a_slab"] >= half_cols // 64 for t in phase[frees[0] + 1:]), phase
        if marks:
            n_split += 1
            before, after = phase[:marks[0] + 1], phase[marks[0] + 1:]
            assert after, "the mark cannot be the last step of a phase"
            assert all(t["col"] + t["n"] <= half_cols for t in before), before
            assert all(t["col"] >= half_cols for t in after), after
            for part in (before, after):
                assert {t["lane"] for t in part} == {0, 1}
                assert len({(t["col"], t["n"]) for t in part}) == 2          # one chunk per issuer
        phase = []
    assert n_split >= 8 and n_free == 8      # the eight trunk phases carry the early-free mark
