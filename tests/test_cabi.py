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


@pytest.mark.parametrize("kw", [dict(feat=256), dict(layers=6), dict(skip_layer=3), dict(num_sem_classes=9, emb_dim=9),
                                dict(mapping=1, num_sem_classes=5, emb_dim=5)])
def test_unsupported_configurations_are_refused(kw):
    base = dict(feat=512, layers=8, skip_layer=4, mapping=0, sem=1, num_sem_classes=3, emb_dim=3, beta=0, t_dim=4)
    base.update(kw)
    cfg = _cabi.NetConfig(**base)
    s = _cabi.NetSizes()
    assert _cabi.lib().spnerf_net_sizes(ctypes.byref(cfg), ctypes.byref(s)) == 2     # SPNERF_ERR_UNSUPPORTED


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
    keys = ("n", "col", "a_slab", "ksteps", "first", "last", "lane", "early")
    return [dict(zip(keys, tab[8 * i:8 * i + 8])) for i in range(n)]


@pytest.mark.parametrize("kw", [dict(mapping=0, sem=1), dict(mapping=1, sem=1), dict(mapping=0, sem=0),
                                dict(mapping=1, sem=1, beta=1)])
@pytest.mark.parametrize("backward", [0, 1])
def test_step_table_issue_order_is_safe(kw, backward):
    """Two MMA issuers may accumulate into the same accumulator columns only in an order the tensor pipe
    preserves (mlp_pack.cu end_phase): an item that shares columns with a chunk's overwriting first item
    and is owned by the other issuer must either wait for the hand-off barrier (early = 3, directly
    behind the first item) or sit >= 4 ring positions later (its ring stage is only refilled after the
    first item has retired).  Early items read activation slabs 0..3 and write columns < 256 only."""
    base = dict(feat=512, layers=8, skip_layer=4, mapping=0, sem=1, num_sem_classes=3, emb_dim=3, beta=0, t_dim=4)
    base.update(kw)
    if not base["sem"]:
        base.update(num_sem_classes=0, emb_dim=0)
    steps = _step_table(_cabi.NetConfig(**base), backward)
    assert steps[-1]["last"] == 1
    n_early = 0
    for p, s in enumerate(steps):
        assert s["lane"] in (0, 1) and s["n"] % 16 == 0 and s["col"] + s["n"] <= 512
        if s["early"]:
            n_early += 1
            assert s["a_slab"] < 4 and s["col"] + s["n"] <= 256
            if s["early"] == 3:
                assert steps[p - 1]["first"] and steps[p - 1]["early"] == 1 and steps[p - 1]["lane"] != s["lane"]
        if s["first"]:
            synced = False      # the other issuer has waited for this item's hand-off (later items follow in its program order)
            for q in range(p + 1, min(p + 4, len(steps))):
                t = steps[q]
                if steps[q - 1]["last"]:
                    break
                overlap = t["col"] < s["col"] + s["n"] and s["col"] < t["col"] + t["n"]
                if overlap and t["lane"] != s["lane"]:
                    if q == p + 1 and t["early"] == 3:
                        synced = True
                    assert synced, (p, q, s, t)
    # early items only ever lead a phase
    for p, s in enumerate(steps):
        if s["early"] and p > 0 and not steps[p - 1]["early"]:
            assert steps[p - 1]["last"] == 1
    assert n_early > 0
