"""CPU: the host-side work distribution of the tensor path (persistent CTAs over a flattened
(row block, column tile) list; slots of the partial records; two-phase split for multi-GPU overlap)."""
import ctypes
import random

import pytest

import debug_lib
from wav2vec_contr_loss_b200 import _cabi
from wav2vec_contr_loss_b200.functional import make_problem


def _sched(lib, T, P, U, cta, rb):
    b, e = ctypes.c_int64(), ctypes.c_int64()
    f, l = ctypes.c_int32(), ctypes.c_int32()
    rc = lib.supcon_debug_sched(T, P, U, cta, rb, ctypes.byref(b), ctypes.byref(e), ctypes.byref(f), ctypes.byref(l))
    assert rc == 0
    return b.value, e.value, f.value, l.value


def test_cta_ranges_partition_the_unit_list(lib_built):
    lib = debug_lib.load()
    rng = random.Random(0)
    for _ in range(40):
        row_blocks, T = rng.randint(1, 40), rng.randint(1, 70)
        U = row_blocks * T
        P = rng.randint(1, min(U, 300))
        prev_end = 0
        owners = [[] for _ in range(row_blocks)]
        for c in range(P):
            b, e, _, _ = _sched(lib, T, P, U, c, 0)
            assert b == prev_end and e >= b            # contiguous, ordered, no gaps
            prev_end = e
            for rb in range(b // T, (e - 1) // T + 1 if e > b else b // T):
                owners[rb].append(c)
        assert prev_end == U
        for rb in range(row_blocks):
            _, _, first, last = _sched(lib, T, P, U, 0, rb)
            assert owners[rb] == list(range(first, last + 1))   # slot index = cta - first is dense


@pytest.mark.parametrize("n,row_offset,n_rows", [(65536, 0, 65536), (65536, 8192, 8192), (4096, 0, 4096),
                                                  (1000, 0, 1000), (16384, 4096, 4096), (2048, 512, 768)])
def test_plan_slots_bound_every_row_block(lib_built, n, row_offset, n_rows):
    lib = debug_lib.load()
    prob = make_problem(n, 256, _cabi.BF16, tau=0.07, similarity=_cabi.COSINE, topk=15, alpha=0.0,
                        row_offset=row_offset, n_rows=n_rows, flags=_cabi.FLAG_UNIT_ROWS)
    out = (ctypes.c_int32 * 16)()
    assert lib.supcon_debug_plan(ctypes.byref(prob), out, 16) == 0
    fP, fT, fslots, bP, bT, bslots, two_phase, lP, rP, lslots, frb, brb, blP, brP, blslots, blT = list(out)
    assert fT == (n + 127) // 128 and bT == (n + 63) // 64
    assert frb == ((n_rows + 127) // 128 + 1) // 2 and brb == (n_rows + 127) // 128
    assert 1 <= fP <= frb * fT and 1 <= bP <= brb * bT
    for P, T, rbs, slots in ((fP, fT, frb, fslots), (bP, bT, brb, bslots)):
        worst = max(_sched(lib, T, P, rbs * T, 0, rb)[3] - _sched(lib, T, P, rbs * T, 0, rb)[2] + 1 for rb in range(rbs))
        assert worst <= slots
    assert two_phase == int(n_rows < n and row_offset % 128 == 0 and n_rows % 128 == 0)
    if two_phase:
        assert lP >= 1 and rP >= 1 and lslots >= 1 and fslots >= lslots + 1
        # backward: the own-column window in 64-column tiles; both phases' slots fit the reserved partial records
        assert blT == n_rows // 64 and blP >= 1 and brP >= 1 and blslots >= 1 and bslots >= blslots + 1
        worst_l = max(_sched(lib, blT, blP, brb * blT, 0, rb)[3] - _sched(lib, blT, blP, brb * blT, 0, rb)[2] + 1
                      for rb in range(brb))
        rT = bT - blT
        worst_r = max(_sched(lib, rT, brP, brb * rT, 0, rb)[3] - _sched(lib, rT, brP, brb * rT, 0, rb)[2] + 1
                      for rb in range(brb))
        assert worst_l <= blslots and worst_l + worst_r <= bslots


def test_plan_rejects_problems_off_the_tensor_path(lib_built):
    lib = debug_lib.load()
    out = (ctypes.c_int32 * 12)()
    prob = make_problem(4096, 256, _cabi.F32, tau=0.07, similarity=_cabi.COSINE)
    assert lib.supcon_debug_plan(ctypes.byref(prob), out, 12) == -2
    prob = make_problem(4096, 128, _cabi.BF16, tau=0.07, similarity=_cabi.COSINE, flags=_cabi.FLAG_UNIT_ROWS)
    assert lib.supcon_debug_plan(ctypes.byref(prob), out, 12) == -2
    # cosine without the caller's unit-rows promise: z is taken as given on the exact path
    prob = make_problem(4096, 256, _cabi.BF16, tau=0.07, similarity=_cabi.COSINE)
    assert lib.supcon_debug_plan(ctypes.byref(prob), out, 12) == -2
    prob = make_problem(4096, 256, _cabi.BF16, tau=0.07, similarity=_cabi.GEODESIC)
    assert lib.supcon_debug_plan(ctypes.byref(prob), out, 12) == 0
