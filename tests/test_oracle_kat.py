"""The oracle against the hand-derived known-answer tests (tests/golden/kats.json).

Each KAT's expected value was derived by hand from the RTL (see tests/golden/make_kats.py); the
oracle must reproduce it with BOTH adders and, inside the hardware limits, with BOTH walkers."""
import numpy as np
import pytest

from helpers import kat_arrays, oracle_cfg, L
from oracle import oracle as O


def _case_ids(k):
    return [c["name"] for c in k]


def test_kats_present(kats):
    assert len(kats["cases"]) >= 16 and len(kats["ring"]) >= 1 and len(kats["lines"]) >= 1


@pytest.mark.parametrize("idx", range(16))
def test_oracle_reproduces_kat(kats, idx):
    assert len(kats["cases"]) == 16, "update the parametrisation"
    c = kats["cases"][idx]
    W, FI, x = kat_arrays(c)
    wl, fl = L.pack_streams(W, FI, c["D"])
    cfg = oracle_cfg(c["D"], c["K"], c["S"], c["missing"], c["F"], W.shape[0])
    want = np.array(c["expect"], dtype=np.uint32)
    got_a = O.scores(cfg, wl, fl, x)
    got_b = O.scores(cfg, wl, fl, x, literal_adder=True)
    assert (got_a == want).all(), (c["name"], got_a, want)
    assert (got_b == want).all(), (c["name"], "literal adder", got_b, want)
    if c.get("rtl_literal", True):
        got_c = O.scores_literal(cfg, wl, fl, x)
        assert (got_c == want).all(), (c["name"], "address-literal walker", got_c, want)
    else:
        # the address-literal walker keeps the RTL's behaviour, which is broken for an early leaf with levels left
        # (DTPU.sv:596,712): it must DIFFER from the build-defined rule on this case — that is the documented decision
        assert (O.scores_literal(cfg, wl, fl, x) != want).any(), c["name"]
    assert (O.scores_blocked(cfg, wl, fl, x) == want).all()
    # threaded path returns the same words
    assert (O.scores(cfg, wl, fl, x, threads=3) == want).all()


def test_ring_kat(kats):
    for r in kats["ring"]:
        got = O.ring_combine([np.array(p, dtype=np.uint32) for p in r["partials"]])
        assert (got == np.array(r["expect"], dtype=np.uint32)).all(), r["name"]


def test_result_line_kat(kats):
    for r in kats["lines"]:
        got = O.result_lines(np.array(r["scores"], dtype=np.uint32))
        assert got.tolist() == r["expect_lines"], r["name"]
        assert L.result_lines(np.array(r["scores"], dtype=np.uint32).view(np.float32)).view(np.uint32).tolist() == r["expect_lines"]


def test_labels_rule():
    s = np.array([0.0, -0.0, 1e-30, -1e-30, 3.5, -2.0], dtype=np.float32).view(np.uint32)
    assert O.labels(s).tolist() == [0, 0, 1, 0, 1, 0]


def test_out_of_contract_refused():
    # a REACHABLE feature index beyond the tuple is refused, not guessed; below an early leaf anything goes
    W = np.zeros((1, 7), dtype=np.uint32)
    x = np.zeros((1, 4), dtype=np.uint32)
    for fis, ok in (([7, 0, 0], False), ([0, 0, 7], False), ([1 << 14, 7, 7], True)):
        FI = np.array([fis], dtype=np.uint16)
        wl, fl = L.pack_streams(W, FI, 2)
        if ok:
            O.scores(oracle_cfg(2, 1, 1, 0, 4, 1), wl, fl, x)
        else:
            with pytest.raises(RuntimeError):
                O.scores(oracle_cfg(2, 1, 1, 0, 4, 1), wl, fl, x)
