"""Synthetic families (workloads/synth.py): index-addressable and deterministic, and every family
is something the reference can actually solve (checked with the oracle on a few problems)."""
import numpy as np
import pytest

from workloads import synth


@pytest.mark.parametrize("name", sorted(synth.WORKLOADS))
def test_any_rank_can_generate_any_range(name):
    a_dom, a_ctx = synth.generate(name, 40, seed=9, first=synth.CHUNK - 17)
    b_dom, b_ctx = synth.generate(name, synth.CHUNK + 23, seed=9, first=0)
    assert np.array_equal(a_dom, b_dom[synth.CHUNK - 17:]) and np.array_equal(a_ctx, b_ctx[synth.CHUNK - 17:])
    c_dom, _ = synth.generate(name, 40, seed=10, first=synth.CHUNK - 17)
    assert not np.array_equal(a_dom, c_dom)


@pytest.mark.parametrize("name", ["sor1d", "cg1", "fimmel", "esced", "test10i", "test12i"])
def test_families_are_solvable(name, port):
    n = 64
    dom, ctx = synth.generate(name, n, seed=3)
    _, st, h, stats = port.bench_dense(0, n, dom, ctx, synth.bignum(name), **synth.options(name))
    assert (st == 0).mean() > 0.9 and stats.pivots > 0
    assert len(np.unique(h)) > 1
