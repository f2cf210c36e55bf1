"""Shared parity check: CUDA scan (through the C ABI) versus the CPU oracle on the same in-memory cohort."""
import numpy as np

from popdel_b200 import api

FLOAT_RTOL = 1e-6          # BASELINE.json: likelihood-derived values within 1e-6 relative
INT_FIELDS = ["initial_length", "iterations", "deletion_length", "filter", "window_position", "position",
              "end_position", "segment"]


def oracle_inputs(samples, rgs):
    pos, dev, off = [], [], [0]
    for s in samples:
        for rg in s.read_groups:
            pos.append(rg.pos.astype(np.uint32))
            dev.append(rg.dev.astype(np.int32))
            off.append(off[-1] + rg.pos.size)
    return (np.concatenate(pos) if pos else np.zeros(0, np.uint32),
            np.concatenate(dev) if dev else np.zeros(0, np.int32), np.array(off, dtype=np.uint64))


def run_oracle(samples, params, rgs, oracle, **kw):
    pos, dev, off = oracle_inputs(samples, rgs)
    return oracle.scan_contig(params.as_dict(), [r.as_dict() for r in rgs], off, pos, dev, len(samples), **kw)


def assert_calls_equal(got_calls, got_ps, ref_calls, ref_ps, rtol=FLOAT_RTOL):
    """Integer fields bit-exact, LR / allele frequency within rtol (relative)."""
    assert len(got_calls) == len(ref_calls), f"{len(got_calls)} calls, oracle has {len(ref_calls)}"
    for f in INT_FIELDS:
        assert np.array_equal(got_calls[f], ref_calls[f]), f"field {f} differs"
    np.testing.assert_allclose(got_calls["lr"], ref_calls["lr"], rtol=rtol, atol=0)
    np.testing.assert_allclose(got_calls["frequency"], ref_calls["frequency"], rtol=rtol, atol=0)
    assert np.array_equal(got_ps, ref_ps), "per-sample PL/LAD/DAD/FL differ"


def compare_scan_with_oracle(samples, oracle, params=None, device=0):
    params = params or api.CallParameters()
    res, rgs = api.scan_cohort(samples, params, device=device)
    ref_calls, ref_ps, n_windows = run_oracle(samples, params, rgs, oracle)
    assert res["n_windows"] == n_windows, f"windows scanned {res['n_windows']} != oracle {n_windows}"
    assert_calls_equal(res["calls"], res["per_sample"], ref_calls, ref_ps)
    return dict(n_calls=int(len(ref_calls)), n_windows=int(n_windows), n_flagged=int(res["n_flagged_windows"]),
                n_candidates=int(res["n_candidates"]), ms_screen=float(res["ms_screen"]),
                ms_genotype=float(res["ms_genotype"]))
