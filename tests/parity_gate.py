"""The north-star parity contract (BASELINE.json) as the tests apply it.

  mean CIEDE2000 <= 0.5 per frame         asserted unmodified by every parity test (`assert_mean_gate`).
  max 8-bit channel error <= 2 per frame  asserted unmodified too (`strict_max_gate`), in tests that are marked
                                          xfail(strict=False) with the reason below - NOT replaced by a looser bound.

Why the max gate cannot be met by ANY implementation that is not bit-identical to one particular torch-CPU run - including
the reference against itself (tests/test_oracle_golden.py::test_strict_max_gate_is_violated_by_the_reference_itself, measured
at BASELINE cfg2 size, same code / weights / input):

    reference fp32, 1 thread  vs  reference fp32, 8 threads :  logit rms 1.0e-6  ->  final frame max err 3,  9 values > 2
    reference fp64            vs  reference fp32            :  logit rms 1.5e-6  ->  final frame max err 4, 33 values > 2

The generator's output goes through a TRUNCATING uint8 quantiser (deoldify/filters.py:67), so any logit perturbation, however
small, flips isolated values by 1; the OpenCV YUV round trip turns a flipped U into +-2..3 of B (B = Y + 2.032 (U-128)), and the
Spline64 resize back to 1080p (negative lobes) overshoots that to 4..5.  The count of values off by more than 2 is linear in
the logit noise (117 at sigma = 1e-5, 1424 at 1e-4, 13470 at 1e-3 of 6.2 M: tools/error_budget notes in DESIGN.md), i.e. the
gate measures bit-identity of the summation order, not arithmetic quality.  What the tests bound instead, next to the mean
gate, is a REGRESSION GUARD on the share of such values at the level measured for each precision policy.
"""

NORTH_STAR_MEAN_DE00 = 0.5
NORTH_STAR_MAX_ERR = 2
XFAIL_REASON = ("max 8-bit error <= 2 needs bit-identical logits: the reference's own fp32 path violates it against itself "
                "(1 vs 8 threads: max 3; fp64 vs fp32: max 4) - see tests/parity_gate.py")


def assert_mean_gate(m, ctx=None):
    assert m["mean_de00"] <= NORTH_STAR_MEAN_DE00, (ctx, m)


def strict_max_gate(m, ctx=None):
    assert m["max_err"] <= NORTH_STAR_MAX_ERR, (ctx, m)


def assert_outlier_guard(m, share, ctx=None):
    """Regression guard (not the gate): values off by more than 2 stay below `share` of the frame."""
    assert m["n_err_gt2"] <= share * m["n_values"], (ctx, m)
