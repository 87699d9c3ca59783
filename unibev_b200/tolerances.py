"""Stated parity tolerances of the encoder output ``fused_bev_embed`` (post-LayerNorm, O(1)) against the oracle.

One definition, imported by the tests, quoted by bench.py's JSON line and DESIGN.md section 4.
BASELINE.json north_star: "encoder BEV-feature output within rtol 1e-3 of the reference"; SURVEY.md 8(d) adds atol 1e-4
for the elements near zero.
"""
TOL_FP32 = dict(rtol=1e-3, atol=1e-4)     # default class 'fp32' (3xTF32 projections, fp32 sampling): north_star's tolerance
TOL_FP16 = dict(rtol=1e-3, atol=5e-3)     # opt-in class 'fp16' (fp16 operands / value maps): 50x looser, does NOT meet north_star


def describe(precision):
    t = TOL_FP32 if precision == 'fp32' else TOL_FP16
    return f"rtol {t['rtol']:g} / atol {t['atol']:g}"
