"""Algorithmic FLOP model of the march (SURVEY 8(d) counting convention) — used for the roofline line of bench.py.

Convention: 1 FLOP per scalar fp32 + - * / sqrt min max and per transcendental; negation, abs, compares, selects,
integer work and loads are free; an FMA counts 2.  Counted on the WGSL as written (src/main_view.wgsl:246-270),
scalarised.  Work the kernel skips is NOT credited: with plane culling on, only the plane tests and the
ray-triangle tests that were actually executed are counted (the kernel reports both).

Per accepted step (as written, adaptive variant): 9+9+7 stage inputs, 18 combine, 9 diff, 1 threshold, 10 blend,
3 step_dir, 3+1 advance, 1 restore test = 71, plus 4 field evaluations of F FLOPs each.
Fixed-step variant (Enhance, or adaptive off): the kernel drops diff / threshold / restore (11 FLOPs): 60 + 4F.
Rejected attempt: 53 + 4F.  Pre-hit step: +15 (normalize + length of step_dir).
Ray-triangle test with e1, e2 precomputed: 45.  Culling: every pre-hit step pays the safe-distance bookkeeping (8: two
L1 norms, margin, compare operand, subtraction); a step whose bound expired ("culling check") additionally pays
normalize + length of step_dir (15), the reach (3), the bound reset (2) and 6 per plane group (3 FMA).  Without culling
every pre-hit step pays 15 + 45 per face.
"""
from __future__ import annotations

# F per predefined function, enum order of src/functions.rs
FIELD_FLOPS = {1: 84, 2: 79, 3: 94, 4: 89, 5: 4, 6: 4, 7: 32, 8: 8, 9: 6}

STEP_ADAPTIVE = 71
STEP_FIXED = 60
STEP_REJECTED = 53
PREHIT_EXTRA = 15      # normalize(step_dir) + length(step_dir), src/main_view.wgsl:260
FACE_TEST = 45
PLANE_TEST = 6
SAFE_STEP = 8          # per pre-hit step with culling on
CHECK_EXTRA = 5        # reach + bound reset, per culling check


# Hand-fused mirage restatements (csrc/device/mirage_fused.cuh, fast math, LINON_FUSED_RK4): executed work per accepted step.
# Stage 1 is the identity (not evaluated); the step's origin contributes once (PREV), stages 2 and 3 evaluate STAGE, stage 4
# is accumulated straight into the Runge-Kutta sum (LAST).  Counted like the rest: FMA = 2, rsqrt / exp2 / sqrt / rcp = 1,
# sign transfer, compares and selects = 0.
#   spherical sigmoid  PREV  6 (scaled origin, 3 FMA) + 5 (q) + 1 (rsqrt) + 2 + 1 (E = exp2(k0 - q rq)) + 1 (b~) + 1 (rcp)
#                            + 2 (g_in) + 1 (D g_in) = 20
#                      STAGE 6 (c' = origin + s h k, 3 FMA) + 5 (q) + 1 (rsqrt) + 3 (E) + 1 (a~) + 1 (rcp) + 2 (eta)
#                            + 5 (k . c') + 1 (d) + 1 (min) + 2 (1 - cr^2) + 1 (eta^2) + 2 (kk) + 1 (sqrt) + 2 (cs) + 1 (cs rq)
#                            + 9 (eta k + csr c') = 44;   LAST: the last 9 become 12 (6 FMA into the sum) = 47
#   spherical linear   E (3) becomes min(1, q rq) (2); eta = (b~ (1 / a~)) g_in is 5 instead of 4 per stage, g_in 4 instead of 5: 18 / 44 / 47
#   plane sigmoid      PREV 2 (y') + 2 (E) + 1 + 1 + 2 + 1 = 9;  STAGE 2 (y') + 2 (E) + 1 (a~) + 1 (rcp) + 2 (eta) + 1 (min) + 2 + 1 + 2
#                      + 1 (sqrt) + 2 (cs) + 4 (eta k + (0, cs, 0)) = 22;  LAST 25
#   plane linear       E (2) becomes min(1, |y'|) (1), eta as above: 7 / 22 / 25
FUSED_FLOPS = {1: (44, 47, 18), 2: (22, 25, 7), 3: (44, 47, 20), 4: (22, 25, 9)}
FUSED_STEP_FIXED = 12 + 2 + 3 + 6    # k1 + 2 k2 + 2 k3 (6 FMA), min3 of the discriminants, / 6, advance (3 FMA); +10 blend when field_weight != 1
SAFE_STEP_V2 = 8                     # step_dir (3), |s|_1 (2), cost (2), budget - cost (1)


QUIET_STEP = 6  # LINON_OPT_QUIET_ZONE: a straight step is the position update alone (3 FMA)


QUIET_STEP_COLLAPSED = 0  # ... and nothing at all where runs are collapsed (LINON_COLLAPSE_RUNS, march_kernels.cuh: advance_coord): the
                          # steps of a run are not executed one by one, and the few real steps and jumps per run are not counted


def fused_march_flops(counters, fn_id: int, field_weight_is_one: bool = True, quiet_steps: int = 0, collapsed_runs: bool = False) -> float:
    """Executed-work model of the fast-mode mirage path in fixed-step variants (Enhance, adaptive off).

    quiet_steps: accepted steps taken as straight steps (counter 15 of linon_debug_profile with LINON_OPT_QUIET_ZONE on).
    They are credited with the position update only; the planning of a run (zone entry, clear steps: ~150 FLOPs, a few
    per ray) is not counted at all, and the per-step safe-distance bookkeeping only for the steps that are not straight."""
    f_stage, f_last, f_prev = FUSED_FLOPS[fn_id]
    step = FUSED_STEP_FIXED + (0 if field_weight_is_one else 10) + 2 * f_stage + f_last + f_prev
    full_steps = counters.accepted_steps - quiet_steps
    flops = full_steps * step + quiet_steps * (QUIET_STEP_COLLAPSED if collapsed_runs else QUIET_STEP) + counters.face_tests * FACE_TEST
    flops += counters.prehit_steps * SAFE_STEP_V2 * (full_steps / max(counters.accepted_steps, 1))
    flops += counters.culling_checks * (PREHIT_EXTRA + CHECK_EXTRA + 3) + counters.plane_tests * PLANE_TEST
    return float(flops)


def march_flops(counters, fn_id: int, adaptive: bool, n_faces: int = 36, field_flops: int | None = None) -> float:
    """counters: object with accepted_steps, prehit_steps, rejected_steps, face_tests, plane_tests."""
    F = FIELD_FLOPS[fn_id] if field_flops is None else field_flops
    step = (STEP_ADAPTIVE if adaptive else STEP_FIXED) + 4 * F
    flops = counters.accepted_steps * step
    flops += counters.rejected_steps * (STEP_REJECTED + 4 * F)
    flops += counters.face_tests * FACE_TEST
    if counters.plane_tests:  # culling on: normalize/length only run on culling checks
        flops += counters.prehit_steps * SAFE_STEP
        flops += counters.culling_checks * (PREHIT_EXTRA + CHECK_EXTRA) + counters.plane_tests * PLANE_TEST
    else:
        flops += counters.prehit_steps * PREHIT_EXTRA
    return float(flops)


def as_written_flops(counters, fn_id: int, n_faces: int = 36) -> float:
    """FLOPs the reference shader would execute for the same step counts (every face on every pre-hit step).
    Reported for context only; never used as a roofline numerator."""
    F = FIELD_FLOPS[fn_id]
    return float(counters.accepted_steps * (STEP_ADAPTIVE + 4 * F) + counters.rejected_steps * (STEP_REJECTED + 4 * F)
                 + counters.prehit_steps * (PREHIT_EXTRA + n_faces * 51))


# Overlay stencil, algorithmic bytes per pixel: mode 1 = 4 colour in + 16 endpoint in + 4 colour out
# + 4 exponent out; mode 0 = 4 in + 4 out.
OVERLAY_BYTES = {0: 8, 1: 28}
MARCH_BYTES_PER_PIXEL = 24  # 4 rgba8 + 16 endpoint + 4 hit id
