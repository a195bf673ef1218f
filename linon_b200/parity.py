"""Parity metrics of the north-star: hit-id agreement (raw and per surface group), colour within
1/255 per channel, endpoint within 1e-3 relative, high-Lyapunov pixels excluded and counted separately.

Pure numpy bookkeeping over two sets of frame buffers; it computes nothing of the path itself.
"""
from __future__ import annotations

import numpy as np


def surface_groups(n_faces: int = 36) -> np.ndarray:
    """Canonical surface id per Cornell-box face.

    The two triangles of a quad are coplanar and share a diagonal, the light quad (faces 8,9) lies in the ceiling
    plane (faces 6,7) and the block footprints (faces 2-5) lie in the floor plane (faces 0,1): which of those wins a
    tie is rounding noise, so hit ids are compared per surface group.  Index -1 (no hit) maps to -1."""
    g = np.arange(n_faces) // 2
    if n_faces == 36:
        g[[2, 3, 4, 5]] = 0   # short / tall block footprints -> floor
        g[[8, 9]] = 3         # ceiling light -> ceiling
    return g


def canonical_hits(hit_id: np.ndarray, n_faces: int = 36) -> np.ndarray:
    g = surface_groups(n_faces)
    out = np.full(hit_id.shape, -1, np.int32)
    m = hit_id >= 0
    out[m] = g[hit_id[m]]
    return out


def high_lyapunov_mask(exponents_scaled: np.ndarray, threshold: float = 0.8) -> np.ndarray:
    """Pixels the reference itself paints as diverging: exp(scaling*lambda - 5) >= 0.8 (src/overlay.wgsl:127-129)."""
    return exponents_scaled >= threshold


def disagreement_mask(a: dict, b: dict, n_faces: int = 36) -> np.ndarray:
    """Pixels where two renderings differ in surface group or by more than 1/255 in a colour channel.  Used to mark
    pixels that the ORACLE ITSELF cannot decide in adaptive mode (fp32 vs fp64 vs a 1-ulp perturbed run): adaptive
    step sequences are driven by rounding noise, so such pixels are excluded and counted, like the
    high-Lyapunov ones."""
    g = canonical_hits(a["hit_id"], n_faces) != canonical_hits(b["hit_id"], n_faces)
    c = np.abs(a["rgba8"].astype(np.int16) - b["rgba8"].astype(np.int16)).max(axis=-1) > 1
    return g | c


def compare_frames(test: dict, ref: dict, exclude: np.ndarray | None = None, sample: np.ndarray | None = None,
                   n_faces: int = 36) -> dict:
    """test / ref: dicts with hit_id [H,W] i32, rgba8 [H,W,4] u8, endpoint [H,W,>=3] f32.

    exclude: bool [H,W] of pixels left out (high Lyapunov exponent); sample: bool [H,W] of pixels that were rendered
    by both sides (sub-sampled oracle runs).  Returns counts and the three north-star fractions."""
    H, W = ref["hit_id"].shape
    use = np.ones((H, W), bool) if sample is None else sample.copy()
    n_sampled = int(use.sum())
    n_excluded = 0
    if exclude is not None:
        n_excluded = int((use & exclude).sum())
        use &= ~exclude
    n = int(use.sum())
    th, rh = test["hit_id"][use], ref["hit_id"][use]
    raw_same = th == rh
    grp_same = canonical_hits(th, n_faces) == canonical_hits(rh, n_faces)
    tc = test["rgba8"][use].astype(np.int16)
    rc = ref["rgba8"][use].astype(np.int16)
    cdiff = np.abs(tc - rc).max(axis=-1)
    te = test["endpoint"][use][..., :3].astype(np.float64)
    re = ref["endpoint"][use][..., :3].astype(np.float64)
    err = np.linalg.norm(te - re, axis=-1)
    rel = err / np.maximum(np.linalg.norm(re, axis=-1), 1e-12)
    finite = np.isfinite(rel)
    both_nonfinite = ~np.isfinite(te).all(axis=-1) & ~np.isfinite(re).all(axis=-1)
    ep_ok = (rel <= 1e-3) | both_nonfinite
    matched = grp_same
    nm = max(int(matched.sum()), 1)
    return dict(
        pixels=n_sampled,
        excluded_high_lyapunov=n_excluded,
        compared=n,
        hit_raw_agree=float(raw_same.mean()) if n else 1.0,
        hit_group_agree=float(grp_same.mean()) if n else 1.0,
        colour_within_1=float((cdiff[matched] <= 1).sum() / nm),
        colour_max_diff=int(cdiff[matched].max()) if matched.any() else 0,
        endpoint_within_1e3=float(ep_ok[matched].sum() / nm),
        endpoint_max_rel=float(rel[matched & finite].max()) if (matched & finite).any() else 0.0,
        endpoint_bit_equal=float((test["endpoint"][use][..., :3] == ref["endpoint"][use][..., :3]).all(axis=-1).mean()) if n else 1.0,
    )


def passes_north_star(m: dict, check_endpoint: bool = True) -> bool:
    ok = m["hit_group_agree"] >= 0.999 and m["colour_within_1"] >= 0.999
    if check_endpoint:
        ok = ok and m["endpoint_within_1e3"] >= 0.999
    return ok
