"""Mirage bodies with EDITED constants (what a user does in the reference's editor, src/application.rs:453-489): members of
the mirage family of csrc/host/mirage_match.hpp, so the fast-math build runs the hand-fused restatement with these constants.
Shared by the fixture generator (tests/golden/make_edited_mirages.py) and the GPU test (tests/test_gpu_units.py)."""


def bodies():
    from linon_b200.main_view import PredefinedFunction as PF
    sph, pla = PF.MirageSpherical.to_code(), PF.MiragePlane.to_code()
    sphs, plas = PF.MirageSphericalSigmoid.to_code(), PF.MiragePlaneSigmoid.to_code()
    return {
        "spherical: hotter source, moved and larger sphere":
            (1, sph.replace("200.0", "260.0").replace("0.25", "0.32").replace("-0.5, 0.5, -0.5", "-0.42, 0.45, -0.6")),
        "plane: higher, thicker layer, cooler source":
            (2, pla.replace("30.0", "24.0").replace("0.01;", "0.015;").replace("0.0, 0.1, 0.0", "0.0, 0.18, 0.0")),
        "spherical sigmoid: retuned sigmoid, colder environment":
            (3, sphs.replace("15.0", "5.0").replace("12.0", "9.0").replace("6.0", "4.5").replace("-0.5, 0.5, -0.5", "-0.55, 0.6, -0.4")),
        "plane sigmoid: hotter source, higher plane, retuned sigmoid":
            (4, plas.replace("30.0", "45.0").replace("0.0, 0.1, 0.0", "0.0, 0.14, 0.0").replace("12.0", "10.0").replace("6.0", "5.5")),
    }
