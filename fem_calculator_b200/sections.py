"""Closed-form section-property front end.

Fast stand-in for the reference's ``calculate_section_properties(section_type,
params, rotate)`` (BeamSolver.py:32-82), which runs a sectionproperties FE warping
analysis per physical group.  sectionproperties is neither installed here nor part
of the hot path: the path's per-group input record is the returned 8-tuple

    (A, I_x, I_y, J, kappa_y, kappa_z, c_y_max, c_z_max)

with I_x = ixx_c (second moment about the picture's horizontal axis), I_y = iyy_c,
kappa_y = A_sx / A, kappa_z = A_sy / A, c_y_max / c_z_max the extreme fibre
distances along the picture's x / y axes, and ``rotate`` swapping the (I, kappa, c)
pairs (BeamSolver.py:73-79).  Same section-type strings and parameter keys as the
seven section dialogs (BeamSolver.py:104-137).  Thin-wall / textbook formulas, so
values agree with the FE analysis to engineering accuracy (J and shear areas to a
few %), not to round-off; unknown types and failures return eight zeros like the
reference (BeamSolver.py:55-57,80-82).
"""
from __future__ import annotations

import math
import warnings


def _rect_J(long_side: float, short_side: float) -> float:
    a, b = max(long_side, short_side), min(long_side, short_side)
    if a <= 0 or b <= 0:
        return 0.0
    return a * b**3 * (1.0 / 3.0 - 0.21 * (b / a) * (1.0 - b**4 / (12.0 * a**4)))


def _props(section_type: str, p: dict):
    if section_type == "rectangular section":
        d, b = p["d"], p["b"]
        A = b * d
        return A, b * d**3 / 12, d * b**3 / 12, _rect_J(d, b), 5 / 6 * A, 5 / 6 * A, b / 2, d / 2
    if section_type == "circular section":
        d = p["d"]
        A = math.pi * d**2 / 4
        I = math.pi * d**4 / 64
        return A, I, I, 2 * I, 6 / 7 * A, 6 / 7 * A, d / 2, d / 2
    if section_type == "hollow circular section":
        d, t = p["d"], p["t"]
        di = d - 2 * t
        A = math.pi * (d**2 - di**2) / 4
        I = math.pi * (d**4 - di**4) / 64
        m = di / d
        ks = 6 * (1 + m**2) ** 2 / (7 * (1 + m**2) ** 2 + 20 * m**2)
        return A, I, I, 2 * I, ks * A, ks * A, d / 2, d / 2
    if section_type == "I section":
        d, b, tf, tw = p["d"], p["b"], p["t_f"], p["t_w"]
        hw = d - 2 * tf
        A = 2 * b * tf + hw * tw
        ixx = (b * d**3 - (b - tw) * hw**3) / 12
        iyy = (2 * tf * b**3 + hw * tw**3) / 12
        J = (2 * b * tf**3 + (d - tf) * tw**3) / 3
        return A, ixx, iyy, J, 5 / 6 * 2 * b * tf, d * tw, b / 2, d / 2
    if section_type == "C section":
        d, b, tf, tw = p["d"], p["b"], p["t_f"], p["t_w"]
        hw = d - 2 * tf
        A = 2 * b * tf + hw * tw
        cx = (2 * b * tf * b / 2 + hw * tw * tw / 2) / A
        ixx = (b * d**3 - (b - tw) * hw**3) / 12
        iyy = 2 * (tf * b**3 / 12 + b * tf * (b / 2 - cx) ** 2) + hw * tw**3 / 12 + hw * tw * (cx - tw / 2) ** 2
        J = (2 * b * tf**3 + (d - tf) * tw**3) / 3
        return A, ixx, iyy, J, 5 / 6 * 2 * b * tf, d * tw, max(cx, b - cx), d / 2
    if section_type == "L section":
        d, b, t = p["d"], p["b"], p["t"]
        A = t * (d + b - t)
        cx = (d * t * t / 2 + (b - t) * t * (t + (b - t) / 2)) / A
        cy = (d * t * d / 2 + (b - t) * t * t / 2) / A
        ixx = t * d**3 / 12 + d * t * (d / 2 - cy) ** 2 + (b - t) * t**3 / 12 + (b - t) * t * (cy - t / 2) ** 2
        iyy = d * t**3 / 12 + d * t * (cx - t / 2) ** 2 + t * (b - t) ** 3 / 12 + (b - t) * t * (t + (b - t) / 2 - cx) ** 2
        J = (d + b - t) * t**3 / 3
        return A, ixx, iyy, J, 5 / 6 * b * t, 5 / 6 * d * t, max(cx, b - cx), max(cy, d - cy)
    if section_type == "hollow box section":
        d, b, t = p["d"], p["b"], p["t"]
        A = b * d - (b - 2 * t) * (d - 2 * t)
        ixx = (b * d**3 - (b - 2 * t) * (d - 2 * t) ** 3) / 12
        iyy = (d * b**3 - (d - 2 * t) * (b - 2 * t) ** 3) / 12
        J = 2 * t * (b - t) ** 2 * (d - t) ** 2 / (b + d - 2 * t)
        return A, ixx, iyy, J, 2 * b * t, 2 * d * t, b / 2, d / 2
    return None


_FILLET_KEYS = ("r", "r_out", "r_r", "r_t")      # fillet radii of the I / C / box / L dialogs (BeamSolver.py:104-137)


def calculate_section_properties(section_type: str, params: dict, rotate: bool = False):
    """Same signature and return record as BeamSolver.py:32.  The closed forms are for sharp corners: a non-zero
    fillet radius is reported (warning) and left out, it is not silently dropped."""
    try:
        rounded = [k for k in _FILLET_KEYS if params.get(k)]
        if rounded:
            warnings.warn(f"closed-form section properties ignore the fillet radii {rounded} of '{section_type}' "
                          f"(sharp-corner formulas; use the reference's sectionproperties front end for rounded corners)",
                          RuntimeWarning, stacklevel=2)
        r = _props(section_type, params)
        if r is None:
            print(f"Warning: Unknown section type '{section_type}'.")
            return (0,) * 8
        A, ixx, iyy, J, asx, asy, cy_, cz_ = r
        ky = asx / A if A > 0 else 0
        kz = asy / A if A > 0 else 0
        I_y, I_z = ixx, iyy
        if rotate:
            I_y, I_z = I_z, I_y
            ky, kz = kz, ky
            cy_, cz_ = cz_, cy_
        return A, I_y, I_z, J, ky, kz, cy_, cz_
    except Exception as e:  # reference swallows every failure (BeamSolver.py:80-82)
        print(f"Error in section properties for {section_type} with params {params}: {e}")
        return (0,) * 8
