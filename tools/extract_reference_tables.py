#!/usr/bin/env python
"""Extracts the colorimetric DATA TABLES (CIE 1931 observer, Smits-style RGB basis spectra, CIE daylight
components, sun/atmosphere curves) from the reference's Haskell sources into a binary npz that the stand-in
loader uses. These are published measurement tables, not code; the reference keeps them as list literals in
Spectrum.hs:509-1139 and SunSky.hs:128-158. Run in the build container only (needs /root/reference):

    python tools/extract_reference_tables.py
"""
import re
import sys
from pathlib import Path

import numpy as np

REF = Path("/root/reference/src/lib/Graphics/Bling")
OUT = Path(__file__).resolve().parent.parent / "bling_b200" / "data" / "spectral_tables.npz"

NUM = r"[-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[eE][-+]?\d+)?"


def list_after(src: str, name: str, start_pat: str = None) -> np.ndarray:
    """first [...] literal after `name =` (or after start_pat)"""
    m = re.search(start_pat or (r"^" + re.escape(name) + r"\s*=.*?\[", ), src, re.S | re.M)
    if not m:
        raise KeyError(name)
    i = src.index("[", m.start())
    j = src.index("]", i)
    return np.array([float(x) for x in re.findall(NUM, src[i + 1:j])], np.float64)


def main():
    spec = (REF / "Spectrum.hs").read_text()
    sky = (REF / "SunSky.hs").read_text()
    d = {}
    for n in ("cieXValues", "cieYValues", "cieZValues"):
        d[n] = list_after(spec, n, r"^" + n + r"\s*=\s*\[")
        assert len(d[n]) == 471, (n, len(d[n]))
    for n in ("rgbIllumWhite", "rgbIllumCyan", "rgbIllumMagenta", "rgbIllumYellow", "rgbIllumRed", "rgbIllumGreen",
              "rgbIllumBlue", "rgbReflWhite", "rgbReflCyan", "rgbReflMagenta", "rgbReflYellow", "rgbReflRed",
              "rgbReflGreen", "rgbReflBlue"):
        d[n] = list_after(spec, n, r"^" + n + r"\s*=\s*rgbFunc")
        assert len(d[n]) == 32, (n, len(d[n]))
    for n in ("cieS0", "cieS1", "cieS2"):
        d[n] = list_after(spec, n, r"^" + n + r"\s*=\s*mkSpd'")
        assert len(d[n]) == 54, (n, len(d[n]))
    d["solCurve"] = list_after(sky, "solCurve", r"^solCurve\s*=\s*mkSpd'")
    assert len(d["solCurve"]) == 38
    m = re.search(r"^koCurve.*?ls\s*=\s*\[(.*?)\].*?as\s*=\s*\[(.*?)\]", sky, re.S | re.M)
    d["koCurve_l"] = np.array([float(x) for x in re.findall(NUM, m.group(1))])
    d["koCurve_a"] = np.array([float(x) for x in re.findall(NUM, m.group(2))])
    assert len(d["koCurve_l"]) == len(d["koCurve_a"]) == 64
    m = re.search(r"^kwaCurve.*?ls\s*=\s*\[(.*?)\].*?as\s*=\s*\[(.*?)\]", sky, re.S | re.M)
    d["kwaCurve_l"] = np.array([float(x) for x in re.findall(NUM, m.group(1))])
    d["kwaCurve_a"] = np.array([float(x) for x in re.findall(NUM, m.group(2))])
    assert len(d["kwaCurve_l"]) == len(d["kwaCurve_a"]) == 13
    d["kgCurve_l"] = np.array([759.0, 760.0, 770.0, 771.0])
    d["kgCurve_a"] = np.array([0.0, 3.0, 0.210, 0.0])
    OUT.parent.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(OUT, **d)
    print("wrote", OUT, {k: v.shape for k, v in d.items()})


if __name__ == "__main__":
    sys.exit(main())
