#!/usr/bin/env python3
"""Pulls the reference's `inline.load [[ ... ]]` C bodies out of its Lua sources, AS THEY ARE
(only `inline`'s %% escape is undone), into oracle/_ref/*.inc so that oracle/Makefile can compile
them against the shim.  TEST INFRASTRUCTURE ONLY; the outputs are build artefacts (git-ignored).

    python3 extract_inline.py /root/reference oracle/_ref
"""
import os
import sys

# (output name, file, first line of the body, last line of the body) -- 1-based, the lines
# strictly between `inline.load [[` / `inline.preamble [[` and the closing `]]`
BLOCKS = [
    ("pp_preamble", "opticalflow_model.lua", 329, 340),          # comp() for qsort
    ("pp_fmax", "opticalflow_model.lua", 343, 385),              # postProcessImage 'max' (mode filter)
    ("pp_fmed", "opticalflow_model.lua", 389, 433),              # postProcessImage 'med'
    ("radial_depth", "test_opticalflow.lua", 151, 188),          # radial(): flow -> depth
    ("enlarge_mask", "depth_estimation_api.lua", 78, 128),       # enlargeMask
    ("c2p_mask", "radial/cartesian2polar.lua", 17, 39),          # getC2PMask buildMask
    ("p2c_mask", "radial/cartesian2polar.lua", 62, 85),          # getP2CMask buildMask
    ("flow2depth", "radial/radial_opticalflow_display.lua", 17, 52),
]
# plain C++ function bodies (not inline.load strings): the lines strictly inside the braces
CPP_BODIES = [
    # ARdroneAPI::computeDepthMapFromFlow (signature on :99, closing brace on :140)
    ("drone_depth", "ardrone/ardrone_api.cpp", 100, 139, "computeDepthMapFromFlow"),
]


def main():
    ref, out = sys.argv[1], sys.argv[2]
    os.makedirs(out, exist_ok=True)
    for name, rel, a, b in BLOCKS:
        lines = open(os.path.join(ref, rel), encoding="latin-1").read().split("\n")
        opener, closer = lines[a - 2], lines[b]
        assert "[[" in opener and "]]" in closer, (name, opener, closer)
        body = "\n".join(lines[a - 1:b]).replace("%%", "%")
        with open(os.path.join(out, name + ".inc"), "w") as f:
            f.write(body + "\n")
    for name, rel, a, b, fn in CPP_BODIES:
        lines = open(os.path.join(ref, rel), encoding="latin-1").read().split("\n")
        assert fn in lines[a - 2] and lines[a - 2].rstrip().endswith("{") and lines[b].strip() == "}", (
            name, lines[a - 2], lines[b])
        with open(os.path.join(out, name + ".inc"), "w") as f:
            f.write("\n".join(lines[a - 1:b]) + "\n")


if __name__ == "__main__":
    main()
