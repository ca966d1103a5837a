#!/usr/bin/env python
"""Regenerates tests/golden/reference_hashes.json FROM THE UNMODIFIED REFERENCE (oracle/_ref, built by
`make -C oracle ref` in the container that has /root/reference).  The reference ships no golden vectors or
tests of its own, so these are outputs of the reference itself on seeded grids
(tests/grids.py), reduced to SHA-256 digests per LOD level; the GPU box, where /root/reference does not exist,
checks the CUDA path and the CPU restatement against them.

    python tests/golden/make_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

import golden_hash  # noqa: E402
import grids  # noqa: E402
import harness  # noqa: E402


def main():
    ref = harness.reference()
    out = {"_comment": "SHA-256 of the reference's own output (see make_golden.py); regenerate, never edit", "grids": {}}
    for name in sorted(list(grids.SMALL) + list(grids.MEDIUM)):
        dist, mat, blend = (grids.SMALL.get(name) or grids.MEDIUM[name])()
        g = ref.grid_from_dense(dist, mat, blend)
        s, _ = ref.polygonize(g)
        entry = {"input_sha256": golden_hash.input_digest(dist, mat, blend), "stats": [int(v) for v in ref.surface_stats(s)], "levels": []}
        for l in range(ref.surface_levels(s)):
            entry["levels"].append(golden_hash.level_digests(ref.surface_level(s, l)))
        out["grids"][name] = entry
        ref.surface_destroy(s); ref.grid_destroy(g)
        print(name, entry["stats"][:4], [lv["counts"] for lv in entry["levels"]])
    with open(os.path.join(HERE, "reference_hashes.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
