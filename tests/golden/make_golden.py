"""Writes tests/golden/kat_mesh_sha256.json: SHA-256 of the oracle's points and cells (LITERAL mode) for the
reference's 19 CTest rows.  The reference asserts only the counts of these meshes (Testing/CuberilleTest01.cxx:
193-204) and ships no golden mesh; the hashes freeze what THIS repository's restatement produces, so that neither
the oracle nor the CUDA path can drift silently.  Run from the repo root: python tests/golden/make_golden.py"""
import hashlib
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from util import KAT, KAT_ARGS, oracle, read_fixture  # noqa: E402


def mesh_digest(points, cells):
    h = hashlib.sha256()
    h.update(points.astype("<f4").tobytes())
    h.update(cells.astype("<u8").tobytes())
    return h.hexdigest()


if __name__ == "__main__":
    O = oracle()
    out = {}
    for name, fixture, iso, n_points, n_cells, tri, proj, max_steps in KAT:
        m = O.cuberille(read_fixture(fixture).data, iso, triangles=tri, project=proj, max_steps=max_steps, **KAT_ARGS)
        assert m.points.shape[0] == n_points and m.cells.shape[0] == n_cells
        out[name] = {"points": n_points, "cells": n_cells, "sha256": mesh_digest(m.points, m.cells)}
    with open(os.path.join(os.path.dirname(__file__), "kat_mesh_sha256.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", len(out), "digests")
