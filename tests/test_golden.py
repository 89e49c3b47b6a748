"""Committed golden digests of the 19 reference CTest meshes (tests/golden/kat_mesh_sha256.json, written by
tests/golden/make_golden.py from the oracle): the oracle on CPU, the CUDA path on the GPU box."""
import json
import os
import sys

import pytest

from util import GOLDEN, KAT, KAT_ARGS, oracle, read_fixture, run_filter

sys.path.insert(0, GOLDEN)
from make_golden import mesh_digest  # noqa: E402

DIGESTS = json.load(open(os.path.join(GOLDEN, "kat_mesh_sha256.json")))


@pytest.mark.parametrize("row", KAT, ids=[r[0] for r in KAT])
def test_oracle_reproduces_golden_digests(row):
    name, fixture, iso, n_points, n_cells, tri, proj, max_steps = row
    m = oracle().cuberille(read_fixture(fixture).data, iso, triangles=tri, project=proj, max_steps=max_steps, **KAT_ARGS)
    assert DIGESTS[name] == {"points": n_points, "cells": n_cells, "sha256": mesh_digest(m.points, m.cells)}


@pytest.mark.gpu
@pytest.mark.parametrize("row", KAT, ids=[r[0] for r in KAT])
def test_cuda_reproduces_golden_digests(row):
    name, fixture, iso, n_points, n_cells, tri, proj, max_steps = row
    mesh = run_filter(read_fixture(fixture), iso, triangles=tri, project=proj, max_steps=max_steps, **KAT_ARGS)
    assert DIGESTS[name]["sha256"] == mesh_digest(mesh.points, mesh.cells)
