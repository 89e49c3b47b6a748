"""CPU: what the documents promise about the library's switches matches the source."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CAPI = os.path.join(ROOT, "midas-journal-740_b200", "csrc", "cuberille_capi.cu")


def test_every_tuning_knob_is_documented_and_read_once():
    src = open(CAPI).read()
    knobs = re.findall(r'env_int\("(CUB_[A-Z0-9_]+)",\s*(-?\d+)', src)
    assert len(knobs) >= 10
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    for name, default in knobs:
        assert f"`{name}`" in doc, f"{name} is read by cub_create but missing from INTEGRATION.md"
    # the environment is read in cub_create only (no getenv on the hot path): env_int's own getenv is the one call
    assert src.count("getenv(") == 1
    create = src[src.index("int cub_create("):]
    create = create[:create.index("\n}\n")]
    for name, _ in knobs:
        assert name in create, f"{name} is not read in cub_create"


def test_fused_kernel_has_no_debug_switch_left():
    src = open(CAPI).read() + open(os.path.join(ROOT, "midas-journal-740_b200", "csrc", "k_fused.cuh")).read()
    assert "CUB_FUSE_DBG" not in src and "dbg" not in src
