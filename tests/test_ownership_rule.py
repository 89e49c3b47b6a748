"""CPU: the closed form of the first-touch ownership rule that k_sweep.cuh evaluates bit-parallel
(csrc/k_sweep.cuh::corner_owners) against its definition, for all 256 block configurations and for every
way the image border can clip the 2x2x2 block."""
import itertools


def activates(i, p):
    """block voxel p (raster order qz,qy,qx) is inside and has an outside face neighbour inside the block:
    vertexHasQuad of that voxel for this corner (txx:164-173)"""
    return bool(i[p] and not (i[p ^ 1] and i[p ^ 2] and i[p ^ 4]))


def definition(i, valid):
    for p in range(8):
        if valid[p] and activates(i, p):
            return p
    return None


def closed_form(i):
    if not i[0]:
        return next((p for p in range(1, 8) if i[p]), None)
    if not (i[1] and i[2] and i[4]):
        return 0
    if not (i[3] and i[5]):
        return 1
    if not i[6]:
        return 2
    if not i[7]:
        return 3
    return None


def test_closed_form_equals_first_activating_voxel():
    for bits in range(256):
        i = [(bits >> p) & 1 for p in range(8)]
        assert closed_form(i) == definition(i, [True] * 8), bits


def test_clipped_blocks_alias_hands_over_to_twin():
    """out-of-image block voxels read the clamped (edge-replicated) value; the kernel evaluates the closed form
    on those 8 values and moves a claim of an out-of-image alias to its in-image twin (only needed on the LOW
    side, where the alias precedes the twin in raster order)"""
    n = 0
    for clip in itertools.product(["none", "low", "high"], repeat=3):  # x, y, z
        def is_valid(p):
            return all(not ((clip[a] == "low" and ((p >> a) & 1) == 0) or (clip[a] == "high" and ((p >> a) & 1) == 1))
                       for a in range(3))

        def twin(p, low_only=False):
            t = p
            for a in range(3):
                if clip[a] == "low":
                    t |= 1 << a
                if clip[a] == "high" and not low_only:
                    t &= ~(1 << a)
            return t

        valid_pos = [p for p in range(8) if is_valid(p)]
        for vals in itertools.product([0, 1], repeat=len(valid_pos)):
            v = dict(zip(valid_pos, vals))
            i = [v[twin(p)] for p in range(8)]
            want = definition(i, [is_valid(p) for p in range(8)])
            got = closed_form(i)
            got = None if got is None else twin(got, low_only=True)  # the kernel's fix-up: low side only
            assert got == want, (clip, vals)
            n += 1
    assert n == 416
