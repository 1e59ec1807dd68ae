"""Shared helpers of the test-suite: fixtures on disk, hashing, golden access."""
import hashlib
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
REF = os.environ.get("GPVIEW_REF", "/root/reference")
HAVE_REF = os.path.exists(os.path.join(REF, "files", "cessna.obj")) and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libgpvref.so"))

_tmp_cache = {}


def mesh_path(name, tmpdir):
    """Path of a fixture mesh.  'cessna' is re-written from tests/golden/cessna_mesh.npz with %.9g (float32 round-trips
    exactly through strtof), so the GPU box -- which has no /root/reference -- loads it through the product's own loader."""
    if name == "cessna":
        p = os.path.join(str(tmpdir), "cessna.obj")
        if not os.path.exists(p):
            from gpview_b200 import meshgen
            z = np.load(os.path.join(GOLD, "cessna_mesh.npz"))
            meshgen.write_obj(p, z["V"], z["F"])
        return p
    for ext in (".obj", ".off"):
        p = os.path.join(GOLD, "meshes", name + ext)
        if os.path.exists(p):
            return p
    raise FileNotFoundError(name)


def golden(case):
    with open(os.path.join(GOLD, case + ".json")) as f:
        info = json.load(f)
    z = np.load(os.path.join(GOLD, case + ".npz"))
    return info, z


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def l1_bits(state_bytes):
    """the packed form the fixtures store: bit-plane of (inside), bit-plane of (boundary)"""
    return np.packbits(np.stack([(state_bytes == 127), (state_bytes == 254)]).astype(np.uint8))


GOLDEN_CASES = [("cessna", 8, 4), ("cessna", 64, 4), ("sphere", 32, 4), ("torus", 32, 4), ("block", 32, 4), ("cad", 32, 4),
                ("cessna", 128, 8), ("cessna", 256, 16)]
