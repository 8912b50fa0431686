"""Regenerates tests/golden/goldens.json.  Run in the build container (needs oracle/ built;
uses oracle/_ref — the reference's own terrain generator compiled from /root/reference —
for the `scene_ref_*` entries and oracle/_ref/libtws_ref_step.so — the reference's own
flowUpdate / flowApply / waterBrush shaders compiled from /root/reference — for the step goldens).
The step goldens are outputs of the REFERENCE SHADERS (checked here to equal the oracle's), stored as
FNV-1a-64 over the little-endian bytes in index order; they give the oracle and the GPU tests fixed,
reference-derived targets on boxes where /root/reference does not exist.
"""
import ctypes as C
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from oracle.oracle_py import Oracle, RefStep, dam_break, new_state  # noqa: E402

o = Oracle(openmp=True)
r = RefStep()          # fails loudly when the reference tree / _ref build is absent: goldens must come from the reference
out = {"step_goldens_source": "reference shaders (oracle/_ref/libtws_ref_step.so built from /root/reference/terrainwatersim/shader/*.comp); "
                              "asserted equal to oracle/tws_oracle.cpp when generated"}


def same(a, b):
    assert all(np.array_equal(x.view(np.uint8), y.view(np.uint8)) for x, y in zip(a, b)), "oracle != reference shaders"


def hashes(t, f, v):
    return {"d": f"{o.fnv1a64(np.ascontiguousarray(t[..., 3])):016x}", "F": f"{o.fnv1a64(f):016x}", "v": f"{o.fnv1a64(v.view(np.uint16)):016x}",
            "volume": float(t[..., 3].sum(dtype=np.float64))}


# per-step constants for the reference defaults (pin 'powf' differences between libms)
c = o.derive_consts(1024.0, 1024)
out["consts_default_hex"] = [float(x).hex() for x in c]

# config 1: 256x256 dam break, walled and open, after 1/10/100/1000 steps
for name, rim in (("dam256_walled", True), ("dam256_open", False)):
    h, d = dam_break(256, rim=rim)
    cc = o.derive_consts(256.0, 256)
    t, f, v = new_state(h, d)
    t2, f2, v2 = new_state(h, d)
    done = 0
    out[name] = {}
    for n in (1, 10, 100, 1000):
        r.step(t, f, v, cc, n - done)
        o.step(t2, f2, v2, cc, n - done)
        same((t, f, v), (t2, f2, v2))
        done = n
        out[name][str(n)] = hashes(t, f, v)

# config 2: reference default scene 1024x1024 (+ brush at (512,512), 100/60 per step)
scene = o.create_scene(1024)
out["scene1024"] = {"h": f"{o.fnv1a64(np.ascontiguousarray(scene[..., 0])):016x}", "d": f"{o.fnv1a64(np.ascontiguousarray(scene[..., 3])):016x}",
                    "h_min": float(scene[..., 0].min()), "h_max": float(scene[..., 0].max()), "d_sum": float(scene[..., 3].sum(dtype=np.float64)),
                    "wet": int((scene[..., 3] > 0).sum()), "d_max": float(scene[..., 3].max())}
ref_so = ROOT / "oracle" / "_ref" / "libtws_ref_terrain.so"
if ref_so.exists():
    rt = C.CDLL(str(ref_so))
    a = np.zeros((1024, 1024, 4), np.float32)
    rt.tws_ref_create_scene(C.c_uint32(231656522), 1024, C.c_float(300.0), a.ctypes.data_as(C.c_void_p))
    out["scene1024_ref"] = {"h": f"{o.fnv1a64(np.ascontiguousarray(a[..., 0])):016x}", "d": f"{o.fnv1a64(np.ascontiguousarray(a[..., 3])):016x}"}
    w = np.zeros(4096, np.float32)
    rt.tws_ref_white_noise(C.c_uint32(231656522), w.ctypes.data_as(C.c_void_p))
    out["white_noise_ref"] = f"{o.fnv1a64(w):016x}"
t = scene.copy()
f = np.zeros((1024, 1024, 4), np.float32)
v = np.zeros((1024, 1024, 2), np.float16)
t2, f2, v2 = t.copy(), f.copy(), v.copy()
done = 0
out["scene1024_brush"] = {}
for n in (1, 10, 100, 1000):
    for _ in range(n - done):
        r.brush(t, 512.0, 512.0, np.float32(100.0 / 60.0), 32.0)
        r.step(t, f, v, c, 1)
        o.brush(t2, 512.0, 512.0, np.float32(100.0 / 60.0), 32.0)
        o.step(t2, f2, v2, c, 1)
    same((t, f, v), (t2, f2, v2))
    done = n
    out["scene1024_brush"][str(n)] = hashes(t, f, v)

# config 3 initial state: 8192x8192 scene hash
s8 = o.create_scene(8192)
out["scene8192"] = {"h": f"{o.fnv1a64(np.ascontiguousarray(s8[..., 0])):016x}", "d": f"{o.fnv1a64(np.ascontiguousarray(s8[..., 3])):016x}",
                    "d_sum": float(s8[..., 3].sum(dtype=np.float64)), "wet": int((s8[..., 3] > 0).sum())}
(Path(__file__).parent / "goldens.json").write_text(json.dumps(out, indent=1, sort_keys=True) + "\n")
print(json.dumps(out, indent=1, sort_keys=True))
