"""Generates the committed golden fixtures.  Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py

1. scene_checksums.json -- particle counts and SHA-256 of the positions produced by the REFERENCE'S OWN
   Source/SceneManager.cpp (compiled unmodified into oracle/_ref/libsf_refscene.so by oracle/Makefile),
   for every scene at several resolutions.  /root/reference does not exist on the GPU box, so these
   checksums are what pins the scene generators there.
2. oracle_dambreak_res12.npz -- outputs of the CPU oracle on a tiny Dambreak (pins the oracle against
   itself across machines / compilers: any drift in libm or flags shows up here).
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_binding as ob  # noqa: E402


def main():
    out = {"source": "/root/reference/Source/SceneManager.cpp compiled unmodified (oracle/_ref)",
           "screenshots": {"DoubleDambreak@24": 23958, "CubeDrop@24": 13824, "SphereDrop@24": 7145}, "scenes": {}}
    for res in (24, 48, 100, 161):
        for name in ob.SCENES:
            p = ob.default_params(res, name)
            pos = ob.ref_scene(p.particleRadius, name)
            assert pos is not None, "oracle/_ref is missing (needs /root/reference)"
            out["scenes"][f"{name}@{res}"] = {"n": int(len(pos)), "sha256": hashlib.sha256(pos.tobytes()).hexdigest()}
    with open(os.path.join(HERE, "scene_checksums.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)

    p = ob.default_params(12, "Dambreak")
    pos = ob.scene(p)
    orc = ob.Oracle(p, pos, boundary_seed=0)
    cnt, ids = orc.neighbors()
    dts = [orc.advance()]
    rho1, acc1, x1, v1, cell1 = orc.density(), orc.accel(), orc.positions(), orc.velocities(), orc.cell_index()
    for _ in range(199):
        dts.append(orc.advance())
    np.savez_compressed(os.path.join(HERE, "oracle_dambreak_res12.npz"), pos0=pos, nbr_count=cnt, nbr_ids=ids,
                        rho1=rho1, acc1=acc1, x1=x1, v1=v1, cell1=cell1, dts=np.array(dts, np.float32),
                        x200=orc.positions(), v200=orc.velocities(), rho200=orc.density())
    print("wrote fixtures; N(res12) =", len(pos))


if __name__ == "__main__":
    main()
