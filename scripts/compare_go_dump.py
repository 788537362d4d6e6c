#!/usr/bin/env python
"""Compares what the unmodified Go reference rendered (go/parity/parity_dump.go) with the CPU oracle, scene by scene:
TPF, every colour byte and every depth bit.  With --pin the reference's hashes are also copied to
tests/golden/go_reference_outputs.json, which tests/test_oracle.py then holds the oracle to — the step that turns
"parity unpinned" into pinned.

    python scripts/compare_go_dump.py DUMPDIR [--pin]
"""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402


def main():
    import scene_defs
    from oracle_binding import Oracle

    dump = sys.argv[1]
    pin = "--pin" in sys.argv[2:]
    ref = json.load(open(os.path.join(dump, "go_reference_outputs.json")))
    orc = Oracle()
    bad = 0
    for name, g in sorted(ref.items()):
        if name not in scene_defs.PINNED:
            print(f"{name}: not a pinned scene of this checkout, skipped")
            continue
        sc = scene_defs.PINNED[name]()
        res = orc.draw(sc.renderer(None), sc.objects, sc.camera)
        hp = hashlib.sha256(res["pixels"].tobytes()).hexdigest()
        hz = hashlib.sha256(res["zbuffer"].tobytes()).hexdigest()
        ok = hp == g["pixels_sha256"] and hz == g["zbuffer_sha256"] and res["tpf"] == g["tpf"]
        msg = "identical" if ok else "DIFFERENT"
        if not ok:
            bad += 1
            px = np.fromfile(os.path.join(dump, name + ".pixels"), np.uint8).reshape(sc.height, sc.width, 4)
            zb = np.fromfile(os.path.join(dump, name + ".zbuffer"), "<f4").reshape(sc.height, sc.width)
            dp = (px != res["pixels"]).any(axis=-1)
            dz = zb.view(np.uint32) != res["zbuffer"].view(np.uint32)
            msg += (f": TPF go {g['tpf']} / oracle {res['tpf']}, {int(dp.sum())} colour pixels and {int(dz.sum())} depth values differ "
                    f"of {dp.size}; max |colour diff| {int(np.abs(px.astype(int) - res['pixels'].astype(int)).max())}")
            if dz.any():
                with np.errstate(divide="ignore", invalid="ignore"):
                    rel = np.abs(zb - res["zbuffer"]) / np.maximum(np.abs(zb), 1e-30)
                msg += f", max relative depth difference {float(np.nanmax(rel[dz])):.3g}"
        print(f"{name:32s} {msg}")
    print(f"{len(ref) - bad} of {len(ref)} scenes identical to the Go reference")
    if pin and bad == 0:
        dst = os.path.join(ROOT, "tests", "golden", "go_reference_outputs.json")
        shutil.copyfile(os.path.join(dump, "go_reference_outputs.json"), dst)
        print("pinned:", dst)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
