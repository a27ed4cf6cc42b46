#!/usr/bin/env python
"""Build-time variants of the flagship kernel / its graph image, A/B-ed on ONE GPU box (same clocks, same neighbours):
   python tests/perf/flagship_variants.py build      (here, no GPU: writes lib/libstmp_<name>.so)
   python tests/perf/flagship_variants.py run        (GPU box: bench.py --no-* with STMP_LIB pointing at each, two rounds)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
LIBDIR = os.path.join(ROOT, "pytorch_geometric_temporal_b200", "lib")
VARIANTS = {
    "pf": ["-DSTMP_TC_PREFETCH=1"],
    "pf_gu1": ["-DSTMP_TC_PREFETCH=1", "-DSTMP_TC_GUNROLL=1"],
}

if sys.argv[1] == "build":
    from pytorch_geometric_temporal_b200 import build
    for name, flags in VARIANTS.items():
        print(build.build(extra_flags=flags, out=os.path.join(LIBDIR, f"libstmp_{name}.so")))
else:
    res = {}
    for rnd in range(2):
        for name in ["base"] + list(VARIANTS):
            env = dict(os.environ)
            if name != "base":
                env["STMP_LIB"] = os.path.join(LIBDIR, f"libstmp_{name}.so")
            out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "30", "--warmup", "5", "--no-cpu", "--no-spmm", "--no-refgpu",
                                  "--no-hostwin", "--no-train"], env=env, capture_output=True, text=True).stdout
            res.setdefault(name, []).append(round(json.loads(out)["value"]))
    print(json.dumps(res))
