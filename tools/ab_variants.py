#!/usr/bin/env python
"""A/B of compile-time variants of the library on one box: builds each variant here (nvcc), runs them there.

    python tools/ab_variants.py build name=DEF1,DEF2 name2=...     # -> no-node-comparison_b200/libnbody_b200_<name>.so
    python tools/ab_variants.py run name name2 ...                # on the GPU box: bench.py --quick for each, alternating
"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
lib = lambda n: os.path.join(ROOT, "no-node-comparison_b200", f"libnbody_b200_{n}.so")
if sys.argv[1] == "build":
    from no_node_comparison_b200.build import build_library
    for spec in sys.argv[2:]:
        name, _, defs = spec.partition("=")
        print(build_library(defines=[d for d in defs.split(",") if d], out=lib(name)))
else:
    names, rounds = sys.argv[2:], 2
    extra = os.environ.get("AB_ARGS", "--quick --steps 30 --warmup 5").split()
    res = {n: [] for n in names}
    for r in range(rounds):
        for n in names:
            env = dict(os.environ, NB_B200_LIBRARY=lib(n)) if n != "default" else dict(os.environ)
            out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + extra, env=env, capture_output=True, text=True)
            try:
                d = json.loads(out.stdout.strip().splitlines()[-1])
                res[n].append((round(d["ms_per_step"], 4), d.get("us_per_launch")))
            except Exception:
                res[n].append(("FAILED", out.stderr[-400:]))
    for n in names:
        print(n, res[n])
