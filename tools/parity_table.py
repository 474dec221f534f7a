#!/usr/bin/env python
"""gpurun_out/r02_parity_errors.jsonl (written by the `parity_log` fixture of the GPU tests) -> profiles/r02_parity_errors.md:
the measured error of every parity comparison next to the tolerance its test asserts."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "r02_parity_errors.jsonl")
dst = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles", "r02_parity_errors.md")
rows, seen = [], set()
for line in open(src):
    r = json.loads(line)
    if r["stage"] in seen:
        continue
    seen.add(r["stage"])
    rows.append(r)
out = ["# Measured parity errors (B200, round 2) — GPU path vs the reference goldens / the pinned CPU oracle", "",
       "`max abs` / `mean abs`: |got - ref| over all compared elements; `ref rms` / `ref max`: magnitude of the reference; `rtol needed`:",
       "smallest rtol that passes at atol = 1e-4 (0 = every element already inside atol); `asserted`: [rtol, atol] of the test.", "",
       "| comparison | elements | max abs | mean abs | ref rms | ref max | rtol needed @ atol 1e-4 | asserted [rtol, atol] |",
       "|---|---:|---:|---:|---:|---:|---:|---|"]
for r in rows:
    out.append(f"| {r['stage']} | {r['n']} | {r['max_abs']:.2e} | {r['mean_abs']:.2e} | {r['ref_rms']:.3f} | {r['ref_max']:.2f} | "
               f"{r['rtol_needed_at_atol1e4']:.1e} | {r['tol'] if r['tol'] else 'recorded only'} |")
open(dst, "w").write("\n".join(out) + "\n")
print(f"{len(rows)} comparisons -> {dst}")
