#!/usr/bin/env python
"""Same-box A/B of step-kernel build variants (how v5 -> v6 and the inlining choices of DESIGN.md §5 were decided).

  python tools/ab_kernel_variants.py build  A="" B="-DRAFTING_MINBLOCKS=6" C="-DRAFTING_NST2=4"
      compiles one librafting_b200.so per variant (extra nvcc flags after the '=') into variants/<name>.so and
      restores the default build afterwards;
  gpurun -- 'python tools/ab_kernel_variants.py run A B C'
      on the GPU box: for two rounds, copies each variant over rafting_b200/librafting_b200.so and runs
      `bench.py --steps 30 --warmup 5 --no-e2e --no-cpu`, printing value and roofline fraction per run, then the
      secondary rates (configs #3/#5) once per variant.  Alternating the variants inside one call removes the
      box-to-box variance (a few per cent) that separate calls have.
variants/ is scratch (git-ignored via *.so; delete it after the run)."""
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "rafting_b200", "librafting_b200.so")
VAR = os.path.join(ROOT, "variants")


def build(specs):
    sys.path.insert(0, ROOT)
    from rafting_b200 import _build
    os.makedirs(VAR, exist_ok=True)
    for spec in specs:
        name, _, flags = spec.partition("=")
        os.environ["RAFTING_NVCC_EXTRA"] = flags
        _build.build(force=True)
        shutil.copy(LIB, os.path.join(VAR, name + ".so"))
        print("built", name, flags or "(default flags)")
    os.environ["RAFTING_NVCC_EXTRA"] = ""
    _build.build(force=True)


def run(names):
    keep = os.path.join(VAR, "_default.so")
    shutil.copy(LIB, keep)
    try:
        for rnd in (1, 2):
            for n in names:
                shutil.copy(os.path.join(VAR, n + ".so"), LIB)
                out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "30", "--warmup", "5", "--no-e2e", "--no-cpu"],
                                     capture_output=True, text=True).stdout
                line = [l for l in out.splitlines() if l.startswith("{")]
                if line:
                    b = json.loads(line[-1])
                    print(f"{n} round {rnd}: {b['value'] / 1e9:.2f} G acks/s, frac {b['roofline']['frac']:.4f}", flush=True)
                else:
                    print(n, "bench failed", flush=True)
        for n in names:
            shutil.copy(os.path.join(VAR, n + ".so"), LIB)
            out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "bench_secondary.py")], capture_output=True, text=True).stdout
            for l in out.splitlines():
                if l.startswith("{"):
                    r = json.loads(l)
                    print(f"{n} {r['config']}: {r['kernel_ms_per_step_median']:.4f} ms/step, frac {r['roofline']['frac']:.4f}", flush=True)
    finally:
        shutil.copy(keep, LIB)


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "build":
        build(sys.argv[2:])
    elif len(sys.argv) >= 3 and sys.argv[1] == "run":
        run(sys.argv[2:])
    else:
        print(__doc__)
