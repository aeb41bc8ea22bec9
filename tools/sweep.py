"""Run quick_perf-style timing of the text workloads for each variant library (dev tool)."""
import glob, os, subprocess, sys
here = os.path.dirname(os.path.abspath(__file__))
libs = [("base", "")] + [(os.path.basename(p)[7:-3], p) for p in sorted(glob.glob(os.path.join(here, "..", "build_variants", "libdfl_*.so")))]
for name, path in libs:
    env = dict(os.environ)
    if path:
        env["DFL_B200_LIB"] = os.path.abspath(path)
    out = subprocess.run([sys.executable, os.path.join(here, "sweep_one.py")], env=env, capture_output=True, text=True)
    print(f"{name:8s} {out.stdout.strip()} {out.stderr.strip()[-200:] if out.returncode else ''}", flush=True)
