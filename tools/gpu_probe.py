"""Run every kernel check in its own subprocess (a trap or hang in one kernel must not poison the others) and print
one diagnostic line per check.  Usage on the GPU box: python tools/gpu_probe.py [name ...] > gpurun_out/probe.log"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(name):
    from tests.kernel_checks import CHECKS
    try:
        res = CHECKS[name]()
        print("PROBE_OK", name, json.dumps(res, default=str))
    except AssertionError as e:
        print("PROBE_FAIL", name, str(e)[:3000])
    except Exception as e:  # noqa
        print("PROBE_ERROR", name, type(e).__name__, str(e)[:3000])


def main():
    if len(sys.argv) >= 3 and sys.argv[1] == "--child":
        child(sys.argv[2])
        return
    from tests.kernel_checks import CHECKS
    names = sys.argv[1:] or list(CHECKS)
    for n in names:
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", n], capture_output=True, text=True,
                               timeout=180, cwd=ROOT)
            lines = [l for l in (r.stdout + r.stderr).splitlines() if l.strip()]
            keep = [l for l in lines if l.startswith("PROBE_") or "stad:" in l or "rror" in l]
            print(f"[{n}] rc={r.returncode}", *(keep[-6:] if keep else lines[-6:]), sep="\n  ", flush=True)
        except subprocess.TimeoutExpired:
            print(f"[{n}] TIMEOUT after 180 s", flush=True)


if __name__ == "__main__":
    main()
