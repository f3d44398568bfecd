"""Run every GPU parity check of tests/gpu_checks.py in its own subprocess (a kernel fault or an mbarrier-watchdog trap
kills only that check), with a timeout, and write gpurun_out/diag.json + one log per failing check.

    python tools/gpu_diag.py [--only name1,name2] [--timeout 180] [--pending]

--pending runs gpu_checks.CHECKS_PENDING instead (checks of code that has not been on a GPU yet; not part of `pytest -m gpu`).
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.join(ROOT, "gpurun_out")


def run_one(name):
    import gpu_checks
    import ref_gpu_checks
    import torch
    t0 = time.time()
    res = {**gpu_checks.CHECKS, **gpu_checks.CHECKS_PENDING, **ref_gpu_checks.CHECKS}[name]()
    torch.cuda.synchronize()
    print("RESULT " + json.dumps(dict(name=name, ok=True, seconds=round(time.time() - t0, 2), result=res)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check")
    ap.add_argument("--only", default="")
    ap.add_argument("--timeout", type=int, default=180)
    ap.add_argument("--pending", action="store_true")
    ap.add_argument("--stop-on-fail", action="store_true", help="stop at the first failing check (a trapping kernel costs its whole watchdog time per check)")
    ap.add_argument("--reference", action="store_true", help="run tests/ref_gpu_checks.CHECKS (parity against the staged reference on the GPU)")
    a = ap.parse_args()
    if a.check:
        return run_one(a.check)
    os.makedirs(OUT, exist_ok=True)
    import gpu_checks
    import ref_gpu_checks
    table = ref_gpu_checks.CHECKS if a.reference else gpu_checks.CHECKS_PENDING if a.pending else gpu_checks.CHECKS
    names = [n for n in table if not a.only or n in a.only.split(",")]
    summary = []
    for n in names:
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--check", n], capture_output=True, text=True, timeout=a.timeout)
            out = p.stdout + p.stderr
            line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
            if p.returncode == 0 and line:
                rec = json.loads(line[-1][7:])
            else:
                tail = "\n".join(out.strip().splitlines()[-6:])
                rec = dict(name=n, ok=False, rc=p.returncode, seconds=round(time.time() - t0, 2), tail=tail)
                open(os.path.join(OUT, f"diag_{n}.log"), "w").write(out)
        except subprocess.TimeoutExpired as e:
            rec = dict(name=n, ok=False, rc="timeout", seconds=a.timeout)
            open(os.path.join(OUT, f"diag_{n}.log"), "w").write((e.stdout or b"").decode(errors="replace") + (e.stderr or b"").decode(errors="replace"))
        summary.append(rec)
        print(json.dumps(rec), flush=True)
        if a.stop_on_fail and not rec["ok"]:
            break
    json.dump(summary, open(os.path.join(OUT, "diag_reference.json" if a.reference else "diag_pending.json" if a.pending else "diag.json"), "w"), indent=1)
    bad = [r["name"] for r in summary if not r["ok"]]
    print(f"{len(summary) - len(bad)}/{len(summary)} checks passed; failing: {bad}")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
