"""Run every tcgen05 cross-check case in its own process (a trap / illegal access must not hide the others)."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import tc_cases
sel = [int(a) for a in sys.argv[1:]] or range(len(tc_cases.CASES))
for i in sel:
    code = "import sys; sys.path.insert(0, %r); from tests import tc_cases as t; import json; r = t.run_case(%d, True); r['bad'] = t.check(r); print('RES ' + json.dumps(r))" % (ROOT, i)
    try:
        p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=180)
        line = [l for l in p.stdout.splitlines() if l.startswith("RES ")]
        if line:
            r = json.loads(line[0][4:])
            print(i, "OK " if not r["bad"] else "BAD", {k: ("%.2e" % v if isinstance(v, float) else v) for k, v in r.items()})
        else:
            print(i, "CRASH rc=%d" % p.returncode, tc_cases.CASES[i][-1], (p.stderr or p.stdout)[-600:].replace("\n", " | "))
    except subprocess.TimeoutExpired:
        print(i, "TIMEOUT", tc_cases.CASES[i][-1])
    sys.stdout.flush()
