"""Run the golden parity cases on the GPU (one subprocess each) and print compact reports."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
CASES = ["c1_test_sem", "c2_train_depth_sem", "guided_test_nosem", "c3_train_guided_mapping_sc"]

def one(name):
    import torch, spnerf_b200
    from parity_common import run_case
    g_meta_mode = None
    rep = run_case(name, with_backward=True)
    rep["watchdog"] = int(spnerf_b200._cabi.lib().spnerf_watchdog_code())
    print("REPORT " + json.dumps(rep))

if __name__ == "__main__":
    if len(sys.argv) > 1:
        one(sys.argv[1]); sys.exit(0)
    allr = {}
    for c in CASES:
        try:
            p = subprocess.run([sys.executable, __file__, c], capture_output=True, text=True, timeout=300)
            lines = [l for l in p.stdout.splitlines() if l.startswith("REPORT ")]
            if lines:
                r = json.loads(lines[-1][7:]); allr[c] = r
                print("==", c, "state_ok", r["state_ok"], "keys_ok", r["keys_ok"], "watchdog", r["watchdog"])
                for k, v in r["out"].items(): print("   out", k, v)
                for k, v in r.get("loss", {}).items(): print("   loss", k, v)
                for k, v in r.get("grad", {}).items(): print("   grad", k, {a: (round(b, 6) if isinstance(b, float) else b) for a, b in v.items()})
            else:
                allr[c] = {"rc": p.returncode, "stderr": p.stderr[-3000:]}
                print("==", c, "FAILED rc", p.returncode); print(p.stderr[-3000:])
        except subprocess.TimeoutExpired:
            allr[c] = {"timeout": True}; print("==", c, "TIMEOUT")
        sys.stdout.flush()
    json.dump(allr, open(os.path.join(ROOT, "gpurun_out", "e2e_check.json"), "w"), indent=1)
